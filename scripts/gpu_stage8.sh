#!/bin/bash
# 8 GPUs: correctness at full ring size, exchange bandwidth, headline bench variants
mkdir -p gpurun_out
EGB_TEST_WORLDS=8 timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 -x -k "p2p_vs_simulator or overlap" 2>&1 | tail -6 > gpurun_out/pytest_multigpu8.txt; cat gpurun_out/pytest_multigpu8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29821 benchmarks/exchange_bw.py --iters 40 --out gpurun_out/exchange_bw_8gpu.json > gpurun_out/exchange8.txt 2>&1; tail -120 gpurun_out/exchange8.txt | grep -v "^\*\|OMP_NUM\|^$"
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 298$((30+RANDOM%60)) bench.py --gpus 8 --steps 40 --warmup 5 "$@" > gpurun_out/bench8_$name.txt 2>&1; tail -1 gpurun_out/bench8_$name.txt | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'saved', d['comm']['messages_saved'], 'clk', d.get('clocks'))
except Exception as e: print('$name FAILED', e)
"; }
run dpsgd_fused --overlap off
run dpsgd_overlap --overlap on --no-e2e
run event --algo event --overlap off --no-e2e
run nccl --impl nccl --no-e2e
