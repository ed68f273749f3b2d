#!/bin/bash
# 4 GPUs: N>=4 code path (split-step auto on) with the final code
mkdir -p gpurun_out
EGB_TEST_WORLDS=4 timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 -x -k "p2p_vs_simulator or overlap" 2>&1 | tail -3
for ov in auto off; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2985$((RANDOM%9)) bench.py --gpus 4 --steps 40 --warmup 5 --overlap $ov > gpurun_out/bench4_$ov.txt 2>&1; tail -1 gpurun_out/bench4_$ov.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=4 overlap=$ov', d['config']['overlap_push'], round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']))"
done
