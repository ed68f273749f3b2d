#!/bin/bash
# round-2 multi-GPU call:  gpurun --gpus N -- 'bash scripts/gpu_r2_multi.sh N [quick]'
N=${1:-2}; MODE=${2:-full}; O=gpurun_out/r2_multi$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29900
run() { name=$1; shift; port=$((port+1)); timeout 600 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/bench_$name.txt 2>&1
  grep '^{"metric"' $O/bench_$name.txt | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'clk', (d.get('clocks') or {}).get('samples'), 'ovl', d['config']['overlap_push'], 'dbuf', d['config']['double_buffer'])
except Exception as e: print('$name FAILED', e)
"; }
if [ "$MODE" != "quick" ]; then
  EGB_TEST_WORLDS=$N timeout 1500 python -m pytest tests/test_multigpu.py -q --timeout 900 > $O/pytest_multi.txt 2>&1; echo "pytest multigpu rc=$?"; tail -6 $O/pytest_multi.txt
fi
run default
run bf16_overlap --dtype bf16 --also '' --no-e2e
run bf16_fused_dbuf --dtype bf16 --also '' --no-e2e --overlap off
run bf16_fused_ack --dtype bf16 --also '' --no-e2e --overlap off --no-double-buffer
run bf16_ce --dtype bf16 --also '' --no-e2e --ce-push
run bf16_event --dtype bf16 --also '' --no-e2e --algo event
run bf16_spevent --dtype bf16 --also '' --no-e2e --algo spevent
run bf16_cent --dtype bf16 --also '' --no-e2e --algo cent
EGB_NVLS=1 run bf16_cent_nvls --dtype bf16 --also '' --no-e2e --algo cent
run bf16_nccl --dtype bf16 --also '' --no-e2e --impl nccl
port=$((port+1)); timeout 600 $TR --master-port $port benchmarks/exchange_bw.py --iters 40 --out $O/exchange_bw.json > $O/exchange.txt 2>&1; echo "exchange rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/exchange_bw.json"))
    for k,v in d.items():
        if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print("exchange parse failed", e)
PY
tail -5 $O/exchange.txt | cut -c1-300
if [ "$MODE" != "quick" ]; then
  port=$((port+1)); timeout 900 $TR --master-port $port bench.py --impl reference --gpus $N --steps 4 --warmup 1 > $O/bench_reference.txt 2>&1; tail -1 $O/bench_reference.txt | cut -c1-260
fi
