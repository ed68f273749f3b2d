#!/bin/bash
# round-2, 2 GPUs (charged 2x): whole multi-GPU test tier at world 2, headline, exchange rows with clean timing
N=2; O=gpurun_out/r2_final2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
EGB_TEST_WORLDS=2 timeout 900 python -m pytest tests/test_multigpu.py -q --timeout 600 > $O/pytest_multi2.txt 2>&1; echo "pytest multigpu2 rc=$?"; tail -3 $O/pytest_multi2.txt | cut -c1-300
show() { grep '^{"metric"' $1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$2', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'ovl', d['config']['overlap_push'], 'clk', d.get('clocks',{}).get('samples'))
except Exception as e: print('$2 FAILED', e)
"; }
timeout 400 $TR --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 --out $O/bench2.json > $O/bench_default.txt 2>&1; show $O/bench_default.txt default
timeout 300 $TR --master-port 29612 benchmarks/exchange_bw.py --iters 40 --out $O/exchange_bw.json > $O/exchange.txt 2>&1; echo "exchange rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/exchange_bw.json"))
    for k,v in d.items():
        if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if 'raw' not in a and 'frac' not in a})
except Exception as e: print("exchange parse failed", e)
PY
