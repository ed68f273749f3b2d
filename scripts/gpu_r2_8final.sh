#!/bin/bash
# round-2 final, 8 GPUs (charged 8x, ~200 s of box time left): headline first, then the exchange rows with clean timing, then file_write
N=8; O=gpurun_out/r2_final8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29681 bench.py --gpus $N --steps 30 --warmup 5 --also bf16 --out $O/bench8.json > $O/bench_default.txt 2>&1
grep '^{"metric"' $O/bench_default.txt | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('N=8', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'clk', d.get('clocks',{}).get('samples'))
except Exception as e: print('bench FAILED', e)"
timeout 110 $TR --master-port 29682 benchmarks/exchange_bw.py --iters 30 --out $O/exchange_bw.json > $O/exchange.txt 2>&1; echo "exchange rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/exchange_bw.json"))
    for k,v in d.items():
        if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if 'raw' not in a and 'frac' not in a and 'hbm' not in a and 'grid' not in a})
except Exception as e: print("exchange parse failed", e)
PY
timeout 60 $TR --master-port 29683 -m eventgrad_b200.cli.mnist_event 1 1 0.9 --log-dir $O/logs_mnist_event --epochs 2 > $O/cli_mnist_event.txt 2>&1; echo "cli mnist_event file_write rc=$?"; tail -2 $O/cli_mnist_event.txt | cut -c1-200
python scripts/compare_logs.py profiles/logs_mnist_event_gloo_r8 $O/logs_mnist_event 2>&1 | tail -1 | cut -c1-300
