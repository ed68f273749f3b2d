#!/bin/bash
N=4; O=gpurun_out/r2_final4; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 30 --warmup 5 --also bf16 --out $O/bench4.json > $O/bench_default.txt 2>&1
grep '^{"metric"' $O/bench_default.txt | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N=4', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'clk', d.get('clocks',{}).get('samples'))"
