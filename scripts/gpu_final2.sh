#!/bin/bash
mkdir -p gpurun_out
for ov in on off; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2986$((RANDOM%9)) bench.py --gpus 2 --steps 40 --warmup 5 --overlap $ov --no-e2e > gpurun_out/bench2_ov_$ov.txt 2>&1; tail -1 gpurun_out/bench2_ov_$ov.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 overlap=$ov', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms')"
done
