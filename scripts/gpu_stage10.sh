#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bn.py tests/test_gpu_trainer.py -m gpu -q --timeout 300 -x 2>&1 | tail -8
for gb in 32 256; do
timeout 600 python bench.py --gpus 1 --global-batch $gb --steps 40 --warmup 5 --no-e2e > gpurun_out/bench1_b$gb.txt 2>&1; tail -1 gpurun_out/bench1_b$gb.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gb $gb', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms')"
done
EGB_BN_FUSED_SMALL=0 timeout 600 python bench.py --gpus 1 --global-batch 32 --steps 40 --warmup 5 --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gb 32 split-BN', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms')"
timeout 300 python bench.py --gpus 1 --global-batch 32 --steps 10 --warmup 5 --no-e2e --profile gpurun_out/profile_b32.txt > /dev/null 2>&1; head -14 gpurun_out/profile_b32.txt | cut -c1-58,150-260
