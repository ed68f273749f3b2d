#!/bin/bash
# 2 GPUs: validate split-step overlap + compare bench fused vs overlap
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 240 -x -k "overlap or decent_bitwise" 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -8 > gpurun_out/pytest_multigpu.txt; cat gpurun_out/pytest_multigpu.txt
for ov in off on; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2981$((RANDOM%9)) bench.py --gpus 2 --steps 30 --warmup 5 --overlap $ov --no-e2e > gpurun_out/bench2_overlap_$ov.txt 2>&1; tail -1 gpurun_out/bench2_overlap_$ov.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('overlap', d['config']['overlap_push'], 'value', d['value'], 'ms', d['ms_per_step'])"
done
