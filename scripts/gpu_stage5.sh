#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --deselect tests/test_multigpu.py 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python benchmarks/kernel_micro.py --out gpurun_out/kernel_micro.json > gpurun_out/kernel_micro.txt 2>&1; cat gpurun_out/kernel_micro.txt
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --profile gpurun_out/e2e_profile.txt > gpurun_out/bench1.txt 2>&1; tail -1 gpurun_out/bench1.txt
EGB_FUSED_BN=0 timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-e2e > gpurun_out/bench1_nofusedbn.txt 2>&1; tail -1 gpurun_out/bench1_nofusedbn.txt
head -60 gpurun_out/e2e_profile.txt
