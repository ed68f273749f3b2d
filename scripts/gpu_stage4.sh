#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 180 -x --deselect tests/test_multigpu.py 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python benchmarks/kernel_micro.py --out gpurun_out/kernel_micro.json > gpurun_out/kernel_micro.txt 2>&1; cat gpurun_out/kernel_micro.txt
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/bench1.txt 2>&1; tail -1 gpurun_out/bench1.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gossip_step -s 6 -c 2 -f -o gpurun_out/prof_async_nofire python benchmarks/kernel_micro.py --mode async_nofire --iters 3 > gpurun_out/ncu_async_nofire.txt 2>&1; tail -2 gpurun_out/ncu_async_nofire.txt
