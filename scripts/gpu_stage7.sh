#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_multigpu.py 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python benchmarks/kernel_micro.py --out gpurun_out/kernel_micro.json > gpurun_out/kernel_micro.txt 2>&1; cat gpurun_out/kernel_micro.txt
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 > gpurun_out/bench1.txt 2>&1; tail -1 gpurun_out/bench1.txt | cut -c1-400; tail -1 gpurun_out/bench1.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e', d.get('e2e'))"
