#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 > gpurun_out/bench1.txt 2>&1; tail -1 gpurun_out/bench1.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), d['e2e']['host_ms'], d['clocks'])"
timeout 120 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3
