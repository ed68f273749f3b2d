#!/bin/bash
# First GPU contact: single-GPU kernel numerics (virtual ranks), smoke, short bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
python -c "import torch;print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))" >> gpurun_out/gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 180 -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; tail -5 gpurun_out/smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench1.txt 2>&1; tail -3 gpurun_out/bench1.txt
