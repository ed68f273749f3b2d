#!/bin/bash
# Two-GPU stage: real IPC windows over NVLink -- numerics vs simulator, exchange bandwidth, bench N=2.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -40 > gpurun_out/pytest_multigpu.txt
cat gpurun_out/pytest_multigpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 benchmarks/exchange_bw.py --out gpurun_out/exchange_bw_2gpu.json > gpurun_out/exchange2.txt 2>&1; tail -80 gpurun_out/exchange2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2.txt 2>&1; tail -2 gpurun_out/bench2.txt
