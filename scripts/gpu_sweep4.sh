#!/bin/bash
mkdir -p gpurun_out
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$T --master-port 29871 benchmarks/message_sweep.py --program cifar_event --horizons 1.0 --out gpurun_out/sweep_cifar_r4.json 2>&1 | grep "^{"
$T --master-port 29872 benchmarks/message_sweep.py --program mnist_event --horizons 1.0,0.9 --out gpurun_out/sweep_mnist_r4.json 2>&1 | grep "^{"
$T --master-port 29873 benchmarks/message_sweep.py --program cifar_spevent --horizons 1.0 --topk 1,10 --epochs 6 --out gpurun_out/sweep_spevent_r4.json 2>&1 | grep "^{"
