#!/bin/bash
mkdir -p gpurun_out
/usr/bin/time -f "mnist_event world=1 wall %es" timeout 300 python -m eventgrad_b200.cli.mnist_event 0 1 1.0 --epochs 1 --train-samples 6400 --test-samples 1000 --device cuda 2>&1 | tail -4
timeout 500 ncu --set full --clock-control none --import-source on -k regex:bn_ -s 200 -c 8 -f -o gpurun_out/prof_bn2 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-graph > gpurun_out/ncu_bn2.txt 2>&1; tail -1 gpurun_out/ncu_bn2.txt | cut -c1-80
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 6 -c 1 -f -o gpurun_out/prof_linear_tc python benchmarks/linear_tc_bench.py > gpurun_out/ncu_lin.txt 2>&1; tail -1 gpurun_out/ncu_lin.txt | cut -c1-80
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gossip_step -s 6 -c 1 -f -o gpurun_out/prof_gossip_async_nofire2 python benchmarks/kernel_micro.py --mode async_nofire --iters 3 > gpurun_out/ncu_g2.txt 2>&1; tail -1 gpurun_out/ncu_g2.txt | cut -c1-80
