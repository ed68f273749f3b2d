#!/bin/bash
# 1 GPU: ncu evidence for profiles/ (never a bench value)
mkdir -p gpurun_out
for mode in async_nofire async_allfire sgd_only; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gossip_step -s 6 -c 1 -f -o gpurun_out/prof_gossip_$mode python benchmarks/kernel_micro.py --mode $mode --iters 3 > gpurun_out/ncu_$mode.txt 2>&1; tail -1 gpurun_out/ncu_$mode.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_ -s 400 -c 8 -f -o gpurun_out/prof_bn python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-graph > gpurun_out/ncu_bn.txt 2>&1; tail -1 gpurun_out/ncu_bn.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file gpurun_out/launches_bench1.csv python bench.py --gpus 1 --steps 6 --warmup 3 --no-e2e --no-graph > gpurun_out/ncu_bench.txt 2>&1; tail -1 gpurun_out/ncu_bench.txt | cut -c1-200
ls -la gpurun_out | head -30
