#!/bin/bash
mkdir -p gpurun_out
timeout 300 python benchmarks/kernel_micro.py --out gpurun_out/kernel_micro.json > gpurun_out/kernel_micro.txt 2>&1; cat gpurun_out/kernel_micro.txt
timeout 200 python benchmarks/loader_timing.py > gpurun_out/loader_timing.txt 2>&1; cat gpurun_out/loader_timing.txt
for mode in async_nofire sync_allfire; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gossip_step -s 6 -c 2 -f -o gpurun_out/prof_$mode python benchmarks/kernel_micro.py --mode $mode --iters 3 > gpurun_out/ncu_$mode.txt 2>&1; tail -3 gpurun_out/ncu_$mode.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file gpurun_out/launches_bench1.csv python bench.py --gpus 1 --steps 6 --warmup 3 --no-e2e --no-graph > gpurun_out/ncu_bench.txt 2>&1; tail -2 gpurun_out/ncu_bench.txt
ls -la gpurun_out
