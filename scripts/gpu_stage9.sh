#!/bin/bash
# 1 GPU: what does the per-GPU step of the 8-GPU configuration (batch 32) spend its time on?
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --global-batch 32 --steps 40 --warmup 5 --profile gpurun_out/profile_b32.txt > gpurun_out/bench1_b32.txt 2>&1; tail -1 gpurun_out/bench1_b32.txt | cut -c1-300
head -70 gpurun_out/profile_b32.txt | cut -c1-60,150-260
timeout 600 python bench.py --gpus 1 --global-batch 32 --steps 40 --warmup 5 --no-graph --no-e2e > gpurun_out/bench1_b32_nograph.txt 2>&1; tail -1 gpurun_out/bench1_b32_nograph.txt | cut -c1-300
