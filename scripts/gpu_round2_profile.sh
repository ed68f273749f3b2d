#!/bin/bash
# Round-2 ncu captures (ONE GPU; run only for the experimental kernels that PASSED scripts/gpu_round2_first.sh):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_profile.sh'
# then here:  python scripts/ncu_summary.py gpurun_out/prof_r2_*.ncu-rep > profiles/round2_ncu_raw.md
# (numbers printed by a run under ncu are never bench values).
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
# BN v2 (bit-mask backward) and cluster BN inside the real step, eager so every kernel is its own launch
EGB_BN_V2=1 timeout 500 $NCU -k regex:bn2_ -s 120 -c 8 -o gpurun_out/prof_r2_bn_v2 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-graph > gpurun_out/ncu_r2_bn_v2.txt 2>&1; tail -1 gpurun_out/ncu_r2_bn_v2.txt | cut -c1-80
EGB_BN_V2=1 EGB_BN_CLUSTER=1 timeout 500 $NCU -k regex:bn_cluster -s 40 -c 8 -o gpurun_out/prof_r2_bn_cluster python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-graph --global-batch 32 > gpurun_out/ncu_r2_bn_cluster.txt 2>&1; tail -1 gpurun_out/ncu_r2_bn_cluster.txt | cut -c1-80
# TMA Linear on the full-batch MNIST GEMM
EGB_TC_LINEAR=tma timeout 300 $NCU -k regex:linear_tma -s 6 -c 1 -o gpurun_out/prof_r2_linear_tma python benchmarks/linear_tc_bench.py > gpurun_out/ncu_r2_lin.txt 2>&1; tail -1 gpurun_out/ncu_r2_lin.txt | cut -c1-80
# default kernels again for the before/after table
timeout 500 $NCU -k regex:bn_bwd -s 120 -c 4 -o gpurun_out/prof_r2_bn_default python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e --no-graph > gpurun_out/ncu_r2_bn_default.txt 2>&1; tail -1 gpurun_out/ncu_r2_bn_default.txt | cut -c1-80
