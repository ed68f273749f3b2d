#!/bin/bash
# first hardware run of the tcgen05 fp32-accuracy conv kernels: numerics vs fp64, then per-layer timing vs cuDNN fp32
O=gpurun_out/r2_conv1; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -s --timeout 280 > $O/pytest.txt 2>&1; echo "pytest rc=$?"
grep -a "rms err\|passed\|failed\|FAILED\|Error\|error" $O/pytest.txt | cut -c1-220 | head -70
timeout 300 python benchmarks/conv_tc_bench.py --batch 256 --out $O/conv_bench_b256.json > $O/conv_bench_b256.txt 2>&1; echo "bench rc=$?"; cut -c1-600 $O/conv_bench_b256.txt | tail -8
