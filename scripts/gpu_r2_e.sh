#!/bin/bash
O=gpurun_out/r2_e; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bn.py -q --timeout 600 -x -s > $O/pytest.txt 2>&1; echo "pytest rc=$?"; grep -a "fp32 gradient error\|fp32 50-step\|passed\|failed\|Error" $O/pytest.txt | tail -12 | cut -c1-300
for t in 1 10; do timeout 200 python benchmarks/step_kernel_profile.py --algo spevent --topk $t --out $O/prof_spevent_$t.txt | head -7; done
