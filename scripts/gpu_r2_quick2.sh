#!/bin/bash
O=gpurun_out/r2_quick2; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q --timeout 280 -s -k "bottleneck or whole_resnet" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt | cut -c1-200; grep -a "vs fp64" $O/pytest.txt | grep -v print | cut -c1-260; grep -a "^E  " $O/pytest.txt | head -8 | cut -c1-200
