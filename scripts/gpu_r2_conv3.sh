#!/bin/bash
# conv_tc: remaining tests + ncu capture of the four kernel instantiations + per-layer bench at batch 256 and 32
O=gpurun_out/r2_conv3; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_trainer.py -q --timeout 500 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.txt | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3 --launch-skip 4 --launch-count 4 -f -o $O/conv_tc python benchmarks/conv_tc_ncu.py > $O/ncu.txt 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu.txt
timeout 300 python benchmarks/conv_tc_bench.py --batch 32 --no-cudnn --out $O/conv_bench_b32.json > $O/conv_bench_b32.txt 2>&1; cut -c1-420 $O/conv_bench_b32.txt | tail -5
