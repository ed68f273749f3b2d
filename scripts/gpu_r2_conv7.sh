#!/bin/bash
# A/B of a conv_tc kernel change: numerics, per-layer timing, headline
O=gpurun_out/r2_conv7; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q --timeout 280 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt | cut -c1-200
timeout 300 python benchmarks/conv_tc_bench.py --batch 256 --no-cudnn --out $O/conv_bench.json > $O/conv_bench.txt 2>&1; cut -c1-300 $O/conv_bench.txt | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --also '' --no-e2e > $O/bench_default.txt 2>&1; grep '^{"metric"' $O/bench_default.txt | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('default img/s', round(d['value']), 'ms', round(d['ms_per_step'],3))"
