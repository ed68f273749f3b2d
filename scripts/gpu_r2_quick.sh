#!/bin/bash
O=gpurun_out/r2_quick; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_trainer.py tests/test_gpu_bn.py -q --timeout 380 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --also '' --no-e2e > $O/bench_default.txt 2>&1; grep '^{"metric"' $O/bench_default.txt | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('default img/s', round(d['value']), 'ms', round(d['ms_per_step'],3))"
