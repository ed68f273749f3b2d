#!/bin/bash
# halo-reuse forward kernel: numerics, per-layer A/B, headline
O=gpurun_out/r2_conv5; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q --timeout 500 > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt | cut -c1-300
timeout 300 python benchmarks/conv_tc_bench.py --batch 256 --no-cudnn --out $O/conv_bench_halo.json > $O/conv_bench_halo.txt 2>&1; cut -c1-330 $O/conv_bench_halo.txt | tail -4
EGB_CONV_HALO=0 timeout 300 python benchmarks/conv_tc_bench.py --batch 256 --no-cudnn --out $O/conv_bench_nohalo.json > $O/conv_bench_nohalo.txt 2>&1; cut -c1-330 $O/conv_bench_nohalo.txt | tail -4
show() { grep '^{"metric"' $1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$2', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'own', d['own_kernels_per_step'])
except Exception as e: print('$2 FAILED', e)
"; }
timeout 600 python bench.py --steps 20 --warmup 5 --also '' --no-e2e > $O/bench_default.txt 2>&1; show $O/bench_default.txt default
timeout 300 python bench.py --steps 20 --warmup 5 --global-batch 32 --also '' --no-e2e > $O/bench_b32.txt 2>&1; show $O/bench_b32.txt b32
