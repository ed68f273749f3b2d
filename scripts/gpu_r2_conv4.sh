#!/bin/bash
# conv_tc iteration: conv/bn/trainer tests, headline bench, batch-32 proxy (with kernel table)
O=gpurun_out/r2_conv4; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bn.py tests/test_gpu_trainer.py -q --timeout 500 -s > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt | cut -c1-300; grep -a "resnet18-ref vs fp64" $O/pytest.txt | grep -v print | head -2
show() { grep '^{"metric"' $1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$2', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'own', d['own_kernels_per_step'], 'loss', d.get('loss'))
except Exception as e: print('$2 FAILED', e)
"; }
timeout 600 python bench.py --steps 20 --warmup 5 --also '' --profile $O/prof_default.txt > $O/bench_default.txt 2>&1; show $O/bench_default.txt default
timeout 300 python bench.py --steps 20 --warmup 5 --global-batch 32 --also '' --profile $O/prof_b32.txt > $O/bench_b32.txt 2>&1; show $O/bench_b32.txt b32
head -24 $O/prof_default.txt | cut -c1-60,150-260
head -40 $O/prof_b32.txt | cut -c1-60,150-260
