#!/bin/bash
# tensor-core fp32 conv integrated into the Trainer: GPU tests, headline bench (+ cuDNN-fp32 / bf16 / tf32 rows), batch-32 proxy, kernel table
O=gpurun_out/r2_conv2; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -k "not multigpu" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.txt | cut -c1-300
show() { grep '^{"metric"' $1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$2', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'own', d['own_kernels_per_step'], 'loss', d.get('loss'))
except Exception as e: print('$2 FAILED', e)
"; }
timeout 600 python bench.py --steps 20 --warmup 5 --profile $O/prof_default.txt > $O/bench_default.txt 2>&1; show $O/bench_default.txt default
timeout 300 python bench.py --steps 20 --warmup 5 --global-batch 32 --no-e2e --also fp32_cudnn > $O/bench_b32.txt 2>&1; show $O/bench_b32.txt b32
head -45 $O/prof_default.txt | cut -c1-60,150-260
