#!/bin/bash
# last single-GPU confirmation of the committed tree: GPU test tier, smoke(), headline bench
O=gpurun_out/r2_last; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -k "not multigpu" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; grep -a "\[smoke\]" $O/smoke.txt | cut -c1-260
timeout 600 python bench.py --out $O/bench1.json > $O/bench_default.txt 2>&1; grep '^{"metric"' $O/bench_default.txt | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('default', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'launches', d['gpu_launches'], 'clk', d.get('clocks'))"
