#!/bin/bash
# A/B: weight gradient on a side stream concurrent with the data gradient (EGB_CONV_PAR_BWD)
O=gpurun_out/r2_conv6; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
EGB_CONV_PAR_BWD=1 timeout 300 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_trainer.py -q --timeout 280 > $O/pytest.txt 2>&1; echo "pytest(par) rc=$?"; tail -2 $O/pytest.txt | cut -c1-200
show() { grep '^{"metric"' $1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$2', 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3))
except Exception as e: print('$2 FAILED', e)
"; }
for par in 0 1; do
  EGB_CONV_PAR_BWD=$par timeout 200 python bench.py --steps 30 --warmup 5 --global-batch 32 --also '' --no-e2e > $O/b32_$par.txt 2>&1; show $O/b32_$par.txt b32_par$par
  EGB_CONV_PAR_BWD=$par timeout 200 python bench.py --steps 30 --warmup 5 --global-batch 64 --also '' --no-e2e > $O/b64_$par.txt 2>&1; show $O/b64_$par.txt b64_par$par
  EGB_CONV_PAR_BWD=$par timeout 200 python bench.py --steps 20 --warmup 5 --also '' --no-e2e > $O/b256_$par.txt 2>&1; show $O/b256_$par.txt b256_par$par
done
