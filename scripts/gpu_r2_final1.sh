#!/bin/bash
# round-2 evidence set, ONE GPU: full GPU test tier, smoke(), headline bench (+ cuDNN-fp32 / bf16 / tf32 rows, e2e), the
# batch-32 proxy of the 8-GPU run, per-layer conv timing, ncu captures (conv_tc x4, sparse x3, allreduce)
O=gpurun_out/r2_final1; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -k "not multigpu" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; grep -a "\[smoke\]" $O/smoke.txt | cut -c1-300
show() { grep '^{"metric"' $1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$2', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'own', d['own_kernels_per_step'], 'clk', d.get('clocks',{}).get('sm_mhz'), d.get('clocks',{}).get('samples'))
except Exception as e: print('$2 FAILED', e)
"; }
timeout 600 python bench.py --steps 30 --warmup 5 --out $O/bench1.json --profile $O/prof_default.txt > $O/bench_default.txt 2>&1; show $O/bench_default.txt default
timeout 300 python bench.py --steps 30 --warmup 5 --global-batch 32 --no-e2e --also fp32_cudnn --out $O/bench1_b32.json > $O/bench_b32.txt 2>&1; show $O/bench_b32.txt b32
timeout 300 python benchmarks/conv_tc_bench.py --batch 256 --out $O/conv_bench_b256.json > $O/conv_bench_b256.txt 2>&1; echo "conv bench rc=$?"
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 600 $NCU -k regex:conv3x3 --launch-skip 0 --launch-count 8 -o $O/conv_tc python benchmarks/conv_tc_ncu.py > $O/ncu_conv.txt 2>&1; echo "ncu conv rc=$?"
timeout 400 $NCU -k regex:sparse_ --launch-skip 15 --launch-count 3 -o $O/sparse python benchmarks/exchange_bw.py --iters 6 --only spevent_1 --skip-nccl > $O/ncu_sparse.txt 2>&1; echo "ncu sparse rc=$?"
timeout 400 $NCU -k regex:allreduce_kernel --launch-skip 5 --launch-count 1 -o $O/allreduce python benchmarks/exchange_bw.py --iters 6 --only allreduce_sgd_resnet --skip-nccl > $O/ncu_allreduce.txt 2>&1; echo "ncu allreduce rc=$?"
ls -la $O | cut -c30-120
