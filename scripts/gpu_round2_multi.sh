#!/bin/bash
# Round-2, N GPUs:  gpurun --gpus N --timeout 1200 -- 'bash scripts/gpu_round2_multi.sh N [quick|full]'
# An N-GPU call is charged N x its box time, so the default is `quick` (~5 min => ~40 GPU-min at N=8): the headline
# with the final code, its experimental variants, the NCCL baseline and the multi-GPU experimental tests.
# `full` adds refport / cent / NVLS / exchange_bw (~11 min).  Validate the experimental paths at N=2 first.
N=${1:-8}; MODE=${2:-quick}; O=gpurun_out/round2_multi$N; mkdir -p $O
run() { name=$1; shift; timeout 420 env $ENVV python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 298$((30+RANDOM%60)) bench.py --gpus $N --steps 40 --warmup 5 "$@" > $O/bench_$name.txt 2>&1; grep '^{"metric"' $O/bench_$name.txt | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'clk', d.get('clocks'))
except Exception as e: print('$name FAILED', e)
"; }
ENVV="EGB_BN_V2=0" run dpsgd_overlap --overlap on
EGB_TEST_WORLDS=$N EGB_EXPERIMENTAL=1 timeout 420 python -m pytest tests/test_gpu_experimental.py -q --timeout 400 -k "double_buffered_decent_multi_gpu or ce_push_multi_gpu" > $O/exp_multi.txt 2>&1; echo "experimental multi rc=$?"; tail -4 $O/exp_multi.txt
ENVV="EGB_BN_V2=0" run dpsgd_overlap_ce --overlap on --ce-push --no-e2e
ENVV="EGB_BN_V2=0" run dpsgd_fused --overlap off --no-e2e
ENVV="EGB_BN_V2=0" run nccl --impl nccl --no-e2e
[ "$MODE" = "full" ] || exit 0
ENVV="EGB_BN_V2=0" run dpsgd_fused_dbuf --overlap off --double-buffer --no-e2e
ENVV="EGB_BN_V2=0" run refport --impl refport --no-e2e
ENVV="EGB_BN_V2=0" run cent --algo cent --no-e2e
EGB_TEST_WORLDS=$N EGB_EXPERIMENTAL=1 timeout 420 python -m pytest tests/test_gpu_experimental.py -q --timeout 400 -k "nvls" > $O/exp_nvls.txt 2>&1; echo "experimental nvls rc=$?"; tail -4 $O/exp_nvls.txt
ENVV="EGB_NVLS=1" run cent_nvls --algo cent --no-e2e
EGB_EXPERIMENTAL=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29821 benchmarks/exchange_bw.py --iters 40 --out $O/exchange_bw.json > $O/exchange.txt 2>&1; tail -60 $O/exchange.txt | grep -v "^\*\|OMP_NUM\|^$"
