#!/bin/bash
# Round-2 kick-off on ONE GPU (gpurun --timeout 2400 -- 'bash scripts/gpu_round2_first.sh'), ~15-20 GPU-min.
# 1. default GPU suite (must stay green)  2. every EXPERIMENTAL kernel written blind at the end of round 1
# (each group under its own timeout so a hang cannot eat the call)  3. A/B benches for the ones that pass.
# Everything lands in gpurun_out/round2_first/.
O=gpurun_out/round2_first; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 1200 python -m pytest tests -m "gpu and not multigpu" -x -q --timeout 600 > $O/pytest_default.txt 2>&1; echo "default suite rc=$?"; tail -3 $O/pytest_default.txt
declare -A PASS
# group:timeout(s) -- kernels that wait on mbarriers / cluster barriers have no device-side timeout, so a deadlock
# costs the whole group timeout: keep those short
for kt in bn_v2_forward_backward:300 bn_v2_matches:120 bn_v2_resnet:240 bn_cluster_forward_backward:240 bn_cluster_resnet:240 \
          double_buffered_decent_bitwise:240 ce_push_split:180 conv_split:240 linear_tc_tma:150 native_loader:120 p2p_file_write:180; do
  k=${kt%%:*}; t=${kt##*:}
  EGB_EXPERIMENTAL=1 timeout $t python -m pytest tests/test_gpu_experimental.py -q --timeout $t -k "$k" > $O/exp_$k.txt 2>&1
  rc=$?; PASS[$k]=$rc
  echo "experimental $k rc=$rc : $(tail -1 $O/exp_$k.txt)"
done
ok() { for k in "$@"; do [ "${PASS[$k]}" = "0" ] || return 1; done; return 0; }   # A/B only what passed its tests
bench() { name=$1; shift; timeout 300 env "$@" python bench.py --gpus 1 --steps 30 --warmup 5 $BARGS > $O/bench_$name.txt 2>&1
  grep '^{"metric"' $O/bench_$name.txt | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'clk', d.get('clocks'))
except Exception as e: print('$name FAILED', e)
"; }
bench default EGB_BN_V2=0
ok bn_v2_forward_backward bn_v2_resnet && bench bn_v2 EGB_BN_V2=1
ok bn_v2_forward_backward bn_cluster_forward_backward bn_cluster_resnet && bench bn_v2_cluster EGB_BN_V2=1 EGB_BN_CLUSTER=1
bench default_again EGB_BN_V2=0
# per-GPU batch 32 (the 8-GPU configuration on one GPU): kernel-count bound, where the split conv backward could pay
BARGS="--global-batch 32 --no-e2e"
bench b32_default EGB_CONV_SPLIT_BWD=0
ok conv_split && bench b32_split_bwd EGB_CONV_SPLIT_BWD=1
ok bn_v2_forward_backward bn_v2_resnet && bench b32_bn_v2 EGB_BN_V2=1
ok bn_v2_forward_backward bn_cluster_forward_backward bn_cluster_resnet && bench b32_bn_cluster EGB_BN_V2=1 EGB_BN_CLUSTER=1
ok conv_split bn_v2_forward_backward bn_cluster_forward_backward bn_cluster_resnet && bench b32_all EGB_CONV_SPLIT_BWD=1 EGB_BN_V2=1 EGB_BN_CLUSTER=1
BARGS=""
timeout 300 python benchmarks/linear_tc_bench.py > $O/linear_default.txt 2>&1; tail -5 $O/linear_default.txt
ok linear_tc_tma && EGB_TC_LINEAR=tma timeout 300 python benchmarks/linear_tc_bench.py > $O/linear_tma.txt 2>&1; tail -5 $O/linear_tma.txt
