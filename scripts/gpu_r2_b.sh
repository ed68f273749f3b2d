#!/bin/bash
# round-2 second call (1 GPU): full default GPU tier with the landed kernels + fp32 default, smoke(), headline bench
# (fp32 + bf16/tf32 rows), NCHW/ATen-BN A/B at fp32, reference arm on this box's CPU.
O=gpurun_out/r2_b; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 1500 python -m pytest tests -m "gpu and not multigpu" -q --timeout 600 -x > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest.txt
timeout 300 python __graft_entry__.py --smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.txt 2>&1; echo "bench rc=$?"; tail -1 $O/bench_default.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-channels-last --also '' --no-e2e > $O/bench_fp32_nchw.txt 2>&1; tail -1 $O/bench_fp32_nchw.txt | cut -c1-400
EGB_FUSED_BN=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --also '' --no-e2e > $O/bench_fp32_nhwc_atenbn.txt 2>&1; tail -1 $O/bench_fp32_nhwc_atenbn.txt | cut -c1-400
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --global-batch 32 --also 'bf16' --no-e2e > $O/bench_b32.txt 2>&1; tail -1 $O/bench_b32.txt | cut -c1-600
nproc; lscpu | grep "Model name"
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > $O/bench_reference.txt 2>&1; tail -1 $O/bench_reference.txt | cut -c1-300
