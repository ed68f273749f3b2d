#!/bin/bash
O=gpurun_out/r2_d; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_bn.py -q --timeout 600 -x > $O/pytest_bn.txt 2>&1; echo "pytest bn rc=$?"; tail -12 $O/pytest_bn.txt | cut -c1-300
for t in 1 10; do timeout 200 python benchmarks/step_kernel_profile.py --algo spevent --topk $t --out $O/prof_spevent_$t.txt | head -14; done
timeout 200 python benchmarks/step_kernel_profile.py --algo decent --out $O/prof_decent.txt | head -8
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --also '' > $O/bench_fp32.txt 2>&1; tail -1 $O/bench_fp32.txt | cut -c1-420
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --also '' --no-e2e --global-batch 32 > $O/bench_fp32_b32.txt 2>&1; tail -1 $O/bench_fp32_b32.txt | cut -c1-420
timeout 300 python __graft_entry__.py --smoke > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.txt
