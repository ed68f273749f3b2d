#!/bin/bash
O=gpurun_out/r2_c; mkdir -p $O
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --no-channels-last --also '' --no-graph --profile $O/prof_fp32_nchw.txt > $O/b1.txt 2>&1; tail -1 $O/b1.txt | cut -c1-300
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --also '' --no-graph --profile $O/prof_fp32_nhwc.txt > $O/b2.txt 2>&1; tail -1 $O/b2.txt | cut -c1-300
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --no-channels-last --global-batch 32 --also '' --no-graph --profile $O/prof_fp32_nchw_b32.txt > $O/b3.txt 2>&1; tail -1 $O/b3.txt | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_trainer.py -q --timeout 600 -x > $O/pytest_sparse.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_sparse.txt
