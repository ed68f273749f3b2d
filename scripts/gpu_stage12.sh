#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_linear_tc.py -m gpu -q --timeout 120 -x 2>&1 | tail -25
