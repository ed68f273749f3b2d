#!/bin/bash
O=gpurun_out/r2_f; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 600 -x -k "spevent or overlap" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt | cut -c1-300
for t in 1 10; do timeout 200 python benchmarks/step_kernel_profile.py --algo spevent --topk $t --out $O/prof_spevent_$t.txt | head -6; done
timeout 200 python benchmarks/exchange_bw.py --iters 20 --only gossip_dense_dbuf,event_async > $O/ex1.txt 2>&1; grep -a "nvlink\|ms\"" $O/ex1.txt | head -12
