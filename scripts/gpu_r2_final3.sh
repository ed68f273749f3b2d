#!/bin/bash
# final single-GPU pass over the committed code: whole GPU test tier + ncu of the sparse and all-reduce kernels
O=gpurun_out/r2_final3; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -k "not multigpu" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt | cut -c1-300
NCU="ncu --set full --import-source on --clock-control none -f"
timeout 400 $NCU -k regex:sparse_ --launch-skip 18 --launch-count 3 -o $O/sparse python benchmarks/step_kernel_profile.py --algo spevent --topk 1 --steps 3 > $O/ncu_sparse.txt 2>&1; echo "ncu sparse rc=$?"; tail -2 $O/ncu_sparse.txt | cut -c1-160
timeout 400 $NCU -k regex:allreduce_kernel --launch-skip 5 --launch-count 1 -o $O/allreduce python benchmarks/step_kernel_profile.py --algo cent --steps 3 > $O/ncu_allreduce.txt 2>&1; echo "ncu allreduce rc=$?"; tail -2 $O/ncu_allreduce.txt | cut -c1-160
timeout 200 python benchmarks/step_kernel_profile.py --algo spevent --topk 1 --out $O/prof_spevent_1.txt | head -6
timeout 200 python benchmarks/step_kernel_profile.py --algo spevent --topk 10 --out $O/prof_spevent_10.txt | head -6
