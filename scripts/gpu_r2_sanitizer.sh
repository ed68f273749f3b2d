#!/bin/bash
# compute-sanitizer over every hand-written kernel incl. the tcgen05 convolutions (one GPU)
O=gpurun_out/r2_sanitizer; mkdir -p $O
python -m eventgrad_b200.build_ext > $O/build.txt 2>&1
timeout 600 python scripts/sanitizer_smoke.py > $O/plain.txt 2>&1; echo "plain rc=$?"; tail -3 $O/plain.txt | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitizer_smoke.py > $O/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; grep -a "ERROR SUMMARY\|SANITIZER_SMOKE_OK\|ok conv" $O/sanitizer_memcheck.txt | tail -4
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitizer_smoke.py > $O/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; grep -a "RACECHECK SUMMARY\|SANITIZER_SMOKE_OK\|ok conv" $O/sanitizer_racecheck.txt | tail -4
