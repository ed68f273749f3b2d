#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers busy/transient (rc 3)
LOG=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1; rc=$?
  if grep -q "status=transient\|status=busy" "$LOG" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc try=$i" >> "$LOG"
