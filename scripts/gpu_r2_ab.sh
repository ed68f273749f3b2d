#!/bin/bash
# A/B of the exchange micro-benchmark on the SAME box: round-1 tree (baseline/_r1, git-ignored) vs current tree
N=${1:-2}; O=gpurun_out/r2_ab$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(cd baseline/_r1 && timeout 400 $TR --master-port 29931 benchmarks/exchange_bw.py --iters 40 --skip-nccl --out /root/repo/$O/r1.json > /root/repo/$O/r1.txt 2>&1); echo "r1 rc=$?"
timeout 400 $TR --master-port 29932 benchmarks/exchange_bw.py --iters 40 --skip-nccl --out $O/cur.json > $O/cur.txt 2>&1; echo "cur rc=$?"
(cd baseline/_r1 && timeout 400 $TR --master-port 29933 benchmarks/exchange_bw.py --iters 40 --skip-nccl --out /root/repo/$O/r1b.json > /root/repo/$O/r1b.txt 2>&1); echo "r1b rc=$?"
python - <<PY
import json
for n in ("r1","cur","r1b"):
    try:
        d=json.load(open("$O/%s.json"%n))
        print(n, {k:round(v.get("ms",v.get("ms_fused",0)),4) for k,v in d.items() if isinstance(v,dict) and ("ms" in v or "ms_fused" in v)})
    except Exception as e: print(n,"failed",e)
PY
bash scripts/ncu_2gpu_gossip.sh gpurun_out/r2_ab$N/ncu2
