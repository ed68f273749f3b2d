#!/bin/bash
# 2 ranks on 2 GPUs, rank 0 under ncu: the fused iter-sync gossip step with NVLink byte counters.
# The handshake counters are monotonic and the pushes idempotent, so ncu's kernel replay on rank 0 is benign: the
# replayed passes find the neighbour's flags already set and rewrite the same bytes into its inbox.
O=${1:-gpurun_out/ncu2}; mkdir -p $O
# NOTE: only metrics that fit ONE pass: a second replay pass restores rank 0's memory -- including the flags its
# neighbour wrote during pass 1 -- and the replayed kernel would then spin for flags that never come again.  The peer
# wait is bounded to 2 s anyway (EGB_PEER_TIMEOUT_S) so a surprise replay cannot eat the call.
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29977 WORLD_SIZE=2 LOCAL_WORLD_SIZE=2 EGB_PEER_TIMEOUT_S=2
M="nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum"
for row in gossip_dense_ack_v256 gossip_dense_dbuf; do
  RANK=1 LOCAL_RANK=1 timeout 300 python benchmarks/exchange_bw.py --iters 6 --only $row --skip-nccl > $O/rank1_$row.txt 2>&1 &
  P1=$!
  RANK=0 LOCAL_RANK=0 timeout 300 ncu --clock-control none --metrics $M -k regex:gossip_step_kernel -s 6 -c 3 --csv --log-file $O/ncu_$row.csv \
      python benchmarks/exchange_bw.py --iters 6 --only $row --skip-nccl > $O/rank0_$row.txt 2>&1
  wait $P1
  echo "== $row"; grep -v "^==" $O/ncu_$row.csv | python -c "
import csv,sys
rows=list(csv.DictReader(sys.stdin))
out={}
for r in rows:
    out.setdefault(r.get('ID'),{})[r.get('Metric Name')]=(r.get('Metric Value'),r.get('Metric Unit'))
for k,v in out.items(): print(k, {a:b for a,b in v.items()})
" 2>/dev/null | head -6
done
