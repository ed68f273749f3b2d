#!/bin/bash
# round-2, 8 GPUs (charged 8x), QUICK: ring correctness, headline, exchange rows, fused-vs-split A/B, short sweeps.
N=8; O=gpurun_out/r2_multi8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29500
run() { name=$1; shift; port=$((port+1)); timeout 300 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/bench_$name.txt 2>&1
  grep '^{"metric"' $O/bench_$name.txt | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'saved', round(d['comm']['messages_saved'],3), 'ovl', d['config']['overlap_push'], 'dbuf', d['config']['double_buffer'])
except Exception as e: print('$name FAILED', e)
"; }
EGB_TEST_WORLDS=8 timeout 300 python -m pytest tests/test_multigpu.py -q --timeout 250 -k "p2p_vs_simulator or nvls" > $O/pytest_multi8.txt 2>&1; echo "pytest multigpu8 rc=$?"; tail -3 $O/pytest_multi8.txt | cut -c1-300
run default
run fused_dbuf --also bf16 --no-e2e --overlap off
run cent --also bf16 --no-e2e --algo cent
run spevent --also '' --no-e2e --algo spevent --topk 1
port=$((port+1)); timeout 300 $TR --master-port $port benchmarks/exchange_bw.py --iters 30 --out $O/exchange_bw.json > $O/exchange.txt 2>&1; echo "exchange rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/exchange_bw.json"))
    for k,v in d.items():
        if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if 'nvlink_raw' not in a})
except Exception as e: print("exchange parse failed", e)
PY
sweep() { prog=$1; mode=$2; shift 2; port=$((port+1)); timeout 300 $TR --master-port $port benchmarks/message_sweep.py --program $prog --sync-mode $mode "$@" --out $O/sweep_${prog}_${mode}.json > $O/sweep_${prog}_${mode}.txt 2>&1; echo "sweep $prog $mode rc=$?"; grep '^{"program"' $O/sweep_${prog}_${mode}.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['program'], d['sync_mode'], 'h', d['horizon'], 'topk', d['topk_percent'], 'events', d['events_total'], '/', d['dense_messages'], 'saved', round(d['messages_saved'],4), 'acc', d['test_acc'], 'train_s', round(d['train_time_s'],1))
"; }
sweep mnist_event iter --horizons 1.0,0.9
sweep mnist_event async --horizons 1.0
sweep cifar_spevent iter --horizons 1.0 --topk 1,10 --epochs 4
sweep cifar_event iter --horizons 1.0,0.9 --epochs 4
