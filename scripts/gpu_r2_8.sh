#!/bin/bash
# round-2, 8 GPUs (charged 8x): everything that needs the full ring, each piece under its own timeout.
N=8; O=gpurun_out/r2_multi8; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29500
run() { name=$1; shift; port=$((port+1)); timeout 400 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 "$@" > $O/bench_$name.txt 2>&1
  grep '^{"metric"' $O/bench_$name.txt | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', d['dtype'], 'img/s', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'other', {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d.get('other_dtypes',{}).items()}, 'saved', round(d['comm']['messages_saved'],3), 'ovl', d['config']['overlap_push'], 'dbuf', d['config']['double_buffer'])
except Exception as e: print('$name FAILED', e)
"; }
# 1. correctness on the full ring (subset: each torchrun launch costs ~10 s x 8 GPUs)
EGB_TEST_WORLDS=8 timeout 600 python -m pytest tests/test_multigpu.py -q --timeout 500 -k "p2p_vs_simulator or nvls or double_buffered or overlap_vs_simulator" > $O/pytest_multi8.txt 2>&1; echo "pytest multigpu8 rc=$?"; tail -4 $O/pytest_multi8.txt
# 2. headline (fp32, default switches, + bf16 / tf32 rows, e2e)
run default
# 3. bf16 A/B of the exchange variants + the other programs
run bf16_overlap_ce --dtype bf16 --also '' --no-e2e --ce-push
run bf16_fused_dbuf --dtype bf16 --also '' --no-e2e --overlap off
run bf16_fused_ack --dtype bf16 --also '' --no-e2e --overlap off --no-double-buffer
run bf16_event --dtype bf16 --also '' --no-e2e --algo event
run bf16_spevent --dtype bf16 --also '' --no-e2e --algo spevent
run bf16_cent --dtype bf16 --also '' --no-e2e --algo cent
EGB_NVLS=1 run bf16_cent_nvls --dtype bf16 --also '' --no-e2e --algo cent
run bf16_nccl --dtype bf16 --also '' --no-e2e --impl nccl
run fp32_event --also '' --no-e2e --algo event
# 4. exchange micro-benchmark (incl. NVLS all-reduce rows)
port=$((port+1)); timeout 500 $TR --master-port $port benchmarks/exchange_bw.py --iters 40 --out $O/exchange_bw.json > $O/exchange.txt 2>&1; echo "exchange rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/exchange_bw.json"))
    for k,v in d.items():
        if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print("exchange parse failed", e)
PY
# 5. message sweeps at R=8, full reference schedules (BASELINE configs 2-5)
sweep() { prog=$1; mode=$2; shift 2; port=$((port+1)); timeout 500 $TR --master-port $port benchmarks/message_sweep.py --program $prog --sync-mode $mode "$@" --out $O/sweep_${prog}_${mode}.json > $O/sweep_${prog}_${mode}.txt 2>&1; echo "sweep $prog $mode rc=$?"; grep '^{"program"' $O/sweep_${prog}_${mode}.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['program'], d['sync_mode'], 'h', d['horizon'], 'topk', d['topk_percent'], 'events', d['events_total'], '/', d['dense_messages'], 'saved', round(d['messages_saved'],4), 'acc', d['test_acc'], 'train_s', round(d['train_time_s'],1))
"; }
sweep mnist_event iter --horizons 1.0,0.9
sweep mnist_event async --horizons 1.0,0.9
sweep cifar_event iter --horizons 1.0,0.9
sweep cifar_event async --horizons 1.0,0.9
sweep cifar_spevent iter --horizons 1.0,0.9 --topk 1,10
sweep cifar_spevent async --horizons 1.0 --topk 1,10
# 6. the MNIST programs at the reference's full batch through their CLIs
port=$((port+1)); timeout 300 $TR --master-port $port -m eventgrad_b200.cli.cent > $O/cli_cent.txt 2>&1; echo "cli cent rc=$?"; tail -4 $O/cli_cent.txt | cut -c1-200
port=$((port+1)); timeout 300 $TR --master-port $port -m eventgrad_b200.cli.decent 1 --log-dir $O/logs_decent > $O/cli_decent.txt 2>&1; echo "cli decent rc=$?"; tail -4 $O/cli_decent.txt | cut -c1-200
port=$((port+1)); timeout 300 $TR --master-port $port -m eventgrad_b200.cli.mnist_event 1 1 0.9 --log-dir $O/logs_mnist_event --epochs 2 > $O/cli_mnist_event.txt 2>&1; echo "cli mnist_event file_write rc=$?"; tail -3 $O/cli_mnist_event.txt | cut -c1-200; ls $O/logs_mnist_event 2>/dev/null | head -30 | tr '\n' ' '
# 7. the unmodified reference on this box's CPU, 8 MPI ranks
port=$((port+1)); timeout 900 $TR --master-port $port bench.py --impl reference --gpus $N --steps 4 --warmup 1 > $O/bench_reference.txt 2>&1; tail -1 $O/bench_reference.txt | cut -c1-260
