#!/bin/bash
mkdir -p gpurun_out
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29841 benchmarks/message_sweep.py --program mnist_event --horizons 1.0,0.95,0.9 --out gpurun_out/sweep_mnist_r2.json 2>&1 | grep "^{"
$T --master-port 29842 benchmarks/message_sweep.py --program cifar_event --horizons 1.0,0.9 --out gpurun_out/sweep_cifar_r2.json 2>&1 | grep "^{"
$T --master-port 29843 benchmarks/message_sweep.py --program cifar_spevent --horizons 1.0 --topk 10 --epochs 8 --out gpurun_out/sweep_spevent_r2.json 2>&1 | grep "^{"
EGB_TEST_WORLDS=2 timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
$T --master-port 29844 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench2.txt 2>&1; tail -1 gpurun_out/bench2.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms e2e', d['e2e']['value'])"
