#!/usr/bin/env python
"""Headline benchmark (driver contract).

Metric (BASELINE.json): images/sec, whole job, device-timed, max over ranks, for the CIFAR-10
ResNet D-PSGD training step -- the dcifar10/event configuration of the reference
(/root/reference/dcifar10/event/event.cpp:29-42, :91, :196-200): ResNet{2,2,2,2} exactly as the
reference builds it (12 BasicBlocks, 86 tensors, 17 444 682 parameters), GLOBAL batch 256 split
over the ranks (strong scaling), SGD lr 1e-2 momentum 0.9, ring gossip with both neighbours every
step.  Default algorithm = dense D-PSGD (every tensor pushed every step -- the most communication
the reference ever does); `--algo event|spevent|cent` times the other programs.

    python bench.py --gpus N --steps K --warmup W         (torchrun launches N ranks for N>1)

One JSON line on rank 0.  `value` is timed on the device (CUDA events) over K whole training steps
(forward, backward, fused exchange+average+SGD) with inputs already resident; `e2e` repeats the
measurement through the public Trainer API with, every step, the H2D copy of that step's uint8
batch from pinned host memory, GPU decode/augmentation, and a D2H read of the loss.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference", "nccl", "refport"],
                   help="ours: fused p2p kernels | nccl: batched torch.distributed baseline | refport: reference-"
                        "structured per-tensor port (host sync per tensor) | reference: the unmodified reference C++ programs (baseline/_ref, CPU + shm MPI shim)")
    p.add_argument("--algo", default="dpsgd", choices=["dpsgd", "event", "spevent", "cent"])
    p.add_argument("--model", default="resnet18")
    p.add_argument("--global-batch", type=int, default=256)
    p.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    p.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "tf32"])
    p.add_argument("--sync-mode", default="iter", choices=["iter", "async"])
    p.add_argument("--horizon", type=float, default=1.0)
    p.add_argument("--topk", type=float, default=10.0)
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--overlap", default="auto", choices=["auto", "on", "off"],
                   help="split step: pushes on a side stream overlapping forward/backward "
                        "(auto = on for N >= 2: measured 2.65 vs 2.83 ms/step at N=2, 1.95 vs 2.22 at N=4, "
                        "1.89 vs 2.09 at N=8)")
    p.add_argument("--ce-push", action="store_true",
                   help="experimental: copy-engine push in the split step (decent + overlap)")
    p.add_argument("--double-buffer", action="store_true", help="experimental: ack-free double-buffered decent step")
    p.add_argument("--no-channels-last", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--out", default="")
    p.add_argument("--profile", default="", help="write a torch.profiler kernel table of a few e2e steps here")
    return p.parse_args()


def reference_arm(args):
    # The unmodified reference, built by baseline/build_ref.py into baseline/_ref (shm MPI + OpenCV stand-ins); none of
    # this repo's package is imported on this path.  Details: baseline/ref_arm.py, baseline/README.md.
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline"))
    import ref_arm
    return ref_arm.run(args)


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from eventgrad_b200.config import preset
    from eventgrad_b200.data import synthetic_source
    from eventgrad_b200.engine.trainer import Trainer
    from eventgrad_b200.utils.clocks import ClockSampler
    from eventgrad_b200.utils.dist import barrier, init_distributed, max_over_ranks, shutdown, sum_over_ranks

    if not torch.cuda.is_available():
        print(json.dumps({"impl": args.impl, "error": "bench.py needs a CUDA device (run it through gpurun / on the B200 box)"}))
        return 2
    env = init_distributed("cuda")
    N = env.world
    if N != args.gpus and env.rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={N}", file=sys.stderr)
    gb = args.global_batch if args.scaling == "strong" else args.global_batch * N
    per_rank = max(1, gb // N)
    algo = {"dpsgd": "decent", "event": "event", "spevent": "spevent", "cent": "cent"}[args.algo]
    backend = {"ours": "p2p", "nccl": "nccl", "refport": "refport"}[args.impl]
    steps_needed = args.warmup + args.steps + 2
    n_train = max(per_rank * N * 8, 4096)
    cfg = preset("cifar_event", algo=algo, model=args.model, backend=backend, device="cuda",
                 dtype=args.dtype, batch_size=gb, batch_mode="global", epochs=10 ** 6,
                 sync_mode="iter" if algo == "decent" else args.sync_mode,
                 horizon=args.horizon, topk_percent=args.topk,
                 overlap_push=(args.overlap == "on") or (args.overlap == "auto" and N >= 2 and backend == "p2p"),
                 channels_last=not args.no_channels_last, cuda_graph=not args.no_graph,
                 ce_push=args.ce_push, double_buffer=args.double_buffer,
                 train_samples=n_train, test_samples=256, quiet=True, augment=True)
    src = synthetic_source("cifar10", n_train).pin()
    tr = Trainer(cfg, env, train_source=src)
    dev = env.device
    table = tr.arena.table

    # ---------------- device-timed: inputs resident on the GPU (a rotating pool of batches) -------
    pool = []
    it = iter(tr.loader)
    for _ in range(4):
        x, y = next(it)
        pool.append((x.clone(), y.clone()))
    torch.cuda.synchronize()
    prewarm = max(args.warmup, Trainer.GRAPH_WARMUP_STEPS + 2)   # eager warm-up + graph capture, untimed
    for i in range(prewarm):
        tr.train_step(*pool[i % len(pool)])
    torch.cuda.synchronize()
    barrier(env)
    sampler = ClockSampler(dev.index or 0, 25).start() if env.rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(env)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        tr.train_step(*pool[i % len(pool)])
    e1.record()
    torch.cuda.synchronize()
    barrier(env)
    ms = max_over_ranks(e0.elapsed_time(e1), env)
    clocks = sampler.stop() if sampler is not None else None
    tr.backend.check_status() if hasattr(tr.backend, "check_status") else None
    ms_per_step = ms / args.steps
    value = gb * args.steps / (ms / 1e3)
    loss_dev = float(tr.last_loss)

    # ---------------- end to end through the public API: H2D every step + D2H loss every step ----
    e2e = None
    if not args.no_e2e:
        it = iter(tr.loader)
        for _ in range(args.warmup):
            x, y = next(it)
            tr.train_step(x, y)
        torch.cuda.synchronize()
        barrier(env)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        d2h = 0
        tl = tt = ti = 0.0
        # the step's result (loss) is read back EVERY step through a pinned host buffer; the read of
        # step k is issued right after its kernels are enqueued and consumed after step k+1 has been
        # launched (standard 1-step-lagged async logging), so the D2H never drains the GPU queue
        host_loss = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        evs = [torch.cuda.Event() for _ in range(2)]
        losses = []
        for i in range(args.steps):
            ta = time.perf_counter()
            try:
                x, y = next(it)
            except StopIteration:
                it = iter(tr.loader)
                x, y = next(it)
            tb = time.perf_counter()
            loss = tr.train_step(x, y)
            host_loss[i & 1].copy_(loss.reshape(1), non_blocking=True)     # D2H of this step's result
            evs[i & 1].record()
            tc = time.perf_counter()
            if i > 0:
                evs[(i - 1) & 1].synchronize()
                losses.append(float(host_loss[(i - 1) & 1][0]))
            td = time.perf_counter()
            tl += tb - ta; tt += tc - tb; ti += td - tc
            d2h += loss.element_size()
        evs[(args.steps - 1) & 1].synchronize()
        losses.append(float(host_loss[(args.steps - 1) & 1][0]))
        lv = losses[-1]
        assert len(losses) == args.steps
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier(env)
        ms2 = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3), env)
        c, h, w = src.sample_shape
        e2e = {"value": gb * args.steps / (ms2 / 1e3), "unit": "images/s",
               "ms_per_step": ms2 / args.steps,
               "h2d_bytes_per_step": int(per_rank * (c * h * w + 8) * N),
               "d2h_bytes_per_step": int(d2h / args.steps * N),
               "result_read": "every step, async D2H into pinned memory, consumed one step later",
               "host_ms": {"loader": tl / args.steps * 1e3, "launch": tt / args.steps * 1e3,
                           "result_wait": ti / args.steps * 1e3},
               "last_loss": lv}

    if args.profile and env.rank == 0:
        from torch.profiler import ProfilerActivity, profile
        it = iter(tr.loader)
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(4):
                x, y = next(it)
                loss = tr.train_step(x, y)
                loss.item()
        os.makedirs(os.path.dirname(os.path.abspath(args.profile)), exist_ok=True)
        with open(args.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40))

    # ---------------- communication accounting ----------------------------------------------------
    be = tr.backend
    total_steps = be.pass_num
    bytes_rank = be.bytes_sent()
    bytes_all = sum_over_ranks(bytes_rank, env)
    events_all = sum_over_ranks(be.num_events(), env)
    dense_msgs = 2 * table.n_tensors * total_steps * N
    push_bytes_step = bytes_rank / max(1, total_steps)
    # kernels of THIS repo launched per step inside the timed region: the exchange/update kernel(s) plus
    # the fused BatchNorm kernels (2 forward + 2 backward per BN layer when the bf16 NHWC path is active)
    from eventgrad_b200.ops.bn_act import FusedBNAct
    n_bn = sum(1 for m in tr.model.modules() if isinstance(m, FusedBNAct))
    bn_native = args.dtype == "bf16" and cfg.channels_last and os.environ.get("EGB_FUSED_BN", "1") != "0"
    step_kernels = {"decent": 1, "event": 1, "cent": 1, "spevent": 11}[algo] \
        + (1 if (cfg.overlap_push and algo in ("decent", "event") and N > 1) else 0)
    kern_per_step = (step_kernels if args.impl == "ours" else 0) + (4 * n_bn if bn_native else 0)
    out = {
        "metric": "images/sec, CIFAR-10 ResNet (reference topology) D-PSGD ring gossip training step",
        "value": value, "unit": "images/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "impl": args.impl,
        "config": {"model": f"{args.model}-ref(12 BasicBlocks, 86 tensors, {table.n_elems} params)"
                            if args.model == "resnet18" else args.model,
                   "global_batch": gb, "per_gpu_batch": per_rank, "seq_len": None, "image": "3x32x32",
                   "parallelism": f"dp{N}-ring-gossip" if algo != "cent" else f"dp{N}-allreduce",
                   "algorithm": args.algo, "backend": backend, "sync_mode": cfg.sync_mode, "overlap_push": cfg.overlap_push,
                   "optimizer": "SGD lr=1e-2 momentum=0.9", "cuda_graph": cfg.cuda_graph,
                   "channels_last": cfg.channels_last,
                   # opt-in experimental paths active in this run (all off = the validated default code)
                   "experimental": {k: v for k, v in {
                       "bn_v2": os.environ.get("EGB_BN_V2") == "1", "bn_cluster": os.environ.get("EGB_BN_CLUSTER") == "1",
                       "conv_split_bwd": os.environ.get("EGB_CONV_SPLIT_BWD") == "1",
                       "nvls": os.environ.get("EGB_NVLS") == "1", "ce_push": bool(cfg.ce_push),
                       "double_buffer": bool(cfg.double_buffer)}.items() if v},
                   "l2_policy": "per-step working set (theta,grad,mom,2 inboxes = "
                                f"{5 * table.n_padded * 4 / 1e6:.0f} MB + activations) exceeds the 126 MB L2; "
                                "no explicit flush"},
        "gpu_launches": int(kern_per_step * args.steps),
        "own_kernels_per_step": {"exchange_update": step_kernels if args.impl == "ours" else 0,
                                 "fused_batchnorm": 4 * n_bn if bn_native else 0},
        "comm": {"bytes_pushed_per_step_per_gpu": push_bytes_step,
                 "events_total": events_all, "dense_messages": dense_msgs,
                 "messages_saved": (1.0 - events_all / dense_msgs) if (algo in ("event", "spevent") and dense_msgs and N > 1) else 0.0,
                 "bytes_pushed_total": bytes_all},
        "loss": loss_dev,
    }
    if clocks is not None:
        out["clocks"] = clocks
    if e2e is not None:
        out["e2e"] = e2e
    if env.rank == 0:
        line = json.dumps(out)
        print(line, flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(line + "\n")
    tr.close()
    shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
