#!/usr/bin/env python
"""Messages-saved measurement -- the reference's only published metric (README.md:4: EventGraD saves
~70 % of messages on MNIST and ~60 % on CIFAR-10 versus dense neighbour gossip).

Runs the real training programs (full reference schedule by default) through the public Trainer on
R GPUs and reports  saved = 1 - total_events / (2 * sz * passes * R)  together with accuracy, for a
sweep of thresholds (adaptive horizon / constant) and, for spevent, top-k percentages.
Data is synthetic (no network), so absolute savings are not comparable digit-for-digit with runs on
the real datasets; the trend against the threshold and the byte accounting are what is measured.

    torchrun --nproc-per-node R benchmarks/message_sweep.py --program mnist_event --epochs 10
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventgrad_b200.config import preset  # noqa: E402
from eventgrad_b200.data import load_source  # noqa: E402
from eventgrad_b200.engine.trainer import Trainer  # noqa: E402
from eventgrad_b200.utils.dist import init_distributed, shutdown  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--program", default="mnist_event", choices=["mnist_event", "cifar_event", "cifar_spevent"])
    ap.add_argument("--epochs", type=int, default=0, help="0 = the reference's epoch count")
    ap.add_argument("--horizons", default="1.0,0.95,0.9")
    ap.add_argument("--topk", default="10")
    ap.add_argument("--sync-mode", default="iter")
    ap.add_argument("--train-samples", type=int, default=0)
    ap.add_argument("--dtype", default="", help="default: bf16 for the CIFAR programs (the 20-epoch schedule x many "
                    "configurations is a GPU-minutes question), fp32 for MNIST")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    env = init_distributed("cuda")
    rows = []
    ds = "mnist" if a.program == "mnist_event" else "cifar10"
    ntr = a.train_samples or (60000 if ds == "mnist" else 50000)
    train = load_source(ds, "synthetic", ntr, True)
    test = load_source(ds, "synthetic", 10000, False)
    topks = [float(x) for x in a.topk.split(",")] if a.program == "cifar_spevent" else [None]
    for hz in [float(x) for x in a.horizons.split(",")]:
        for tk in topks:
            kw = dict(backend="p2p", device="cuda", horizon=hz, thres_type=1, quiet=True, sync_mode=a.sync_mode,
                      train_samples=ntr, dtype=a.dtype or ("bf16" if ds == "cifar10" else "fp32"))
            if a.epochs:
                kw["epochs"] = a.epochs
            if tk is not None:
                kw["topk_percent"] = tk
            cfg = preset(a.program, **kw)
            tr = Trainer(cfg, env, train_source=train, test_source=test)
            t0 = time.perf_counter()
            tr.fit()
            res = tr.finalize(evaluate=True)
            tr.backend.check_status()
            row = {"program": a.program, "world": env.world, "horizon": hz, "topk_percent": tk,
                   "sync_mode": a.sync_mode, "dtype": cfg.dtype,
                   "epochs": cfg.epochs, "steps": res["steps"], "events_total": res["events_total"],
                   "dense_messages": res["dense_messages"], "messages_saved": res["messages_saved"],
                   "bytes_sent_rank0": res["bytes_sent_rank"], "train_acc_last_epoch": getattr(tr, "last_train_acc", None),
                   "test_acc": res.get("test_acc"), "train_time_s": res["train_time_s"],
                   "wall_s": time.perf_counter() - t0}
            rows.append(row)
            if env.rank == 0:
                print(json.dumps(row), flush=True)
            tr.close()
    if env.rank == 0 and a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(rows, open(a.out, "w"), indent=1)
    shutdown()


if __name__ == "__main__":
    main()
