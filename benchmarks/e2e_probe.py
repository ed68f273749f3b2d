#!/usr/bin/env python
"""Isolate what makes the end-to-end step slower than the device-timed step (1 GPU)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventgrad_b200.config import preset
from eventgrad_b200.data import synthetic_source
from eventgrad_b200.engine.trainer import Trainer
from eventgrad_b200.utils.dist import init_distributed

env = init_distributed("cuda")
cfg = preset("cifar_event", algo="decent", backend="p2p", device="cuda", dtype="bf16", batch_size=256,
             epochs=10**6, channels_last=True, cuda_graph=True, train_samples=8192, quiet=True)
src = synthetic_source("cifar10", 8192).pin()
tr = Trainer(cfg, env, train_source=src)
it = iter(tr.loader)
pool = [tuple(t.clone() for t in next(it)) for _ in range(4)]
for i in range(6):
    tr.train_step(*pool[i % 4])
torch.cuda.synchronize()
K = 30
res = {"threads": torch.get_num_threads()}

def run(name, get, read):
    host = [torch.zeros(1).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    for i in range(3):
        tr.train_step(*get(i))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        loss = tr.train_step(*get(i))
        if read == "lag":
            host[i & 1].copy_(loss.reshape(1), non_blocking=True); evs[i & 1].record()
            if i: evs[(i - 1) & 1].synchronize()
        elif read == "sync":
            loss.item()
    torch.cuda.synchronize()
    res[name] = (time.perf_counter() - t0) / K * 1e3

state = {"it": iter(tr.loader)}
def from_loader(i):
    try:
        return next(state["it"])
    except StopIteration:
        state["it"] = iter(tr.loader)
        return next(state["it"])

run("A_pool_noread", lambda i: pool[i % 4], None)
run("B_pool_lagread", lambda i: pool[i % 4], "lag")
run("B2_pool_syncread", lambda i: pool[i % 4], "sync")
run("C_loader_noread", from_loader, None)
run("D_loader_lagread", from_loader, "lag")
tr.loader.prefetch = False
state["it"] = iter(tr.loader)
run("E_loader_noprefetch_lagread", from_loader, "lag")
tr.loader.prefetch = True
torch.set_num_threads(2)
state["it"] = iter(tr.loader)
run("F_loader_2threads_lagread", from_loader, "lag")
# loader alone
state["it"] = iter(tr.loader)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(K): from_loader(i)
torch.cuda.synchronize(); res["loader_alone"] = (time.perf_counter() - t0) / K * 1e3
print(json.dumps(res, indent=1))
