#!/usr/bin/env python
"""Where does the end-to-end step time go? Times the input pipeline pieces on one GPU."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventgrad_b200.data import BatchLoader, ShardSampler, synthetic_source

def t(fn, n=30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

src = synthetic_source("cifar10", 4096).pin()
print("pinned:", src.images.is_pinned(), "threads:", torch.get_num_threads())
res = {}
for B in (128, 256):
    idx = torch.randperm(4096)[:B]
    stage = torch.empty(B, 3, 32, 32, dtype=torch.uint8).pin_memory()
    res[f"index_select_out_B{B}"] = t(lambda: torch.index_select(src.images, 0, idx, out=stage))
    res[f"fancy_index_B{B}"] = t(lambda: src.images[idx])
    flat = src.images.view(4096, -1); sflat = stage.view(B, -1)
    res[f"index_select_flat_B{B}"] = t(lambda: torch.index_select(flat, 0, idx, out=sflat))
    res[f"h2d_B{B}"] = t(lambda: stage.to("cuda", non_blocking=True))
    samp = ShardSampler(4096, 1, 0)
    ld = BatchLoader(src, samp, B, "cuda", augment=True, channels_last=True)
    def one_epoch():
        for x, y in ld: pass
    res[f"loader_per_batch_B{B}"] = t(one_epoch, 3) / len(ld)
print(json.dumps(res, indent=1))
