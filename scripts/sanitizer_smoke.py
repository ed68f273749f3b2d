#!/usr/bin/env python
"""Tiny end-to-end exercise of every hand-written kernel, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitizer_smoke.py     (shared-memory hazards)

Two virtual ranks on one GPU (ops.local_world) drive the cross-rank protocol of the exchange kernels;
the BN / Linear / augment kernels run on small tensors.  Exits non-zero on any numerical mismatch.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventgrad_b200.config import TrainConfig  # noqa: E402
from eventgrad_b200.engine.simulator import RingSimulator  # noqa: E402
from eventgrad_b200.models import build_model  # noqa: E402
from eventgrad_b200.ops.local_world import LocalWorld  # noqa: E402
from eventgrad_b200.parallel.trigger import TriggerConfig  # noqa: E402


def exchange(algo, **kw):
    cfg = TrainConfig(algo=algo, dataset="mnist", model="cnn2", lr=0.05, momentum=0.9, sync_mode="iter",
                      initial_comm_passes=2, topk_percent=10.0, **kw).validate()
    w = LocalWorld(cfg, 2, lambda: build_model("cnn2"), grid_cap=4, timeout_ns=20_000_000_000)
    t = w.arenas[0].table
    sim = RingSimulator(2, w.arenas[0].theta.cpu(), t, algo, TriggerConfig.from_train(cfg), lr=cfg.lr,
                        momentum=cfg.momentum, topk_percent=10.0, serial_skip=False)
    mask = torch.zeros(t.n_padded, device="cuda")
    for o, n in zip(t.offsets, t.numels):
        mask[o:o + n] = 1
    for s in range(3):
        fires = [be.fire.clone().bool() for be in w.backends] if algo in ("event", "spevent") else None
        g = [torch.randn(t.n_padded, device="cuda", generator=torch.Generator("cuda").manual_seed(10 * s + r)) * 0.05 * mask
             for r in range(2)]
        w.step(g)
        if algo != "cent":
            sim.step([x.cpu() for x in g], fires=fires)
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        if algo != "cent":
            assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), (algo, r)
    if algo != "cent":
        w.final_average()
    w.close()
    print("ok", algo, kw)


def main():
    for algo in ("decent", "event", "spevent", "cent"):
        exchange(algo)
    exchange("decent", overlap_push=True)
    # fused BN (+res)(+relu), both paths
    from eventgrad_b200.ops.bn_act import FusedBNAct, _workspace
    for fused in (0, 1):
        x = torch.randn(4, 128, 8, 8, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        r = torch.randn_like(x)
        _workspace(x.device)["fused"] = fused
        bn = FusedBNAct(128).cuda().train()
        xa = x.clone().requires_grad_(True)
        ra = r.clone().requires_grad_(True)
        y = bn(xa, residual=ra, relu=True)
        y.backward(torch.randn_like(y))
        assert torch.isfinite(xa.grad.float()).all()
    _workspace(x.device)["fused"] = 0
    print("ok bn")
    # tcgen05 linear
    from eventgrad_b200.ops.linear_tc import linear_tc_forward
    xx = torch.randn(200, 80, device="cuda").to(torch.bfloat16)
    ww = (torch.randn(32, 80, device="cuda") * 0.1).to(torch.bfloat16)
    bb = torch.randn(32, device="cuda")
    out = linear_tc_forward(xx, ww, bb, True, torch.float32)
    ref = (xx.float() @ ww.float().t() + bb).relu()
    assert torch.allclose(out, ref, rtol=2e-3, atol=2e-3)
    print("ok linear_tc")
    # decode/augment
    from eventgrad_b200.data.augment import decode_augment_torch, draw_augment_params
    from eventgrad_b200.ops.augment import decode_augment
    u = torch.randint(0, 256, (9, 3, 32, 32), dtype=torch.uint8, device="cuda")
    p = draw_augment_params(9, 4, "cuda")
    assert torch.allclose(decode_augment(u, 1.0, 0.0, 1.0, p), decode_augment_torch(u, 1.0, 0.0, 1.0, p))
    print("ok augment")
    print("SANITIZER_SMOKE_OK")


if __name__ == "__main__":
    main()
