"""Multi-process worker: run one algorithm for a few steps on a real process group and compare
every rank's parameters with the single-process simulator (bitwise in iter-sync mode).

    torchrun --nproc-per-node R tests/dist_worker.py --algo event --backend p2p|nccl|gloo ...
Used by tests/test_gloo.py (CPU, gloo) and tests/test_multigpu.py (NCCL bootstrap + p2p kernels).
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from eventgrad_b200.config import TrainConfig  # noqa: E402
from eventgrad_b200.engine.simulator import RingSimulator  # noqa: E402
from eventgrad_b200.models import build_model  # noqa: E402
from eventgrad_b200.parallel import ParamArena, Ring, make_backend  # noqa: E402
from eventgrad_b200.parallel.trigger import TriggerConfig  # noqa: E402
from eventgrad_b200.utils.dist import init_distributed, shutdown  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algo", default="event")
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--model", default="cnn2")
    ap.add_argument("--dataset", default="mnist")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--momentum", type=float, default=0.9)
    ap.add_argument("--horizon", type=float, default=1.0)
    ap.add_argument("--thres-type", type=int, default=1)
    ap.add_argument("--constant", type=float, default=0.0)
    ap.add_argument("--topk", type=float, default=10.0)
    ap.add_argument("--warm", type=int, default=4)
    ap.add_argument("--sync-mode", default="iter")
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--double-buffer", action=argparse.BooleanOptionalAction, default=None)
    ap.add_argument("--fresh-replicas", action="store_true")
    ap.add_argument("--ce-push", action="store_true")
    a = ap.parse_args()
    dev_pref = "cpu" if (a.backend == "gloo" or os.environ.get("EGB_WORKER_CPU") == "1") else "cuda"
    env = init_distributed(dev_pref)
    cfg = TrainConfig(algo=a.algo, dataset=a.dataset, model=a.model, lr=0.05, momentum=a.momentum,
                      horizon=a.horizon, thres_type=a.thres_type, constant=a.constant,
                      topk_percent=a.topk, initial_comm_passes=a.warm, backend=a.backend,
                      sync_mode=a.sync_mode, overlap_push=a.overlap,
                      double_buffer=a.double_buffer, ce_push=a.ce_push).validate()
    torch.manual_seed(0)
    model = build_model(a.model)
    extra = [build_model(a.model) for _ in range(3)] if a.fresh_replicas else None
    ring = Ring(env.rank, env.world)
    if a.backend == "p2p":
        from eventgrad_b200.parallel.p2p import P2PBackend, preallocate_arena_buffers
        theta, grad, symm = preallocate_arena_buffers(model, cfg, env)
        arena = ParamArena(model, env.device, theta=theta, grad=grad)
        be = P2PBackend(cfg, arena, ring, env, symm=symm, timeout_ns=10_000_000_000)
    else:
        arena = ParamArena(model, env.device)
        be = make_backend(cfg, arena, ring, env)
    t = arena.table
    theta0 = arena.theta.detach().cpu().clone()
    sparse_init = None
    if extra is not None:
        packed = [arena.pack(m) for m in extra]
        be.set_sparse_init(*packed)
        sparse_init = tuple(x.detach().cpu().clone() for x in packed)
    mask = torch.zeros(t.n_padded)
    for o, n in zip(t.offsets, t.numels):
        mask[o:o + n] = 1

    def grad_of(step, rank):
        g = torch.Generator().manual_seed(step * 1000 + rank)
        return torch.randn(t.n_padded, generator=g) * 0.05 * mask

    fires = []
    for s in range(a.steps):
        if a.backend == "p2p" and a.algo in ("event", "spevent"):
            fires.append(be.fire.clone().bool().cpu())
        arena.grad.copy_(grad_of(s, env.rank).to(env.device))
        be.step()
    be.synchronize()
    if hasattr(be, "check_status"):
        be.check_status()
    # ---- gather and compare on rank 0 -----------------------------------------------------------
    W = env.world
    mine = arena.theta.detach().clone()
    allth = [torch.empty_like(mine) for _ in range(W)]
    if W > 1:
        dist.all_gather(allth, mine)
    else:
        allth = [mine]
    ev = torch.tensor([be.num_events(), be.bytes_sent()], dtype=torch.int64, device=env.device)
    allev = [torch.empty_like(ev) for _ in range(W)]
    if W > 1:
        dist.all_gather(allev, ev)
    else:
        allev = [ev]
    allf = None
    if fires:
        f = torch.stack(fires).to(torch.uint8).to(env.device)
        allf = [torch.empty_like(f) for _ in range(W)]
        if W > 1:
            dist.all_gather(allf, f)
        else:
            allf = [f]
    ok = True
    if env.rank == 0:
        sim = RingSimulator(W, theta0, t, a.algo, TriggerConfig.from_train(cfg), lr=cfg.lr, momentum=cfg.momentum,
                            topk_percent=a.topk, serial_skip=(a.dataset == "cifar10"), sparse_init=sparse_init)
        for s in range(a.steps):
            fo = [allf[r][s].bool().cpu() for r in range(W)] if allf is not None else None
            sim.step([grad_of(s, r) for r in range(W)], fires=fo)
        for r in range(W):
            got = allth[r].cpu()
            if a.sync_mode == "async":
                # reference RMA semantics: neighbours' values may be one step stale or newer, so the
                # trajectory is timing dependent -- only sanity is checked (finite, near the oracle)
                same = bool(torch.isfinite(got).all()) and float((got - sim.theta[r]).abs().max()) < 0.5
            elif a.algo != "cent" and a.backend != "nccl":
                same = torch.equal(got, sim.theta[r])       # (CUDA eager div_(3) multiplies by 1/3: nccl is ~1 ulp off)
            else:
                same = torch.allclose(got, sim.theta[r], rtol=1e-5, atol=1e-6)
            if not same:
                ok = False
                print(f"MISMATCH rank {r}: max abs diff {(got - sim.theta[r]).abs().max().item():.3e}")
            if a.algo in ("event", "spevent") and a.sync_mode == "iter":
                if int(allev[r][0]) != sim.events[r]:
                    ok = False
                    print(f"EVENTS rank {r}: {int(allev[r][0])} vs {sim.events[r]}")
                if int(allev[r][1]) != sim.bytes[r]:
                    ok = False
                    print(f"BYTES rank {r}: {int(allev[r][1])} vs {sim.bytes[r]}")
        print("WORKER_OK" if ok else "WORKER_FAIL", f"algo={a.algo} backend={a.backend} world={W} "
              f"events={sum(int(e[0]) for e in allev)} dense={sim.dense_messages()} "
              f"nvls={int(getattr(be, 'nvls', False))} nvls_step={int(getattr(be, 'nvls_step', False))} "
              f"dbuf={int(getattr(be, 'dbuf', False))} ce_push={int(getattr(be, 'ce_push', False))}", flush=True)
    # final averaging must agree across ranks
    if a.algo != "cent":
        be.final_average()
        be.synchronize()
        avg = arena.theta.detach().clone()
        alla = [torch.empty_like(avg) for _ in range(W)]
        if W > 1:
            dist.all_gather(alla, avg)
            if env.rank == 0:
                ref = torch.stack([x.cpu() for x in allth]).double().mean(0).float()
                for r in range(W):
                    if not torch.allclose(alla[r].cpu(), ref, rtol=1e-5, atol=1e-6):
                        ok = False
                        print(f"FINAL_AVG mismatch rank {r}")
                print("FINAL_OK" if ok else "FINAL_FAIL", flush=True)
    be.close()
    shutdown()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
