"""R virtual ranks of the p2p backend inside ONE process on ONE GPU.

Each virtual rank has its own window, arena, backend and CUDA stream; "peer-mapped" pointers
are simply the other virtual ranks' device pointers.  The complete cross-rank protocol of the
fused kernels (pushes, per-group release/acquire flags, acks, the all-reduce barriers) runs
unchanged -- the kernels of the R ranks are launched on R streams and spin on each other, so
the persistent grid per rank is capped to keep all R grids co-resident.  This is how the
`gpu`-marked tests validate multi-rank behaviour on the single-GPU test tier, and it is also a
handy deterministic debugging harness.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List

import torch

from ..config import TrainConfig
from ..parallel.arena import ParamArena
from ..parallel.p2p import P2PBackend, preallocate_arena_buffers
from ..parallel.topology import Ring
from ..parallel.window import LocalBootstrap


class LocalWorld:
    def __init__(self, cfg: TrainConfig, world: int, model_fn, device="cuda:0", grid_cap: int = 8,
                 **backend_kw):
        self.cfg, self.world = cfg, world
        self.device = torch.device(device)
        self.boot = LocalBootstrap(world)
        self.models, self.arenas, self.backends, self.streams = [], [], [], []
        for r in range(world):
            env = SimpleNamespace(rank=r, world=world, local_rank=0, device=self.device, backend="local")
            torch.manual_seed(cfg.seed)
            model = model_fn()
            theta, grad, symm = preallocate_arena_buffers(model, cfg, env, bootstrap=self.boot.for_rank(r))
            arena = ParamArena(model, self.device, theta=theta, grad=grad)
            be = P2PBackend(cfg, arena, Ring(r, world), env, symm=symm, grid_cap=grid_cap,
                            defer_connect=True, **backend_kw)
            self.models.append(model)
            self.arenas.append(arena)
            self.backends.append(be)
            self.streams.append(torch.cuda.Stream(self.device))
        torch.cuda.synchronize(self.device)
        for be in self.backends:
            be.connect()
        torch.cuda.synchronize(self.device)

    def step(self, grads: List[torch.Tensor]) -> None:
        """One synchronous training step with externally supplied flat gradients."""
        cur = torch.cuda.current_stream(self.device)
        for r in range(self.world):
            self.arenas[r].grad.copy_(grads[r])
        for st in self.streams:
            st.wait_stream(cur)
        for r in range(self.world):
            with torch.cuda.stream(self.streams[r]):
                self.backends[r].step()
        for st in self.streams:
            cur.wait_stream(st)

    def final_average(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)
        for r in range(self.world):
            with torch.cuda.stream(self.streams[r]):
                self.backends[r].final_average_nocheck()
        for st in self.streams:
            cur.wait_stream(st)
        torch.cuda.synchronize(self.device)
        for be in self.backends:
            be.check_status()

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        for be in self.backends:
            be.close()
