"""Single-process, R-virtual-rank simulator of the four algorithms -- the golden model.

Written from the behavioural specification, independent of the
distributed backends: all R ranks' arenas live in one process and "communication" is a
tensor copy, executed with iteration-synchronous semantics.  The CPU/gloo backend, the NCCL
baseline and the fused CUDA kernels are all tested against this.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from ..parallel.arena import TensorTable
from ..parallel.trigger import (TriggerConfig, TriggerState, mix3_, sgd_, topk_select,
                                trigger_step)


class RingSimulator:
    def __init__(self, world: int, theta0: torch.Tensor, table: TensorTable, algo: str,
                 tcfg: Optional[TriggerConfig] = None, lr: float = 1e-2, momentum: float = 0.0,
                 topk_percent: float = 10.0, serial_skip: bool = True, sparse_init=None):
        assert algo in ("cent", "decent", "event", "spevent")
        self.R, self.table, self.algo = world, table, algo
        self.tcfg = tcfg or TriggerConfig()
        self.lr, self.mu = lr, momentum
        self.serial_skip = serial_skip
        self.theta = [theta0.clone() for _ in range(world)]
        self.mom = [torch.zeros_like(theta0) for _ in range(world)]
        self.state = [TriggerState(table.n_tensors, self.tcfg.sent_history) for _ in range(world)]
        self.events = [0] * world
        self.bytes = [0] * world
        self.pass_num = 0
        self.fire_history: List[List[torch.Tensor]] = []
        if algo == "spevent":
            self.k = table.topk_counts(topk_percent)
            # default: everything starts from theta_0; sparse_init=(prev, left, right) mirrors the reference's
            # three extra randomly initialised networks (quirk Q8, identical on every rank: same seed)
            p0, l0, r0 = sparse_init if sparse_init is not None else (theta0, theta0, theta0)
            self.prev = [p0.clone() for _ in range(world)]
            self.rep_l = [l0.clone() for _ in range(world)]
            self.rep_r = [r0.clone() for _ in range(world)]
        elif algo in ("decent", "event"):
            self.inbox_l = [torch.zeros_like(theta0) for _ in range(world)]
            self.inbox_r = [torch.zeros_like(theta0) for _ in range(world)]

    # ------------------------------------------------------------------
    def _flat(self, buf, i):
        t = self.table
        return buf[t.offsets[i]: t.offsets[i] + t.numels[i]]

    def _norms(self, buf):
        return torch.stack([torch.linalg.vector_norm(self._flat(buf, i))
                            for i in range(self.table.n_tensors)]).float()

    def left(self, r):
        return (r - 1) % self.R

    def right(self, r):
        return (r + 1) % self.R

    @torch.no_grad()
    def step(self, grads: Sequence[torch.Tensor], fires: Optional[Sequence[torch.Tensor]] = None,
             norms: Optional[Sequence[torch.Tensor]] = None) -> None:
        """`fires` (optional, one bool mask per rank) overrides the trigger decisions -- used to
        compare the data path of a kernel bit-for-bit under the kernel's own decisions; `norms`
        (optional, one [sz] tensor per rank) replaces the oracle's own norm computation so the FSM
        logic can be compared exactly while norm accuracy is tested separately."""
        R, t = self.R, self.table
        self.pass_num += 1
        if self.algo == "cent":
            g = torch.stack(list(grads)).sum(0) / R if R > 1 else grads[0]
            for r in range(R):
                sgd_(self.theta[r], g, self.mom[r], self.lr, self.mu)
                self.bytes[r] += t.n_elems * 4 if R > 1 else 0
            return
        if R == 1 and self.serial_skip:
            sgd_(self.theta[0], grads[0], self.mom[0], self.lr, self.mu)
            return
        # ---- phase 1: every rank decides and "Puts" (uses pre-mix theta_k) --------------
        if self.algo == "decent":
            fires = [torch.ones(t.n_tensors, dtype=torch.bool) for _ in range(R)]
        elif fires is not None:
            fires = [f.clone().bool().cpu() for f in fires]
        else:
            fires = [trigger_step(self.state[r],
                                  self._norms(self.theta[r]) if norms is None else norms[r].float().cpu(),
                                  self.pass_num, self.tcfg)
                     for r in range(R)]
        self.fire_history.append([f.clone() for f in fires])
        snap = [th.clone() for th in self.theta]
        for r in range(R):
            L, Rn = self.left(r), self.right(r)
            for i in range(t.n_tensors):
                if not fires[r][i]:
                    continue
                if self.algo != "decent":
                    self.events[r] += 2
                src = self._flat(snap[r], i)
                if self.algo == "spevent":
                    pv = self._flat(self.prev[r], i)
                    v, ix = topk_select(src, pv, self.k[i])
                    pv[ix] = v
                    self._flat(self.rep_r[L], i)[ix] = v     # I am my left neighbour's right
                    self._flat(self.rep_l[Rn], i)[ix] = v    # and my right neighbour's left
                    self.bytes[r] += 2 * 2 * self.k[i] * 4
                else:
                    self._flat(self.inbox_r[L], i).copy_(src)
                    self._flat(self.inbox_l[Rn], i).copy_(src)
                    self.bytes[r] += 2 * t.numels[i] * 4
        # ---- phase 2: mix with whatever the inboxes hold, then SGD -----------------------
        for r in range(R):
            if self.algo == "spevent":
                mix3_(self.theta[r], self.rep_l[r], self.rep_r[r])
            else:
                mix3_(self.theta[r], self.inbox_l[r], self.inbox_r[r])
            sgd_(self.theta[r], grads[r], self.mom[r], self.lr, self.mu)

    def final_average(self) -> torch.Tensor:
        return torch.stack(self.theta).sum(0) / self.R

    def total_events(self) -> int:
        return int(sum(self.events))

    def dense_messages(self) -> int:
        """Denominator of 'messages saved': 2 * sz * passes * R."""
        return 2 * self.table.n_tensors * self.pass_num * self.R
