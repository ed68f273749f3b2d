"""EventGraD trigger + adaptive threshold state machine -- specification oracle.

This is the vectorised (one lane per parameter tensor) PyTorch statement of the
per-tensor scalar logic in /root/reference/dmnist/event/event.cpp:325-392 and
/root/reference/dcifar10/event/event.cpp:300-365.  The CUDA kernels in
csrc/gossip.cu implement the same arithmetic on the device; tests compare them to this.

Arithmetic types mirror the reference: all state is fp32 (`calloc(sz, 4)`), the slope
average is accumulated in double (`auto slope_avg = 0.0`) and rounded to fp32 when it
becomes the new threshold.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class TriggerConfig:
    thres_type: int = 1          # 1 adaptive, 0 constant
    horizon: float = 1.0
    constant: float = 0.0
    sent_history: int = 2
    initial_comm_passes: int = 30

    @staticmethod
    def from_train(cfg) -> "TriggerConfig":
        return TriggerConfig(cfg.thres_type, cfg.horizon, cfg.constant,
                             cfg.sent_history, cfg.initial_comm_passes)


class TriggerState:
    """thres, last_norm, last_iter [sz]; slopes [sz, H]; all fp32, zero-initialised
    (event.cpp:202-220)."""

    def __init__(self, n_tensors: int, history: int = 2, device="cpu"):
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.thres = z(n_tensors)
        self.last_norm = z(n_tensors)
        self.last_iter = z(n_tensors)
        self.slopes = z(n_tensors, history)

    def clone(self) -> "TriggerState":
        c = TriggerState.__new__(TriggerState)
        c.thres, c.last_norm = self.thres.clone(), self.last_norm.clone()
        c.last_iter, c.slopes = self.last_iter.clone(), self.slopes.clone()
        return c

    def state_dict(self):
        return {k: getattr(self, k).cpu().clone() for k in ("thres", "last_norm", "last_iter", "slopes")}

    def load_state_dict(self, sd):
        for k in ("thres", "last_norm", "last_iter", "slopes"):
            getattr(self, k).copy_(sd[k].to(getattr(self, k).device))


@torch.no_grad()
def trigger_step(state: TriggerState, curr_norm: torch.Tensor, pass_num: int,
                 cfg: TriggerConfig) -> torch.Tensor:
    """Advance the FSM one training step. Returns the boolean fire mask [sz].
    No host synchronisation: everything stays on curr_norm's device."""
    curr = curr_norm.to(torch.float32)
    value_diff = (curr - state.last_norm).abs()                       # fabs(float)
    iter_diff = torch.tensor(float(pass_num), dtype=torch.float32,
                             device=curr.device) - state.last_iter    # int - float -> float
    if cfg.thres_type == 1:
        state.thres.mul_(float(torch.tensor(cfg.horizon, dtype=torch.float32)))  # thres * (float)horizon
    else:
        state.thres.fill_(float(torch.tensor(cfg.constant, dtype=torch.float32)))
    fire = value_diff >= state.thres
    if pass_num < cfg.initial_comm_passes:
        fire = torch.ones_like(fire)
    new_slope = value_diff / iter_diff
    shifted = torch.cat([state.slopes[:, 1:], new_slope[:, None]], dim=1)
    avg = shifted.double().sum(dim=1) / float(cfg.sent_history)
    state.slopes.copy_(torch.where(fire[:, None], shifted, state.slopes))
    if cfg.thres_type == 1:
        state.thres.copy_(torch.where(fire, avg.float(), state.thres))
    state.last_norm.copy_(torch.where(fire, curr, state.last_norm))
    state.last_iter.copy_(torch.where(fire, torch.full_like(curr, float(pass_num)),
                                      state.last_iter))
    return fire


# ---------------------------------------------------------------------------------
# Functional pieces of the per-step update, shared by the collective backends and the
# single-process simulator (the CUDA kernels fuse all of them).
# ---------------------------------------------------------------------------------
@torch.no_grad()
def mix3_(theta: torch.Tensor, left: torch.Tensor, right: torch.Tensor) -> torch.Tensor:
    """theta <- (theta + L + R)/3 exactly as add_, add_, div_(3)
    (/root/reference/dcifar10/event/event.cpp:459-461)."""
    return theta.add_(left).add_(right).div_(3)


@torch.no_grad()
def sgd_(theta: torch.Tensor, grad: torch.Tensor, mom, lr: float, mu: float) -> None:
    """torch::optim::SGD step, dampening 0, no nesterov, no weight decay
    (dcifar10/event/event.cpp:196-200; zero momentum buffer == first-step clone)."""
    if mu != 0.0:
        mom.mul_(mu).add_(grad)
        theta.add_(mom, alpha=-lr)
    else:
        theta.add_(grad, alpha=-lr)


@torch.no_grad()
def topk_select(theta_flat: torch.Tensor, prev_flat: torch.Tensor, k: int):
    """k indices with the largest |theta - prev| (spevent.cpp:346-349). Ties are broken
    towards the LOWEST index (a deterministic refinement; torch.topk leaves it unspecified).
    Returns (values = theta[idx], idx int64) sorted by (diff desc, index asc)."""
    diff = (theta_flat - prev_flat).abs()
    _, order = torch.sort(diff, descending=True, stable=True)
    idx = order[:k]
    return theta_flat[idx], idx
