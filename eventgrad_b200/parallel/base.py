"""Communication-backend interface.

A backend owns everything that happens between backward() and the next forward():
the exchange (all-reduce or neighbour gossip), the consensus average and the optimizer
update.  Three implementations:

  p2p        fused sm_100a kernels over peer-mapped memory (the product)       parallel/p2p.py
  nccl/gloo  torch.distributed collectives + eager elementwise ops (baseline /
             GPU-free plumbing)                                                parallel/collective.py
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch


@dataclass
class StepLog:
    """Per-step per-tensor records that back the reference debug files."""
    pass_num: int
    curr_norm: torch.Tensor          # [sz] sender norm
    thres: torch.Tensor              # [sz] threshold used in the test
    fired: torch.Tensor              # [sz] bool
    left_norm: Optional[torch.Tensor] = None
    right_norm: Optional[torch.Tensor] = None
    left_new: Optional[torch.Tensor] = None
    right_new: Optional[torch.Tensor] = None


class CommBackend:
    name = "base"

    def __init__(self, cfg, arena, ring):
        self.cfg = cfg
        self.arena = arena
        self.ring = ring
        self.pass_num = 0
        self.logs: List[StepLog] = []
        self.want_logs = bool(cfg.file_write)

    # -- one training step (after backward) -------------------------------------------
    def step(self) -> None:
        raise NotImplementedError

    # -- end of training ----------------------------------------------------------------
    def final_average(self) -> None:
        raise NotImplementedError

    def num_events(self) -> int:
        return 0

    def total_events(self) -> int:
        return 0

    def bytes_sent(self) -> int:
        return 0

    def set_sparse_init(self, prev, rep_l, rep_r) -> None:
        """spevent only, reference quirk Q8 (spevent.cpp:123-136): start the 'previously sent' copy and the two
        neighbour replicas from the given flat tensors (three more random networks in the reference)
        instead of theta_0.  Must be called before the first step."""
        if self.cfg.algo != "spevent" or not hasattr(self, "prev"):
            raise RuntimeError("set_sparse_init applies to the spevent algorithm only")
        if self.pass_num != 0:
            raise RuntimeError("set_sparse_init must be called before the first step")
        self.prev.copy_(prev.to(self.prev.device))
        self.rep_l.copy_(rep_l.to(self.rep_l.device))
        self.rep_r.copy_(rep_r.to(self.rep_r.device))

    def drain_logs(self) -> List[StepLog]:
        out, self.logs = self.logs, []
        return out

    def synchronize(self) -> None:
        pass

    def state_dict(self) -> Dict:
        return {"pass_num": self.pass_num}

    def load_state_dict(self, sd: Dict) -> None:
        self.pass_num = int(sd.get("pass_num", 0))

    def close(self) -> None:
        pass

    # helpers ---------------------------------------------------------------------------
    @property
    def comm_enabled(self) -> bool:
        """CIFAR programs skip all communication when R == 1 (event.cpp:281); the MNIST event
        program does not (it Puts to itself).  `serial_skip` carries that distinction."""
        if self.ring.world > 1:
            return True
        return not getattr(self.cfg, "dataset", "cifar10") == "cifar10"
