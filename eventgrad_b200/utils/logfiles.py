"""Reference-compatible debug files, enabled by file_write == 1.

  send<r>.txt   one line per step; per tensor "<curr_norm>,  <thres>,  <1|0>,  "
                (/root/reference/dmnist/event/event.cpp:337-339, :385-391)
  recv<r>.txt   one line per step; per tensor "[<1|0>,  ]<left_norm>,  [<1|0>,  ]<right_norm>,  "
                MNIST writes the flag only when it is 1 (:418-426); CIFAR always
                (/root/reference/dcifar10/event/event.cpp:400-412)
  train<r>.txt  "<pass_num>, <loss>" per step (CIFAR, event.cpp:271-273)
  values<r>.txt "<epoch>, <loss>" (cent / decent, /root/reference/dmnist/decent/decent.cpp:165-167)

Numbers use C++ iostream default formatting (== printf %g).  Unlike the reference, nothing
is written from inside the hot loop: backends buffer per-step records (on the device for the
p2p backend) and the trainer drains them at epoch boundaries.
"""
from __future__ import annotations

import os
from typing import Iterable

from ..parallel.base import StepLog


def _g(x) -> str:
    return f"{float(x):g}"


class RefLogWriter:
    def __init__(self, log_dir: str, rank: int, algo: str, dataset: str, enabled: bool):
        self.enabled = enabled
        self.mnist_style = dataset == "mnist"
        self.fps = self.fpr = self.fpt = self.fpv = None
        if not enabled:
            return
        os.makedirs(log_dir, exist_ok=True)
        op = lambda stem: open(os.path.join(log_dir, f"{stem}{rank}.txt"), "w")
        if algo in ("event", "spevent"):
            self.fps, self.fpr = op("send"), op("recv")
            if dataset == "cifar10":
                self.fpt = op("train")
        else:
            self.fpv = op("values")

    def write_steps(self, logs: Iterable[StepLog]) -> None:
        if not self.enabled or self.fps is None:
            return
        for lg in logs:
            n = lg.curr_norm.numel()
            s = []
            for i in range(n):
                s.append(f"{_g(lg.curr_norm[i])},  {_g(lg.thres[i])},  {1 if bool(lg.fired[i]) else 0},  ")
            self.fps.write("".join(s) + "\n")
            if lg.left_norm is not None:
                r = []
                for i in range(n):
                    for new, val in ((lg.left_new[i], lg.left_norm[i]), (lg.right_new[i], lg.right_norm[i])):
                        if bool(new):
                            r.append("1,  ")
                        elif not self.mnist_style:
                            r.append("0,  ")
                        r.append(f"{_g(val)},  ")
                self.fpr.write("".join(r) + "\n")

    def write_train(self, pass_num: int, loss: float) -> None:
        if self.enabled and self.fpt is not None:
            self.fpt.write(f"{pass_num}, {_g(loss)}\n")

    def write_value(self, epoch: int, loss: float) -> None:
        if self.enabled and self.fpv is not None:
            self.fpv.write(f"{epoch}, {_g(loss)}\n")

    def close(self) -> None:
        for f in (self.fps, self.fpr, self.fpt, self.fpv):
            if f is not None:
                f.close()
