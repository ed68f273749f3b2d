"""Device-side timing (CUDA events on the launching stream) + max-over-ranks reduction,
replacing the reference's single MPI_Wtime() bracket
(/root/reference/dmnist/event/event.cpp:267, :503-505).  Also optional NVTX ranges."""
from __future__ import annotations

import contextlib
import time
from collections import defaultdict
from typing import Dict

import torch


class PhaseTimer:
    """Accumulates per-phase device time. On CPU falls back to perf_counter."""

    def __init__(self, device, enabled: bool = True):
        self.cuda = torch.device(device).type == "cuda"
        self.enabled = enabled
        self._pending = []
        self.totals: Dict[str, float] = defaultdict(float)
        self.counts: Dict[str, int] = defaultdict(int)

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            yield
            return
        if self.cuda:
            torch.cuda.nvtx.range_push(name)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            try:
                yield
            finally:
                e.record()
                torch.cuda.nvtx.range_pop()
                self._pending.append((name, s, e))
        else:
            t0 = time.perf_counter()
            try:
                yield
            finally:
                self.totals[name] += (time.perf_counter() - t0) * 1e3
                self.counts[name] += 1

    def flush(self) -> None:
        if self.cuda and self._pending:
            torch.cuda.synchronize()
            for name, s, e in self._pending:
                self.totals[name] += s.elapsed_time(e)
                self.counts[name] += 1
            self._pending.clear()

    def summary_ms(self) -> Dict[str, float]:
        self.flush()
        return {k: self.totals[k] / max(1, self.counts[k]) for k in self.totals}
