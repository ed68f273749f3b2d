"""Python face of the native host prefetcher (csrc/host_loader.h).

A C++ worker thread gathers each batch of the epoch's index order into a small ring of staging slots
(pinned on GPU runs) while the trainer is busy; `next()` releases the GIL while it waits.  This is the
host half of the input pipeline -- the equivalent of the reference's LibTorch DataLoader worker
(/root/reference/dcifar10/event/event.cpp:93-105) -- the device half is csrc/augment.cu.
"""
from __future__ import annotations

from typing import Iterator, Tuple

import torch

from .sources import DataSource


def native_available() -> bool:
    try:
        from ..ops import ext
        return hasattr(ext(), "HostPrefetcher")
    except Exception:  # noqa: BLE001
        return False


class NativeHostBatches:
    def __init__(self, source: DataSource, batch: int, n_slots: int = 3, pinned: bool = False):
        from ..ops import ext
        self.src, self.batch, self.n_slots = source, int(batch), int(n_slots)
        imgs = source.images
        if not (imgs.dtype == torch.uint8 and imgs.is_contiguous() and source.labels.dtype == torch.int64
                and source.labels.is_contiguous() and not imgs.is_cuda):
            raise TypeError("native loader needs contiguous host uint8 images and int64 labels")
        c, h, w = source.sample_shape
        self.sample_bytes = c * h * w
        mk = (lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()) if pinned else \
            (lambda *s, dtype: torch.empty(*s, dtype=dtype))
        self.slot_x = [mk(batch, c, h, w, dtype=torch.uint8) for _ in range(n_slots)]
        self.slot_y = [mk(batch, dtype=torch.int64) for _ in range(n_slots)]
        self._h = ext().HostPrefetcher(imgs.data_ptr(), source.labels.data_ptr(), len(source), self.sample_bytes,
                                       self.batch, [t.data_ptr() for t in self.slot_x],
                                       [t.data_ptr() for t in self.slot_y])
        self._order = None

    def start_epoch(self, order: torch.Tensor) -> int:
        self._order = order.to(torch.int64).contiguous()          # keep alive: the worker copies it at start
        self._h.start_epoch(self._order.data_ptr(), self._order.numel())
        return int(self._h.num_batches())

    def next(self) -> Tuple[int, int]:
        slot, n = self._h.next()
        if slot < 0 and n < 0:
            raise IndexError("native loader: sample index out of range")
        return int(slot), int(n)

    def release(self, slot: int) -> None:
        self._h.release(int(slot))

    def batches(self, order: torch.Tensor) -> Iterator[Tuple[torch.Tensor, torch.Tensor, int]]:
        """Yield (x_u8[:n], y[:n], slot); the caller must release(slot) when done with the views."""
        nb = self.start_epoch(order)
        for _ in range(nb):
            slot, n = self.next()
            yield self.slot_x[slot][:n], self.slot_y[slot][:n], slot

    def close(self) -> None:
        self._h.stop()
