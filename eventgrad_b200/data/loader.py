"""Batch loader: host gather -> pinned staging -> async H2D of raw uint8 -> GPU decode/augment.

Replaces the reference's LibTorch DataLoader + OpenCV decode + per-sample CPU transforms
(/root/reference/dcifar10/event/event.cpp:93-105).  Only raw bytes cross PCIe (3 KB per
CIFAR image instead of 12 KB of fp32), double-buffered on a copy stream so the transfer of
batch k+1 overlaps the training step of batch k.
"""
from __future__ import annotations

from typing import Iterator, Tuple

import torch

from .augment import decode_augment_torch, draw_augment_params
from .sampler import ShardSampler
from .sources import DataSource


class BatchLoader:
    def __init__(self, source: DataSource, sampler: ShardSampler, batch: int, device,
                 augment: bool = False, out_dtype: torch.dtype = torch.float32,
                 channels_last: bool = False, seed: int = 0, prefetch: bool = True, native: str = "auto"):
        self.src, self.sampler, self.batch = source, sampler, batch
        self.device = torch.device(device)
        self.augment, self.out_dtype, self.channels_last = augment, out_dtype, channels_last
        self.cuda = self.device.type == "cuda"
        self.prefetch = prefetch and self.cuda
        self.gen = torch.Generator(device=self.device).manual_seed(seed * 7919 + sampler.rank)
        c, h, w = source.sample_shape
        self._stage = []
        self._slot_ev = [None, None]      # event of the last H2D copy that READ each pinned slot
        if self.cuda:
            for _ in range(2):
                xi = torch.empty(batch, c, h, w, dtype=torch.uint8).pin_memory()
                yi = torch.empty(batch, dtype=torch.int64).pin_memory()
                self._stage.append((xi, yi))
            self.copy_stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes_per_batch = batch * (c * h * w + 8)
        # native host prefetcher (C++ worker thread, csrc/host_loader.h): default on the CPU path, opt-in
        # ("on") on the CUDA path until it has been validated on hardware
        self.native = None
        want_native = (native == "on") or (native == "auto" and not self.cuda)
        if want_native:
            from .native_loader import NativeHostBatches, native_available
            if native_available():
                try:
                    self.native = NativeHostBatches(source, batch, n_slots=4, pinned=self.cuda)
                except TypeError:
                    self.native = None

    def __len__(self) -> int:
        return self.sampler.num_batches(self.batch)

    # ------------------------------------------------------------------
    def _decode(self, x_u8: torch.Tensor) -> torch.Tensor:
        params = draw_augment_params(x_u8.shape[0], 4, x_u8.device, self.gen) if self.augment else None
        if self.cuda:
            from ..ops import augment as aug_op
            x = aug_op.decode_augment(x_u8, self.src.scale, self.src.mean, self.src.std, params,
                                      pad=4, out_dtype=self.out_dtype,
                                      channels_last=self.channels_last)
        else:
            x = decode_augment_torch(x_u8, self.src.scale, self.src.mean, self.src.std, params,
                                     pad=4, out_dtype=self.out_dtype)
            if self.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
        return x

    def _host_batch(self, idx: torch.Tensor, slot: int):
        if not self.cuda:
            return self.src.images[idx], self.src.labels[idx]
        xi, yi = self._stage[slot]
        # the async H2D copy that last read this pinned slot must have EXECUTED before the host overwrites it: a
        # stream-side wait_event is not enough, the host runs ahead of the GPU (no syncs in the step, CUDA graphs)
        if self._slot_ev[slot] is not None:
            self._slot_ev[slot].synchronize()
        n = idx.numel()
        torch.index_select(self.src.images, 0, idx, out=xi[:n])
        torch.index_select(self.src.labels, 0, idx, out=yi[:n])
        return xi[:n], yi[:n]

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        order = self.sampler.indices()
        nb = len(self)
        if not self.cuda:
            if self.native is not None:
                for xs, ys, slot in self.native.batches(order):
                    x, y = self._decode(xs), ys.clone()
                    self.native.release(slot)
                    yield x, y
                return
            for b in range(nb):
                idx = order[b * self.batch:(b + 1) * self.batch]
                x, y = self._host_batch(idx, 0)
                yield self._decode(x), y
            return
        if self.native is not None:
            yield from self._iter_cuda_native(order)
            return
        cur = torch.cuda.current_stream(self.device)
        pending = None

        def issue(b):
            idx = order[b * self.batch:(b + 1) * self.batch]
            slot = b & 1
            xh, yh = self._host_batch(idx, slot)
            with torch.cuda.stream(self.copy_stream):
                xd = xh.to(self.device, non_blocking=True)
                yd = yh.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            self._slot_ev[slot] = ev
            return xd, yd, ev

        if nb:
            pending = issue(0)
        for b in range(nb):
            xd, yd, ev = pending
            cur.wait_event(ev)
            xd.record_stream(cur)
            yd.record_stream(cur)
            x = self._decode(xd)
            if self.prefetch and b + 1 < nb:
                # the pinned slot (b+1)&1 was last read by batch b-1's copy: _host_batch host-waits on its event
                pending = issue(b + 1)
            yield x, yd
            if not self.prefetch and b + 1 < nb:
                pending = issue(b + 1)


    def _iter_cuda_native(self, order):
        """CUDA path fed by the native prefetcher: the worker stages batches into pinned slots ahead of
        time; here only the async H2D + decode are issued.  A slot is handed back once its copy has run."""
        cur = torch.cuda.current_stream(self.device)
        inflight = []                                         # (slot, event) of issued H2D copies
        for xs, ys, slot in self.native.batches(order):
            with torch.cuda.stream(self.copy_stream):
                xd = xs.to(self.device, non_blocking=True)
                yd = ys.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            inflight.append((slot, ev))
            while len(inflight) > 2:                          # keep <= 2 copies pending, recycle older slots
                s0, e0 = inflight.pop(0)
                e0.synchronize()
                self.native.release(s0)
            cur.wait_event(ev)
            xd.record_stream(cur)
            yd.record_stream(cur)
            yield self._decode(xd), yd
        for s0, e0 in inflight:
            e0.synchronize()
            self.native.release(s0)


def eval_batches(source: DataSource, batch: int, device, out_dtype=torch.float32,
                 channels_last: bool = False):
    """Sequential, un-augmented batches for rank-0 evaluation (event.cpp:107-112)."""
    dev = torch.device(device)
    n = len(source)
    for s in range(0, n, batch):
        x = source.images[s:s + batch].to(dev, non_blocking=True)
        y = source.labels[s:s + batch].to(dev, non_blocking=True)
        if dev.type == "cuda":
            from ..ops import augment as aug_op
            xf = aug_op.decode_augment(x, source.scale, source.mean, source.std, None,
                                       out_dtype=out_dtype, channels_last=channels_last)
        else:
            xf = decode_augment_torch(x, source.scale, source.mean, source.std, None, out_dtype=out_dtype)
        yield xf, y
