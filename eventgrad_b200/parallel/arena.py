"""Flat fp32 parameter arena.

Every parameter tensor of a model lives in ONE contiguous fp32 buffer, padded per
tensor to a whole number of tiles (TILE elements) so that each tile belongs to exactly
one tensor.  nn.Parameters (and their .grad) become views into the arena, so a single
fused kernel can stream the whole model: push -> (wait) -> mix -> SGD -> norm-on-write.
The arena layout is identical on all ranks, which is what makes a tensor's offset in a
neighbour's inbox equal to its local offset -- the reference's running displacement
`disp` (/root/reference/dcifar10/event/event.cpp:278, :464).

The sparse (spevent) record layout mirrors the reference window
(/root/reference/dcifar10/spevent/spevent.cpp:161, :368-379, :545): per tensor a
[values(k_i) | indices(k_i)] record at displacement 2*sum_{j<i} k_j, with
k_i = ceil(p/100 * numel_i) (:148).  Indices travel as int32 (Q9) in the same 4 bytes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

TILE = 2048  # elements; must match EG_TILE in csrc/api.h


@dataclass
class TensorTable:
    names: List[str]
    shapes: List[torch.Size]
    numels: List[int]
    offsets: List[int]        # element offset of tensor i in the arena (tile aligned)
    tile_start: List[int]     # first tile of tensor i
    tile_count: List[int]
    n_elems: int              # sum(numels): the reference's num_elem_param
    n_padded: int             # arena length in elements
    n_tiles: int
    tile: int = TILE

    @property
    def n_tensors(self) -> int:
        return len(self.numels)

    def tile_to_tensor(self) -> torch.Tensor:
        out = torch.empty(self.n_tiles, dtype=torch.int32)
        for i, (s, c) in enumerate(zip(self.tile_start, self.tile_count)):
            out[s:s + c] = i
        return out

    def topk_counts(self, percent: float) -> List[int]:
        # spevent.cpp:148  (int)ceil((topk_percent/100.0) * numel)
        return [int(math.ceil((percent / 100.0) * n)) for n in self.numels]

    @staticmethod
    def from_named(named, tile: int = TILE) -> "TensorTable":
        names, shapes, numels, offsets, tstart, tcount = [], [], [], [], [], []
        off = 0
        for n, p in named:
            names.append(n)
            shapes.append(p.shape)
            numels.append(p.numel())
            offsets.append(off)
            nt = max(1, -(-p.numel() // tile))
            tstart.append(off // tile)
            tcount.append(nt)
            off += nt * tile
        return TensorTable(names, shapes, numels, offsets, tstart, tcount,
                           sum(numels), off, off // tile, tile)


class ParamArena:
    """Owns theta / grad / momentum flat buffers and re-points a model's parameters at them."""

    def __init__(self, model: nn.Module, device: torch.device | str = "cpu", tile: int = TILE,
                 theta: Optional[torch.Tensor] = None, grad: Optional[torch.Tensor] = None,
                 with_momentum: bool = True, channels_last: bool = False):
        self.model = model
        self.device = torch.device(device)
        # channels_last: 4-D (conv) parameters are stored physically as [O,kh,kw,I] so that the
        # views handed to cuDNN are already NHWC -- the gossip kernels are layout-agnostic.
        self.channels_last = channels_last
        named = [(n, p) for n, p in model.named_parameters()]
        for n, p in named:
            if p.dtype != torch.float32:
                raise TypeError(f"arena is fp32; parameter {n} is {p.dtype}")
        self.table = TensorTable.from_named(named, tile)
        n = self.table.n_padded

        def _buf(given):
            if given is not None:
                if given.numel() < n or given.dtype != torch.float32:
                    raise ValueError("provided arena buffer too small / wrong dtype")
                b = given.view(-1)[:n]
                b.zero_()
                return b
            return torch.zeros(n, dtype=torch.float32, device=self.device)

        self.theta = _buf(theta)
        self.grad = _buf(grad)
        self.mom = torch.zeros(n, dtype=torch.float32, device=self.device) if with_momentum else None
        self.params: List[nn.Parameter] = []
        self.table_mode = False
        self.compute: List[torch.Tensor] = []
        self.shadow: Optional[torch.Tensor] = None
        with torch.no_grad():
            for i, (_, p) in enumerate(named):
                v = self.view(self.theta, i)
                v.copy_(p.detach().to(self.theta.device))
                p.data = v
                p.grad = self.view(self.grad, i)
                self.params.append(p)
        # BN running stats etc. stay outside the arena (never exchanged: SURVEY.md Q6)
        for b in model.buffers():
            b.data = b.data.to(self.device)

    # ------------------------------------------------------------------
    def enable_table_mode(self, shadow: bool) -> None:
        """Gradient-table mode of the fused step kernel: autograd's gradient tensors are read in
        place, so parameters carry NO pre-set .grad.  With `shadow`, Conv/Linear weights additionally
        get a bf16 leaf copy (view of self.shadow, same physical layout as the master) that the
        forward pass uses; `compute[i]` is the leaf whose .grad belongs to arena tensor i."""
        from ..ops.shadow import ShadowConv2d, ShadowLinear
        self.table_mode = True
        self.compute = list(self.params)
        self.shadow = None
        for p in self.params:
            p.grad = None
        if shadow:
            self.shadow = torch.zeros(self.table.n_padded, dtype=torch.bfloat16, device=self.device)
            idx = {id(p): i for i, p in enumerate(self.params)}
            for m in self.model.modules():
                if isinstance(m, (ShadowConv2d, ShadowLinear)):
                    i = idx[id(m.weight)]
                    m.w16 = self.view(self.shadow, i).requires_grad_(True)
                    self.compute[i] = m.w16
                    self.params[i].requires_grad_(False)
                    if m.bias is not None:
                        j = idx[id(m.bias)]
                        m.b16 = self.view(self.shadow, j).requires_grad_(True)
                        self.compute[j] = m.b16
                        self.params[j].requires_grad_(False)

    def clear_compute_grads(self) -> None:
        for c in self.compute:
            c.grad = None

    # ------------------------------------------------------------------
    def view(self, buf: torch.Tensor, i: int) -> torch.Tensor:
        t = self.table
        flat = buf[t.offsets[i]: t.offsets[i] + t.numels[i]]
        shp = t.shapes[i]
        if self.channels_last and len(shp) == 4:
            o, c, h, w = shp
            return flat.view(o, h, w, c).permute(0, 3, 1, 2)
        return flat.view(shp)

    def pack(self, other: nn.Module) -> torch.Tensor:
        """Flat copy (arena layout, padding lanes zero) of ANOTHER model with the same architecture."""
        out = torch.zeros_like(self.theta)
        named = list(other.named_parameters())
        if [tuple(p.shape) for _, p in named] != [tuple(s) for s in self.table.shapes]:
            raise ValueError("pack(): model does not match the arena's tensor table")
        with torch.no_grad():
            for i, (_, p) in enumerate(named):
                self.view(out, i).copy_(p.detach().to(out.device))
        return out

    def flat(self, buf: torch.Tensor, i: int) -> torch.Tensor:
        t = self.table
        return buf[t.offsets[i]: t.offsets[i] + t.numels[i]]

    def zero_grad(self) -> None:
        self.grad.zero_()

    def tensor_sumsq(self, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Per-tensor sum of squares (float64 accumulate) -- oracle for the norm-on-write path."""
        buf = self.theta if buf is None else buf
        return torch.stack([self.flat(buf, i).double().pow(2).sum()
                            for i in range(self.table.n_tensors)])

    def tensor_norms(self, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        buf = self.theta if buf is None else buf
        return torch.stack([torch.linalg.vector_norm(self.flat(buf, i))
                            for i in range(self.table.n_tensors)]).float()

    def state_dict(self) -> dict:
        return {"theta": self.theta.detach().cpu().clone(),
                "mom": None if self.mom is None else self.mom.detach().cpu().clone(),
                "n_padded": self.table.n_padded, "numels": list(self.table.numels),
                # theta is stored in PHYSICAL layout: conv weights are [O,kh,kw,I] under channels_last
                "channels_last": bool(self.channels_last), "tile": int(TILE)}

    def load_state_dict(self, sd: dict) -> None:
        if list(sd["numels"]) != list(self.table.numels):
            raise ValueError("checkpoint tensor table does not match this model")
        if "channels_last" in sd and (bool(sd["channels_last"]) != bool(self.channels_last)
                                      or int(sd.get("tile", TILE)) != int(TILE)):
            raise ValueError(f"checkpoint was written with channels_last={sd['channels_last']} tile={sd.get('tile')}: "
                             f"the arena stores conv weights in physical layout, resume with the same setting "
                             f"(this run: channels_last={self.channels_last} tile={TILE})")
        self.theta.copy_(sd["theta"].to(self.theta.device))
        if self.mom is not None and sd.get("mom") is not None:
            self.mom.copy_(sd["mom"].to(self.mom.device))
