"""GPU-side decode + augmentation: uint8 NCHW batch -> normalised float batch, with the
reference's CIFAR training transform ConstantPad(4) -> RandomHorizontalFlip(0.5) ->
RandomCrop(32x32) (/root/reference/dcifar10/event/event.cpp:94-98,
/root/reference/dcifar10/common/transform.hpp:68-101) folded into ONE gather:

    out[b,c,y,x] = padded[b,c, y+oy_b, flip_b ? (W+2p-1) - (x+ox_b) : x+ox_b]

Crop offsets are drawn from randint(0, H_pad - h) which EXCLUDES the maximal offset
(transform.hpp:14-16, :90-101) -- mirrored.  On CUDA this runs as the hand-written kernel
`augment_u8` (csrc/augment.cu); the torch implementation below is the CPU fallback and the
numerics oracle for that kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def draw_augment_params(batch: int, pad: int, device, generator: Optional[torch.Generator] = None):
    """(oy, ox, flip) int32 tensors [B] -- offsets in [0, 2*pad) as the reference draws them."""
    hi = max(1, 2 * pad)
    oy = torch.randint(0, hi, (batch,), device=device, generator=generator, dtype=torch.int32)
    ox = torch.randint(0, hi, (batch,), device=device, generator=generator, dtype=torch.int32)
    flip = (torch.rand(batch, device=device, generator=generator) < 0.5).to(torch.int32)
    return oy, ox, flip


def decode_augment_torch(x_u8: torch.Tensor, scale: float, mean: float, std: float,
                         params: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                         pad: int = 4, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Reference implementation (any device)."""
    B, C, H, W = x_u8.shape
    x = x_u8.to(torch.float32)
    if params is not None:
        oy, ox, flip = [p.long() for p in params]
        ys = torch.arange(H, device=x.device)[None, :] + oy[:, None] - pad          # [B,H] source row
        xs = torch.arange(W, device=x.device)[None, :] + ox[:, None]               # padded column
        xs = torch.where(flip[:, None].bool(), (W + 2 * pad - 1) - xs, xs) - pad   # source column
        vy = (ys >= 0) & (ys < H)
        vx = (xs >= 0) & (xs < W)
        b = torch.arange(B, device=x.device)[:, None, None]
        g = x[b, :, ys.clamp(0, H - 1)[:, :, None], xs.clamp(0, W - 1)[:, None, :]]  # [B,H,W,C]
        g = g * (vy[:, :, None] & vx[:, None, :])[..., None]
        x = g.permute(0, 3, 1, 2)
    x = (x * scale - mean) / std
    return x.to(out_dtype).contiguous()
