"""decode_augment: uint8 NCHW batch -> normalised float batch with the reference CIFAR transform
(pad 4, random flip, random crop; /root/reference/dcifar10/common/transform.hpp:68-101) fused
into one gather kernel (csrc/augment.cu).  Oracle: data/augment.py:decode_augment_torch."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ext


def decode_augment(x_u8: torch.Tensor, scale: float, mean: float, std: float,
                   params: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None,
                   pad: int = 4, out_dtype: torch.dtype = torch.float32,
                   channels_last: bool = False) -> torch.Tensor:
    if not x_u8.is_cuda:
        raise RuntimeError("ops.augment.decode_augment is the CUDA path; use decode_augment_torch on CPU")
    assert x_u8.dtype == torch.uint8 and x_u8.dim() == 4 and x_u8.is_contiguous()
    assert out_dtype in (torch.float32, torch.bfloat16)
    B, C, H, W = x_u8.shape
    if channels_last:
        out = torch.empty((B, C, H, W), dtype=out_dtype, device=x_u8.device,
                          memory_format=torch.channels_last)
    else:
        out = torch.empty((B, C, H, W), dtype=out_dtype, device=x_u8.device)
    if params is not None:
        oy, ox, fl = [p.to(torch.int32).contiguous() for p in params]
        po, px, pf = oy.data_ptr(), ox.data_ptr(), fl.data_ptr()
    else:
        po = px = pf = 0
    with torch.cuda.device(x_u8.device):
        ext().decode_augment(x_u8.data_ptr(), out.data_ptr(), po, px, pf, B, C, H, W, pad,
                             float(scale), float(mean), 1.0 / float(std),
                             1 if out_dtype == torch.bfloat16 else 0, 1 if channels_last else 0,
                             torch.cuda.current_stream(x_u8.device).cuda_stream)
    return out
