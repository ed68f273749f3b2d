"""bf16 shadow weights: Conv2d / Linear whose forward reads a bf16 copy of the fp32 master weight.

The fused step kernel writes the bf16 copy (`GossipParams.shadow`) in the same pass that updates
the fp32 master in the arena, so autocast never launches a cast kernel per weight, and autograd's
weight gradients (bf16, freshly allocated by cuDNN's wgrad) are consumed IN PLACE by the step
kernel through a pointer table -- no AccumulateGrad add, no fp32 grad arena, no zeroing.
The registered nn.Parameters stay the fp32 masters (names / order / state_dict unchanged).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


_SIDE = {}


def _side_stream(device) -> "torch.cuda.Stream":
    st = _SIDE.get(device)
    if st is None:
        st = _SIDE[device] = torch.cuda.Stream(device)
    return st


class _ConvSplitBwd(torch.autograd.Function):
    """EXPERIMENTAL (EGB_CONV_SPLIT_BWD=1): convolution whose backward issues cuDNN's wgrad on a side stream
    while dgrad runs on the main stream.  At the per-GPU batch of the 8-GPU configuration (32 images) each of
    the two kernels under-fills 148 SMs, and inside the whole-step CUDA graph the fork/join becomes two
    parallel branches.  Same library kernels, same arithmetic as the default path."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, dilation, groups):
        ctx.save_for_backward(x, w)
        ctx.conf = (stride, padding, dilation, groups)
        return torch.ops.aten.convolution(x, w, None, stride, padding, dilation, False, [0, 0], groups)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, padding, dilation, groups = ctx.conf
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        bw = torch.ops.aten.convolution_backward
        dx = dw = None
        if need_dx and need_dw:
            cur = torch.cuda.current_stream(dy.device)
            side = _side_stream(dy.device)
            side.wait_stream(cur)                       # dy (and everything before it) is ready
            with torch.cuda.stream(side):
                dw = bw(dy, x, w, None, stride, padding, dilation, False, [0, 0], groups, [False, True, False])[1]
            dx = bw(dy, x, w, None, stride, padding, dilation, False, [0, 0], groups, [True, False, False])[0]
            cur.wait_stream(side)                       # join: later consumers of dw run on `cur`
            dw.record_stream(cur)                       # allocated on `side`, consumed (step kernel) on `cur`
        elif need_dx or need_dw:
            dx, dw, _ = bw(dy, x, w, None, stride, padding, dilation, False, [0, 0], groups, [need_dx, need_dw, False])
        return dx, dw, None, None, None, None


def _split_bwd_enabled() -> bool:
    import os
    return os.environ.get("EGB_CONV_SPLIT_BWD", "0") == "1"


class ShadowConv2d(nn.Conv2d):
    w16 = None
    b16 = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.w16 is not None and x.is_cuda and torch.is_autocast_enabled():
            if self.b16 is None and self.padding_mode == "zeros" and _split_bwd_enabled():
                xb = x if x.dtype == self.w16.dtype else x.to(self.w16.dtype)
                return _ConvSplitBwd.apply(xb, self.w16, list(self.stride), list(self.padding), list(self.dilation),
                                           self.groups)
            return self._conv_forward(x, self.w16, self.b16)
        return super().forward(x)


class ShadowLinear(nn.Linear):
    w16 = None
    b16 = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.w16 is not None and x.is_cuda and torch.is_autocast_enabled():
            return F.linear(x, self.w16, self.b16)
        return super().forward(x)
