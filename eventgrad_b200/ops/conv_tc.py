"""3x3 / stride-1 / pad-1 convolution on the tcgen05 tensor cores at fp32 accuracy: front-end of csrc/conv_tc.cu.

The reference's convolutions are fp32 (/root/reference/dcifar10/common/resnet.hpp:3-9, event.cpp:259-276).  fp32 has
no tensor-core path in cuDNN (SIMT kernels, 74 TFLOP/s peak), so the forward, the data gradient and the weight
gradient of every eligible conv run here instead: each fp32 tensor is split into three bf16 planes (x = x0+x1+x2,
24 mantissa bits) and every product is rebuilt from six bf16 tensor-core MMAs accumulated in fp32 -- fp32 storage,
fp32 accuracy (measured against fp64 in tests/test_gpu_conv_tc.py), tensor-core throughput.

Eligibility: CUDA, fp32, channels_last activations AND weights, 3x3 kernel, stride 1, padding 1, dilation 1,
groups 1, no bias, Cin % 64 == 0, Cout % 64 == 0, W in {4, 8, 16, 32, 64} with whole tiles (see conv_tc_supported).
Everything else (the 3-channel stem, strided and 1x1 convs) stays on cuDNN.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

_SM = {}


def _sm(device) -> int:
    n = _SM.get(device)
    if n is None:
        n = _SM[device] = torch.cuda.get_device_properties(device).multi_processor_count
    return n


def enabled() -> bool:
    return os.environ.get("EGB_CONV_TC", "1") != "0"


def split3(x: torch.Tensor) -> torch.Tensor:
    """fp32 tensor (dense in memory, any dim order) -> bf16 planes [3, numel] in the SAME memory order."""
    from . import ext
    n = x.numel()
    assert x.dtype == torch.float32 and n % 8 == 0
    out = torch.empty(3, n, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        ext().split3(x.data_ptr(), out.data_ptr(), n, torch.cuda.current_stream(x.device).cuda_stream)
    return out


def fprop_planes(ap: torch.Tensor, wp: torch.Tensor, N: int, H: int, W: int, Ca: int, Cb: int) -> torch.Tensor:
    """ap: planes of NHWC activations [3, N*H*W*Ca]; wp: planes of [Cb][9][Ca] weights -> fp32 NHWC [N,H,W,Cb]."""
    from . import ext
    y = torch.empty(N, H, W, Cb, dtype=torch.float32, device=ap.device)
    with torch.cuda.device(ap.device):
        ext().conv3x3_fprop(ap.data_ptr(), wp.data_ptr(), y.data_ptr(), N, H, W, Ca, Cb, _sm(ap.device),
                            torch.cuda.current_stream(ap.device).cuda_stream)
    return y


def wgrad_planes(xp: torch.Tensor, gp: torch.Tensor, N: int, H: int, W: int, Ca: int, Cb: int) -> torch.Tensor:
    """xp: planes of the conv input (NHWC, Ca channels); gp: planes of dY (NHWC, Cb channels) -> dW [Cb,3,3,Ca]."""
    from . import ext
    C = ext()
    splits = C.conv_wgrad_splits(N, H, W, Ca, Cb, _sm(xp.device))
    ws = torch.empty(splits, 9 * Ca, Cb, dtype=torch.float32, device=xp.device)
    dw = torch.empty(Cb, 3, 3, Ca, dtype=torch.float32, device=xp.device)
    with torch.cuda.device(xp.device):
        C.conv3x3_wgrad(xp.data_ptr(), gp.data_ptr(), ws.data_ptr(), dw.data_ptr(), N, H, W, Ca, Cb, splits,
                        torch.cuda.current_stream(xp.device).cuda_stream)
    return dw


def eligible(x: torch.Tensor, weight: torch.Tensor, stride, padding, dilation, groups) -> bool:
    if not (enabled() and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 4):
        return False
    if tuple(weight.shape[2:]) != (3, 3) or tuple(stride) != (1, 1) or tuple(padding) != (1, 1) \
            or tuple(dilation) != (1, 1) or groups != 1:
        return False
    if not x.is_contiguous(memory_format=torch.channels_last):
        return False
    from . import ext
    N, C, H, W = x.shape
    return bool(ext().conv_tc_supported(N, H, W, C, weight.shape[0]))


class _Conv3x3TcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight):
        N, Ca, H, W = x.shape
        Cb = weight.shape[0]
        xp = split3(x)                                                  # NHWC order (x is channels_last)
        w_ohwi = weight.permute(0, 2, 3, 1).contiguous()               # no copy when the weight is channels_last
        y = fprop_planes(xp, split3(w_ohwi), N, H, W, Ca, Cb)
        ctx.save_for_backward(xp, w_ohwi)
        ctx.geom = (N, H, W, Ca, Cb)
        return y.permute(0, 3, 1, 2)                                    # logical NCHW, channels_last memory

    @staticmethod
    def backward(ctx, dy):
        xp, w_ohwi = ctx.saved_tensors
        N, H, W, Ca, Cb = ctx.geom
        if dy.dtype != torch.float32 or not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.to(torch.float32).contiguous(memory_format=torch.channels_last)
        gp = split3(dy)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            # dX[p, ci] = sum_{t', co} dY[p + off(t'), co] * W[co, 8 - t', ci]: the forward kernel on flipped,
            # transposed weights [ci][t'][co]
            wt = w_ohwi.reshape(Cb, 9, Ca).flip(1).permute(2, 1, 0).contiguous()
            dx = fprop_planes(gp, split3(wt), N, H, W, Cb, Ca).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dw = wgrad_planes(xp, gp, N, H, W, Ca, Cb).permute(0, 3, 1, 2)   # OIHW logical, OHWI memory
        return dx, dw


def conv3x3_tc(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    return _Conv3x3TcFn.apply(x, weight)


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias, stride, padding, dilation, groups) -> torch.Tensor:
    """F.conv2d with the tensor-core fp32 path for eligible shapes."""
    if bias is None and eligible(x, weight, stride, padding, dilation, groups):
        return conv3x3_tc(x, weight)
    return F.conv2d(x, weight, bias, stride, padding, dilation, groups)
