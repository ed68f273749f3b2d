"""tcgen05 fp32-accuracy 3x3 convolution (csrc/conv_tc.cu) against an fp64 convolution.

The claim under test: the tensor-core path (three bf16 planes per fp32 tensor, six MMAs per product, fp32
accumulation) is as accurate as cuDNN's native fp32 kernels.  Every case computes the TRUTH in fp64, measures
cuDNN-fp32's error against it, and requires ours to be within a small factor of that (and tiny in absolute terms).
Reference op: torch::nn::Conv2d(3x3, stride 1, pad 1, no bias) of /root/reference/dcifar10/common/resnet.hpp:3-9.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)

from eventgrad_b200.ops import conv_tc  # noqa: E402

CASES = [  # N, H, W, Cin, Cout
    (4, 32, 32, 64, 64),
    (4, 16, 16, 128, 128),
    (4, 16, 16, 64, 128),
    (6, 8, 8, 256, 256),
    (16, 4, 4, 512, 512),
    (5, 4, 4, 128, 64),      # 4x4 images, batch not a multiple of the 8-image tile: zero-filled tail
    (3, 8, 8, 64, 64),       # odd number of 2-image tiles
]


def _rel(a, truth):
    return float((a.double() - truth).abs().max() / truth.abs().max())


def _rms(a, truth):
    return float(((a.double() - truth) ** 2).mean().sqrt() / (truth ** 2).mean().sqrt())


def _bias(a, truth):
    """signed shrink towards zero, relative: > 0 means |a| < |truth| on average (tensor-core truncation shows here)"""
    return float(((truth - a.double()) * truth.sign()).mean() / truth.abs().mean())


def _mk(N, H, W, Ca, Cb, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, Ca, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cb, Ca, 3, 3, device="cuda", generator=g) * (2.0 / (9 * Ca)) ** 0.5).contiguous(
        memory_format=torch.channels_last)
    dy = torch.randn(N, Cb, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    return x, w, dy


def test_split3_is_exact_to_2pow_minus_24():
    x = torch.randn(1 << 16, device="cuda") * torch.logspace(-6, 6, 1 << 16, device="cuda")
    p = conv_tc.split3(x).float()
    back = (p[2].double() + p[1].double() + p[0].double())
    err = ((back - x.double()).abs() / x.double().abs().clamp_min(1e-30)).max()
    assert float(err) <= 2.0 ** -23, float(err)
    assert torch.equal(p[0], x.to(torch.bfloat16).float())


@pytest.mark.parametrize("N,H,W,Ca,Cb", CASES)
def test_forward_matches_fp64(N, H, W, Ca, Cb):
    torch.backends.cudnn.allow_tf32 = False
    x, w, _ = _mk(N, H, W, Ca, Cb)
    assert conv_tc.eligible(x, w, (1, 1), (1, 1), (1, 1), 1)
    truth = F.conv2d(x.double(), w.double(), padding=1)
    ours = conv_tc.conv3x3_tc(x, w)
    ref32 = F.conv2d(x, w, padding=1)
    assert ours.shape == ref32.shape and ours.is_contiguous(memory_format=torch.channels_last)
    e_ours, e_ref = _rms(ours, truth), _rms(ref32, truth)
    print(f"fwd {N}x{H}x{W} {Ca}->{Cb}: rms err ours {e_ours:.3e} cudnn-fp32 {e_ref:.3e} | max ours {_rel(ours, truth):.3e} "
          f"cudnn {_rel(ref32, truth):.3e} | shrink ours {_bias(ours, truth):.2e} cudnn {_bias(ref32, truth):.2e}")
    assert e_ours < 3e-7 and e_ours <= 4 * e_ref + 1e-8
    assert _rel(ours, truth) < 2e-6


@pytest.mark.parametrize("N,H,W,Ca,Cb", CASES)
def test_backward_matches_fp64(N, H, W, Ca, Cb):
    torch.backends.cudnn.allow_tf32 = False
    x, w, dy = _mk(N, H, W, Ca, Cb, seed=1)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(xd, wd, padding=1).backward(dy.double())
    x1, w1 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    conv_tc.conv3x3_tc(x1, w1).backward(dy)
    x2, w2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    F.conv2d(x2, w2, padding=1).backward(dy)
    for name, ours, ref32, truth in (("dx", x1.grad, x2.grad, xd.grad), ("dw", w1.grad, w2.grad, wd.grad)):
        e_ours, e_ref = _rms(ours, truth), _rms(ref32, truth)
        print(f"{name} {N}x{H}x{W} {Ca}->{Cb}: rms err ours {e_ours:.3e} cudnn-fp32 {e_ref:.3e} | max ours "
              f"{_rel(ours, truth):.3e} cudnn {_rel(ref32, truth):.3e} | shrink ours {_bias(ours, truth):.2e} "
              f"cudnn {_bias(ref32, truth):.2e}")
        assert e_ours < 5e-7 and e_ours <= 4 * e_ref + 1e-8, name
        assert _rel(ours, truth) < 4e-6, name
    assert w1.grad.stride() == w1.stride()


def test_wgrad_is_bitwise_reproducible():
    x, w, dy = _mk(4, 16, 16, 128, 128, seed=2)
    outs = []
    for _ in range(3):
        w1 = w.clone().requires_grad_(True)
        conv_tc.conv3x3_tc(x, w1).backward(dy)
        outs.append(w1.grad.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_resnet_block_uses_the_tensor_core_conv():
    from eventgrad_b200.models.resnet import BasicBlock
    from eventgrad_b200.ops import ext
    blk = BasicBlock(64, 64).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(8, 64, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
    before = dict(ext().launch_counts())["conv"]
    blk(x).sum().backward()
    assert dict(ext().launch_counts())["conv"] > before
    assert x.grad is not None and torch.isfinite(x.grad).all()
