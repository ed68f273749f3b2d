"""CPU check of the MATH used by the experimental csrc/bn_act_v2.cu (the kernels themselves need a
GPU: tests/test_gpu_experimental.py).  Emulates, with NumPy, exactly what the v2 kernels compute --
the [slice][row][8-byte] bit-mask layout, dz = dy & mask on packed pairs, dgamma = invstd * sum dz*(x-mean),
dx = a*dz + b*x + c -- and compares with PyTorch autograd of bn -> (+res) -> relu."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _emulate(x, res, gamma, beta, dy, eps, relu):
    M, C = x.shape
    mean = x.mean(0)
    var = x.var(0)
    invstd = 1.0 / np.sqrt(var + eps)
    sc = gamma * invstd
    sh = beta - mean * sc
    v = x * sc + sh + (res if res is not None else 0.0)
    y = np.maximum(v, 0.0) if relu else v
    slices = C // 64
    mask = np.zeros((slices, M, 8), dtype=np.uint8)          # byte (slice,row,tx): bit e = channel slice*64+tx*8+e
    if relu:
        for s in range(slices):
            for tx in range(8):
                for e in range(8):
                    mask[s, :, tx] |= ((v[:, s * 64 + tx * 8 + e] > 0).astype(np.uint8) << e)
    # backward: unpack the mask exactly like apply_mask()
    dz = np.array(dy)
    if relu:
        for s in range(slices):
            for tx in range(8):
                for e in range(8):
                    keep = (mask[s, :, tx] >> e) & 1
                    dz[:, s * 64 + tx * 8 + e] *= keep
    s1 = dz.sum(0)
    s2 = (dz * (x - mean)).sum(0)
    dbeta, dgamma = s1, s2 * invstd
    k1, k2 = dbeta / M, dgamma / M
    a = gamma * invstd
    b = -a * invstd * k2
    c = -a * (k1 - mean * invstd * k2)
    dx = a * dz + b * x + c
    return y, dx, dgamma, dbeta, dz


@pytest.mark.parametrize("relu,has_res", [(True, True), (True, False), (False, False)])
def test_v2_math_matches_autograd(relu, has_res):
    g = torch.Generator().manual_seed(3)
    M, C = 96, 128
    x = (torch.randn(M, C, generator=g, dtype=torch.float64) * 1.7 + 0.4).requires_grad_()
    res = torch.randn(M, C, generator=g, dtype=torch.float64).requires_grad_() if has_res else None
    gamma = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).requires_grad_()
    beta = torch.randn(C, generator=g, dtype=torch.float64).requires_grad_()
    dy = torch.randn(M, C, generator=g, dtype=torch.float64)
    eps = 1e-5
    y = F.batch_norm(x, None, None, gamma, beta, True, 0.1, eps)
    if has_res:
        y = y + res
    if relu:
        y = F.relu(y)
    y.backward(dy)
    ye, dxe, dge, dbe, dze = _emulate(x.detach().numpy(), res.detach().numpy() if has_res else None,
                                      gamma.detach().numpy(), beta.detach().numpy(), dy.numpy(), eps, relu)
    np.testing.assert_allclose(ye, y.detach().numpy(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(dxe, x.grad.numpy(), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(dge, gamma.grad.numpy(), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(dbe, beta.grad.numpy(), rtol=1e-8, atol=1e-9)
    if has_res:
        np.testing.assert_allclose(dze, res.grad.numpy(), rtol=1e-9, atol=1e-12)
