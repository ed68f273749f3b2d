"""Fused BatchNorm(+add)(+ReLU) kernels vs a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

from eventgrad_b200.ops.bn_act import FusedBNAct, _eligible, bn_act_reference

pytestmark = pytest.mark.gpu


def _mk(N, C, H, W, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(N, C, H, W, generator=g, device="cuda") * 1.5 + 0.3).to(torch.bfloat16)
    x = x.contiguous(memory_format=torch.channels_last)
    r = torch.randn(N, C, H, W, generator=g, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(N, C, H, W, generator=g, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    return x, r, dy


@pytest.mark.parametrize("shape", [(32, 64, 32, 32), (7, 128, 16, 16), (5, 256, 8, 8), (3, 512, 4, 4), (2, 2048, 4, 4),
                                   (1, 64, 3, 5), (128, 64, 32, 32), (64, 128, 16, 16)])   # last two: split path
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
def test_fused_bn_act_forward_backward(shape, relu, res):
    N, C, H, W = shape
    x, r, dy = _mk(*shape)
    bn = FusedBNAct(C).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    ref = FusedBNAct(C).cuda().train()
    ref.load_state_dict(bn.state_dict())
    xa = x.clone().requires_grad_(True)
    ra = r.clone().requires_grad_(True) if res else None
    assert _eligible(xa, ra)
    y = bn(xa, residual=ra, relu=relu)
    y.backward(dy)
    # fp32 reference from the same bf16 inputs
    xb = x.float().requires_grad_(True)
    rb = r.float().requires_grad_(True) if res else None
    yb = bn_act_reference(xb, ref.weight, ref.bias, ref.running_mean, ref.running_var, rb, True, ref.momentum,
                          ref.eps, relu)
    # mask the reference backward exactly like the kernel does (on the bf16-rounded output sign)
    yb.backward(dy.float())
    torch.testing.assert_close(y.float(), yb, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(bn.running_mean, ref.running_mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(bn.running_var, ref.running_var, rtol=1e-4, atol=1e-5)
    scale = float(xb.grad.abs().max()) + 1e-6
    # a handful of elements sit exactly on the ReLU boundary (|y| ~ 1e-7) where the bf16 kernel and the
    # fp32 reference may disagree on the mask; everything else must agree to bf16 precision
    bad = ((xa.grad.float() - xb.grad).abs() / scale > 3e-2).float().mean()
    assert float(bad) < 2e-5, float(bad)
    gs = float(ref.weight.grad.abs().max()) + 1e-6
    assert float((bn.weight.grad - ref.weight.grad).abs().max()) / gs < 2e-2
    bs = float(ref.bias.grad.abs().max()) + 1e-6
    assert float((bn.bias.grad - ref.bias.grad).abs().max()) / bs < 2e-2
    if res:
        rs = float(rb.grad.abs().max()) + 1e-6
        badr = ((ra.grad.float() - rb.grad).abs() / rs > 2e-2).float().mean()
        assert float(badr) < 2e-5, float(badr)
    assert int(bn.num_batches_tracked) == 1
    from eventgrad_b200.ops.bn_act import bn_status
    assert bn_status(x.device) == 0


def test_fused_bn_eval_mode():
    x, r, _ = _mk(4, 128, 8, 8)
    bn = FusedBNAct(128).cuda()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 2.0)
    bn.eval()
    y = bn(x, residual=r, relu=True)
    yb = bn_act_reference(x.float(), bn.weight, bn.bias, bn.running_mean, bn.running_var, r.float(), False, 0.1,
                          bn.eps, True)
    torch.testing.assert_close(y.float(), yb, rtol=2e-2, atol=2e-2)


def test_resnet_fused_vs_fallback_one_step():
    """Whole flagship model, bf16 autocast NHWC: the fused kernels must track an fp32 reference run at
    least as well as the ATen bf16 path does (both are bf16 approximations of the same maths)."""
    import os
    from eventgrad_b200.models import build_model
    torch.manual_seed(0)
    ms = [build_model("resnet18").cuda().train() for _ in range(3)]
    for m in ms[1:]:
        m.load_state_dict(ms[0].state_dict())
    x = torch.randn(64, 3, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last)
    yl = torch.randint(0, 10, (64,), device="cuda")
    res = []
    for m, flag, amp in ((ms[0], "1", True), (ms[1], "0", True), (ms[2], "0", False)):
        os.environ["EGB_FUSED_BN"] = flag
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            out = m(x)
        loss = torch.nn.functional.cross_entropy(out.float(), yl)
        loss.backward()
        res.append((loss.item(), torch.cat([p.grad.flatten() for p in m.parameters()])))
    os.environ["EGB_FUSED_BN"] = "1"
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a, b, dim=0))
    c_fused, c_aten = cos(res[0][1], res[2][1]), cos(res[1][1], res[2][1])
    assert abs(res[0][0] - res[2][0]) < 5e-2
    assert c_fused > c_aten - 0.02, (c_fused, c_aten)
    assert c_fused > 0.8, c_fused


def test_single_launch_variant_matches_split(monkeypatch):
    """EGB_BN_FUSED_SMALL=1 (one launch, per-slice epoch flags) must give the same numbers."""
    import eventgrad_b200.ops.bn_act as B
    x, r, dy = _mk(32, 64, 32, 32)
    outs = []
    for fused in (0, 1):
        ws = B._workspace(x.device)
        ws["fused"] = fused
        bn = FusedBNAct(64).cuda().train()
        xa = x.clone().requires_grad_(True)
        ra = r.clone().requires_grad_(True)
        y = bn(xa, residual=ra, relu=True)
        y.backward(dy)
        outs.append((y.detach().float(), xa.grad.float(), bn.weight.grad.clone(), bn.running_var.clone()))
    B._workspace(x.device)["fused"] = 0
    assert B.bn_status(x.device) == 0
    # both variants round their results to bf16; fp32 statistics may differ in the last bit
    for a, b in zip(outs[0], outs[1]):
        torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------------------------
# fp32 activations (the reference's precision, /root/reference/dcifar10/event/event.cpp:259-276): same kernels
# instantiated for float, compared against plain PyTorch fp32 at fp32 tolerances
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(32, 64, 32, 32), (7, 128, 16, 16), (5, 256, 8, 8), (3, 512, 4, 4), (1, 64, 3, 5),
                                   (256, 64, 32, 32), (64, 128, 16, 16)])   # last two: split path
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
@pytest.mark.parametrize("fused_small", [0, 1])
def test_fused_bn_act_fp32(shape, relu, res, fused_small):
    _fp32_case(shape, relu, res, fused_small, torch.channels_last)


# fp32 NCHW kernels (csrc/bn_nchw.cu) -- the layout of the reference-precision default path.  Includes channel counts
# that are not a multiple of 64 and an H*W whose float4 count is not a power of two (division path).
@pytest.mark.parametrize("shape", [(32, 64, 32, 32), (7, 128, 16, 16), (5, 256, 8, 8), (3, 512, 4, 4), (4, 10, 6, 6),
                                   (256, 64, 32, 32), (64, 128, 16, 16), (1, 3, 2, 2), (2, 2048, 2, 2), (9, 20, 12, 12)])
@pytest.mark.parametrize("relu,res", [(True, False), (True, True), (False, False), (False, True)])
def test_fused_bn_act_fp32_nchw(shape, relu, res):
    _fp32_case(shape, relu, res, 0, torch.contiguous_format)


def test_fused_bn_nchw_eval_mode():
    x = torch.randn(4, 128, 8, 8, device="cuda")
    r = torch.randn(4, 128, 8, 8, device="cuda")
    bn = FusedBNAct(128).cuda()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2)
        bn.running_var.uniform_(0.5, 2.0)
    bn.eval()
    y = bn(x, residual=r, relu=True)
    yb = bn_act_reference(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, r, False, 0.1, bn.eps, True)
    torch.testing.assert_close(y, yb, rtol=1e-5, atol=1e-5)


def _fp32_case(shape, relu, res, fused_small, fmt):
    import eventgrad_b200.ops.bn_act as B
    N, C, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    mk = lambda s, b: (torch.randn(N, C, H, W, generator=g, device="cuda") * s + b).contiguous(memory_format=fmt)
    x, r, dy = mk(1.5, 0.3), mk(1.0, 0.0), mk(1.0, 0.0)
    B._workspace(x.device)["fused"] = fused_small
    try:
        bn = FusedBNAct(C).cuda().train()
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.5, 0.5)
        ref = torch.nn.BatchNorm2d(C).cuda().train()
        ref.load_state_dict(bn.state_dict())
        xa = x.clone().requires_grad_(True)
        ra = r.clone().requires_grad_(True) if res else None
        assert _eligible(xa, ra) if fmt == torch.channels_last else B._eligible_nchw(xa, ra)
        n0 = B._workspace(x.device)["C"].launch_counts()["bn"]
        y = bn(xa, residual=ra, relu=relu)
        assert B._workspace(x.device)["C"].launch_counts()["bn"] > n0          # the CUDA path ran, not the fallback
        assert y.dtype == torch.float32 and y.is_contiguous(memory_format=fmt)
        y.backward(dy)
        xb = x.clone().requires_grad_(True)
        rb = r.clone().requires_grad_(True) if res else None
        yb = ref(xb)
        if res:
            yb = yb + rb
        if relu:
            yb = torch.relu(yb)
        yb.backward(dy)
    finally:
        B._workspace(x.device)["fused"] = 0
    torch.testing.assert_close(y, yb, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(bn.running_mean, ref.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var, ref.running_var, rtol=1e-5, atol=1e-6)
    # elements exactly on the ReLU boundary may flip between two correct fp32 evaluations; everything else is tight
    scale = float(xb.grad.abs().max()) + 1e-6
    bad = ((xa.grad - xb.grad).abs() / scale > 1e-4).float().mean()
    assert float(bad) < 2e-5, float(bad)
    # dgamma / dbeta are fp32 statistics accumulated in double by the kernel: tight check
    gs = float(ref.weight.grad.abs().max()) + 1e-6
    assert float((bn.weight.grad - ref.weight.grad).abs().max()) / gs < 2e-4
    bs = float(ref.bias.grad.abs().max()) + 1e-6
    assert float((bn.bias.grad - ref.bias.grad).abs().max()) / bs < 2e-4
    if res:
        rs = float(rb.grad.abs().max()) + 1e-6
        assert float(((ra.grad - rb.grad).abs() / rs > 1e-5).float().mean()) < 2e-5
    assert B.bn_status(x.device) == 0


def _plain_run(steps, batches, theta0_model, channels_last, lr, mu):
    """Plain PyTorch fp32 training (ATen BN, torch.optim.SGD) on the given batches; returns the loss list."""
    import os
    import torch.nn.functional as F
    from eventgrad_b200.models import build_model
    m = build_model("resnet18").cuda().train()
    m.load_state_dict(theta0_model)
    if channels_last:
        m = m.to(memory_format=torch.channels_last)
    opt = torch.optim.SGD(m.parameters(), lr=lr, momentum=mu)
    out = []
    os.environ["EGB_FUSED_BN"] = "0"
    try:
        for x, y in batches[:steps]:
            xx = x.contiguous(memory_format=torch.channels_last) if channels_last else x.contiguous()
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(m(xx), y)
            loss.backward()
            opt.step()
            out.append(loss.detach())
    finally:
        os.environ["EGB_FUSED_BN"] = "1"
    return torch.stack(out).cpu()


@pytest.mark.parametrize("nhwc", [False, True])
def test_fp32_first_step_gradients_vs_fp64_truth(nhwc):
    """Same weights, same batch.  Ground truth = the same network in float64 (plain PyTorch).  The fused fp32 BN kernels
    (NCHW = default fp32 layout, and NHWC) must be as close to it as plain PyTorch fp32 (cuDNN BN + ATen add/ReLU) is:
    two fp32 evaluations differ by ReLU-boundary flips (an element whose pre-activation is +1e-7 in one run and -1e-7 in
    the other flips a whole gradient term), so fp32-vs-fp32 is the wrong yardstick -- distance to fp64 is the right one."""
    import os
    import torch.nn.functional as F
    from eventgrad_b200.models import build_model
    from eventgrad_b200.ops import ext
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    a = build_model("resnet18").cuda().train()
    b = build_model("resnet18").cuda().train()
    t = build_model("resnet18").cuda().train()
    b.load_state_dict(a.state_dict())
    t.load_state_dict(a.state_dict())
    t = t.double()
    if nhwc:
        a = a.to(memory_format=torch.channels_last)
    x = torch.randn(32, 3, 32, 32, device="cuda")
    y = torch.randint(0, 10, (32,), device="cuda")
    n0 = ext().launch_counts()["bn"]
    la = F.cross_entropy(a(x.contiguous(memory_format=torch.channels_last) if nhwc else x), y)
    la.backward()
    assert ext().launch_counts()["bn"] - n0 == 4 * 28          # every BN layer took the fused CUDA path
    os.environ["EGB_FUSED_BN"] = "0"
    try:
        lb = F.cross_entropy(b(x), y)
        lb.backward()
        lt = F.cross_entropy(t(x.double()), y)
        lt.backward()
    finally:
        os.environ["EGB_FUSED_BN"] = "1"
    cat = lambda m: torch.cat([p.grad.detach().double().flatten() for p in m.parameters()])
    ga, gb, gt = cat(a), cat(b), cat(t)
    err_ours = float((ga - gt).norm() / gt.norm())
    err_plain = float((gb - gt).norm() / gt.norm())
    print(f"fp32 gradient error vs fp64 truth: fused BN ({'NHWC' if nhwc else 'NCHW'}) {err_ours:.3e}, plain PyTorch {err_plain:.3e}; "
          f"loss error {abs(float(la) - float(lt)):.2e} vs {abs(float(lb) - float(lt)):.2e}")
    assert abs(float(la.detach()) - float(lt.detach())) < 1e-5
    assert err_ours <= 3.0 * err_plain + 1e-5, (err_ours, err_plain)


@pytest.mark.parametrize("fmt", [torch.contiguous_format, torch.channels_last])
@pytest.mark.parametrize("shape", [(32, 512, 4, 4), (32, 64, 32, 32), (16, 128, 16, 16)])
def test_fp32_bn_statistics_vs_fp64(shape, fmt):
    """No ReLU (no mask discontinuity): y, dx and the fp32 statistics dgamma / dbeta / mean / var of the fused fp32
    kernels against a float64 evaluation of the same op -- errors at the fp32 rounding level (1e-6), tighter than
    what cuDNN's own fp32 BN achieves on the same inputs."""
    N, C, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    mk = lambda s, b: (torch.randn(N, C, H, W, generator=g, device="cuda") * s + b).contiguous(memory_format=fmt)
    x, r, dy = mk(1.5, 0.7), mk(1.0, 0.0), mk(1.0, 0.1)
    bn = FusedBNAct(C).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    ref = torch.nn.BatchNorm2d(C).cuda().train().double()
    ref.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in bn.state_dict().items()})
    xa, ra = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    y = bn(xa, residual=ra, relu=False)
    y.backward(dy)
    xb, rb = x.double().requires_grad_(True), r.double().requires_grad_(True)
    yb = ref(xb) + rb
    yb.backward(dy.double())
    rel = lambda u, v: float((u.double() - v).norm() / v.norm())
    assert rel(y, yb) < 2e-6 and rel(xa.grad, xb.grad) < 5e-6
    assert rel(bn.weight.grad, ref.weight.grad) < 2e-6, rel(bn.weight.grad, ref.weight.grad)
    assert rel(bn.bias.grad, ref.bias.grad) < 2e-6, rel(bn.bias.grad, ref.bias.grad)
    assert rel(bn.running_mean, ref.running_mean) < 2e-6 and rel(bn.running_var, ref.running_var) < 2e-6
    assert torch.equal(ra.grad, dy)                              # dres is dz itself


def test_fp32_training_tracks_plain_pytorch_50_steps():
    """Default product path on a GPU (fp32, fused BN kernels, whole-step CUDA graph, fused SGD step kernel) against plain
    PyTorch fp32 + torch.optim.SGD from the same seed and batches over 50 steps.  Early ResNet training at lr 1e-2 /
    momentum 0.9 amplifies fp32 rounding differences between cuDNN algorithms by orders of magnitude per step, so the
    yardstick is measured, not assumed: two PLAIN PyTorch runs that differ only in memory format (NCHW vs NHWC) give
    the noise floor D0; our trajectory must stay within max(1e-3, 3*D0) of the plain run, and within 1e-3 on the first
    two steps (before amplification)."""
    from eventgrad_b200.config import preset
    from eventgrad_b200.data import synthetic_source
    from eventgrad_b200.engine.trainer import Trainer
    from eventgrad_b200.utils.dist import DistEnv
    dev = torch.device("cuda", 0)
    cfg = preset("cifar_event", algo="decent", backend="p2p", device="cuda", dtype="fp32", train_samples=512,
                 test_samples=64, batch_size=32, epochs=100, quiet=True, augment=False, sampler="sequential",
                 cudnn_benchmark=False)
    tr = Trainer(cfg, DistEnv(0, 1, 0, dev, "none"), train_source=synthetic_source("cifar10", 512).pin())
    assert tr.cfg.cuda_graph                                     # the fast path IS the default on a GPU
    theta0 = {k: v.detach().clone().contiguous() for k, v in tr.model.state_dict().items()}
    batches, l_ours = [], []
    it = iter(tr.loader)
    for s in range(50):
        try:
            x, y = next(it)
        except StopIteration:
            it = iter(tr.loader)
            x, y = next(it)
        batches.append((x.detach().clone().contiguous(), y.clone()))
        l_ours.append(tr.train_step(x, y).clone())
    tr.backend.check_status()
    a = torch.stack(l_ours).cpu()
    tr.close()
    p_nchw = _plain_run(50, batches, theta0, False, cfg.lr, cfg.momentum)
    p_nhwc = _plain_run(50, batches, theta0, True, cfg.lr, cfg.momentum)
    d0 = float((p_nchw - p_nhwc).abs().max())
    d_ours = min(float((a - p_nchw).abs().max()), float((a - p_nhwc).abs().max()))
    print(f"fp32 50-step trajectory: ours-vs-plain {d_ours:.3e}, plain NCHW-vs-NHWC noise floor {d0:.3e}; "
          f"first steps {[float(v) for v in (a - p_nchw).abs()[:3]]}")
    assert torch.isfinite(a).all()
    assert float((a - p_nchw).abs()[:2].max()) < 1e-3
    assert d_ours <= max(1e-3, 3.0 * d0), (d_ours, d0)
    assert abs(float(a[-1]) - float(p_nchw[-1])) < max(1e-2, 3.0 * d0)
