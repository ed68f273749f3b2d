"""CPU-side checks of the gradient-table / shadow-weight plumbing (the kernels are tested on GPU)."""
import torch

from eventgrad_b200.models import build_model
from eventgrad_b200.ops.shadow import ShadowConv2d, ShadowLinear
from eventgrad_b200.parallel.arena import ParamArena


def test_shadow_modules_fall_back_to_plain_layers_on_cpu():
    m = build_model("lenet").eval()          # eval: Dropout2d off, so two forwards are comparable
    assert isinstance(m.conv1, ShadowConv2d) and isinstance(m.fc1, ShadowLinear)
    x = torch.randn(2, 3, 32, 32)
    y1 = m(x)
    m.conv1.w16 = torch.zeros_like(m.conv1.weight, dtype=torch.bfloat16)      # ignored off-GPU / without autocast
    torch.testing.assert_close(m(x), y1)


def test_table_mode_compute_tensors_and_grad_flow():
    torch.manual_seed(0)
    m = build_model("cnn2")
    a = ParamArena(m)
    a.enable_table_mode(shadow=False)
    assert a.table_mode and a.shadow is None and len(a.compute) == a.table.n_tensors
    assert all(c is p for c, p in zip(a.compute, a.params))
    assert all(p.grad is None for p in a.params)
    out = m(torch.randn(4, 1, 28, 28))
    out.sum().backward()
    for c, p in zip(a.compute, a.params):
        assert c.grad is not None and c.grad.shape == p.shape and c.grad.stride() == p.stride()
    a.clear_compute_grads()
    assert all(c.grad is None for c in a.compute)


def test_arena_channels_last_views_alias_flat_memory():
    m = build_model("resnet18")
    a = ParamArena(m, channels_last=True)
    w = m.layer1[0].conv1.weight
    assert w.shape == (64, 64, 3, 3) and w.is_contiguous(memory_format=torch.channels_last)
    i = [n for n, _ in m.named_parameters()].index("layer1.0.conv1.weight")
    flat = a.flat(a.theta, i)
    with torch.no_grad():
        flat.zero_()
    assert float(w.abs().sum()) == 0.0                     # the parameter IS the arena memory
    assert a.table.n_tensors == 86 and a.table.n_elems == 17444682
