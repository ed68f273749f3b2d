"""Model inventories must match the reference exactly: tensor count, element
count and named_parameters() order define the arena layout (= the reference's `disp`)."""
import pytest
import torch

from eventgrad_b200.models import build_model, param_inventory, outputs_log_probs

EXPECTED = {"mlp": (4, 101770), "cnn1": (8, 38390), "cnn2": (8, 27480), "lenet": (10, 62006),
            "resnet18": (86, 17444682)}


@pytest.mark.parametrize("name", list(EXPECTED))
def test_inventory(name):
    n, e, _ = param_inventory(build_model(name))
    assert (n, e) == EXPECTED[name]


def test_resnet_canonical_and_quirk():
    n, e, _ = param_inventory(build_model("resnet18", resnet_variant="canonical"))
    assert (n, e) == (62, 11173962)
    m = build_model("resnet18")                       # Q1: blocks+1 per stage -> 12 BasicBlocks
    assert sum(len(getattr(m, f"layer{i}")) for i in range(1, 5)) == 12
    hist = sorted(p.numel() for p in m.parameters())
    assert hist[0] == 10 and hist[-1] == 2359296 and hist.count(2359296) == 5


def test_cnn2_tensor_shapes_are_the_message_shapes():
    _, _, items = param_inventory(build_model("cnn2"))
    assert [n for _, n in items] == [90, 10, 1800, 20, 25000, 50, 500, 10]   # BASELINE.md call-site table


@pytest.mark.parametrize("name,shape", [("mlp", (5, 1, 28, 28)), ("cnn1", (5, 1, 28, 28)),
                                        ("cnn2", (5, 1, 28, 28)), ("lenet", (5, 3, 32, 32)),
                                        ("resnet18", (2, 3, 32, 32)), ("resnet50", (2, 3, 32, 32))])
def test_forward_shapes(name, shape):
    m = build_model(name).eval()
    out = m(torch.randn(*shape))
    assert out.shape == (shape[0], 10)
    if outputs_log_probs(name):
        assert torch.allclose(out.exp().sum(1), torch.ones(shape[0]), atol=1e-5)


def test_mlp_relu_on_logits_quirk():
    m = build_model("mlp")
    assert (m(torch.randn(16, 784)) >= 0).all()       # Q11
