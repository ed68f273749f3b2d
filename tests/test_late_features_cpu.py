"""CPU checks of the host-side logic around the p2p backend options (their kernels are exercised by the gpu tier)."""
import pytest
import torch

from eventgrad_b200.config import TrainConfig, parse_cli
from eventgrad_b200.models import build_model
from eventgrad_b200.parallel import ParamArena
from eventgrad_b200.parallel.arena import TensorTable
from eventgrad_b200.parallel.p2p import _wants_dbuf, build_layout


def test_cli_flags_of_backend_options_parse():
    cfg = parse_cli("decent", ["0", "--double-buffer", "--ce-push", "--overlap-push"])
    assert cfg.double_buffer and cfg.ce_push and cfg.overlap_push
    cfg = parse_cli("cifar_spevent", ["0", "1", "1.0", "10", "--fresh-replicas"])
    assert cfg.spevent_fresh_replicas and cfg.topk_percent == 10.0
    assert parse_cli("decent", ["0"]).double_buffer is None                 # auto
    assert parse_cli("decent", ["0", "--no-double-buffer"]).double_buffer is False
    cfg = parse_cli("cifar_event", ["0", "1", "1.0", "--no-cuda-graph", "--no-channels-last", "--peer-timeout-s", "5"])
    assert cfg.cuda_graph is False and cfg.channels_last is False and cfg.peer_timeout_s == 5.0


def test_execution_switches_resolve_to_the_fast_path_on_cuda_only():
    """device=cuda implies NHWC + fused BN, whole-step CUDA graph and (N>=2, decent/event) overlapped pushes unless
    turned off; on the CPU everything stays off (bench config == CLI default config)."""
    c = parse_cli("cifar_event", ["0", "1", "1.0"])
    assert c.channels_last is None and c.cuda_graph is None and c.overlap_push is None
    g8 = c.resolved("cuda", 8)
    # fp32: tensor-core convolutions at fp32 accuracy (csrc/conv_tc.cu) on NHWC activations ...
    assert g8.conv_tc and g8.channels_last and g8.cuda_graph and g8.overlap_push
    # ... unless switched off: cuDNN's fp32 convs are NCHW kernels, so that path stays NCHW
    off_tc = c.replace(conv_tc=False).resolved("cuda", 8)
    assert not off_tc.conv_tc and not off_tc.channels_last
    assert not c.replace(dtype="bf16").resolved("cuda", 8).conv_tc
    assert c.replace(dtype="bf16").resolved("cuda", 8).channels_last and c.replace(dtype="tf32").resolved("cuda", 1).channels_last
    g1 = c.resolved("cuda", 1)
    assert g1.cuda_graph and not g1.overlap_push
    cpu = c.resolved("cpu", 2)
    assert not cpu.channels_last and not cpu.cuda_graph and not cpu.overlap_push and not cpu.conv_tc
    off = parse_cli("cifar_event", ["0", "1", "1.0", "--no-overlap-push", "--no-cuda-graph"]).resolved("cuda", 8)
    assert off.overlap_push is False and off.cuda_graph is False
    assert not parse_cli("cifar_spevent", ["0", "1", "1.0", "10"]).resolved("cuda", 8).overlap_push
    assert not parse_cli("cent", []).resolved("cuda", 8).overlap_push


def test_double_buffer_applies_to_dense_iter_sync_fused_decent_only():
    base = dict(dataset="mnist", model="cnn2", double_buffer=True)
    assert _wants_dbuf(TrainConfig(algo="decent", **base).validate())
    assert not _wants_dbuf(TrainConfig(algo="event", **base).validate())
    with pytest.raises(ValueError):                      # decent is lock-step by definition
        TrainConfig(algo="decent", sync_mode="async", **base).validate()
    assert not _wants_dbuf(TrainConfig(algo="decent", overlap_push=True, **base).validate())
    t = TensorTable.from_named(list(build_model("cnn2").named_parameters()), 2048)
    assert _wants_dbuf(TrainConfig(algo="decent", dataset="mnist", model="cnn2").validate())     # default: on
    one = build_layout(t, TrainConfig(algo="decent", dataset="mnist", model="cnn2", double_buffer=False).validate(), 4, 64)
    two = build_layout(t, TrainConfig(algo="decent", **base).validate(), 4, 64)
    assert two.nbytes("inbox_l") == 2 * one.nbytes("inbox_l") == 2 * t.n_padded * 4
    assert two.offset("inbox_r") % 256 == 0 and two.size > one.size


@pytest.mark.parametrize("channels_last", [False, True])
def test_arena_pack_uses_the_arena_layout(channels_last):
    torch.manual_seed(0)
    a, b = build_model("lenet"), build_model("lenet")
    arena = ParamArena(a, channels_last=channels_last)
    flat = arena.pack(b)
    for i, (_, p) in enumerate(b.named_parameters()):
        assert torch.equal(arena.view(flat, i), p.detach())
    # padding lanes stay zero
    t = arena.table
    used = torch.zeros(t.n_padded, dtype=torch.bool)
    for o, n in zip(t.offsets, t.numels):
        used[o:o + n] = True
    assert float(flat[~used].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        arena.pack(build_model("cnn2"))


def test_set_sparse_init_guards():
    from eventgrad_b200.parallel.collective import CollectiveBackend
    from eventgrad_b200.parallel.topology import Ring
    m = build_model("cnn2")
    arena = ParamArena(m)
    z = torch.zeros_like(arena.theta)
    dense = CollectiveBackend(TrainConfig(algo="event", dataset="mnist", model="cnn2").validate(), arena, Ring(0, 1))
    with pytest.raises(RuntimeError):
        dense.set_sparse_init(z, z, z)
    sp = CollectiveBackend(TrainConfig(algo="spevent", dataset="mnist", model="cnn2").validate(), arena, Ring(0, 1))
    sp.set_sparse_init(z + 1, z + 2, z + 3)
    assert float(sp.prev[0]) == 1 and float(sp.rep_l[0]) == 2 and float(sp.rep_r[0]) == 3
    sp.pass_num = 1
    with pytest.raises(RuntimeError):
        sp.set_sparse_init(z, z, z)
