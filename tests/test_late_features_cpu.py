"""CPU checks of the host-side logic behind the late round-1 additions (their kernels need a GPU:
tests/test_gpu_experimental.py)."""
import pytest
import torch

from eventgrad_b200.config import TrainConfig, parse_cli
from eventgrad_b200.models import build_model
from eventgrad_b200.parallel import ParamArena
from eventgrad_b200.parallel.arena import TensorTable
from eventgrad_b200.parallel.p2p import _wants_dbuf, build_layout


def test_cli_flags_of_experimental_paths_parse():
    cfg = parse_cli("decent", ["0", "--double-buffer", "--ce-push", "--overlap-push"])
    assert cfg.double_buffer and cfg.ce_push and cfg.overlap_push
    cfg = parse_cli("cifar_spevent", ["0", "1", "1.0", "10", "--fresh-replicas"])
    assert cfg.spevent_fresh_replicas and cfg.topk_percent == 10.0
    assert not parse_cli("decent", ["0"]).double_buffer


def test_double_buffer_applies_to_dense_iter_sync_fused_decent_only():
    base = dict(dataset="mnist", model="cnn2", double_buffer=True)
    assert _wants_dbuf(TrainConfig(algo="decent", **base).validate())
    assert not _wants_dbuf(TrainConfig(algo="event", **base).validate())
    with pytest.raises(ValueError):                      # decent is lock-step by definition
        TrainConfig(algo="decent", sync_mode="async", **base).validate()
    assert not _wants_dbuf(TrainConfig(algo="decent", overlap_push=True, **base).validate())
    t = TensorTable.from_named(list(build_model("cnn2").named_parameters()), 2048)
    one = build_layout(t, TrainConfig(algo="decent", dataset="mnist", model="cnn2").validate(), 4, 64)
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


def test_bn_cluster_plan_covers_every_row_and_fits_shared_memory():
    """Host-side plan of csrc/bn_act_cluster.cu (pure host code in the extension, runs without a GPU)."""
    from eventgrad_b200.ops import ext
    C = ext()
    for which, slabs, cap in ((0, 1, 1536), (2, 2, 768)):
        for M in list(range(1, 600, 7)) + [1024, 2048, 4095, 4096, 8192, 12288, 12289, 16384, 24576, 24577, 32768, 262144]:
            cs, rows, smem = C.bn_cluster_plan(M, which)
            if cs == 0:
                assert M > 16 * cap - 16 * 31, (M, which)        # only genuinely too-large slices are refused
                continue
            assert cs in (1, 2, 4, 8, 16)
            assert rows % 32 == 0 and rows <= cap
            assert cs * rows >= M                                   # every row has an owner
            assert (cs - 1) * rows < M + rows                       # no more than one (partially) idle tail CTA chain
            assert smem == rows * 128 * slabs <= 196608
    assert C.bn_cluster_plan(2048, 0)[0] == 8                       # batch-32 stage 3: one slice over 8 SMs
    assert C.bn_cluster_plan(8192, 2)[0] == 16                      # batch-32 stage 2 backward needs the 16-CTA cluster
