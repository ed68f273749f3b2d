import torch
import torch.nn.functional as F

from eventgrad_b200.config import TrainConfig
from eventgrad_b200.data import (BatchLoader, ShardSampler, decode_augment_torch, draw_augment_params,
                                 per_rank_batch, synthetic_source)


def test_shards_are_disjoint_floor_sized_and_replayed():
    for mode in ("random", "sequential"):
        parts = [ShardSampler(1003, 4, r, mode).indices() for r in range(4)]
        assert all(len(p) == 250 for p in parts)                     # floor(N/R), allow_duplicates=false
        assert len(torch.cat(parts).unique()) == 1000
    s = ShardSampler(100, 2, 1, "random")
    a = s.indices(); s.set_epoch(3); b = s.indices()
    assert torch.equal(a, b)                                          # reference never calls set_epoch
    assert torch.equal(ShardSampler(100, 2, 1, "sequential").indices(), torch.arange(50, 100))


def test_batch_modes():
    assert per_rank_batch(TrainConfig(batch_mode="global", batch_size=256), 8, 6250) == 32
    assert per_rank_batch(TrainConfig(batch_mode="per_rank", batch_size=64), 8, 7500) == 64
    assert per_rank_batch(TrainConfig(batch_mode="full"), 4, 15000) == 15000


def test_augment_equals_pad_flip_crop():
    x = torch.randint(0, 256, (16, 3, 32, 32), dtype=torch.uint8)
    p = draw_augment_params(16, 4, "cpu", torch.Generator().manual_seed(1))
    assert int(p[0].max()) <= 7 and int(p[1].max()) <= 7              # randint excludes the max offset
    out = decode_augment_torch(x, 1.0, 0.0, 1.0, p)
    for b in range(16):
        im = F.pad(x[b].float(), (4, 4, 4, 4))
        if p[2][b]:
            im = im.flip(-1)
        assert torch.equal(im[:, p[0][b]:p[0][b] + 32, p[1][b]:p[1][b] + 32], out[b])


def test_loader_keeps_partial_batch_and_normalises():
    src = synthetic_source("mnist", 1000)
    ld = BatchLoader(src, ShardSampler(1000, 3, 0, "sequential"), 64, "cpu")
    sizes = [x.shape[0] for x, _ in ld]
    assert sum(sizes) == 333 and sizes[-1] == 333 % 64 and len(ld) == 6
    x, y = next(iter(ld))
    ref = (src.images[:64].float() / 255.0 - 0.1307) / 0.3081
    assert torch.allclose(x, ref, atol=1e-6) and torch.equal(y, src.labels[:64])


def test_synthetic_is_deterministic_and_learnable_shape():
    a, b = synthetic_source("cifar10", 64), synthetic_source("cifar10", 64)
    assert torch.equal(a.images, b.images) and a.images.shape == (64, 3, 32, 32)
    assert a.scale == 1.0 and a.mean == 0.0                           # Q7: raw 0..255, un-normalised
