"""Data sharding = the data-parallel split.

Mirrors LibTorch's distributed samplers as used by the reference:
  DistributedRandomSampler(size, R, rank, allow_duplicates=false)   cent / CIFAR programs
  DistributedSequentialSampler(size, R, rank, false)                decent / dmnist/event
Each rank owns floor(size/R) samples: the contiguous block [rank*local, (rank+1)*local) of
either the identity order (sequential) or a permutation seeded by the sampler epoch
(random).  The reference never calls set_epoch(), so the same order is replayed every epoch;
`reshuffle=True` opts into per-epoch reshuffling.  The partial last batch is kept
(DataLoader drop_last=false).
"""
from __future__ import annotations

import torch


class ShardSampler:
    def __init__(self, size: int, world: int, rank: int, mode: str = "random",
                 seed: int = 0, reshuffle: bool = False):
        if mode not in ("random", "sequential"):
            raise ValueError("mode must be random|sequential")
        self.size, self.world, self.rank, self.mode = size, world, rank, mode
        self.seed, self.reshuffle = seed, reshuffle
        self.local_size = size // world
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def indices(self) -> torch.Tensor:
        b, e = self.rank * self.local_size, (self.rank + 1) * self.local_size
        if self.mode == "sequential":
            return torch.arange(b, e)
        ep = self.epoch if self.reshuffle else 0
        g = torch.Generator().manual_seed(self.seed + ep)
        return torch.randperm(self.size, generator=g)[b:e]

    def num_batches(self, batch: int) -> int:
        return -(-self.local_size // batch)


def per_rank_batch(cfg, world: int, local_size: int) -> int:
    """Batch size on one rank for the three batch modes of the reference programs:
    global/R (dcifar10/event/event.cpp:91), fixed per rank (dmnist/event/event.cpp:145),
    or the whole shard (dmnist/cent/cent.cpp:62-65)."""
    if cfg.batch_mode == "full":
        return local_size
    if cfg.batch_mode == "global":
        return max(1, cfg.batch_size // world)
    return cfg.batch_size
