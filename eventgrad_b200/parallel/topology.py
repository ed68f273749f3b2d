"""Logical ring topology (rank order), including the degenerate sizes.

Reference: left=(rank-1) mod N, right=(rank+1) mod N
(/root/reference/dmnist/event/event.cpp:114-122); the CIFAR programs only set it
when numranks>1 (/root/reference/dcifar10/event/event.cpp:70-81).

On NVSwitch every peer is one hop away, so the ring is purely logical: no
topology-aware placement is needed and ring order == rank order.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Ring:
    rank: int
    world: int

    def __post_init__(self):
        if not (0 <= self.rank < self.world):
            raise ValueError(f"rank {self.rank} outside world {self.world}")

    @property
    def left(self) -> int:
        return (self.rank - 1) % self.world

    @property
    def right(self) -> int:
        return (self.rank + 1) % self.world

    @property
    def degenerate_pair(self) -> bool:
        """R == 2: left == right, the same peer is counted twice -> (theta + 2 theta')/3
        (SURVEY.md A.4 / Q15)."""
        return self.world == 2

    @property
    def serial(self) -> bool:
        return self.world == 1

    def neighbours(self):
        return self.left, self.right

    def mixing_row(self):
        """Row of the doubly-stochastic mixing matrix W for this rank (weights 1/3)."""
        row = [0.0] * self.world
        row[self.rank] += 1.0 / 3.0
        row[self.left] += 1.0 / 3.0
        row[self.right] += 1.0 / 3.0
        return row
