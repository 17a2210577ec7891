"""N-sharding of observations across ranks.

The statistics G, r, q, ℓ are sums over observations, so each rank accumulates its contiguous block of columns
and ONE sum-allreduce of the packed (D² + D + 3)-double buffer combines them; the D x D factorisation then runs
replicated on every rank (bit-identical results, no broadcast).  Prediction shards by test point with no
collective at all.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous, balanced partition of N observations over `world` ranks (sizes differ by at most 1 chunk of
    `align` observations; boundaries are multiples of `align` so device tiles stay aligned)."""

    N: int
    world: int
    align: int = 16

    def bounds(self, rank: int) -> Tuple[int, int]:
        if not (0 <= rank < self.world):
            raise ValueError("rank out of range")
        units = -(-self.N // self.align)
        base, extra = divmod(units, self.world)
        lo_u = rank * base + min(rank, extra)
        hi_u = lo_u + base + (1 if rank < extra else 0)
        return min(lo_u * self.align, self.N), min(hi_u * self.align, self.N)

    def size(self, rank: int) -> int:
        lo, hi = self.bounds(rank)
        return hi - lo

    def all_bounds(self):
        return [self.bounds(r) for r in range(self.world)]


def packed_len(D: int) -> int:
    return D * D + D + 3
