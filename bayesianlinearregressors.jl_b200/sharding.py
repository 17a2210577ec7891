"""N-sharding of observations across ranks.

The statistics G, r, q, ℓ are sums over observations, so each rank accumulates its contiguous block of columns
and ONE sum-allreduce of the packed (D² + D + 3)-double buffer combines them; the D x D factorisation then runs
replicated on every rank (bit-identical results, no broadcast).  Prediction shards by test point with no
collective at all.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib as L


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


def pack_stats(G: np.ndarray, r: np.ndarray, q: float, ℓ: float, n: float) -> np.ndarray:
    D = r.shape[0]
    out = np.empty(packed_len(D))
    out[: D * D] = np.asarray(G, dtype=np.float64).reshape(-1, order="F")
    out[D * D : D * D + D] = r
    out[D * D + D :] = (q, ℓ, n)
    return out


def unpack_stats(p: np.ndarray, D: int):
    return p[: D * D].reshape(D, D, order="F"), p[D * D : D * D + D], p[D * D + D], p[D * D + D + 1], p[D * D + D + 2]


def distributed_infer(local_stats: Callable[[int, int], np.ndarray], N: int, D: int, allreduce: Callable[[np.ndarray], np.ndarray],
                      rank: int, world: int, solve: Callable[[np.ndarray], object], align: int = 16):
    """Host-side control flow of the sharded path, with the three device stages injected:

        local_stats(lo, hi) -> packed statistics of observations [lo, hi)   (blr_stats_accumulate)
        allreduce(packed)   -> elementwise sum over ranks                    (blr_stats_allreduce / NCCL)
        solve(packed)       -> posterior + logpdf from reduced statistics    (blr_infer_from_stats)

    In production the stages are the C-ABI calls named on the right (see model._infer, which runs them inside
    blr_infer); the indirection exists so the N > 1 control flow is testable under gloo on CPU boxes.
    """
    lo, hi = ShardPlan(N, world, align).bounds(rank)
    packed = local_stats(lo, hi)
    if packed.shape[0] != packed_len(D):
        raise ValueError("packed statistics have the wrong length")
    return solve(allreduce(packed))
