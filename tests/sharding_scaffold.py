"""TEST INFRASTRUCTURE: the host-side control flow of the N-sharded path with the three device stages injected, so that the
N > 1 flow (partition -> per-rank packed statistics -> ONE sum-allreduce -> replicated solve) is testable under gloo on a CPU
box (tests/test_sharding_gloo.py).  The product never calls this: model._infer runs the same three stages inside blr_infer
(blr_stats_accumulate / ncclAllReduce / blr_infer_from_stats)."""
from typing import Callable

import numpy as np

from blr_b200.sharding import ShardPlan, packed_len


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

    In production the stages are the C-ABI calls named on the right (model._infer runs them inside blr_infer).
    """
    lo, hi = ShardPlan(N, world, align).bounds(rank)
    packed = local_stats(lo, hi)
    if packed.shape[0] != packed_len(D):
        raise ValueError("packed statistics have the wrong length")
    return solve(allreduce(packed))
