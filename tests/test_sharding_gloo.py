"""CPU, world_size 2 over gloo: the host-side control flow of the N-sharded path --
partition, per-rank packed statistics, ONE sum-allreduce, replicated solve.  The device stages are stood in
for by the oracle (this container has no GPU); on the GPU box the same flow runs inside blr_infer with
blr_stats_accumulate / ncclAllReduce / blr_infer_from_stats (tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blr_b200.sharding import ShardPlan, packed_len
from tests.sharding_scaffold import distributed_infer, pack_stats, unpack_stats
from oracle import blr_oracle as ref


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(D=24, N=1000):
    rng = np.random.default_rng(42)
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N))
    mw = rng.standard_normal(D)
    B = rng.standard_normal((D, D))
    Λ = B @ B.T + np.eye(D)
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    return X, y, σ2, mw, Λ


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, y, σ2, mw, Λ = _problem()
    D, N = X.shape

    def local_stats(lo, hi):
        G, r, qq, ℓ = ref.gram_stats(X[:, lo:hi], y[lo:hi], σ2[lo:hi], mw)
        return pack_stats(G, r, qq, ℓ, hi - lo)

    def allreduce(p):
        t = torch.from_numpy(p.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    def solve(p):
        G, r, qq, ℓ, n = unpack_stats(p, D)
        assert n == N
        return ref.infer_from_stats(mw, Λ, G, r, qq, ℓ, N)

    lp, m, T = distributed_infer(local_stats, N, D, allreduce, rank, world, solve)
    q.put((rank, lp, m, T))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_plan_properties():
    for N in (0, 1, 15, 16, 17, 1000, 1 << 24):
        for world in (1, 2, 3, 8):
            b = ShardPlan(N, world).all_bounds()
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(lo % 16 == 0 for lo, _ in b if lo < N)
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 32 or N < 16 * world
    with pytest.raises(ValueError):
        ShardPlan(10, 2).bounds(2)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    G, r = rng.standard_normal((5, 5)), rng.standard_normal(5)
    p = pack_stats(G, r, 1.5, -2.5, 77)
    assert p.shape == (packed_len(5),)
    G2, r2, q, ℓ, n = unpack_stats(p, 5)
    assert np.array_equal(G, G2) and np.array_equal(r, r2) and (q, ℓ, n) == (1.5, -2.5, 77)


def test_world2_sharded_inference_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, y, σ2, mw, Λ = _problem()
    fx = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2)
    lp_ref, post = ref.logpdf(fx, y), ref.posterior(fx, y)
    for _, lp, m, T in res:
        assert abs(lp - lp_ref) <= 1e-11 * abs(lp_ref)
        assert np.linalg.norm(m - post.mw) <= 1e-11 * np.linalg.norm(post.mw)
        assert np.linalg.norm(T.T @ T - ref.dense(post.Λw)) <= 1e-12 * np.linalg.norm(ref.dense(post.Λw))
    # replicated solve: every rank holds bit-identical results
    assert res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])
