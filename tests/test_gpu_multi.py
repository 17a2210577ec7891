"""GPU, 2 ranks (skipped on a 1-GPU box): N-sharded posterior+logpdf with the library-owned NCCL allreduce
must equal the single-GPU result and the oracle; every rank must hold bit-identical posteriors."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %(root)r)
    import torch, torch.distributed as dist
    import blr_b200 as blr
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = blr.Context(local); blr.set_default_context(ctx); ctx.init_comm_from_torch()
    rng = np.random.default_rng(3)
    D, N = 192, 20011
    X = rng.standard_normal((D, N)); s2 = np.exp(rng.standard_normal(N)); mw = rng.standard_normal(D)
    B = rng.standard_normal((D, D)); Lam = B @ B.T + np.eye(D)
    y = X.T @ rng.standard_normal(D) + np.sqrt(s2) * rng.standard_normal(N)
    lo, hi = blr.ShardPlan(N, world).bounds(rank)
    f = blr.BayesianLinearRegressor(mw, Lam)
    post, lp = blr.posterior_and_logpdf(f(blr.ColVecs(X[:, lo:hi]), s2[lo:hi]), y[lo:hi])
    np.savez(os.path.join(%(out)r, f"rank{rank}.npz"), lp=lp, m=post.mw, L=post.Λw.dense())
    dist.barrier(); dist.destroy_process_group()
    """
)


def test_two_rank_sharded_inference(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert r0["lp"] == r1["lp"] and np.array_equal(r0["m"], r1["m"]) and np.array_equal(r0["L"], r1["L"])

    from oracle import blr_oracle as ref

    rng = np.random.default_rng(3)
    D, N = 192, 20011
    X = rng.standard_normal((D, N)); s2 = np.exp(rng.standard_normal(N)); mw = rng.standard_normal(D)
    B = rng.standard_normal((D, D)); Lam = B @ B.T + np.eye(D)
    y = X.T @ rng.standard_normal(D) + np.sqrt(s2) * rng.standard_normal(N)
    fx = ref.BayesianLinearRegressor(mw, Lam)(ref.ColVecs(X), s2)
    lp, post = ref.logpdf(fx, y), ref.posterior(fx, y)
    assert abs(r0["lp"] - lp) <= 1e-9 * abs(lp)
    assert np.linalg.norm(r0["m"] - post.mw) <= 1e-9 * np.linalg.norm(post.mw)
    assert np.linalg.norm(r0["L"] - ref.dense(post.Λw)) <= 1e-9 * np.linalg.norm(ref.dense(post.Λw))


def test_single_process_two_devices():
    """One process driving two GPUs (the shape of a Julia session): blr_comm_init_all + blr_stats_allreduce_all.
    Each context accumulates its shard, the grouped all-reduce sums them, every device solves the same posterior."""
    import ctypes as C

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import blr_b200 as blr
    from blr_b200 import _lib as L
    from oracle import blr_oracle as ref

    rng = np.random.default_rng(5)
    D, N = 256, 30011
    X = rng.standard_normal((D, N))
    s2 = np.exp(rng.standard_normal(N))
    mw = rng.standard_normal(D)
    y = X.T @ rng.standard_normal(D) + np.sqrt(s2) * rng.standard_normal(N)
    ctxs = [blr.Context(0), blr.Context(1)]
    blr.Context.init_comm_all(ctxs)
    stats, keep = [], []
    for r, ctx in enumerate(ctxs):
        lo, hi = blr.ShardPlan(N, 2).bounds(r)
        Xd = blr.DeviceMatrix.upload(ctx, X[:, lo:hi], 0)
        yd, sd = blr.DeviceVector.upload(ctx, y[lo:hi]), blr.DeviceVector.upload(ctx, s2[lo:hi])
        st = blr.Stats(ctx, D)
        noise = L.Noise(L.NOISE_VECTOR, 0.0, sd.handle, None, 0)
        ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, st.handle, np.ascontiguousarray(mw).ctypes.data_as(C.c_void_p), Xd.handle,
                                               yd.handle, C.byref(noise)))
        stats.append(st)
        keep.append((Xd, yd, sd))
    blr.Stats.allreduce_all(stats)
    f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(D)))
    outs = []
    for ctx, st in zip(ctxs, stats):
        prior, kp = f._prior_struct()
        lp = C.c_double()
        m = np.empty(D)
        ctx.check(ctx.lib.blr_infer_from_stats(ctx.handle, C.byref(prior), st.handle, C.byref(lp), m.ctypes.data_as(C.c_void_p), None, None, None))
        outs.append((lp.value, m))
    lpo, mo, _ = ref.infer_streaming(mw, ref.Diagonal(np.ones(D)), X, y, s2)
    assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])
    assert abs(outs[0][0] - lpo) <= 1e-9 * abs(lpo)
    assert np.linalg.norm(outs[0][1] - mo) / np.linalg.norm(mo) < 1e-9
