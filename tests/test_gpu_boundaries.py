"""GPU: every dispatch threshold of the library, at the threshold and one either side, against the oracle.

The path has a kernel family per regime -- thread-per-observation (D <= 8), fused warp kernel (D <= 16), per-warp TMA ring (even
D <= 64, aligned ColVecs, N >= 64), register-fed warp kernel (odd D, RowVecs), tiled TMA kernel (even D >= 64 / RowVecs D >= 128),
generic kernel (odd D > 64); marginals: tiny / small / TMA (D >= 128 even, N* >= 32) / generic; rand: generic / single-group /
two-group (D >= 64 even, N* >= 128), 64-sample blocks -- and the eligibility rules meet at D = 8, 16, 64, 128, at even / odd D and
at N = 32, 64, 128.  A parity bug at such a seam is invisible to tests that sit in the middle of a regime.  Here every seam is
crossed: posterior mean / precision, logpdf, marginal mean / variance and rand with supplied draws must all agree with the
reference's op sequence (oracle) to 1e-9, for both input layouts and for zero and non-zero prior means."""
import math

import numpy as np
import pytest

import blr_b200 as blr
from oracle import blr_oracle as ref

pytestmark = pytest.mark.gpu
RTOL = 1e-9

D_SEAMS = [7, 8, 9, 15, 16, 17, 18, 31, 32, 33, 34, 50, 62, 63, 64, 65, 66, 126, 127, 128, 129, 130]
N_SEAMS = [1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 257]


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def make(D, N, seed, zero_mean):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N))
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    mw = np.zeros(D) if zero_mean else rng.standard_normal(D)
    B = rng.standard_normal((D, D)) / math.sqrt(D)
    return X, σ2, y, mw, B @ B.T + np.eye(D)


@pytest.mark.parametrize("layout", ["col", "row"])
@pytest.mark.parametrize("D", D_SEAMS)
def test_inference_across_dispatch_seams(D, layout):
    """posterior + logpdf for every N seam (device-resident inputs, so the device kernels -- not the host-streaming chunker --
    see exactly these shapes)."""
    ctx = blr.default_context()
    worst = {}
    for N in N_SEAMS:
        for zero_mean in (False, True):
            X, σ2, y, mw, Λ = make(D, N, seed=1000 * D + N, zero_mean=zero_mean)
            Xd = blr.DeviceMatrix.upload(ctx, X if layout == "col" else np.ascontiguousarray(X.T), 0 if layout == "col" else 1)
            x = blr.ColVecs(Xd) if layout == "col" else blr.RowVecs(Xd)
            post, lp = blr.posterior_and_logpdf(blr.BayesianLinearRegressor(mw, Λ)(x, blr.DeviceVector.upload(ctx, σ2)),
                                                blr.DeviceVector.upload(ctx, y))
            fo = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2)
            po = ref.posterior(fo, y)
            e = {"logpdf": abs(lp - ref.logpdf(fo, y)) / abs(lp), "mean": relerr(post.mw, po.mw),
                 "precision": relerr(post.Λw.dense(), ref.dense(po.Λw))}
            for k, v in e.items():
                worst[k] = max(worst.get(k, 0.0), v)
                assert v < RTOL, (D, N, layout, zero_mean, k, v)
    print(f"[seams] D={D} {layout}: worst " + " ".join(f"{k} {v:.1e}" for k, v in worst.items()))


@pytest.mark.parametrize("D", D_SEAMS)
def test_prediction_across_dispatch_seams(D):
    """marginals and rand (supplied draws, S at and around the 64-sample block) for every N* seam, ColVecs and RowVecs."""
    rng = np.random.default_rng(D)
    Nfit = 3 * D + 5
    X, σ2, y, mw, Λ = make(D, Nfit, seed=77 * D, zero_mean=False)
    post = blr.posterior(blr.BayesianLinearRegressor(mw, Λ)(blr.ColVecs(X), σ2), y)
    po = ref.posterior(ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2), y)
    for Nt in N_SEAMS:
        Xt = rng.standard_normal((D, Nt))
        σt = np.exp(rng.standard_normal(Nt))
        mo, vo = ref.mean_and_var(po(ref.ColVecs(Xt), σt))
        for layout in ("col", "row"):
            xt = blr.ColVecs(Xt) if layout == "col" else blr.RowVecs(np.ascontiguousarray(Xt.T))
            m, v = blr.mean_and_var(post(xt, σt))
            assert relerr(m, mo) < RTOL and relerr(v, vo) < RTOL, (D, Nt, layout, relerr(m, mo), relerr(v, vo))
        for S in ((1, 64, 65) if Nt in (127, 128, 129, 257) else (3,)):
            Zw, Zy = rng.standard_normal((D, S)), rng.standard_normal((Nt, S))
            Y = blr.rand_with_draws(post(blr.ColVecs(Xt), σt), Zw, Zy)
            Yo = ref.rand(po(ref.ColVecs(Xt), σt), Zw, Zy)
            assert relerr(Y, Yo) < RTOL, (D, Nt, S, relerr(Y, Yo))


@pytest.mark.parametrize("D", [66, 68, 72, 74, 80, 82, 88, 90, 96])
@pytest.mark.parametrize("N", [64, 100, 1000, 40_003])
def test_mid_ring_kernel_statistics(D, N):
    """64 < D <= 96 (even, aligned ColVecs): the team-of-two-warps ring kernel (csrc/gram_mid.cu).  Every block-grid width
    (9 .. 12 block rows, with and without a split leftover row), D on and off a multiple of 8, dense and strided observations
    (ld = D and ld = D + 6), heteroscedastic and scalar noise, zero and non-zero prior mean, N from a single stage to several
    ring revolutions per team with a ragged tail -- statistics against the oracle's Gram form, and against the same library with
    the kernel switched off (BLR_MID_RING=0: one padded tile of K1)."""
    import ctypes as C
    import os

    from blr_b200 import _lib as L

    rng = np.random.default_rng(D * 1000 + N)
    ctx = blr.default_context()
    os.environ["BLR_MID_RING"] = "0"
    try:
        ctx_off = blr.Context(0)
    finally:
        del os.environ["BLR_MID_RING"]
    for ld, scalar, zero_mean in ((D, False, False), (D + 6, False, True), (D, True, False)):
        Xp = np.zeros((ld, N), order="F")
        Xp[:D] = rng.standard_normal((D, N))
        X = Xp[:D]
        σ2 = np.full(N, 0.37) if scalar else np.exp(rng.standard_normal(N))
        y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
        mw = np.zeros(D) if zero_mean else rng.standard_normal(D)
        Go, ro, qo, lo = ref.gram_stats(X, y, σ2, mw)
        out = []
        for c in (ctx, ctx_off):
            Xd = blr.DeviceMatrix.upload(c, Xp, 0)                       # ld x N on the device ...
            ptr, _ = Xd.device_ptr()
            xh = C.c_void_p()                                             # ... viewed as D x N with leading dimension ld
            c.check(c.lib.blr_x_wrap_device(c.handle, C.c_void_p(ptr), D, N, ld, L.COLVECS, C.byref(xh)))
            st = blr.Stats(c, D)
            yv = blr.DeviceVector.upload(c, y)
            s2v = blr.DeviceVector.upload(c, σ2)
            noise = L.Noise(L.NOISE_SCALAR, 0.37, None, None, 0) if scalar else L.Noise(L.NOISE_VECTOR, 0.0, s2v.handle, None, 0)
            c.check(c.lib.blr_stats_accumulate(c.handle, st.handle, mw.ctypes.data_as(C.c_void_p), xh, yv.handle, C.byref(noise)))
            G, r, q, ell, n = st.unpack()
            c.check(c.lib.blr_x_free(c.handle, xh))
            assert n == N and np.array_equal(G, G.T)
            assert relerr(G, Go) < 1e-12 and relerr(r, ro) < 1e-11, (D, N, ld, relerr(G, Go), relerr(r, ro))
            assert abs(q - qo) <= 1e-11 * abs(qo) and abs(ell - lo) <= 1e-11 * max(abs(lo), 1.0)
            out.append((G, r, q, ell))
        assert relerr(out[0][0], out[1][0]) < 1e-12 and relerr(out[0][1], out[1][1]) < 1e-11
