"""GPU parity: the CUDA path (through the C ABI, via the host mirror) against the CPU oracle on identical inputs.

Tolerance: BASELINE.json's north_star asks for relative error <= 1e-9 on posterior mean, precision, logpdf and
marginal variances (norm-wise for arrays, plain relative for the scalar), and the same for rand under identical
standard-normal draws.  RTOL below is that bound; most quantities land near 1e-13.
"""
import math
import os

import numpy as np
import pytest

import blr_b200 as blr
from oracle import blr_oracle as ref
from tests.toy import make_phi

pytestmark = pytest.mark.gpu

RTOL = 1e-9
EPS = np.finfo(np.float64).eps


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def problem(D, N, seed=0, dense=True, zero_mean=False, scalar_noise=False):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((D, N))
    mw = np.zeros(D) if zero_mean else rng.standard_normal(D)
    if dense:
        B = rng.standard_normal((D, D))
        Λ = B @ B.T + np.eye(D)
    else:
        Λ = None
    σ2 = 0.37 if scalar_noise else np.exp(rng.standard_normal(N))
    w = rng.standard_normal(D)
    y = X.T @ w + np.sqrt(σ2) * rng.standard_normal(N)
    return X, mw, Λ, σ2, y


def both_priors(mw, Λ, D):
    if Λ is None:
        return blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(D))), ref.BayesianLinearRegressor(mw, ref.Diagonal(np.ones(D)))
    return blr.BayesianLinearRegressor(mw, Λ), ref.BayesianLinearRegressor(mw, Λ)


# ------------------------------------------------------------------------------------------------ statistics (K0/K1)
STAT_CASES = [
    # D, N, zero_mean, scalar_noise, layout
    (64, 16, True, False, "col"),
    (64, 1000, False, False, "col"),
    (128, 4096, True, True, "col"),
    (192, 1003, False, True, "col"),     # scalar noise: unscaled Gram kernel (1/σ² applied by the reduction), partial last stage
    (1024, 2049, True, True, "col"),
    (130, 515, False, False, "col"),     # D tail inside a 128 tile, N tail inside a stage
    (256, 20000, False, False, "col"),
    (320, 7777, True, False, "col"),
    (1024, 5000, False, False, "col"),
    (256, 3001, False, False, "row"),    # RowVecs, odd N -> transposed staging
    (320, 4096, False, False, "row"),    # RowVecs, even N -> native feature-major TMA ring
    (130, 516, True, False, "row"),      # native, narrow last tile
    (257, 1000, False, True, "row"),     # native, odd D
    (1024, 2048, False, False, "row"),
    (2, 10, True, False, "col"),         # generic path
    (3, 11, False, False, "row"),
    (7, 13, False, True, "col"),
    (33, 257, False, False, "col"),
    (63, 1000, False, False, "row"),
    (65, 300, False, False, "col"),      # odd D >= 64 -> generic
]


@pytest.mark.parametrize("D,N,zero_mean,scalar_noise,layout", STAT_CASES)
def test_sufficient_statistics(D, N, zero_mean, scalar_noise, layout):
    X, mw, _, σ2, y = problem(D, N, seed=D * 7 + N, zero_mean=zero_mean, scalar_noise=scalar_noise)
    ctx = blr.default_context()
    Xd = blr.DeviceMatrix.upload(ctx, X if layout == "col" else np.ascontiguousarray(X.T), 0 if layout == "col" else 1)
    yd = blr.DeviceVector.upload(ctx, y)
    from blr_b200.runtime import make_noise
    import ctypes as C

    noise, keep = make_noise(ctx, σ2, N)
    st = blr.Stats(ctx, D)
    mwc = np.ascontiguousarray(mw)
    for rep in range(2):  # accumulate twice: statistics must add up (streamed conditioning)
        ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, st.handle, mwc.ctypes.data_as(C.c_void_p), Xd.handle, yd.handle,
                                               C.byref(noise)))
    G, r, q, ℓ, n = st.unpack()
    Go, ro, qo, ℓo = ref.gram_stats(X, y, σ2, mw, chunk=4096)
    assert n == 2 * N
    assert relerr(G, 2 * Go) < 1e-12
    assert np.array_equal(G, G.T)
    assert relerr(r, 2 * ro) < 1e-11
    assert abs(q - 2 * qo) <= 1e-11 * abs(2 * qo)
    assert abs(ℓ - 2 * ℓo) <= 1e-11 * max(abs(2 * ℓo), 1.0)


@pytest.mark.parametrize("D", [18, 24, 32, 40, 48, 50, 56, 62, 64])
@pytest.mark.parametrize("N,zero_mean,scalar_noise,pad", [(1003, False, False, 0), (40013, True, True, 0), (5000, False, False, 6)])
def test_small_d_ring_kernel(D, N, zero_mean, scalar_noise, pad, monkeypatch):
    """K1s with the per-warp TMA ring (16 < D <= 64, D even, aligned ColVecs): every block width ceil(D / 8) = 3..8, D not a multiple of 8 (the reference's D = 50 example shape), ragged N
    (partial last stage, warps without work), dense columns (one bulk copy per stage) and a padded leading dimension (one copy
    per observation), against the oracle -- and against the register-fed kernel it replaces (BLR_SMALL_RING=0)."""
    import ctypes as C

    import torch

    from blr_b200.runtime import make_noise

    X, mw, _, σ2, y = problem(D, N, seed=D * 11 + N, zero_mean=zero_mean, scalar_noise=scalar_noise)
    Go, ro, qo, ℓo = ref.gram_stats(X, y, σ2, mw, chunk=4096)
    got = []
    for ring in ("1", "0"):
        monkeypatch.setenv("BLR_SMALL_RING", ring)
        ctx = blr.Context(0)
        if pad:
            Xt = torch.zeros((N, D + pad), dtype=torch.float64, device="cuda")  # ColVecs with ld = D + pad (even)
            Xt[:, :D] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
            Xd = blr.DeviceMatrix.wrap_torch(ctx, Xt[:, :D], 0)
        else:
            Xd = blr.DeviceMatrix.upload(ctx, X, 0)
        yd = blr.DeviceVector.upload(ctx, y)
        noise, keep = make_noise(ctx, σ2, N)
        st = blr.Stats(ctx, D)
        mwc = np.ascontiguousarray(mw)
        ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, st.handle, mwc.ctypes.data_as(C.c_void_p), Xd.handle, yd.handle, C.byref(noise)))
        G, r, q, ℓ, n = st.unpack()
        assert n == N and np.array_equal(G, G.T)
        assert relerr(G, Go) < 1e-12 and relerr(r, ro) < 1e-11
        assert abs(q - qo) <= 1e-11 * abs(qo) and abs(ℓ - ℓo) <= 1e-11 * max(abs(ℓo), 1.0)
        got.append((G, r))
    assert relerr(got[0][0], got[1][0]) < 1e-13 and relerr(got[0][1], got[1][1]) < 1e-12


def test_statistics_bit_reproducible():
    X, mw, _, σ2, y = problem(256, 30011, seed=5)
    ctx = blr.default_context()
    outs = []
    for _ in range(2):
        f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(256)))
        post, lp = blr.posterior_and_logpdf(f(blr.ColVecs(X), σ2), y)
        outs.append((post.mw.copy(), lp))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]


@pytest.mark.parametrize("D,N,period", [(256, 20011, 512), (384, 9000, 256), (1024, 6000, 128), (130, 5000, 64)])
def test_periodic_schedule_many_periods(D, N, period, monkeypatch):
    """The Gram kernel repeats its stream-K cut every `period` observations (soft barrier between periods, partial
    tiles parked in the workspace by CTAs that switch tiles).  Force tiny periods so that many of them, the
    read-modify-write flushes and the partial last period are all exercised at test sizes."""
    X, mw, _, σ2, y = problem(D, N, seed=D + period)
    monkeypatch.setenv("BLR_GRAM_PERIOD_OBS", str(period))
    ctx = blr.Context(0)  # reads the environment at creation
    monkeypatch.delenv("BLR_GRAM_PERIOD_OBS")
    f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(D)))
    fx = f(blr.ColVecs(X), σ2)
    fx.ctx = ctx
    post, lp = blr.posterior_and_logpdf(fx, y)
    fxo = ref.BayesianLinearRegressor(mw, ref.Diagonal(np.ones(D)))(ref.ColVecs(X), σ2)
    assert abs(lp - ref.logpdf(fxo, y)) <= RTOL * abs(lp)
    po = ref.posterior(fxo, y)
    assert relerr(post.mw, po.mw) < RTOL and relerr(post.Λw.dense(), ref.dense(po.Λw)) < RTOL
    post2, lp2 = blr.posterior_and_logpdf(fx, y)  # bit-reproducible
    assert lp2 == lp and np.array_equal(post.mw, post2.mw)


@pytest.mark.parametrize("D,N,period", [(256, 20011, 0), (130, 5000, 0), (320, 7777, 0), (1024, 6000, 0), (384, 9000, 256)])
@pytest.mark.parametrize("cs", [0, 1])
def test_gram_consumer_tilings(D, N, period, cs, monkeypatch):
    """Both consumer tilings of the Gram fast path (2 x 4 warps with 64 x 32 warp tiles; column strips = 1 x 8 warps with
    128 x 16 warp tiles, half the B-fragment scalings) against the oracle, including a D tail inside a tile, a narrow last
    tile row and the periodic schedule's read-modify-write flushes; results are bit-reproducible for either."""
    X, mw, _, σ2, y = problem(D, N, seed=3 * D + cs)
    monkeypatch.setenv("BLR_GRAM_CS", str(cs))
    if period:
        monkeypatch.setenv("BLR_GRAM_PERIOD_OBS", str(period))
    ctx = blr.Context(0)  # reads the environment at creation
    monkeypatch.delenv("BLR_GRAM_CS")
    if period:
        monkeypatch.delenv("BLR_GRAM_PERIOD_OBS")
    Xd = blr.DeviceMatrix.upload(ctx, X, 0)
    yd = blr.DeviceVector.upload(ctx, y)
    sd = blr.DeviceVector.upload(ctx, σ2)
    from blr_b200 import _lib as L
    import ctypes as C

    noise = L.Noise(L.NOISE_VECTOR, 0.0, sd.handle, None, 0)
    outs = []
    for rep in range(2):
        st = blr.Stats(ctx, D)
        ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, st.handle, np.ascontiguousarray(mw).ctypes.data_as(C.c_void_p), Xd.handle,
                                               yd.handle, C.byref(noise)))
        outs.append(st.unpack())
    G, r, q, ℓ, n = outs[0]
    Go, ro, qo, ℓo = ref.gram_stats(X, y, σ2, mw, chunk=4096)
    assert n == N and np.array_equal(G, G.T)
    assert relerr(G, Go) < 1e-12 and relerr(r, ro) < 1e-11
    assert abs(q - qo) <= 1e-11 * abs(qo) and abs(ℓ - ℓo) <= 1e-11 * max(abs(ℓo), 1.0)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


# ------------------------------------------------------------------------------------------------ posterior + logpdf
INFER_CASES = [
    # D, N, dense prior, zero mean, scalar noise
    (2, 10, False, True, False),
    (7, 13, True, False, False),
    (33, 257, True, False, True),
    (64, 500, True, False, False),
    (70, 400, True, False, False),
    (100, 999, False, False, False),
    (130, 515, True, False, False),
    (200, 3000, True, False, False),
    (256, 20000, False, True, False),
    (512, 6000, True, False, False),
    (1024, 4096, True, False, False),
]


@pytest.mark.parametrize("D,N,dense,zero_mean,scalar_noise", INFER_CASES)
@pytest.mark.parametrize("Tx", ["ColVecs", "RowVecs"])
def test_posterior_and_logpdf_match_oracle(D, N, dense, zero_mean, scalar_noise, Tx):
    if Tx == "RowVecs" and D > 256:
        pytest.skip("RowVecs covered at the smaller sizes")
    X, mw, Λ, σ2, y = problem(D, N, seed=11 * D + N, dense=dense, zero_mean=zero_mean, scalar_noise=scalar_noise)
    f, fo = both_priors(mw, Λ, D)
    x = blr.ColVecs(X) if Tx == "ColVecs" else blr.RowVecs(np.ascontiguousarray(X.T))
    fx, fxo = f(x, σ2), fo(ref.ColVecs(X), σ2)
    lp_o = ref.logpdf(fxo, y)
    post_o = ref.posterior(fxo, y)
    lp = blr.logpdf(fx, y)
    post = blr.posterior(fx, y)
    post2, lp2 = blr.posterior_and_logpdf(fx, y)
    assert abs(lp - lp_o) <= RTOL * abs(lp_o), (lp, lp_o)
    assert lp2 == lp
    assert relerr(post.mw, post_o.mw) < RTOL
    assert np.array_equal(post.mw, post2.mw)
    assert relerr(post.Λw.dense(), ref.dense(post_o.Λw)) < RTOL
    assert isinstance(post.Λw, blr.Symmetric)  # src/bayesian_linear_regression.jl:92


@pytest.mark.parametrize("D,N", [(7, 13), (256, 5000), (320, 2049)])
def test_device_resident_inputs(D, N):
    """blr_infer on handles that already live on the device (DeviceMatrix / DeviceVector / CUDA tensors): the path the
    benchmark times.  numpy inputs take the host-streaming entry point instead; both must agree bit for bit."""
    import torch

    X, mw, Λ, σ2, y = problem(D, N, seed=D + N)
    ctx = blr.default_context()
    f = blr.BayesianLinearRegressor(mw, Λ)
    post_h, lp_h = blr.posterior_and_logpdf(f(blr.ColVecs(X), σ2), y)
    Xd = blr.DeviceMatrix.upload(ctx, X, 0)
    post_d, lp_d = blr.posterior_and_logpdf(f(blr.ColVecs(Xd), blr.DeviceVector.upload(ctx, σ2)), blr.DeviceVector.upload(ctx, y))
    Xt = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()  # (N, D) row-major == D x N column-major
    post_t, lp_t = blr.posterior_and_logpdf(f(blr.ColVecs(Xt), torch.from_numpy(σ2).cuda()), torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    fxo = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2)
    assert abs(lp_d - ref.logpdf(fxo, y)) <= RTOL * abs(lp_d)
    assert lp_d == lp_t and np.array_equal(post_d.mw, post_t.mw)
    if N <= (1 << 16):  # a single host chunk sees the same kernel launch as the resident path
        assert lp_h == lp_d and np.array_equal(post_h.mw, post_d.mw)
    assert relerr(post_d.mw, ref.posterior(fxo, y).mw) < RTOL


def test_resident_then_streamed_on_one_context():
    """A device-resident inference at D >= 1024 (precision download overlapped on the copy stream) followed by a host-
    streamed one on the same context: both share the copy stream and its hand-over events."""
    D, N = 1024, 3000
    X, mw, _, σ2, y = problem(D, N, seed=21, dense=False)
    ctx = blr.Context(0)
    f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(D)))
    fx = f(blr.ColVecs(blr.DeviceMatrix.upload(ctx, X, 0)), blr.DeviceVector.upload(ctx, σ2))
    fx.ctx = ctx
    post_d, lp_d = blr.posterior_and_logpdf(fx, blr.DeviceVector.upload(ctx, y))
    post_s, lp_s = blr.posterior_and_logpdf_streamed(f, X, y, σ2, chunk=1024, ctx=ctx)
    post_d2, lp_d2 = blr.posterior_and_logpdf(fx, blr.DeviceVector.upload(ctx, y))
    lpo, mo, To = ref.infer_streaming(mw, ref.Diagonal(np.ones(D)), X, y, σ2)
    for post, lp in ((post_d, lp_d), (post_s, lp_s), (post_d2, lp_d2)):
        assert abs(lp - lpo) <= RTOL * abs(lpo)
        assert relerr(post.mw, mo) < RTOL and relerr(post.Λw.dense(), To.T @ To) < RTOL
    assert lp_d2 == lp_d and np.array_equal(post_d.Λw.dense(), post_d2.Λw.dense())


def test_large_rowvecs_blockwise_staging():
    """A device-resident RowVecs matrix with more than 2^18 observations is transposed block by block."""
    D, N = 64, (1 << 18) + 4099
    X, mw, _, σ2, y = problem(D, N, seed=77, dense=False)
    ctx = blr.default_context()
    f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(D)))
    Xr = blr.DeviceMatrix.upload(ctx, np.ascontiguousarray(X.T), 1)
    post, lp = blr.posterior_and_logpdf(f(blr.RowVecs(Xr), blr.DeviceVector.upload(ctx, σ2)), blr.DeviceVector.upload(ctx, y))
    lpo, mo, To = ref.infer_streaming(mw, ref.Diagonal(np.ones(D)), X, y, σ2)
    assert abs(lp - lpo) <= RTOL * abs(lpo)
    assert relerr(post.mw, mo) < RTOL and relerr(post.Λw.dense(), To.T @ To) < RTOL


def test_pdmat_closure_and_factor():
    """src/bayesian_linear_regression.jl:93 + test/bayesian_linear_regression.jl:71-113."""
    X, mw, Λ, σ2, y = problem(96, 700, seed=3)
    f_pd = blr.BayesianLinearRegressor(mw, blr.PDMat(Λ))
    f_sym = blr.BayesianLinearRegressor(mw, blr.Symmetric(Λ))
    p_pd, p_sym = blr.posterior(f_pd(X, σ2), y), blr.posterior(f_sym(X, σ2), y)
    assert isinstance(p_pd.Λw, blr.PDMat) and isinstance(p_sym.Λw, blr.Symmetric)
    T = p_pd.Λw.U
    assert np.allclose(T, np.triu(T)) and np.all(np.diag(T) > 0)
    fxo = ref.BayesianLinearRegressor(mw, ref.PDMat.from_matrix(Λ))(ref.ColVecs(X), σ2)
    assert relerr(T, ref.posterior_factor_T(fxo, y)) < RTOL
    assert relerr(T.T @ T, p_pd.Λw.mat) < 1e-12
    Xp = np.random.default_rng(1).standard_normal((96, 50))
    assert relerr(blr.mean(p_pd(Xp, σ2[:50])), blr.mean(p_sym(Xp, σ2[:50]))) < 1e-12


def test_repeated_conditioning():
    """test/bayesian_linear_regression.jl:49-70 with diagonal noise: conditioning in two steps == once."""
    D, N = 48, 400
    X, mw, Λ, σ2, y = problem(D, N, seed=9)
    f = blr.BayesianLinearRegressor(mw, Λ)
    N1 = N - 133
    f1 = blr.posterior(f(X[:, :N1], σ2[:N1]), y[:N1])
    f2 = blr.posterior(f1(X[:, N1:], σ2[N1:]), y[N1:])
    fp = blr.posterior(f(X, σ2), y)
    Xp = np.random.default_rng(2).standard_normal((D, 64))
    m1, v1 = blr.mean_and_var(fp(Xp, 0.1))
    m2, v2 = blr.mean_and_var(f2(Xp, 0.1))
    assert relerr(m1, m2) < 1e-9 and relerr(v1, v2) < 1e-9
    assert relerr(blr.cov(fp(Xp, 0.1)), blr.cov(f2(Xp, 0.1))) < 1e-9


def test_posterior_low_noise():
    """test/bayesian_linear_regression.jl:40-48: noise eps() -> interpolation and vanishing covariance."""
    D, N = 7, 13
    rng = np.random.default_rng(123456)
    X = rng.standard_normal((D, N))
    B = rng.standard_normal((D, D))
    f = blr.BayesianLinearRegressor(rng.standard_normal(D), B @ B.T + np.eye(D))
    y = blr.rand(rng, f(X, EPS))
    fp = blr.posterior(f(X, EPS), y)
    m = blr.mean(fp(X, EPS))
    assert np.linalg.norm(m - y) <= math.sqrt(EPS) * max(np.linalg.norm(m), np.linalg.norm(y))
    assert np.all(blr.cov(fp(X, EPS)) < 1000 * EPS)


# ------------------------------------------------------------------------------------------------ prediction / sampling
PRED_CASES = [(2, 1000, False), (7, 13, True), (64, 300, True), (130, 515, True), (256, 2000, True), (512, 1111, True),
              (768, 700, True), (1024, 333, True), (128, 31, True)]


@pytest.mark.parametrize("D,N,dense", PRED_CASES)
@pytest.mark.parametrize("Tx", ["ColVecs", "RowVecs"])
def test_mean_var_cov_match_oracle(D, N, dense, Tx):
    X, mw, Λ, σ2, _ = problem(D, N, seed=31 * D + N, dense=dense)
    f, fo = both_priors(mw, Λ, D)
    x = blr.ColVecs(X) if Tx == "ColVecs" else blr.RowVecs(np.ascontiguousarray(X.T))
    fx, fxo = f(x, σ2), fo(ref.ColVecs(X), σ2)
    m, v = blr.mean_and_var(fx)
    mo, vo = ref.mean_and_var(fxo)
    assert relerr(m, mo) < RTOL and relerr(v, vo) < RTOL
    # mean alone runs the streaming GEMV kernel, mean_and_var the fused one: same numbers up to summation order
    assert relerr(blr.mean(fx), m) < 1e-13 and np.array_equal(blr.var(fx), v)
    mm, ss = blr.marginals(fx)
    assert relerr(ss, np.sqrt(vo)) < RTOL
    if N <= 600:
        Cg = blr.cov(fx)
        assert relerr(Cg, ref.cov(fxo)) < RTOL
        assert relerr(np.diag(Cg), v) < 1e-12


@pytest.mark.parametrize("D,N,S", [(2, 10, 5), (7, 13, 1), (64, 300, 8), (200, 1000, 64), (256, 513, 3), (512, 2000, 70),
                                   (130, 129, 65)])
def test_rand_matches_oracle_with_same_draws(D, N, S):
    X, mw, Λ, σ2, _ = problem(D, N, seed=17 * D + S)
    f, fo = both_priors(mw, Λ, D)
    rng = np.random.default_rng(77)
    Zw, Zy = rng.standard_normal((D, S)), rng.standard_normal((N, S))
    Y = blr.rand_with_draws(f(blr.ColVecs(X), σ2), Zw, Zy)
    Yo = ref.rand(fo(ref.ColVecs(X), σ2), Zw, Zy)
    assert Y.shape == (N, S)
    assert relerr(Y, Yo) < RTOL
    # draw order contract (src/bayesian_linear_regression.jl:51-52): Zw first, then Zy, column-major fills
    r1, r2 = np.random.default_rng(5), np.random.default_rng(5)
    Y1 = blr.rand(r1, f(blr.ColVecs(X), σ2), S)
    Zw2 = np.asfortranarray(r2.standard_normal((S, D)).T)
    Zy2 = np.asfortranarray(r2.standard_normal((S, N)).T)
    assert relerr(Y1, ref.rand(fo(ref.ColVecs(X), σ2), Zw2, Zy2)) < RTOL
    assert blr.rand(np.random.default_rng(5), f(blr.ColVecs(X), σ2)).shape == (N,)


def test_rand_device_draws_same_stream_across_kernels(monkeypatch):
    """The three K7 paths -- single-group kernel with the draws in its epilogue, two-group kernel with the draws in its epilogue,
    two-group kernel fed by the stand-alone draw pass chunk by chunk (the default) -- use the same Philox counters: for one seed
    they must produce the same N x S sample matrix (ragged N and S, more than one chunk of draws)."""
    D, N, S = 128, 40000, 70
    X, mw, Λ, σ2, _ = problem(D, N, seed=5)
    outs = []
    for pp, unf in (("0", "0"), ("2", "0"), ("2", "1")):
        monkeypatch.setenv("BLR_RAND_PP", pp)
        monkeypatch.setenv("BLR_RAND_UNFUSED", unf)
        ctx = blr.Context(0)
        f = blr.BayesianLinearRegressor(mw, Λ)
        fx = f(blr.ColVecs(blr.DeviceMatrix.upload(ctx, X, 0)), blr.DeviceVector.upload(ctx, σ2))
        fx.ctx = ctx
        outs.append(blr.rand(blr.DeviceRNG(11), fx, S))
    assert np.isfinite(outs[0]).all() and outs[0].shape == (N, S)
    assert relerr(outs[1], outs[0]) < 1e-13 and relerr(outs[2], outs[0]) < 1e-13


def test_rand_device_rng_moments():
    """test/bayesian_linear_regression.jl:11-21 on the device generator (Philox): moments of 2e5 samples."""
    D, N, S = 3, 11, 200_000
    X, mw, Λ, σ2, _ = problem(D, N, seed=4)
    f, fo = both_priors(mw, Λ, D)
    Y = blr.rand(blr.DeviceRNG(1234), f(blr.ColVecs(X), σ2), S)
    fxo = fo(ref.ColVecs(X), σ2)
    m_emp = Y.mean(axis=1)
    Yc = Y - m_emp[:, None]
    np.testing.assert_allclose(ref.mean(fxo), m_emp, atol=2.5e-2, rtol=2.5e-2)
    np.testing.assert_allclose(ref.cov(fxo), Yc @ Yc.T / S, atol=6e-2, rtol=2.5e-2)


def test_function_samples():
    """test/sampling_functions.jl:3-47."""
    D, N = 5, 11
    X, mw, Λ, σ2, _ = problem(D, N, seed=8)
    f, fo = both_priors(mw, Λ, D)
    rng = np.random.default_rng(123456)
    z = np.random.default_rng(123456).standard_normal((1, D)).T
    g = blr.rand(rng, f)
    assert isinstance(g, blr.BLRFunctionSample)
    assert relerr(g.w, ref.rand_weights(fo, z[:, 0])) < RTOL
    assert np.array_equal(g(X), g(X))
    assert np.array_equal(g(X), g(blr.ColVecs(X)))
    assert relerr(g(blr.RowVecs(np.ascontiguousarray(X.T))), g(X)) < 1e-14
    assert relerr(g(X), X.T @ g.w) < 1e-13
    gs = blr.rand(rng, f, 30, 40)
    assert gs.shape == (30, 40)
    A = np.empty((4, 5), dtype=object)
    assert blr.rand_into(rng, A, f) is A and isinstance(A[3, 4], blr.BLRFunctionSample)
    W = np.stack([h.w for h in blr.rand(np.random.default_rng(3), f, 20000).reshape(-1)], axis=1)
    Yf = X.T @ W
    Yc = Yf - Yf.mean(axis=1, keepdims=True)
    fxo = fo(ref.ColVecs(X), σ2)
    np.testing.assert_allclose(ref.mean(fxo), Yf.mean(axis=1), atol=5e-2, rtol=5e-2)
    np.testing.assert_allclose(ref.cov(fxo), Yc @ Yc.T / 20000 + np.diag(σ2), atol=0.15, rtol=6e-2)


# ------------------------------------------------------------------------------------------------ BasisFunctionRegressor
def test_doctest_golden_vector_on_gpu():
    """src/basis_function_regression.jl:11-28: var(bfr(x)) == [2.0, 1.25, 1.0, 1.25, 2.0]."""
    ϕ = make_phi(blr)
    x = blr.RowVecs(np.linspace(-1.0, 1.0, 5)[:, None])
    bfr = blr.BasisFunctionRegressor(blr.BayesianLinearRegressor(np.zeros(2), blr.Diagonal(np.ones(2))), ϕ)
    np.testing.assert_allclose(blr.var(bfr(x)), [2.0, 1.25, 1.0, 1.25, 2.0], rtol=1e-15, atol=0)


@pytest.mark.parametrize("Tx", ["Matrix", "ColVecs", "RowVecs"])
def test_bfr_consistent_with_blr(Tx):
    """test/basis_function_regression.jl:13-28."""
    ϕ, ϕo = make_phi(blr), make_phi(ref)
    rng = np.random.default_rng(123456)
    N, D = 11, 2
    Xm = rng.standard_normal((D, N))
    B = rng.standard_normal((D, D))
    mw, Λ, σ2 = rng.standard_normal(D), B @ B.T + np.eye(D), np.exp(rng.standard_normal(N))
    X = {"Matrix": Xm, "ColVecs": blr.ColVecs(Xm), "RowVecs": blr.RowVecs(np.ascontiguousarray(Xm.T))}[Tx]
    f = blr.BayesianLinearRegressor(mw, Λ)
    f_bf = blr.BasisFunctionRegressor(f, ϕ)
    y = blr.rand(rng, f_bf(X, σ2))
    lp_bf, lp = blr.logpdf(f_bf(X, σ2), y), blr.logpdf(f(ϕ(X), σ2), y)
    assert lp_bf == lp
    fo_bf = ref.BasisFunctionRegressor(ref.BayesianLinearRegressor(mw, Λ), ϕo)
    assert abs(lp_bf - ref.logpdf(fo_bf(ref.ColVecs(Xm), σ2), y)) <= RTOL * abs(lp)
    post_bf = blr.posterior(f_bf(X, σ2), y)
    assert isinstance(post_bf, blr.BasisFunctionRegressor)
    assert relerr(blr.mean(post_bf(X)), blr.mean(blr.posterior(f(ϕ(X), σ2), y)(ϕ(X)))) < 1e-12
    assert relerr(blr.mean(post_bf(X)), ref.mean(ref.posterior(fo_bf(ref.ColVecs(Xm), σ2), y)(ref.ColVecs(Xm)))) < RTOL


def test_rff_features_resident_on_device():
    """BASELINE config 5 in miniature: ϕ(x) = sqrt(2/D) cos(Wx + b) evaluated on the device, output never leaves it."""
    rng = np.random.default_rng(21)
    din, D, N = 8, 192, 3000
    x = rng.standard_normal((din, N))
    W, b = rng.standard_normal((D, din)), rng.uniform(0, 2 * np.pi, D)
    σ2 = np.exp(rng.standard_normal(N))
    Φ = np.sqrt(2.0 / D) * np.cos(W @ x + b[:, None])
    y = Φ.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    rff = blr.RandomFourierFeatures(W, b)
    bfr = blr.BasisFunctionRegressor(blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D))), rff)
    post, lp = blr.posterior_and_logpdf(bfr(blr.ColVecs(x), σ2), y)
    fo = ref.BayesianLinearRegressor(np.zeros(D), ref.Diagonal(np.ones(D)))(ref.ColVecs(Φ), σ2)
    assert abs(lp - ref.logpdf(fo, y)) <= RTOL * abs(lp)
    assert relerr(post.blr.mw, ref.posterior(fo, y).mw) < RTOL
    m, v = blr.mean_and_var(post(blr.ColVecs(x[:, :100]), 0.5))
    mo, vo = ref.mean_and_var(ref.posterior(fo, y)(ref.ColVecs(Φ[:, :100]), 0.5))
    assert relerr(m, mo) < RTOL and relerr(v, vo) < RTOL


# ------------------------------------------------------------------------------------------------ README toy + golden fixture
def test_readme_toy():
    """BASELINE config 1 (README.md:44-86): D=2, N=10 ColVecs, heteroscedastic noise, 1000 plot points, noise eps()."""
    rng = np.random.default_rng(123456)
    N, Np = 10, 1000
    X = np.vstack([np.linspace(-5.0, 5.0, N), np.ones(N)])
    σ2 = np.exp(rng.standard_normal(N))
    f = blr.BayesianLinearRegressor(np.zeros(2), blr.Diagonal(np.ones(2)))
    fo = ref.BayesianLinearRegressor(np.zeros(2), ref.Diagonal(np.ones(2)))
    y = blr.rand(rng, f(blr.ColVecs(X), blr.Diagonal(σ2)))
    assert abs(blr.logpdf(f(blr.ColVecs(X), σ2), y) - ref.logpdf(fo(ref.ColVecs(X), σ2), y)) <= RTOL * 50
    fp, fpo = blr.posterior(f(blr.ColVecs(X), σ2), y), ref.posterior(fo(ref.ColVecs(X), σ2), y)
    lp_post = blr.logpdf(fp(blr.ColVecs(X), σ2), y)
    assert abs(lp_post - ref.logpdf(fpo(ref.ColVecs(X), σ2), y)) <= RTOL * abs(lp_post)
    Xp = np.vstack([np.linspace(-6.0, 6.0, Np), np.ones(Np)])
    m, s = blr.marginals(fp(blr.ColVecs(Xp), EPS))
    mo, so = ref.marginals(fpo(ref.ColVecs(Xp), EPS))
    assert relerr(m, mo) < RTOL and relerr(s, so) < RTOL
    assert blr.rand(rng, fp(blr.ColVecs(Xp), EPS), 10).shape == (Np, 10)


GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "blr_golden.npz"))


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in GOLD.files if "/" in k}))
def test_against_committed_golden_fixture(case):
    g = lambda k: GOLD[f"{case}/{k}"]  # noqa: E731
    D = g("mw").shape[0]
    Λw = g("Lambda") if bool(g("dense")) else blr.Diagonal(np.diag(g("Lambda")).copy())
    σ2 = g("sigma2") if g("sigma2").ndim else float(g("sigma2"))
    f = blr.BayesianLinearRegressor(g("mw"), Λw)
    fx = f(blr.ColVecs(g("X")), σ2)
    assert relerr(blr.rand_with_draws(fx, g("Zw"), g("Zy")), g("rand")) < RTOL
    post, lp = blr.posterior_and_logpdf(fx, g("y"))
    assert abs(lp - g("logpdf")) <= RTOL * abs(g("logpdf"))
    assert relerr(post.mw, g("m_post")) < RTOL
    assert relerr(post.Λw.dense(), g("Lambda_post")) < RTOL
    m, v = blr.mean_and_var(post(blr.ColVecs(g("Xt")), EPS))
    assert relerr(m, g("mean_t")) < RTOL and relerr(v, g("var_t")) < RTOL
    assert relerr(blr.cov(post(blr.ColVecs(g("Xt")), EPS)), g("cov_t")) < RTOL
    pm, pv = blr.mean_and_var(fx)
    assert relerr(pm, g("prior_mean")) < RTOL and relerr(pv, g("prior_var")) < RTOL
    assert D == post.mw.shape[0]


# ------------------------------------------------------------------------------------------------ error behaviour
def test_error_contracts():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((5, 11))
    f = blr.BayesianLinearRegressor(np.zeros(5), blr.Diagonal(np.ones(5)))
    with pytest.raises(blr.BLRError):  # src/bayesian_linear_regression.jl:74
        blr.logpdf(f(X, 0.1), np.zeros(10))
    with pytest.raises(blr.DimensionMismatch):
        blr.posterior(f(X, 0.1), np.zeros(12))
    with pytest.raises(RuntimeError):  # :26-31, test/bayesian_linear_regression.jl:116-122
        blr.rand(rng, f([row for row in X.T], 0.1))
    bad = blr.BayesianLinearRegressor(np.zeros(2), np.array([[1.0, 2.0], [2.0, 1.0]]))
    with pytest.raises(blr.PosDefException) as ei:
        blr.var(bad(np.ones((2, 3)), 0.1))
    assert ei.value.info == 2
    with pytest.raises(blr.PosDefException):
        blr.logpdf(bad(np.ones((2, 3)), 0.1), np.zeros(3))
    # inputs are not mutated (the reference copies X before trsm)
    Xc, yc = X.copy(), rng.standard_normal(11)
    yk = yc.copy()
    blr.posterior(f(X, 0.1), yc)
    assert np.array_equal(X, Xc) and np.array_equal(yc, yk)


def test_empty_and_single_observation():
    f = blr.BayesianLinearRegressor(np.array([0.5, -1.0]), np.array([[2.0, 0.3], [0.3, 1.0]]))
    fo = ref.BayesianLinearRegressor(np.array([0.5, -1.0]), np.array([[2.0, 0.3], [0.3, 1.0]]))
    X1, y1 = np.array([[1.0], [2.0]]), np.array([0.7])
    assert abs(blr.logpdf(f(X1, 0.3), y1) - ref.logpdf(fo(ref.ColVecs(X1), 0.3), y1)) < 1e-12
    post = blr.posterior(f(np.zeros((2, 0)), 0.3), np.zeros(0))  # no data: posterior == prior
    assert relerr(post.mw, f.mw) < 1e-15 and relerr(post.Λw.dense(), f.Λw) < 1e-15
    assert blr.logpdf(f(np.zeros((2, 0)), 0.3), np.zeros(0)) == 0.0
