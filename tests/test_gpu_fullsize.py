"""Full-size checks at BASELINE.json's shapes, where the CPU oracle cannot run (cfg3: X is 128 GiB).

The CUDA path is checked through size-independent properties of the domain, each evaluated by an independent
fp64 computation on the same device-resident data (torch / cuBLAS used as the *checker*, never as the product):

  * Freivalds projections of the Gram statistics:  G U = X (s ⊙ (X'U))  for random U (D x 4) -- every entry of G
    enters the identity, and a wrong tile / dropped stage / double-counted stream-K segment breaks it;
    r = X (s ⊙ δ), q = Σ s δ², ℓ = Σ log σ², n = N with δ = y - X'mw  (src/bayesian_linear_regression.jl:81-86);
  * additivity: statistics of the two halves of the data sum to the statistics of the whole (different stream-K
    cuts, the property sequential conditioning rests on, test/bayesian_linear_regression.jl:49-70);
  * the D x D phase against the closed form evaluated on the host from the same statistics (scipy LAPACK):
    Λ' = Λw + G, m' = mw + Λ'⁻¹ r, logpdf = -½ (N log 2π + ℓ + logdet Λ' - logdet Λw + q - r'Λ'⁻¹r), T'T = Λ';
  * marginals / rand at cfg4's shape: sampled blocks of test points against x'm', x'Λ'⁻¹x + σ² and
    X'(m' + Uw⁻¹Zw) + σ Zy computed by torch from the host posterior.

Sizes default to the BASELINE shapes (cfg3: N = 2^24, D = 1024; cfg4: D = 512, N* = 2^24 per GPU, S = 64) and are
halved until the inputs fit in 85 % of the device's free memory (a B200's 180 GB hold cfg3's 128 GiB); BLR_FULLSIZE_LOG2N overrides the cfg3 size.
Tolerances: 1e-9 relative (north_star); the observed errors are ~1e-13.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

import blr_b200 as blr
from blr_b200 import _lib as L

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(autouse=True)
def _release_device_memory():
    """These tests hold most of the device's memory in torch tensors: give it back to the driver afterwards so the
    library's own stream-ordered pool (and the tests that follow) are not starved by torch's caching allocator."""
    yield
    import gc

    import torch

    gc.collect()
    blr.default_context().sync()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _fit_log2n(log2n, bytes_per_obs, torch):
    free, _ = torch.cuda.mem_get_info()
    while log2n > 10 and (1 << log2n) * bytes_per_obs > 0.85 * free:
        log2n -= 1
    return log2n


def _synth(ctx, torch, D, N, seed):
    """Workload of bench.py: X ~ N(0,1), σ² = exp(N(0,1)), y = X'w* + σ ε, generated in place by
    the library into torch-owned buffers so that the checker can read the very same bytes."""
    Xt = torch.empty((N, D), dtype=torch.float64, device="cuda")  # (N, D) row-major == D x N column-major (ColVecs)
    st = torch.empty(N, dtype=torch.float64, device="cuda")
    yt = torch.empty(N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    X = blr.DeviceMatrix.wrap_torch(ctx, Xt, L.COLVECS).synth_(seed)
    s2 = blr.DeviceVector.wrap_torch(ctx, st)
    y = blr.DeviceVector.wrap_torch(ctx, yt)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2.handle, seed, 0))
    ctx.check(ctx.lib.blr_vec_synth_targets(ctx.handle, X.handle, s2.handle, seed, 0, y.handle))
    ctx.sync()
    return (Xt, st, yt), (X, s2, y)


def _accumulate(ctx, st, mw, X, y, s2):
    noise = L.Noise(L.NOISE_VECTOR, 0.0, s2.handle, None, 0)
    ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, st.handle, mw.ctypes.data_as(C.c_void_p), X.handle, y.handle, C.byref(noise)))


def test_cfg3_full_size_statistics_and_posterior():
    import scipy.linalg as sl
    import torch

    ctx = blr.default_context()
    D = 1024
    log2n = _fit_log2n(int(os.environ.get("BLR_FULLSIZE_LOG2N", "24")), 8 * (D + 8), torch)
    N = 1 << log2n
    print(f"[fullsize] cfg3 check at N = 2^{log2n}, D = {D} ({N * D * 8 / 2**30:.0f} GiB)")
    (Xt, st_t, yt), (X, s2, y) = _synth(ctx, torch, D, N, seed=0)
    rng = np.random.default_rng(11)
    mw = 0.05 * rng.standard_normal(D)  # non-zero prior mean: δ = y - X'mw is formed on the device (K0 reads X)

    # ---- product: sufficient statistics of the whole data set, and of its two halves
    whole = blr.Stats(ctx, D)
    _accumulate(ctx, whole, mw, X, y, s2)
    G, r, q, ell, n = whole.unpack()
    halves = blr.Stats(ctx, D)
    h = N // 2 + 4096 + 2  # uneven cut, not a multiple of a stage
    for a, b in ((0, h), (h, N)):
        Xa = blr.DeviceMatrix.wrap_torch(ctx, Xt[a:b], L.COLVECS)
        _accumulate(ctx, halves, mw, Xa, blr.DeviceVector.wrap_torch(ctx, yt[a:b]), blr.DeviceVector.wrap_torch(ctx, st_t[a:b]))
    Gh, rh, qh, ellh, nh = halves.unpack()
    assert n == N and nh == N
    assert np.array_equal(G, G.T)
    assert _rel(Gh, G) < 1e-12 and _rel(rh, r) < 1e-11
    assert abs(qh - q) <= 1e-11 * abs(q) and abs(ellh - ell) <= 1e-11 * max(abs(ell), 1.0)

    # ---- checker: Freivalds projections, r, q, ℓ by torch fp64 over L2-sized blocks (each block is read from HBM once)
    K = 4
    U = rng.standard_normal((D, K))
    Ue = torch.from_numpy(np.concatenate([U, mw[:, None]], axis=1)).cuda()  # D x (K + 1)
    GU = torch.zeros((D, K + 1), dtype=torch.float64, device="cuda")
    qs = torch.zeros((), dtype=torch.float64, device="cuda")
    ls = torch.zeros((), dtype=torch.float64, device="cuda")
    blk = 1 << 13
    for a in range(0, N, blk):
        Xc = Xt[a : a + blk]                    # blk x D
        A = Xc @ Ue                             # blk x (K + 1): X'U and X'mw
        sc = 1.0 / st_t[a : a + blk]
        d = yt[a : a + blk] - A[:, K]
        B = torch.cat([A[:, :K] * sc[:, None], (sc * d)[:, None]], dim=1)
        GU += Xc.T @ B
        qs += (sc * d * d).sum()
        ls += st_t[a : a + blk].log().sum()
    torch.cuda.synchronize()
    GUh = GU.cpu().numpy()
    assert _rel(G @ U, GUh[:, :K]) < RTOL, _rel(G @ U, GUh[:, :K])
    assert _rel(r, GUh[:, K]) < RTOL
    assert abs(q - float(qs)) <= RTOL * abs(q)
    assert abs(ell - float(ls)) <= RTOL * max(abs(ell), 1.0)

    # ---- D x D phase (K3/K4) against the closed form on the host, dense prior precision
    Bm = rng.standard_normal((D, D)) / math.sqrt(D)
    Lw = Bm @ Bm.T + np.eye(D)
    prior, keep = blr.BayesianLinearRegressor(mw, Lw)._prior_struct()
    lp = C.c_double()
    m_post, T_post, L_post = np.empty(D), np.empty((D, D), order="F"), np.empty((D, D), order="F")
    ctx.check(ctx.lib.blr_infer_from_stats(ctx.handle, C.byref(prior), whole.handle, C.byref(lp), m_post.ctypes.data_as(C.c_void_p),
                                           T_post.ctypes.data_as(C.c_void_p), L_post.ctypes.data_as(C.c_void_p), None))
    Lp = Lw + G
    cf = sl.cho_factor(Lp, lower=True)
    z = sl.cho_solve(cf, r)
    logdet_p = 2.0 * np.log(np.diag(cf[0])).sum()
    logdet_w = 2.0 * np.log(np.diag(np.linalg.cholesky(Lw))).sum()
    lp_o = -0.5 * (N * math.log(2 * math.pi) + ell + logdet_p - logdet_w + q - r @ z)
    assert _rel(L_post, Lp) < RTOL
    assert _rel(m_post, mw + z) < RTOL
    assert abs(lp.value - lp_o) <= RTOL * abs(lp_o), (lp.value, lp_o)
    Tu = np.triu(T_post)
    assert _rel(Tu.T @ Tu, Lp) < RTOL

    # ---- the one-call path the benchmark times gives the same answer as the staged calls above
    f = blr.BayesianLinearRegressor(mw, Lw)
    post, lp2 = blr.posterior_and_logpdf(f(blr.ColVecs(X), s2), y)
    assert lp2 == lp.value and np.array_equal(post.mw, m_post)


def test_cfg4_full_size_marginals_and_rand():
    import scipy.linalg as sl
    import torch

    ctx = blr.default_context()
    D, S = 512, 64
    # posterior from a modest fit (what matters here is the prediction side)
    (_, _, _), (Xf, s2f, yf) = _synth(ctx, torch, D, 1 << 18, seed=3)
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    post, _ = blr.posterior_and_logpdf(f(blr.ColVecs(Xf), s2f), yf)
    del Xf, s2f, yf
    Lp = post.Λw.dense()
    cf = np.linalg.cholesky(Lp)  # Λ' = cf cf'
    dpost = post._device(ctx)

    log2n = _fit_log2n(24, 8 * (D + 4), torch)
    Nt = 1 << log2n
    print(f"[fullsize] cfg4 check at N* = 2^{log2n}, D = {D}, S = {S}")
    Xt = torch.empty((Nt, D), dtype=torch.float64, device="cuda")
    sig = torch.empty(Nt, dtype=torch.float64, device="cuda")
    mv = torch.empty((2, Nt), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    X = blr.DeviceMatrix.wrap_torch(ctx, Xt, L.COLVECS).synth_(5)
    sv = blr.DeviceVector.wrap_torch(ctx, sig)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, sv.handle, 5, 0))
    noise = L.Noise(L.NOISE_VECTOR, 0.0, sv.handle, None, 0)
    ctx.check(ctx.lib.blr_mean_var_dev(ctx.handle, dpost.handle, X.handle, C.byref(noise), C.c_void_p(mv[0].data_ptr()),
                                       C.c_void_p(mv[1].data_ptr())))
    ctx.sync()

    cft = torch.from_numpy(cf).cuda()
    mt = torch.from_numpy(post.mw).cuda()
    rng = np.random.default_rng(2)
    blk = 8192
    starts = [0, Nt - blk] + [int(a) for a in rng.integers(0, Nt - blk, 14)]
    for a in starts:
        Xc = Xt[a : a + blk]                                                     # blk x D
        al = torch.linalg.solve_triangular(cft, Xc.T, upper=False)              # α = L⁻¹ X   (:41)
        v_o = (al * al).sum(0) + sig[a : a + blk]
        m_o = Xc @ mt
        assert float((mv[1, a : a + blk] - v_o).norm() / v_o.norm()) < RTOL
        assert float((mv[0, a : a + blk] - m_o).norm() / m_o.norm()) < RTOL
    # var >= σ² everywhere, finite everywhere (cheap whole-array properties)
    assert bool(torch.isfinite(mv).all()) and bool((mv[1] > sig).all())
    del mv

    # ---- rand, S = 64 function samples on the first 2^22 points with supplied draws (Zw host, Zy device)
    Nr = min(Nt, 1 << 22)
    g = torch.Generator(device="cuda").manual_seed(9)
    Zy = torch.randn((S, Nr), dtype=torch.float64, device="cuda", generator=g)   # == Nr x S column-major
    Zw = np.asfortranarray(rng.standard_normal((D, S)))
    Y = torch.empty((S, Nr), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    Xr = blr.DeviceMatrix.wrap_torch(ctx, Xt[:Nr], L.COLVECS)
    svr = blr.DeviceVector.wrap_torch(ctx, sig[:Nr])
    noise_r = L.Noise(L.NOISE_VECTOR, 0.0, svr.handle, None, 0)
    ctx.check(ctx.lib.blr_rand_finite_dev(ctx.handle, dpost.handle, Xr.handle, C.byref(noise_r), S, Zw.ctypes.data_as(C.c_void_p),
                                          C.c_void_p(Zy.data_ptr()), 0, C.c_void_p(Y.data_ptr())))
    ctx.sync()
    Wsamp = post.mw[:, None] + sl.solve_triangular(cf.T, Zw, lower=False)       # mw + Uw⁻¹ Zw   (:52)
    Wt = torch.from_numpy(Wsamp).cuda()
    for a in [0, Nr - blk] + [int(a) for a in rng.integers(0, Nr - blk, 6)]:
        Yo = (Xt[a : a + blk] @ Wt).T + sig[a : a + blk].sqrt()[None, :] * Zy[:, a : a + blk]
        assert float((Y[:, a : a + blk] - Yo).norm() / Yo.norm()) < RTOL
    assert bool(torch.isfinite(Y).all())
