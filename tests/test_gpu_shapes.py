"""GPU parity at the BASELINE shapes the round-1 suite never reached (VERDICT r01 "missing" #1):

  * D = 2048 and D = 4096 (cfg5's regime: 528 lower tiles on 148 CTAs -> multi-segment CTAs with read-modify-write
    flushes, 64-panel Cholesky, 64-block triangular solves) against the streaming oracle, ColVecs and RowVecs, diagonal
    and dense prior, Symmetric and PDMat result kinds (src/bayesian_linear_regression.jl:55-93);
  * cfg2 at its true size N = 2^20, D = 256 against the streaming oracle;
  * cfg5's per-GPU share (N = 2^19, D = 4096, random Fourier features evaluated on the device,
    src/basis_function_regression.jl:41) through Freivalds projections computed by torch from the raw inputs -- the
    checker evaluates sqrt(2/D) cos(Wx + b) itself, so a wrong feature map fails exactly like a wrong Gram tile;
  * the plain-C client (examples/minimal_client.c) built and RUN on the device, its printed numbers against the oracle.

Tolerance 1e-9 relative (north_star); observed errors are printed.
"""
import ctypes as C
import math
import os
import re
import subprocess

import numpy as np
import pytest

import blr_b200 as blr
from blr_b200 import _lib as L
from oracle import blr_oracle as ref

pytestmark = pytest.mark.gpu
RTOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(autouse=True)
def _release_device_memory():
    yield
    import gc

    import torch

    gc.collect()
    blr.default_context().sync()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


_DATA = {}


def big_problem(D, N):
    """X ~ N(0,1), heteroscedastic noise, non-zero prior mean; cached per shape (1 GiB at D = 4096, N = 2^15)."""
    key = (D, N)
    if key not in _DATA:
        rng = np.random.default_rng(1000 + D)
        X = np.empty((D, N), order="F")
        for a in range(0, N, 4096):
            X[:, a : a + 4096] = rng.standard_normal((D, min(4096, N - a)))
        σ2 = np.exp(rng.standard_normal(N))
        mw = 0.1 * rng.standard_normal(D)
        y = X.T @ (rng.standard_normal(D) / math.sqrt(D)) + np.sqrt(σ2) * rng.standard_normal(N)
        B = rng.standard_normal((D, D)) / math.sqrt(D)
        Λ = B @ B.T + np.eye(D)
        G, r, q, ℓ = ref.gram_stats(X, y, σ2, mw, chunk=4096)
        _DATA[key] = (X, σ2, mw, y, Λ, (G, r, q, ℓ))
    return _DATA[key]


@pytest.mark.parametrize("D,N", [(2048, 1 << 14), (4096, 1 << 15)])
@pytest.mark.parametrize("prior", ["diagonal", "dense", "pdmat"])
@pytest.mark.parametrize("Tx", ["ColVecs", "RowVecs"])
def test_large_D_posterior_and_logpdf(D, N, prior, Tx):
    if Tx == "RowVecs" and prior == "pdmat":
        pytest.skip("result kind does not depend on the input layout")
    X, σ2, mw, y, Λ, (G, r, q, ℓ) = big_problem(D, N)
    if prior == "diagonal":
        lam = np.linspace(0.5, 2.0, D)
        f, Λo = blr.BayesianLinearRegressor(mw, blr.Diagonal(lam)), ref.Diagonal(lam)
    elif prior == "dense":
        f, Λo = blr.BayesianLinearRegressor(mw, Λ), Λ
    else:
        f, Λo = blr.BayesianLinearRegressor(mw, blr.PDMat(Λ)), ref.PDMat.from_matrix(Λ)
    lp_o, m_o, T_o = ref.infer_from_stats(mw, Λo, G, r, q, ℓ, N)
    ctx = blr.default_context()
    Xd = blr.DeviceMatrix.upload(ctx, X if Tx == "ColVecs" else np.ascontiguousarray(X.T), L.COLVECS if Tx == "ColVecs" else L.ROWVECS)
    x = blr.ColVecs(Xd) if Tx == "ColVecs" else blr.RowVecs(Xd)
    post, lp = blr.posterior_and_logpdf(f(x, blr.DeviceVector.upload(ctx, σ2)), blr.DeviceVector.upload(ctx, y))
    e = {"logpdf": abs(lp - lp_o) / abs(lp_o), "mean": relerr(post.mw, m_o), "precision": relerr(post.Λw.dense(), T_o.T @ T_o)}
    if prior == "pdmat":
        assert isinstance(post.Λw, blr.PDMat)
        e["T"] = relerr(np.triu(post.Λw.U), T_o)
    else:
        assert isinstance(post.Λw, blr.Symmetric)
    print(f"[shapes] D={D} N={N} {prior} {Tx}: " + ", ".join(f"{k} {v:.1e}" for k, v in e.items()))
    assert all(v < RTOL for v in e.values()), e
    # prediction side at this D from the cached device posterior: marginals and rand with supplied draws on 300 points
    if Tx == "ColVecs":
        rng = np.random.default_rng(5)
        Xt = rng.standard_normal((D, 300))
        m, v = blr.mean_and_var(post(blr.ColVecs(Xt), 0.3))
        Lp = np.linalg.cholesky(T_o.T @ T_o)
        import scipy.linalg as sl

        α = sl.solve_triangular(Lp, Xt, lower=True)
        assert relerr(m, Xt.T @ m_o) < RTOL and relerr(v, (α * α).sum(0) + 0.3) < RTOL
        Zw, Zy = rng.standard_normal((D, 3)), rng.standard_normal((300, 3))
        Y = blr.rand_with_draws(post(blr.ColVecs(Xt), 0.3), Zw, Zy)
        Yo = Xt.T @ (m_o[:, None] + sl.solve_triangular(Lp.T, Zw, lower=False)) + math.sqrt(0.3) * Zy
        assert relerr(Y, Yo) < RTOL


def test_cfg2_true_size():
    """BASELINE config 2 at its real size: N = 2^20, D = 256 (X is 2 GiB on the host), diagonal prior, zero prior mean and a
    non-zero one, against the streaming Gram-form oracle (validated against the literal form in tests/test_oracle_streaming.py)."""
    D, N = 256, 1 << 20
    rng = np.random.default_rng(42)
    X = np.empty((D, N), order="F")
    for a in range(0, N, 1 << 16):
        X[:, a : a + (1 << 16)] = rng.standard_normal((D, 1 << 16))
    σ2 = np.exp(rng.standard_normal(N))
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    ctx = blr.default_context()
    Xd, yd, sd = blr.DeviceMatrix.upload(ctx, X, L.COLVECS), blr.DeviceVector.upload(ctx, y), blr.DeviceVector.upload(ctx, σ2)
    for mw in (np.zeros(D), 0.3 * rng.standard_normal(D)):
        f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(D)))
        post, lp = blr.posterior_and_logpdf(f(blr.ColVecs(Xd), sd), yd)
        lp_o, m_o, T_o = ref.infer_streaming(mw, ref.Diagonal(np.ones(D)), X, y, σ2, chunk=1 << 16)
        e = (abs(lp - lp_o) / abs(lp_o), relerr(post.mw, m_o), relerr(post.Λw.dense(), T_o.T @ T_o))
        print(f"[shapes] cfg2 N=2^20 D=256 mw{'=0' if not mw.any() else '!=0'}: logpdf {e[0]:.1e} mean {e[1]:.1e} precision {e[2]:.1e}")
        assert max(e) < RTOL, e
        # host-streamed entry point (chunked H2D) over the same data: same statistics, different stream-K cuts
        post_s, lp_s = blr.posterior_and_logpdf_streamed(f, X, y, σ2, chunk=1 << 17, ctx=ctx)
        assert abs(lp_s - lp_o) <= RTOL * abs(lp_o) and relerr(post_s.mw, m_o) < RTOL


def test_cfg5_per_gpu_share_rff_freivalds():
    """cfg5's share of one GPU on an 8-GPU box: N = 2^19 inputs of d_in = 32, ϕ = RFF with D = 4096 (ϕ(x) is 16 GiB and never
    leaves the device).  The statistics of BasisFunctionRegressor's posterior+logpdf are checked by Freivalds projections
    that torch evaluates from x, W, b directly."""
    import scipy.linalg as sl
    import torch

    ctx = blr.default_context()
    D, din = 4096, 32
    free, _ = torch.cuda.mem_get_info()
    log2n = 19
    while log2n > 12 and (1 << log2n) * (D + 64) * 8 * 1.3 > 0.8 * free:
        log2n -= 1
    N = 1 << log2n
    print(f"[shapes] cfg5 share at N = 2^{log2n}, D = {D}")
    xt = torch.empty((N, din), dtype=torch.float64, device="cuda")   # = d_in x N column-major
    st = torch.empty(N, dtype=torch.float64, device="cuda")
    yt = torch.empty(N, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    xin = blr.DeviceMatrix.wrap_torch(ctx, xt, L.COLVECS).synth_(3)
    s2, y = blr.DeviceVector.wrap_torch(ctx, st), blr.DeviceVector.wrap_torch(ctx, yt)
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, s2.handle, 5, 0))
    ctx.check(ctx.lib.blr_vec_synth_noise(ctx.handle, y.handle, 6, 0))
    ctx.sync()
    yt.log_()  # targets ~ N(0,1) (synth_noise draws exp(z))
    torch.cuda.synchronize()
    rng = np.random.default_rng(0)
    W, b = rng.standard_normal((D, din)), rng.uniform(0, 2 * np.pi, D)
    rff = blr.RandomFourierFeatures(W, b, ctx)
    lam = np.linspace(0.5, 1.5, D)
    bfr = blr.BasisFunctionRegressor(blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(lam)), rff)

    # (a) the feature map itself on sampled columns against numpy
    Φd = rff(blr.ColVecs(xin)).X
    ptr, ld = Φd.device_ptr()
    cols = [0, 1, N // 3, N - 1]
    xs = xt[cols].cpu().numpy().T                                       # d_in x 4
    Φo = math.sqrt(2.0 / D) * np.cos(W @ xs + b[:, None])
    cudart = C.CDLL("libcudart.so.12")
    for j, c in enumerate(cols):
        col = np.empty(D)
        rc = cudart.cudaMemcpy(col.ctypes.data_as(C.c_void_p), C.c_void_p(ptr + 8 * ld * c), C.c_size_t(8 * D), C.c_int(2))
        assert rc == 0
        assert relerr(col, Φo[:, j]) < 1e-12
    # (b) statistics through the public BasisFunctionRegressor path vs Freivalds projections of torch's own ϕ
    stats = blr.Stats(ctx, D)
    noise = L.Noise(L.NOISE_VECTOR, 0.0, s2.handle, None, 0)
    mw0 = np.zeros(D)
    ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, stats.handle, mw0.ctypes.data_as(C.c_void_p), Φd.handle, y.handle, C.byref(noise)))
    G, r, q, ell, n = stats.unpack()
    del Φd
    K = 3
    U = rng.standard_normal((D, K))
    Wt, bt, Ut = torch.from_numpy(W).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(U).cuda()
    GU = torch.zeros((D, K + 1), dtype=torch.float64, device="cuda")
    qs = torch.zeros((), dtype=torch.float64, device="cuda")
    blk = 1 << 13
    for a in range(0, N, blk):
        Φ = math.sqrt(2.0 / D) * torch.cos(xt[a : a + blk] @ Wt.T + bt)   # blk x D
        sc = 1.0 / st[a : a + blk]
        Bm = torch.cat([(Φ @ Ut) * sc[:, None], (sc * yt[a : a + blk])[:, None]], dim=1)
        GU += Φ.T @ Bm
        qs += (sc * yt[a : a + blk] ** 2).sum()
    torch.cuda.synchronize()
    GUh = GU.cpu().numpy()
    e = {"GU": relerr(G @ U, GUh[:, :K]), "r": relerr(r, GUh[:, K]), "q": abs(q - float(qs)) / abs(q),
         "ell": abs(ell - float(st.log().sum())) / max(abs(ell), 1.0)}
    print("[shapes] cfg5 Freivalds:", {k: f"{v:.1e}" for k, v in e.items()})
    assert n == N and np.array_equal(G, G.T) and all(v < RTOL for v in e.values()), e
    # (c) the D x D phase at D = 4096 (64 panels) from those statistics against the host closed form, and the one-call path
    post, lp = blr.posterior_and_logpdf(bfr(blr.ColVecs(xin), s2), y)
    Lp = np.diag(lam) + G
    cf = sl.cho_factor(Lp, lower=True)
    z = sl.cho_solve(cf, r)
    lp_o = -0.5 * (N * math.log(2 * math.pi) + ell + 2.0 * np.log(np.diag(cf[0])).sum() - np.log(lam).sum() + q - r @ z)
    e2 = (abs(lp - lp_o) / abs(lp_o), relerr(post.blr.mw, z), relerr(post.blr.Λw.dense(), Lp))
    print(f"[shapes] cfg5 D x D phase: logpdf {e2[0]:.1e} mean {e2[1]:.1e} precision {e2[2]:.1e}")
    assert max(e2) < RTOL, e2


def test_c_client_runs_on_device(tmp_path):
    """examples/minimal_client.c compiled as C99 against include/blr_cuda.h and RUN on the B200; its printed logpdf,
    posterior mean, predictive means and variances against the oracle on the same toy problem (README.md:44-60)."""
    libdir = os.path.dirname(L.LIB_PATH)
    exe = tmp_path / "minimal_client"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "minimal_client.c"), "-L", libdir, "-lblr_cuda", f"-Wl,-rpath,{libdir}",
                    "-o", str(exe)], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout
    lp = float(re.search(r"logpdf = (\S+)", out).group(1))
    mpost = [float(v) for v in re.search(r"posterior mean = \[(\S+), (\S+)\]", out).groups()]
    pred = [(float(a), float(b)) for a, b in re.findall(r"mean (\S+), var (\S+)", out)]
    N = 10
    xs = np.array([-5.0 + 10.0 * n / (N - 1) for n in range(N)])
    X = np.vstack([xs, np.ones(N)])
    y = 0.5 * xs - 1.0
    fo = ref.BayesianLinearRegressor(np.zeros(2), ref.Diagonal(np.ones(2)))
    fxo = fo(ref.ColVecs(X), 0.1)
    po = ref.posterior(fxo, y)
    Xt = np.array([[-6.0, 0.0, 6.0], [1.0, 1.0, 1.0]])
    mo, vo = ref.mean_and_var(po(ref.ColVecs(Xt), 0.1))
    assert abs(lp - ref.logpdf(fxo, y)) <= RTOL * abs(lp)
    assert relerr(np.array(mpost), po.mw) < RTOL
    assert relerr(np.array([p[0] for p in pred]), mo) < RTOL and relerr(np.array([p[1] for p in pred]), vo) < RTOL


def test_maximum_dimension_and_beyond():
    """The library's largest D (16 384: blr_ctx::small holds D-vectors of that length) and one past it.  At D = 16 384 the Gram
    matrix is 2 GiB, the lower triangle 8 256 tiles, the fused D x D kernel 32 896 tile tasks.  N < D and a Diagonal prior keep the
    host check cheap: by the Woodbury identity the posterior mean and the log marginal likelihood follow from the N x N system
    S = Σy + X' Λw⁻¹ X (the naive Gaussian of test/bayesian_linear_regression.jl:22-38): m' = mw + Λw⁻¹ X S⁻¹ δ,
    logpdf = -½ (N log 2π + logdet S + δ' S⁻¹ δ); the precision is checked entrywise on sampled rows of Λw + X Σy⁻¹ X'."""
    import scipy.linalg as sl

    ctx = blr.default_context()
    D, N = 16384, 700
    rng = np.random.default_rng(16384)
    X = np.asfortranarray(rng.standard_normal((D, N)))
    σ2 = np.exp(rng.standard_normal(N))
    mw = 0.05 * rng.standard_normal(D)
    lam = np.linspace(0.5, 2.0, D)
    y = X.T @ (rng.standard_normal(D) / math.sqrt(D)) + np.sqrt(σ2) * rng.standard_normal(N)
    post, lp = blr.posterior_and_logpdf(blr.BayesianLinearRegressor(mw, blr.Diagonal(lam))(blr.ColVecs(X), σ2), y)
    δ = y - X.T @ mw
    S = (X.T / lam) @ X + np.diag(σ2)
    cS = sl.cho_factor(S, lower=True)
    a = sl.cho_solve(cS, δ)
    lp_o = -0.5 * (N * math.log(2 * math.pi) + 2.0 * np.log(np.diag(cS[0])).sum() + δ @ a)
    m_o = mw + (X @ a) / lam
    assert abs(lp - lp_o) / abs(lp_o) < RTOL, (lp, lp_o)
    assert relerr(post.mw, m_o) < RTOL
    rows = rng.choice(D, 64, replace=False)
    P = post.Λw.dense()
    Po = (X[rows] / σ2) @ X.T
    Po[np.arange(64), rows] += lam[rows]
    assert relerr(P[rows], Po) < RTOL
    assert np.array_equal(P, P.T)
    # marginals from the cached device factor on a few points: var = x' Λ'⁻¹ x + σ², Λ'⁻¹ x = Λw⁻¹ x - Λw⁻¹ X S⁻¹ X' Λw⁻¹ x
    Xt = rng.standard_normal((D, 40))
    m, v = blr.mean_and_var(post(blr.ColVecs(Xt), 0.3))
    U = Xt / lam[:, None]
    V = U - (X @ sl.cho_solve(cS, X.T @ U)) / lam[:, None]
    assert relerr(m, Xt.T @ m_o) < RTOL and relerr(v, (Xt * V).sum(0) + 0.3) < RTOL
    del post, P
    # one past the maximum: a clean error, not a crash (the reference has no limit; the C ABI reports BLR_E_INVALID)
    D2 = 16385
    with pytest.raises(blr.BLRError):
        blr.posterior(blr.BayesianLinearRegressor(np.zeros(D2), blr.Diagonal(np.ones(D2)))(blr.ColVecs(np.zeros((D2, 4))), 1.0), np.zeros(4))
