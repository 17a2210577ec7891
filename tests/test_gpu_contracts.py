"""GPU: error-behaviour and stream-ordering contracts of the C ABI that round 1 left open:

  * non-positive observation-noise variances -> PosDefException(info) wherever the reference factorises Σy
    (posterior / logpdf, src/bayesian_linear_regression.jl:79; rand :52) and NOT in mean / var / cov (:37,:42);
  * size(X, 1) != length(mw) on the device-resident path -> DimensionMismatch, checked by the library itself (blr_prior.D);
  * repeated blr_stats_accumulate_host calls from pinned memory on one context reuse the two staging slots safely;
  * borrowed torch tensors are ordered after their producer stream (blr_ctx_wait_stream);
  * the fused D x D kernel against the round-1 multi-launch path (BLR_DXD=legacy), bit-reproducible run to run.
"""
import ctypes as C

import numpy as np
import pytest

import blr_b200 as blr
from blr_b200 import _lib as L
from oracle import blr_oracle as ref

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def small_problem(D=96, N=700, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N))
    mw = rng.standard_normal(D)
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    return X, σ2, mw, y


@pytest.mark.parametrize("D", [2, 12, 40, 96, 300])  # tiny / small fused / small / TMA Gram kernels
@pytest.mark.parametrize("bad", [-0.3, 0.0, float("nan")])
def test_nonpositive_noise_vector_raises_posdef(D, bad):
    X, σ2, mw, y = small_problem(D, 500, seed=D)
    f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(D)))
    σ2b = σ2.copy()
    σ2b[37] = bad
    σ2b[411] = bad
    ctx = blr.default_context()
    # host-streamed path and device-resident path
    for fx, yy in ((f(blr.ColVecs(X), σ2b), y),
                   (f(blr.ColVecs(blr.DeviceMatrix.upload(ctx, X, L.COLVECS)), blr.DeviceVector.upload(ctx, σ2b)),
                    blr.DeviceVector.upload(ctx, y))):
        with pytest.raises(blr.PosDefException) as ei:
            blr.posterior(fx, yy)
        assert ei.value.info == 38  # cholesky(Diagonal(σ²)) fails at the first offending entry (1-based)
        with pytest.raises(blr.PosDefException):
            blr.logpdf(fx, yy)
    # rand factorises Σy as well (:52); mean / var only add diag(Σy) and do not throw, as in the reference
    with pytest.raises(blr.PosDefException) as ei:
        blr.rand(np.random.default_rng(0), f(blr.ColVecs(X), σ2b), 2)
    assert ei.value.info == 38
    m, v = blr.mean_and_var(f(blr.ColVecs(X), σ2b))
    assert m.shape == (500,) and np.isfinite(v[0])
    # the context is still usable afterwards
    lp = blr.logpdf(f(blr.ColVecs(X), σ2), y)
    fo = ref.BayesianLinearRegressor(mw, ref.Diagonal(np.ones(D)))
    assert abs(lp - ref.logpdf(fo(ref.ColVecs(X), σ2), y)) <= RTOL * abs(lp)


@pytest.mark.parametrize("bad", [-1.0, 0.0])
def test_nonpositive_scalar_noise_raises_posdef(bad):
    X, σ2, mw, y = small_problem()
    f = blr.BayesianLinearRegressor(mw, blr.Diagonal(np.ones(96)))
    ctx = blr.default_context()
    for fx, yy in ((f(blr.ColVecs(X), bad), y), (f(blr.ColVecs(blr.DeviceMatrix.upload(ctx, X, L.COLVECS)), bad), blr.DeviceVector.upload(ctx, y))):
        with pytest.raises(blr.PosDefException) as ei:
            blr.posterior(fx, yy)
        assert ei.value.info == 1
    with pytest.raises(blr.PosDefException):
        blr.rand(np.random.default_rng(0), f(blr.ColVecs(X), bad), 2)
    blr.mean_and_var(f(blr.ColVecs(X), bad))  # no factorisation of Σy: no exception


def test_dimension_mismatch_on_device_resident_path():
    """A DeviceMatrix whose D differs from length(mw): the reference throws DimensionMismatch from X'mw (:33)."""
    X, σ2, mw, y = small_problem(96, 300)
    ctx = blr.default_context()
    Xd = blr.DeviceMatrix.upload(ctx, X, L.COLVECS)
    yd, sd = blr.DeviceVector.upload(ctx, y), blr.DeviceVector.upload(ctx, σ2)
    for Dw in (64, 128):
        f = blr.BayesianLinearRegressor(np.zeros(Dw), blr.Diagonal(np.ones(Dw)))
        with pytest.raises(blr.DimensionMismatch):
            blr.posterior(f(blr.ColVecs(Xd), sd), yd)
        # and at the C ABI itself (what a Julia / C caller hits): blr_infer validates prior.D
        prior, keep = f._prior_struct()
        noise = L.Noise(L.NOISE_VECTOR, 0.0, sd.handle, None, 0)
        lp = C.c_double()
        rc = ctx.lib.blr_infer(ctx.handle, C.byref(prior), Xd.handle, yd.handle, C.byref(noise), C.byref(lp), None, None, None, None)
        assert rc == L.E_DIM


def test_accumulate_host_twice_from_pinned_memory():
    """Two back-to-back blr_stats_accumulate_host calls on one context (the additive API invites it) from page-locked
    memory, where the H2D copies are truly asynchronous: the second call must not overwrite a staging slot the first
    call's Gram kernel is still reading.  Compared with a single call and with the oracle; repeated to give a race a chance."""
    import torch

    D, N = 256, 40000
    rng = np.random.default_rng(3)
    Xh = torch.empty((N, D), dtype=torch.float64, pin_memory=True)
    yh = torch.empty(N, dtype=torch.float64, pin_memory=True)
    sh = torch.empty(N, dtype=torch.float64, pin_memory=True)
    Xh.numpy()[:] = rng.standard_normal((N, D))
    sh.numpy()[:] = np.exp(rng.standard_normal(N))
    yh.numpy()[:] = rng.standard_normal(N)
    X, y, σ2 = Xh.numpy().T, yh.numpy(), sh.numpy()  # X: D x N Fortran-ordered view of the pinned buffer
    mw = 0.1 * rng.standard_normal(D)
    ctx = blr.Context(0)
    h = N // 2 + 16
    Go, ro, qo, ℓo = ref.gram_stats(X, y, σ2, mw)

    def acc(st, a, b):
        Xs = X[:, a:b]
        ctx.check(ctx.lib.blr_stats_accumulate_host(ctx.handle, st.handle, mw.ctypes.data_as(C.c_void_p),
                                                    C.c_void_p(Xs.ctypes.data), D, b - a, D, L.COLVECS,
                                                    C.c_void_p(y[a:b].ctypes.data), L.NOISE_VECTOR, 0.0,
                                                    C.c_void_p(σ2[a:b].ctypes.data), 4096))

    for rep in range(5):
        two, one = blr.Stats(ctx, D), blr.Stats(ctx, D)
        acc(two, 0, h)
        acc(two, h, N)  # no sync in between
        acc(one, 0, N)
        G2, r2, q2, l2, n2 = two.unpack()
        G1, r1, q1, l1, n1 = one.unpack()
        assert n2 == N and n1 == N
        assert relerr(G2, Go) < 1e-12 and relerr(r2, ro) < 1e-11, rep
        assert relerr(G1, Go) < 1e-12 and abs(q2 - qo) <= 1e-11 * abs(qo) and abs(l2 - ℓo) <= 1e-11 * abs(ℓo)


def test_borrowed_torch_tensors_are_ordered_after_their_producer():
    """wrap_torch makes the context's private stream wait for torch's current stream: inputs built by (slow) torch
    kernels immediately before the call are complete when the library reads them -- no torch.cuda.synchronize()."""
    import torch

    D, N = 128, 1 << 17
    g = torch.Generator(device="cuda").manual_seed(5)
    base = torch.randn((N, D), dtype=torch.float64, device="cuda", generator=g)
    s2 = torch.rand(N, dtype=torch.float64, device="cuda", generator=g) + 0.5
    yv = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    torch.cuda.synchronize()
    f = blr.BayesianLinearRegressor(np.zeros(D), blr.Diagonal(np.ones(D)))
    Xref = (base * 3.0 + 1.0)
    torch.cuda.synchronize()
    lp_ref = blr.logpdf(f(blr.ColVecs(Xref), s2), yv)
    for _ in range(3):
        Xt = base.clone()
        for _ in range(20):  # keep torch's stream busy so that the last writes land late
            Xt = Xt * 1.0
        Xt = Xt * 3.0 + 1.0
        lp = blr.logpdf(f(blr.ColVecs(Xt), s2), yv)  # no synchronize
        assert lp == lp_ref


@pytest.mark.parametrize("D,N,dense", [(2, 10, False), (7, 13, True), (64, 300, True), (65, 300, True), (130, 515, True),
                                       (256, 2000, False), (1000, 1500, True), (1024, 1200, False), (2048, 2100, True)])
def test_fused_dxd_phase_matches_legacy_and_oracle(D, N, dense, monkeypatch):
    """The fused tiled D x D kernel (one cooperative launch: Cholesky + border solve + back-solve + finalize) against the
    round-1 multi-launch path on the same statistics, and against the oracle; and bit-reproducible run to run (its task
    graph has a fixed summation order whatever the timing)."""
    rng = np.random.default_rng(D + N)
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N))
    mw = rng.standard_normal(D)
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    if dense:
        B = rng.standard_normal((D, D)) / np.sqrt(D)
        Λ = B @ B.T + np.eye(D)
        f, Λo = blr.BayesianLinearRegressor(mw, blr.PDMat(Λ)), ref.PDMat.from_matrix(Λ)
    else:
        lam = np.linspace(0.5, 2.0, D)
        f, Λo = blr.BayesianLinearRegressor(mw, blr.Diagonal(lam)), ref.Diagonal(lam)
    lp_o, m_o, T_o = ref.infer_streaming(mw, Λo, X, y, σ2)
    outs = {}
    for mode in ("fused", "fused", "legacy"):
        if mode == "legacy":
            monkeypatch.setenv("BLR_DXD", "legacy")
        ctx = blr.Context(0)
        monkeypatch.delenv("BLR_DXD", raising=False)
        fx = f(blr.ColVecs(X), σ2)
        fx.ctx = ctx
        post, lp = blr.posterior_and_logpdf(fx, y)
        assert abs(lp - lp_o) <= RTOL * abs(lp_o), (mode, lp, lp_o)
        assert relerr(post.mw, m_o) < RTOL and relerr(post.Λw.dense(), T_o.T @ T_o) < RTOL
        if dense:
            T = post.Λw.U
            assert np.array_equal(T, np.triu(T)) and relerr(T, T_o) < RTOL
        outs.setdefault(mode, []).append((lp, post.mw.copy()))
    (lp_a, m_a), (lp_b, m_b) = outs["fused"]
    assert lp_a == lp_b and np.array_equal(m_a, m_b)
    lp_l, m_l = outs["legacy"][0]
    assert abs(lp_a - lp_l) <= 1e-12 * abs(lp_l) and relerr(m_a, m_l) < 1e-11


def test_fused_dxd_not_positive_definite_info():
    """LAPACK's info (order of the first non-positive leading minor) survives the out-of-order tile schedule."""
    D = 300
    rng = np.random.default_rng(1)
    B = rng.standard_normal((D, D))
    Λ = B @ B.T + np.eye(D)
    Λ[200, 200] = -5.0  # leading minor of order 201 is the first that fails
    bad = blr.BayesianLinearRegressor(np.zeros(D), Λ)
    with pytest.raises(blr.PosDefException) as ei:
        blr.var(bad(np.ones((D, 3)), 0.1))
    assert ei.value.info == 201
    with pytest.raises(blr.PosDefException) as ei:
        blr.logpdf(bad(np.ones((D, 3)), 0.1), np.zeros(3))
    assert ei.value.info == 201
