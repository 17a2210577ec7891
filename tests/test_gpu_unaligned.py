"""GPU: inputs that break the 16-byte rules of the TMA kernels take the tensor-core path through an aligned staging buffer.

A dense ColVecs matrix with an ODD number of features (a bias feature next to 2^k learned ones: D = 129, 257, 1025), an odd leading
dimension or a base address that is not 16-byte aligned cannot be addressed by bulk / tiled TMA copies.  Such inputs used to fall
to the DFMA-fed generic kernels (a 7x cliff between D = 127 and D = 128).  Now blocks of observations are repacked into an aligned
buffer (Gram: odd D becomes D + 1 with a zero feature whose row / column the reduction drops; marginals: the tensor map zero-fills
beyond D) -- gram.cu `repack_colvecs`, predict.cu staging branch.  Checked here: statistics against an independent torch fp64
evaluation over the same device bytes, across several staging blocks and for every way of being unaligned; marginals for RowVecs
and unaligned ColVecs device inputs against the aligned evaluation of the same points; posterior / logpdf against the oracle."""
import ctypes as C

import numpy as np
import pytest

import blr_b200 as blr
from blr_b200 import _lib as L
from oracle import blr_oracle as ref

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _torch_stats(torch, Xv, y, s2, mw):
    """G, r, q, ℓ by torch fp64 (the checker) from views of the very same device buffers; Xv is (N, D)."""
    D = Xv.shape[1]
    G = torch.zeros((D, D), dtype=torch.float64, device="cuda")
    r = torch.zeros(D, dtype=torch.float64, device="cuda")
    q = torch.zeros((), dtype=torch.float64, device="cuda")
    mwt = torch.from_numpy(mw).cuda()
    blk = 1 << 16
    for a in range(0, Xv.shape[0], blk):
        Xc = Xv[a : a + blk]
        sc = 1.0 / s2[a : a + blk]
        d = y[a : a + blk] - Xc @ mwt
        G += Xc.T @ (Xc * sc[:, None])
        r += Xc.T @ (sc * d)
        q += (sc * d * d).sum()
    return G.cpu().numpy(), r.cpu().numpy(), float(q), float(s2.log().sum())


@pytest.mark.parametrize("D,N,how", [
    (67, (1 << 21) + 333, "odd-D"),          # three staging blocks of 2^27 / 68 observations, the last one ragged (K1m on 68 features)
    (81, 50_001, "odd-D"),                   # staged, then the team ring kernel (64 < D + 1 <= 96) on 82 stored features
    (95, 10_000, "odd-D"),                   # D + 1 = 96: the widest block grid of that kernel
    (129, 300_001, "odd-D"),                 # two tile rows, D + 1 = 130
    (255, 70_000, "odd-D"),                  # D + 1 = 256 fills the tile row exactly
    (1025, 20_000, "odd-D"),                 # a bias feature next to 1024 learned ones: 9 tile rows
    (96, 200_000, "odd-ld"),                 # even D, leading dimension 97
    (128, 200_000, "misaligned-base"),       # even D and ld, base address 8 bytes off a 16-byte boundary
])
def test_unaligned_colvecs_statistics_match_torch(D, N, how):
    import torch

    ctx = blr.default_context()
    g = torch.Generator(device="cuda").manual_seed(D)
    ld = D + 1 if how == "odd-ld" else D
    off = 1 if how == "misaligned-base" else 0
    flat = torch.randn(N * ld + 2, dtype=torch.float64, device="cuda", generator=g)
    Xfull = flat[off : off + N * ld].view(N, ld)   # (N, ld) row-major == ld x N column-major
    Xv = Xfull[:, :D]
    s2 = torch.exp(torch.randn(N, dtype=torch.float64, device="cuda", generator=g))
    y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    torch.cuda.synchronize()
    assert (Xv.data_ptr() % 16 != 0) == (how == "misaligned-base")
    xh = C.c_void_p()
    ctx.check(ctx.lib.blr_x_wrap_device(ctx.handle, C.c_void_p(Xv.data_ptr()), D, N, ld, L.COLVECS, C.byref(xh)))
    try:
        for zero_mean in (True, False):
            mw = np.zeros(D) if zero_mean else 0.1 * np.random.default_rng(D).standard_normal(D)
            st = blr.Stats(ctx, D)
            yv = blr.DeviceVector.wrap_torch(ctx, y)
            s2v = blr.DeviceVector.wrap_torch(ctx, s2)
            noise = L.Noise(L.NOISE_VECTOR, 0.0, s2v.handle, None, 0)
            ctx.check(ctx.lib.blr_stats_accumulate(ctx.handle, st.handle, mw.ctypes.data_as(C.c_void_p), xh, yv.handle, C.byref(noise)))
            G, r, q, ell, n = st.unpack()
            Go, ro, qo, lo = _torch_stats(torch, Xv, y, s2, mw)
            assert n == N and np.array_equal(G, G.T)
            assert _rel(G, Go) < RTOL and _rel(r, ro) < RTOL, (how, _rel(G, Go), _rel(r, ro))
            assert abs(q - qo) <= RTOL * abs(qo) and abs(ell - lo) <= RTOL * max(abs(lo), 1.0)
    finally:
        ctx.check(ctx.lib.blr_x_free(ctx.handle, xh))


@pytest.mark.parametrize("D", [129, 200, 257, 513])
def test_unaligned_inference_and_marginals_match_oracle(D):
    """Host arrays of odd D go through upload (ld = D): posterior, logpdf, marginals against the oracle, both layouts."""
    rng = np.random.default_rng(D)
    N, Nt = 3 * D + 7, 1000
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N))
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    mw = rng.standard_normal(D)
    B = rng.standard_normal((D, D)) / np.sqrt(D)
    Λ = B @ B.T + np.eye(D)
    fo = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2)
    po, lpo = ref.posterior(fo, y), ref.logpdf(fo, y)
    Xt, σt = rng.standard_normal((D, Nt)), np.exp(rng.standard_normal(Nt))
    mo, vo = ref.mean_and_var(po(ref.ColVecs(Xt), σt))
    ctx = blr.default_context()
    for layout in ("col", "row"):
        if layout == "col":
            x = blr.ColVecs(blr.DeviceMatrix.upload(ctx, X, 0))
            xt = blr.ColVecs(blr.DeviceMatrix.upload(ctx, Xt, 0))
        else:
            x = blr.RowVecs(blr.DeviceMatrix.upload(ctx, np.ascontiguousarray(X.T), 1))
            xt = blr.RowVecs(blr.DeviceMatrix.upload(ctx, np.ascontiguousarray(Xt.T), 1))
        post, lp = blr.posterior_and_logpdf(blr.BayesianLinearRegressor(mw, Λ)(x, blr.DeviceVector.upload(ctx, σ2)),
                                            blr.DeviceVector.upload(ctx, y))
        assert abs(lp - lpo) / abs(lpo) < RTOL
        assert _rel(post.mw, po.mw) < RTOL and _rel(post.Λw.dense(), ref.dense(po.Λw)) < RTOL
        m, v = blr.mean_and_var(post(xt, blr.DeviceVector.upload(ctx, σt)))
        assert _rel(m, mo) < RTOL and _rel(v, vo) < RTOL, (D, layout, _rel(m, mo), _rel(v, vo))


def test_staged_marginals_many_blocks():
    """D = 131, N* = 2.2M unaligned test points: three staging blocks (2^27 / 132 points each) against the aligned evaluation of the
    same points (ld = 132), and the same for a feature-major (RowVecs) copy."""
    import torch

    ctx = blr.default_context()
    D, N = 131, 2_200_003
    rng = np.random.default_rng(5)
    B = rng.standard_normal((D, D)) / np.sqrt(D)
    post = blr.BayesianLinearRegressor(rng.standard_normal(D), B @ B.T + np.eye(D))
    g = torch.Generator(device="cuda").manual_seed(1)
    Xa = torch.zeros((N, D + 1), dtype=torch.float64, device="cuda")          # aligned: ld = 132
    Xa[:, :D] = torch.randn((N, D), dtype=torch.float64, device="cuda", generator=g)
    Xu = Xa[:, :D].contiguous()                                               # dense: ld = 131
    Xr = Xu.T.contiguous()                                                    # feature-major, N odd: D x N row-major
    torch.cuda.synchronize()
    dpost = post._device(ctx)
    noise, keep = blr.runtime.make_noise(ctx, 0.25, N)
    out = {}
    for tag, ptr, ld, layout in (("aligned", Xa.data_ptr(), D + 1, L.COLVECS), ("dense", Xu.data_ptr(), D, L.COLVECS),
                                 ("rowvecs", Xr.data_ptr(), N, L.ROWVECS)):
        xh = C.c_void_p()
        ctx.check(ctx.lib.blr_x_wrap_device(ctx.handle, C.c_void_p(ptr), D, N, ld, layout, C.byref(xh)))
        mv = torch.empty(2 * N, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ctx.check(ctx.lib.blr_mean_var_dev(ctx.handle, dpost.handle, xh, C.byref(noise), C.c_void_p(mv.data_ptr()),
                                           C.c_void_p(mv.data_ptr() + 8 * N)))
        ctx.sync()
        out[tag] = mv.cpu().numpy()
        ctx.check(ctx.lib.blr_x_free(ctx.handle, xh))
    for tag in ("dense", "rowvecs"):
        assert _rel(out[tag][:N], out["aligned"][:N]) < 1e-13 and _rel(out[tag][N:], out["aligned"][N:]) < 1e-13, tag
    # and the aligned evaluation itself against torch on a sample of points
    idx = np.random.default_rng(0).choice(N, 4096, replace=False)
    Xs = Xu[torch.from_numpy(idx).cuda()].cpu().numpy().T                      # D x 4096
    mo, vo = ref.mean_and_var(ref.BayesianLinearRegressor(post.mw, post.Λw.dense() if hasattr(post.Λw, "dense") else post.Λw)(
        ref.ColVecs(Xs), 0.25))
    assert _rel(out["aligned"][:N][idx], mo) < RTOL and _rel(out["aligned"][N:][idx], vo) < RTOL


@pytest.mark.parametrize("D,Nt", [(65, 128), (129, 1000), (257, 70_001), (130, 300)])
def test_unaligned_rand_matches_oracle_and_aligned_device_draws(D, Nt):
    """rand on test points the tensor map cannot address (dense odd D, RowVecs): supplied draws against the oracle; device draws
    (Philox counters are a function of the point's index in the WHOLE problem) identical to the aligned evaluation of the same points."""
    rng = np.random.default_rng(D + Nt)
    B = rng.standard_normal((D, D)) / np.sqrt(D)
    mw, Λ = rng.standard_normal(D), B @ B.T + np.eye(D)
    f, fo = blr.BayesianLinearRegressor(mw, Λ), ref.BayesianLinearRegressor(mw, Λ)
    Xt, σt = rng.standard_normal((D, Nt)), np.exp(rng.standard_normal(Nt))
    ctx = blr.default_context()
    Xpad = np.zeros((D + 2 - D % 2, Nt), order="F")       # aligned copy: even leading dimension
    Xpad[:D] = Xt
    inputs = {
        "dense": blr.ColVecs(blr.DeviceMatrix.upload(ctx, Xt, 0)),
        "rowvecs": blr.RowVecs(blr.DeviceMatrix.upload(ctx, np.ascontiguousarray(Xt.T), 1)),
    }
    σd = blr.DeviceVector.upload(ctx, σt)
    for S in (3, 64, 70):
        Zw, Zy = rng.standard_normal((D, S)), rng.standard_normal((Nt, S))
        Yo = ref.rand(fo(ref.ColVecs(Xt), σt), Zw, Zy)
        for tag, x in inputs.items():
            assert _rel(blr.rand_with_draws(f(x, σd), Zw, Zy), Yo) < RTOL, (tag, S)
    Yd = {tag: blr.rand(blr.DeviceRNG(7), f(x, σd), 5) for tag, x in inputs.items()}
    assert _rel(Yd["rowvecs"], Yd["dense"]) < 1e-12
    # the same draws as a plain numpy input gets (host arrays are uploaded as ColVecs with ld = D: the aligned path when D is even)
    assert _rel(blr.rand(blr.DeviceRNG(7), f(blr.ColVecs(Xt), σt), 5), Yd["dense"]) < 1e-12
