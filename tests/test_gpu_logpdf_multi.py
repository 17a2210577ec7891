"""GPU: one-pass `logpdf(fx, Y::AbstractMatrix)` (blr_logpdf_multi, csrc/rhs_multi.cu) -- the form the reference's
conformance test calls (test/bayesian_linear_regression.jl:7-9), which the reference evaluates as one full
__compute_inference_quantities (src/bayesian_linear_regression.jl:72-89) per column.  Every column must equal the oracle's
single-vector logpdf to 1e-9, for each kernel regime of the first column's Gram pass (tiny / small / TMA), both layouts, zero and
non-zero prior mean (the shared X'mw pass), scalar and vector noise, diagonal and dense priors, and column counts that are not
multiples of the kernel's column block (8)."""
import math

import numpy as np
import pytest

import blr_b200 as blr
from oracle import blr_oracle as ref

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def problem(D, N, k, seed, zero_mean=False, scalar_noise=False, dense_prior=True):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((D, N))
    σ2 = 0.6 if scalar_noise else np.exp(rng.standard_normal(N))
    sd = math.sqrt(σ2) if scalar_noise else np.sqrt(σ2)[:, None]
    Y = X.T @ rng.standard_normal((D, k)) + sd * rng.standard_normal((N, k))
    Y[:, k // 2] *= 50.0  # columns of very different scale: the per-column terms q_j, z_j'z_j differ by orders of magnitude
    mw = np.zeros(D) if zero_mean else rng.standard_normal(D)
    if dense_prior:
        B = rng.standard_normal((D, D)) / math.sqrt(D)
        Λ = B @ B.T + np.eye(D)
        return X, σ2, Y, mw, Λ, Λ
    lam = np.exp(rng.standard_normal(D))
    return X, σ2, Y, mw, blr.Diagonal(lam), ref.Diagonal(lam)


CASES = [
    # D, N, k, layout, zero_mean, scalar_noise, dense_prior
    (2, 300, 3, "col", False, False, True),       # tiny fused Gram
    (12, 700, 9, "row", False, False, False),     # small fused Gram, k not a multiple of 8
    (40, 1000, 8, "col", True, True, True),       # small Gram + prep, mw = 0 (no X'mw pass), scalar noise
    (130, 900, 5, "col", False, False, True),     # TMA Gram, narrow tile, D not a multiple of the row block
    (256, 4000, 17, "col", False, True, False),   # TMA Gram, unit-noise variant, three column blocks
    (192, 2048, 4, "row", False, False, True),    # feature-major ring + RowVecs staging of the skinny pass
    (700, 1500, 2, "row", True, False, True),     # three row blocks, ragged last one
    (64, 50, 6, "col", False, False, True),       # N < D, one partial observation tile
]


@pytest.mark.parametrize("D,N,k,layout,zero_mean,scalar_noise,dense_prior", CASES)
def test_matrix_logpdf_matches_oracle_per_column(D, N, k, layout, zero_mean, scalar_noise, dense_prior):
    X, σ2, Y, mw, Λ, Λo = problem(D, N, k, seed=D + N + k, zero_mean=zero_mean, scalar_noise=scalar_noise, dense_prior=dense_prior)
    x = blr.ColVecs(X) if layout == "col" else blr.RowVecs(np.ascontiguousarray(X.T))
    fx = blr.BayesianLinearRegressor(mw, Λ)(x, σ2)
    lps = blr.logpdf(fx, Y)
    assert lps.shape == (k,)
    fo = ref.BayesianLinearRegressor(mw, Λo)(ref.ColVecs(X), σ2)
    want = np.array([ref.logpdf(fo, Y[:, j]) for j in range(k)])
    err = np.max(np.abs(lps - want) / np.abs(want))
    print(f"[logpdf_multi] D={D} N={N} k={k} {layout}: max rel err {err:.1e}")
    assert err < RTOL
    # and the vector method on the same device data gives the same numbers (it shares column 0's path exactly)
    assert blr.logpdf(fx, Y[:, 0]) == lps[0]
    assert abs(blr.logpdf(fx, np.ascontiguousarray(Y[:, k - 1])) - lps[k - 1]) <= 1e-12 * abs(lps[k - 1])


def test_matrix_logpdf_device_resident_inputs_and_launch_count():
    """Y as a CUDA tensor, X as a device matrix: nothing visits the host, and the Gram kernel runs once, not k times."""
    import torch

    D, N, k = 256, 6000, 12
    X, σ2, Y, mw, Λ, Λo = problem(D, N, k, seed=5)
    ctx = blr.default_context()
    Xd = blr.DeviceMatrix.upload(ctx, X, 0)
    s2d = blr.DeviceVector.upload(ctx, σ2)
    fx = blr.BayesianLinearRegressor(mw, Λ)(blr.ColVecs(Xd), s2d)
    Yt = torch.from_numpy(Y).cuda()
    n0 = ctx.launch_count()
    lps = blr.logpdf(fx, Yt)
    n_multi = ctx.launch_count() - n0
    n0 = ctx.launch_count()
    lp0 = blr.logpdf(fx, Yt[:, 0].contiguous())
    n_single = ctx.launch_count() - n0
    fo = ref.BayesianLinearRegressor(mw, Λo)(ref.ColVecs(X), σ2)
    want = np.array([ref.logpdf(fo, Y[:, j]) for j in range(k)])
    assert np.max(np.abs(lps - want) / np.abs(want)) < RTOL
    assert lp0 == lps[0]
    # one inference + [X'mw, skinny pass, 2 reductions, diagonal-block inverses] + 2 launches per extra column
    assert n_multi <= n_single + 5 + 2 * (k - 1), (n_multi, n_single)


def test_matrix_logpdf_dense_noise_falls_back_to_columns():
    rng = np.random.default_rng(3)
    D, N, k = 5, 9, 3
    X = rng.standard_normal((D, N))
    A = rng.standard_normal((N, N))
    Σ = A @ A.T + np.eye(N)
    Y = rng.standard_normal((N, k))
    mw = rng.standard_normal(D)
    Λ = np.eye(D) * 2.0
    lps = blr.logpdf(blr.BayesianLinearRegressor(mw, Λ)(blr.ColVecs(X), Σ), Y)
    fo = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), Σ)
    want = np.array([ref.logpdf(fo, Y[:, j]) for j in range(k)])
    assert np.max(np.abs(lps - want) / np.abs(want)) < RTOL


def test_matrix_logpdf_contracts():
    rng = np.random.default_rng(4)
    D, N = 6, 20
    X = rng.standard_normal((D, N))
    f = blr.BayesianLinearRegressor(np.zeros(D), np.eye(D))
    assert blr.logpdf(f(blr.ColVecs(X), 0.1), np.empty((N, 0))).shape == (0,)
    with pytest.raises(blr.DimensionMismatch):            # src/bayesian_linear_regression.jl:74
        blr.logpdf(f(blr.ColVecs(X), 0.1), rng.standard_normal((N + 1, 2)))
    with pytest.raises(blr.PosDefException):              # cholesky(Σy) of a non-positive variance (:79)
        blr.logpdf(f(blr.ColVecs(X), -0.1), rng.standard_normal((N, 2)))
    with pytest.raises(blr.PosDefException):              # prior precision not positive definite (:78)
        blr.logpdf(blr.BayesianLinearRegressor(np.zeros(D), -np.eye(D))(blr.ColVecs(X), 0.1), rng.standard_normal((N, 2)))
