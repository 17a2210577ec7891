"""GPU: the reference's own test-suite (dense Σy fixtures of test/test_utils.jl:4-20), restated against the CUDA path.

Same shapes and tolerances as the reference (`≈` = rtol sqrt(eps), norm-wise); each test cites the reference test it
restates.  The dense-noise side path whitens X and y with chol(Σy) on the device and
then runs the same Gram / Cholesky kernels as the diagonal path.
"""
import math

import numpy as np
import pytest
import scipy.linalg as sl

import blr_b200 as blr
from oracle import blr_oracle as ref
from tests.toy import as_matrix, generate_toy_problem, make_phi, take

pytestmark = pytest.mark.gpu
RTOL = math.sqrt(np.finfo(np.float64).eps)
TX = ["Matrix", "ColVecs", "RowVecs"]
ϕ = make_phi(blr)


def isapprox(a, b, rtol=RTOL, atol=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= max(atol, rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


def rng0():
    return np.random.default_rng(123456)


def oracle_twin(f, X, Σy):
    fo = ref.BayesianLinearRegressor(f.mw, f.Λw if isinstance(f.Λw, np.ndarray) else f.Λw.dense())
    return fo(ref.ColVecs(as_matrix(X, blr)), Σy)


@pytest.mark.parametrize("Tx", TX)
def test_public_interface_consistency(Tx):
    """test/bayesian_linear_regression.jl:3-10 (N=11, D=3, dense Σy)."""
    rng = rng0()
    N, D = 11, 3
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, blr)
    fx = f(X, Σy)
    m, C = blr.mean_and_cov(fx)
    assert m.shape == (N,) and C.shape == (N, N)
    assert np.allclose(C, C.T, atol=1e-12) and np.linalg.eigvalsh(C).min() > -1e-12
    m2, v = blr.mean_and_var(fx)
    assert isapprox(m, m2, rtol=1e-14) and isapprox(v, np.diag(C))
    fxo = oracle_twin(f, X, Σy)
    assert isapprox(C, ref.cov(fxo), rtol=1e-9) and isapprox(v, ref.var(fxo), rtol=1e-9) and isapprox(m, ref.mean(fxo), rtol=1e-9)
    Y = blr.rand(rng, fx, 4)
    assert Y.shape == (N, 4) and blr.rand(rng, fx).shape == (N,)
    lp = blr.logpdf(fx, Y[:, 0])
    assert isinstance(lp, float) and np.isfinite(lp)
    lps = blr.logpdf(fx, Y)  # logpdf(fx, Y::Matrix) of AbstractGPs: one density per column
    assert lps.shape == (4,) and lps[0] == lp
    assert blr.rand(None, fx).shape == (N,)  # rand(fx) with the default RNG
    assert isinstance(blr.posterior(fx, Y[:, 0]), blr.BayesianLinearRegressor)


@pytest.mark.parametrize("Tx", TX)
def test_rand_with_dense_noise_matches_oracle(Tx):
    """src/bayesian_linear_regression.jl:52 with a dense Σy: X'w .+ Σy.U' * Zy under the same draws."""
    rng = rng0()
    N, D, S = 11, 3, 6
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, blr)
    Zw, Zy = rng.standard_normal((D, S)), rng.standard_normal((N, S))
    Y = blr.rand_with_draws(f(X, Σy), Zw, Zy)
    assert isapprox(Y, ref.rand(oracle_twin(f, X, Σy), Zw, Zy), rtol=1e-9)


@pytest.mark.parametrize("Tx", TX)
def test_rand_moments(Tx):
    """test/bayesian_linear_regression.jl:11-21 (device Philox draws, 2e5 samples)."""
    rng = rng0()
    N, D, S = 11, 3, 200_000
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, blr)
    Y = blr.rand(blr.DeviceRNG(7), f(X, Σy), S)
    m_emp = Y.mean(axis=1)
    Yc = Y - m_emp[:, None]
    np.testing.assert_allclose(blr.mean(f(X, Σy)), m_emp, atol=2.5e-2, rtol=2.5e-2)
    np.testing.assert_allclose(blr.cov(f(X, Σy)), Yc @ Yc.T / S, atol=6e-2, rtol=2.5e-2)


@pytest.mark.parametrize("Tx", TX)
def test_logpdf_vs_naive_gaussian(Tx):
    """test/bayesian_linear_regression.jl:22-38 -- known answer by construction (N=13, D=7, dense Σy)."""
    rng = rng0()
    N, D = 13, 7
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, blr)
    y = blr.rand(rng, f(X, Σy))
    Xm = as_matrix(X, blr)
    m = Xm.T @ f.mw
    Σ = Xm.T @ sl.cho_solve(sl.cho_factor(f.Λw), Xm) + Σy
    δ = y - m
    _, logdet = np.linalg.slogdet(Σ)
    expect = -(N * math.log(2 * math.pi) + logdet + δ @ np.linalg.solve(Σ, δ)) / 2
    assert blr.logpdf(f(X, Σy), y) == pytest.approx(expect, rel=RTOL)
    assert blr.logpdf(f(X, Σy), y) == pytest.approx(ref.logpdf(oracle_twin(f, X, Σy), y), rel=1e-9)


@pytest.mark.parametrize("Tx", TX)
def test_posterior_low_noise(Tx):
    """test/bayesian_linear_regression.jl:40-48."""
    rng = rng0()
    N, D = 13, 7
    eps = np.finfo(np.float64).eps
    X, f, _ = generate_toy_problem(rng, N, D, Tx, blr)
    y = blr.rand(rng, f(X, eps))
    fp = blr.posterior(f(X, eps), y)
    assert isapprox(blr.mean(fp(X, eps)), y)
    assert np.all(blr.cov(fp(X, eps)) < 1000 * eps)


@pytest.mark.parametrize("Tx", TX)
def test_posterior_repeated_conditioning(Tx):
    """test/bayesian_linear_regression.jl:49-70 (block-diagonal dense noise)."""
    rng = rng0()
    N, D = 13, 7
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, blr)
    Xp = rng.standard_normal((D, N))
    y = blr.rand(rng, f(X, Σy))
    N1 = N - 3
    Σ1, Σ2 = Σy[:N1, :N1], Σy[N1:, N1:]
    Σyp = np.block([[Σ1, np.zeros((N1, N - N1))], [np.zeros((N - N1, N1)), Σ2]])
    X1, X2 = take(X, slice(0, N1), blr), take(X, slice(N1, N), blr)
    f1 = blr.posterior(f(X1, Σ1), y[:N1])
    f2 = blr.posterior(f1(X2, Σ2), y[N1:])
    fp = blr.posterior(f(X, Σyp), y)
    assert isapprox(blr.mean(fp(Xp, Σy)), blr.mean(f2(Xp, Σy)))
    assert isapprox(blr.cov(fp(Xp, Σy)), blr.cov(f2(Xp, Σy)))
    po = ref.posterior(oracle_twin(f, X, Σyp), y)
    assert isapprox(fp.mw, po.mw, rtol=1e-9) and isapprox(fp.Λw.dense(), ref.dense(po.Λw), rtol=1e-9)


def test_pdmat_closure():
    """test/bayesian_linear_regression.jl:71-113."""
    rng = rng0()
    N, D = 13, 7
    X, Xp = rng.standard_normal((D, N)), rng.standard_normal((D, N))
    U = np.triu(rng.standard_normal((D, D)))
    C = 0.1 * rng.standard_normal((N, N))
    mw, Σy = rng.standard_normal(D), C @ C.T + np.eye(N)
    Λ = U.T @ U + np.eye(D)
    f_pd, f_sym = blr.BayesianLinearRegressor(mw, blr.PDMat(Λ)), blr.BayesianLinearRegressor(mw, blr.Symmetric(Λ))
    y = blr.rand(rng, f_pd(X, Σy))
    fp_pd, fp_sym = blr.posterior(f_pd(X, Σy), y), blr.posterior(f_sym(X, Σy), y)
    assert isinstance(fp_pd.Λw, blr.PDMat) and isinstance(fp_sym.Λw, blr.Symmetric)
    assert isapprox(blr.mean(fp_pd(Xp, Σy)), blr.mean(fp_sym(Xp, Σy)))
    assert isapprox(blr.cov(fp_pd(Xp, Σy)), blr.cov(fp_sym(Xp, Σy)))


def test_dense_noise_not_positive_definite():
    X = np.ones((2, 3))
    f = blr.BayesianLinearRegressor(np.zeros(2), blr.Diagonal(np.ones(2)))
    bad = np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    with pytest.raises(blr.PosDefException):
        blr.logpdf(f(X, bad), np.zeros(3))
    with pytest.raises(blr.DimensionMismatch):
        blr.logpdf(f(X, np.eye(4)), np.zeros(3))


@pytest.mark.parametrize("Tx", TX)
def test_bfr_consistency_with_blr(Tx):
    """test/basis_function_regression.jl:13-28 (dense Σy)."""
    rng = rng0()
    N, D = 11, 2
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, blr)
    f_bf = blr.BasisFunctionRegressor(f, ϕ)
    y = blr.rand(rng, f_bf(X, Σy))
    assert blr.logpdf(f(ϕ(X), Σy), y) == pytest.approx(blr.logpdf(f_bf(X, Σy), y), rel=RTOL)
    f_bf_post, f_post = blr.posterior(f_bf(X, Σy), y), blr.posterior(f(ϕ(X), Σy), y)
    assert isapprox(blr.mean(f_bf_post(X)), blr.mean(f_post(ϕ(X))))


@pytest.mark.parametrize("D,N", [(64, 200), (96, 513)])
def test_dense_noise_larger_problem_matches_oracle(D, N):
    """Dense Σy through the fast Gram path (D >= 64): whitened Ã is a RowVecs matrix."""
    rng = np.random.default_rng(D + N)
    X = rng.standard_normal((D, N))
    B, Cn = rng.standard_normal((D, D)), 0.1 * rng.standard_normal((N, N))
    mw, Λ, Σy = rng.standard_normal(D), B @ B.T + np.eye(D), Cn @ Cn.T + np.eye(N)
    y = X.T @ rng.standard_normal(D) + rng.standard_normal(N)
    f, fo = blr.BayesianLinearRegressor(mw, Λ), ref.BayesianLinearRegressor(mw, Λ)
    post, lp = blr.posterior_and_logpdf(f(blr.ColVecs(X), Σy), y)
    fxo = fo(ref.ColVecs(X), Σy)
    assert lp == pytest.approx(ref.logpdf(fxo, y), rel=1e-9)
    po = ref.posterior(fxo, y)
    assert isapprox(post.mw, po.mw, rtol=1e-9) and isapprox(post.Λw.dense(), ref.dense(po.Λw), rtol=1e-9)
