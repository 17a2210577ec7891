"""Pins the CPU oracle against everything the reference's own suite holds for this path.

Each test cites the reference test it restates.  Shapes and tolerances are the reference's
(`≈` = rtol sqrt(eps) ~ 1.5e-8); moment tests use fewer samples with a tolerance scaled by
1/sqrt(samples) so the whole CPU suite stays in the minutes range.
"""
import math

import numpy as np
import pytest
import scipy.linalg as sl

from oracle import blr_oracle as ref
from tests.toy import as_matrix, generate_toy_problem, make_phi, take

RTOL = math.sqrt(np.finfo(np.float64).eps)
TX = ["Matrix", "ColVecs", "RowVecs"]
ϕ = make_phi(ref)


def isapprox(a, b, rtol=RTOL, atol=0.0):
    """Julia's `≈` for arrays: norm(a-b) <= max(atol, rtol*max(norm(a), norm(b)))."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= max(atol, rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


def rng0():
    return np.random.default_rng(123456)


def test_doctest_golden_vector():
    """src/basis_function_regression.jl:11-28 -- the only literal golden vector in the reference."""
    x = ref.RowVecs(np.linspace(-1.0, 1.0, 5)[:, None])
    blr = ref.BayesianLinearRegressor(np.zeros(2), ref.Diagonal(np.ones(2)))
    bfr = ref.BasisFunctionRegressor(blr, ϕ)
    assert np.array_equal(ref.var(bfr(x)), np.array([2.0, 1.25, 1.0, 1.25, 2.0]))


@pytest.mark.parametrize("Tx", TX)
def test_public_interface_consistency(Tx):
    """AbstractGPs.TestUtils.test_finitegp_primary_and_secondary_public_interface as used at
    test/bayesian_linear_regression.jl:3-10 (N=11, D=3): shapes, cov symmetric + PSD,
    var == diag(cov), mean_and_* agree, logpdf is a real, posterior is a regressor."""
    rng = rng0()
    N, D = 11, 3
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, ref)
    fx = f(X, Σy)
    m, C = ref.mean_and_cov(fx)
    assert m.shape == (N,) and C.shape == (N, N)
    assert np.allclose(C, C.T, atol=1e-12)
    assert np.linalg.eigvalsh(C).min() > -1e-12
    m2, v = ref.mean_and_var(fx)
    assert np.array_equal(m, m2)
    np.testing.assert_allclose(v, np.diag(C), rtol=RTOL)
    np.testing.assert_allclose(ref.var(fx), v, rtol=0, atol=0)
    mm, ss = ref.marginals(fx)
    np.testing.assert_allclose(ss**2, v, rtol=1e-14)
    Y = ref.rand(fx, rng.standard_normal((D, 4)), rng.standard_normal((N, 4)))
    assert Y.shape == (N, 4)
    y = Y[:, 0]
    lp = ref.logpdf(fx, y)
    assert isinstance(lp, float) and np.isfinite(lp)
    # logpdf(fx, Y::Matrix) of AbstractGPs goes through mean_and_cov: same numbers column by column.
    cf = sl.cho_factor(C)
    for j in range(4):
        δ = Y[:, j] - m
        naive = -(N * ref.LOG2PI + 2 * np.sum(np.log(np.diag(cf[0]))) + δ @ sl.cho_solve(cf, δ)) / 2
        assert ref.logpdf(fx, Y[:, j]) == pytest.approx(naive, rel=RTOL)
    assert isinstance(ref.posterior(fx, y), ref.BayesianLinearRegressor)


@pytest.mark.parametrize("Tx", TX)
def test_rand_moments(Tx):
    """test/bayesian_linear_regression.jl:11-21 (1e6 samples, atol=rtol=1e-2 there; 2e5 here at 2.5e-2)."""
    rng = rng0()
    N, D, S = 11, 3, 200_000
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, ref)
    fx = f(X, Σy)
    Y = ref.rand(fx, rng.standard_normal((D, S)), rng.standard_normal((N, S)))
    m_emp = Y.mean(axis=1)
    Yc = Y - m_emp[:, None]
    Σ_emp = Yc @ Yc.T / S
    np.testing.assert_allclose(ref.mean(fx), m_emp, atol=2.5e-2, rtol=2.5e-2)
    np.testing.assert_allclose(ref.cov(fx), Σ_emp, atol=6e-2, rtol=2.5e-2)


@pytest.mark.parametrize("Tx", TX)
def test_logpdf_vs_naive_gaussian(Tx):
    """test/bayesian_linear_regression.jl:22-38 -- known answer by construction (N=13, D=7)."""
    rng = rng0()
    N, D = 13, 7
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, ref)
    fx = f(X, Σy)
    y = ref.rand(fx, rng.standard_normal((D, 1)), rng.standard_normal((N, 1)))[:, 0]
    Xm = as_matrix(X, ref)
    m = Xm.T @ f.mw
    Σ = Xm.T @ sl.cho_solve(sl.cho_factor(ref.dense(f.Λw)), Xm) + Σy
    δ = y - m
    _, logdet = np.linalg.slogdet(Σ)
    expect = -(N * math.log(2 * math.pi) + logdet + δ @ np.linalg.solve(Σ, δ)) / 2
    assert ref.logpdf(fx, y) == pytest.approx(expect, rel=RTOL)


@pytest.mark.parametrize("Tx", TX)
def test_posterior_low_noise(Tx):
    """test/bayesian_linear_regression.jl:40-48 -- noise eps(): interpolation, tiny covariance."""
    rng = rng0()
    N, D = 13, 7
    eps = np.finfo(np.float64).eps
    X, f, _ = generate_toy_problem(rng, N, D, Tx, ref)
    y = ref.rand(f(X, eps), rng.standard_normal((D, 1)), rng.standard_normal((N, 1)))[:, 0]
    fp = ref.posterior(f(X, eps), y)
    assert isapprox(ref.mean(fp(X, eps)), y)
    assert np.all(ref.cov(fp(X, eps)) < 1000 * eps)


@pytest.mark.parametrize("Tx", TX)
def test_posterior_repeated_conditioning(Tx):
    """test/bayesian_linear_regression.jl:49-70."""
    rng = rng0()
    N, D = 13, 7
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, ref)
    Xp = rng.standard_normal((D, N))
    y = ref.rand(f(X, Σy), rng.standard_normal((D, 1)), rng.standard_normal((N, 1)))[:, 0]
    N1 = N - 3
    Σ1, Σ2 = Σy[:N1, :N1], Σy[N1:, N1:]
    Σyp = np.block([[Σ1, np.zeros((N1, N - N1))], [np.zeros((N - N1, N1)), Σ2]])
    X1, X2 = take(X, slice(0, N1), ref), take(X, slice(N1, N), ref)
    y1, y2 = y[:N1], y[N1:]
    f1 = ref.posterior(f(X1, Σ1), y1)
    f2 = ref.posterior(f1(X2, Σ2), y2)
    fp = ref.posterior(f(X, Σyp), y)
    assert isapprox(ref.mean(fp(Xp, Σy)), ref.mean(f2(Xp, Σy)))
    assert isapprox(ref.cov(fp(Xp, Σy)), ref.cov(f2(Xp, Σy)))


def test_pdmat_closure():
    """test/bayesian_linear_regression.jl:71-113 -- PDMat prior => PDMat posterior, Symmetric => Symmetric."""
    rng = rng0()
    N, D = 13, 7
    X = rng.standard_normal((D, N))
    Xp = rng.standard_normal((D, N))
    U = np.triu(rng.standard_normal((D, D)))
    C = 0.1 * rng.standard_normal((N, N))
    mw, Σy = rng.standard_normal(D), C @ C.T + np.eye(N)
    Λ = U.T @ U + np.eye(D)
    f_pd = ref.BayesianLinearRegressor(mw, ref.PDMat.from_matrix(Λ))
    f_sym = ref.BayesianLinearRegressor(mw, ref.Symmetric(Λ))
    fx_pd, fx_sym = f_pd(X, Σy), f_sym(X, Σy)
    y = ref.rand(fx_pd, rng.standard_normal((D, 1)), rng.standard_normal((N, 1)))[:, 0]
    fp_pd, fp_sym = ref.posterior(fx_pd, y), ref.posterior(fx_sym, y)
    assert isinstance(fp_pd.Λw, ref.PDMat)
    assert isinstance(fp_sym.Λw, ref.Symmetric)
    assert isapprox(ref.mean(fp_pd(Xp, Σy)), ref.mean(fp_sym(Xp, Σy)))
    assert isapprox(ref.cov(fp_pd(Xp, Σy)), ref.cov(fp_sym(Xp, Σy)))


def test_unrecognised_abstract_vector():
    """test/bayesian_linear_regression.jl:116-122 -- a vector of rows is rejected with an ErrorException."""
    rng = rng0()
    N, D = 11, 5
    x = [row for row in rng.standard_normal((N, D))]
    _, f, Σy = generate_toy_problem(rng, N, D, "ColVecs", ref)
    with pytest.raises(RuntimeError):
        ref.rand(f(x, Σy), rng.standard_normal((D, 1)), rng.standard_normal((N, 1)))


def test_length_mismatch_raises():
    """src/bayesian_linear_regression.jl:74."""
    rng = rng0()
    X, f, Σy = generate_toy_problem(rng, 11, 3, "ColVecs", ref)
    with pytest.raises(RuntimeError):
        ref.logpdf(f(X, Σy), np.zeros(10))


def test_not_positive_definite_raises():
    f = ref.BayesianLinearRegressor(np.zeros(2), np.array([[1.0, 2.0], [2.0, 1.0]]))
    with pytest.raises(ref.PosDefException):
        ref.var(f(np.ones((2, 3)), 0.1))


@pytest.mark.parametrize("Tx", TX)
def test_bfr_consistency_with_blr(Tx):
    """test/basis_function_regression.jl:13-28."""
    rng = rng0()
    N, D = 11, 2
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, ref)
    f_bf = ref.BasisFunctionRegressor(f, ϕ)
    y = ref.rand(f_bf(X, Σy), rng.standard_normal((D, 1)), rng.standard_normal((N, 1)))[:, 0]
    assert ref.logpdf(f(ϕ(X), Σy), y) == pytest.approx(ref.logpdf(f_bf(X, Σy), y), rel=RTOL)
    f_bf_post = ref.posterior(f_bf(X, Σy), y)
    f_post = ref.posterior(f(ϕ(X), Σy), y)
    assert isinstance(f_bf_post, ref.BasisFunctionRegressor)
    assert isapprox(ref.mean(f_bf_post(X)), ref.mean(f_post(ϕ(X))))


@pytest.mark.parametrize("Tx", TX)
def test_bfr_rand_moments(Tx):
    """test/basis_function_regression.jl:29-41."""
    rng = rng0()
    N, D, S = 11, 2, 200_000
    X, f, Σy = generate_toy_problem(rng, N, D, Tx, ref)
    fx = ref.BasisFunctionRegressor(f, ϕ)(X, Σy)
    Y = ref.rand(fx, rng.standard_normal((D, S)), rng.standard_normal((N, S)))
    m_emp = Y.mean(axis=1)
    Yc = Y - m_emp[:, None]
    np.testing.assert_allclose(ref.mean(fx), m_emp, atol=2.5e-2, rtol=2.5e-2)
    np.testing.assert_allclose(ref.cov(fx), Yc @ Yc.T / S, atol=6e-2, rtol=2.5e-2)


def test_function_samples():
    """test/sampling_functions.jl:3-47: sample is a fixed function; Matrix / ColVecs / RowVecs agree
    exactly; rand(rng, f, s1, s2) has shape (s1, s2); moments of the sampled functions."""
    rng = rng0()
    N, D = 11, 5
    X, f, Σy = generate_toy_problem(rng, N, D, "Matrix", ref)
    g = ref.rand_function(f, rng.standard_normal(D))
    assert np.array_equal(g(X), g(X))
    assert np.array_equal(g(X), g(ref.ColVecs(X)))
    np.testing.assert_allclose(g(X), g(ref.RowVecs(np.ascontiguousarray(X.T))), rtol=1e-15, atol=1e-15)
    s1, s2 = 300, 400
    gs = ref.rand_functions(f, rng.standard_normal((D, s1 * s2)), (s1, s2))
    assert gs.shape == (s1, s2)
    W = np.stack([h.w for h in gs.reshape(-1, order="F")], axis=1)
    Y = X.T @ W
    m_emp = Y.mean(axis=1)
    Yc = Y - m_emp[:, None]
    Σ_emp = Yc @ Yc.T / (s1 * s2)
    np.testing.assert_allclose(ref.mean(f(X, Σy)), m_emp, atol=2e-2, rtol=2e-2)
    np.testing.assert_allclose(ref.cov(f(X, Σy)), Σ_emp + Σy, atol=6e-2, rtol=3e-2)


def test_function_samples_bfr():
    """test/sampling_functions.jl:49-66."""
    rng = rng0()
    N, D, S = 11, 2, 100_000
    X, f, Σy = generate_toy_problem(rng, N, D, "ColVecs", ref)
    f_bf = ref.BasisFunctionRegressor(f, ϕ)
    gs = ref.rand_functions(f_bf, rng.standard_normal((D, S)), (S,))
    Y = np.stack([h(X) for h in gs[:2000]], axis=1)
    W = np.stack([h.w for h in gs], axis=1)
    Yall = ϕ(X).X.T @ W
    np.testing.assert_allclose(Y, Yall[:, :2000], rtol=1e-13, atol=1e-13)
    m_emp = Yall.mean(axis=1)
    Yc = Yall - m_emp[:, None]
    np.testing.assert_allclose(ref.mean(f_bf(X, Σy)), m_emp, atol=2e-2, rtol=2e-2)
    np.testing.assert_allclose(ref.cov(f_bf(X, Σy)), Yc @ Yc.T / S + Σy, atol=8e-2, rtol=3e-2)
