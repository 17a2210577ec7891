"""The streaming Gram-form oracle (used at sizes the literal form cannot allocate) must agree with
the literal restatement of src/bayesian_linear_regression.jl:55-89."""
import numpy as np
import pytest

from oracle import blr_oracle as ref


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("D,N,dense_prior,zero_mean", [(7, 13, True, False), (64, 5000, True, False), (256, 20000, False, True)])
def test_streaming_matches_literal(D, N, dense_prior, zero_mean):
    rng = np.random.default_rng(7)
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N))
    mw = np.zeros(D) if zero_mean else rng.standard_normal(D)
    if dense_prior:
        B = rng.standard_normal((D, D))
        Λw = B @ B.T + np.eye(D)
    else:
        Λw = ref.Diagonal(np.ones(D))
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    fx = ref.BayesianLinearRegressor(mw, Λw)(ref.ColVecs(X), σ2)
    lp_lit = ref.logpdf(fx, y)
    post = ref.posterior(fx, y)
    T_lit = ref.posterior_factor_T(fx, y)
    lp, m, T = ref.infer_streaming(mw, Λw, X, y, σ2, chunk=1024)
    assert abs(lp - lp_lit) / abs(lp_lit) < 1e-11
    assert relerr(m, post.mw) < 1e-11
    assert relerr(T.T @ T, ref.dense(post.Λw)) < 1e-12
    assert relerr(T, T_lit) < 1e-11
    # T == chol(Λw + G).U: the factor the device path computes directly.
    G, _, _, _ = ref.gram_stats(X, y, σ2, mw, chunk=1024)
    assert relerr(ref.chol_upper(ref.dense(Λw) + G), T_lit) < 1e-11
