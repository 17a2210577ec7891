"""Pin the oracle -- and, on the GPU box, the CUDA library -- against an independent 50-digit truth (VERDICT r01 "weak" #3).

The reference's own suite checks logpdf against the naive N x N Gaussian only to sqrt(eps) (test/bayesian_linear_regression.jl:
22-38, `≈`).  Here the same naive Gaussian and the closed-form posterior are evaluated with mpmath at 50 significant digits
(tests/highprec.py: no shared code, no shared rounding with either implementation) and both the oracle's literal restatement
and the device path must agree with it to 1e-12 on well-conditioned problems, dense and diagonal Σy.  This removes "parity
unpinned beyond sqrt(eps)" for the oracle's arithmetic: what remains unpinned is only Julia's seed-specific RNG streams.
The extended-precision (longdouble) evaluator used by tests/test_gpu_illcond.py is pinned against mpmath here as well.
"""
import numpy as np
import pytest

from oracle import blr_oracle as ref
from tests import highprec as hp

TOL = 1e-12


def toy(D, N, seed, dense_noise):
    """test/test_utils.jl:4-10: no structure in mean, precision or noise."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((D, N))
    B = rng.standard_normal((D, D))
    mw, Λw = rng.standard_normal(D), B @ B.T + np.eye(D)
    if dense_noise:
        Cn = 0.1 * rng.standard_normal((N, N))
        Σy = Cn @ Cn.T + np.eye(N)
    else:
        Σy = np.exp(rng.standard_normal(N))
    y = X.T @ rng.standard_normal(D) + rng.standard_normal(N)
    Xt = rng.standard_normal((D, 9))
    return X, mw, Λw, Σy, y, Xt


CASES = [(7, 13, True), (7, 13, False), (3, 11, True), (24, 60, False), (16, 40, True)]
_TRUTH = {}


def truth(D, N, dense):
    key = (D, N, dense)
    if key not in _TRUTH:
        X, mw, Λw, Σy, y, Xt = toy(D, N, 100 * D + N, dense)
        t = hp.mp_truth(X, y, Σy, mw, Λw)
        t["mean_t"], t["var_t"] = hp.mp_predict(Xt, t["m_post"], t["Lambda_post"], 0.25)
        _TRUTH[key] = ((X, mw, Λw, Σy, y, Xt), t)
    return _TRUTH[key]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b)) if a.ndim else float(abs(a - b) / abs(b))


@pytest.mark.parametrize("D,N,dense", CASES)
def test_oracle_against_mpmath(D, N, dense):
    (X, mw, Λw, Σy, y, Xt), t = truth(D, N, dense)
    fx = ref.BayesianLinearRegressor(mw, Λw)(ref.ColVecs(X), Σy)
    lp, post = ref.logpdf(fx, y), ref.posterior(fx, y)
    m, v = ref.mean_and_var(post(ref.ColVecs(Xt), 0.25))
    assert rel(lp, t["logpdf"]) < TOL
    assert rel(post.mw, t["m_post"]) < TOL
    assert rel(ref.dense(post.Λw), t["Lambda_post"]) < TOL
    assert rel(m, t["mean_t"]) < TOL and rel(v, t["var_t"]) < TOL


@pytest.mark.parametrize("D,N", [(7, 13), (24, 60)])
def test_longdouble_evaluator_against_mpmath(D, N):
    (X, mw, Λw, Σy, y, Xt), t = truth(D, N, False)
    ld = hp.ld_truth(hp.ld_stats(X, y, Σy, mw), mw, Λw, Xt, 0.25)
    for k in ("logpdf", "m_post", "Lambda_post", "mean_t", "var_t"):
        assert hp.rel(t[k], ld[k]) < 5e-16, k  # mp results are rounded to Float64: agreement to the last bit or two


@pytest.mark.gpu
@pytest.mark.parametrize("D,N,dense", CASES)
def test_gpu_against_mpmath(D, N, dense):
    import blr_b200 as blr

    (X, mw, Λw, Σy, y, Xt), t = truth(D, N, dense)
    f = blr.BayesianLinearRegressor(mw, Λw)
    post, lp = blr.posterior_and_logpdf(f(blr.ColVecs(X), Σy), y)
    m, v = blr.mean_and_var(post(blr.ColVecs(Xt), 0.25))
    assert rel(lp, t["logpdf"]) < TOL
    assert rel(post.mw, t["m_post"]) < TOL
    assert rel(post.Λw.dense(), t["Lambda_post"]) < TOL
    assert rel(m, t["mean_t"]) < TOL and rel(v, t["var_t"]) < TOL
