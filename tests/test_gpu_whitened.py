"""GPU: the reference's literal numerical form on the device (`Context.set_form("whitened")`, BLR_FORM_WHITENED, csrc/whitened.cu).

With this form the library evaluates the inference quantities in the reference's own order -- `Λεy = chol(Uw⁻ᵀ G Uw⁻¹ + I)`
(src/bayesian_linear_regression.jl:81,86), `mεy = Λεy \\ (Bt'δy)` (:64), `T = Λεy.U * Uw` (:67), `m' = mw + Uw \\ mεy` (:68) -- and
applies factors by triangular SOLVES (`Uw' \\ X` in var / cov, :36,:41; `Uw \\ randn` in rand, :51); no inverse is formed anywhere.
Everything the default (direct) form is tested for must hold here too: 1e-9 against the oracle for every output, for dense,
PDMat and Diagonal priors, both input layouts, D on either side of the 64-row block of the solver and right-hand-side counts on either
side of its 32-column slab; the ill-conditioned regimes of tests/test_gpu_illcond.py under the same well-posed criterion; and
agreement with the direct form to rounding on well-conditioned problems."""
import math

import numpy as np
import pytest

import blr_b200 as blr
from oracle import blr_oracle as ref
from tests.test_gpu_illcond import CASES as ILL_CASES
from tests.test_gpu_illcond import test_ill_conditioned_against_extended_precision as _ill_case
from tests.test_gpu_illcond import test_sequential_conditioning_posterior_as_prior as _seq_case

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture
def whitened():
    ctx = blr.default_context()
    ctx.set_form("whitened")
    assert ctx.form() == "whitened"
    yield ctx
    ctx.set_form("direct")
    assert ctx.form() == "direct"


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def problem(D, N, seed, diagonal=False):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N))
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    mw = rng.standard_normal(D)
    if diagonal:
        return X, σ2, y, mw, np.exp(rng.standard_normal(D))
    B = rng.standard_normal((D, D)) / math.sqrt(D)
    return X, σ2, y, mw, B @ B.T + np.eye(D)


@pytest.mark.parametrize("layout", ["col", "row"])
@pytest.mark.parametrize("D,N", [(1, 5), (2, 10), (7, 40), (33, 200), (63, 100), (64, 300), (65, 300), (100, 50), (128, 700),
                                 (200, 900), (256, 1500), (513, 1200), (1024, 2100)])
def test_whitened_inference_and_prediction_match_oracle(whitened, D, N, layout):
    X, σ2, y, mw, Λ = problem(D, N, seed=31 * D + N)
    wrap = (lambda A: blr.ColVecs(A)) if layout == "col" else (lambda A: blr.RowVecs(np.ascontiguousarray(A.T)))
    post, lp = blr.posterior_and_logpdf(blr.BayesianLinearRegressor(mw, blr.PDMat(Λ))(wrap(X), σ2), y)
    fxo = ref.BayesianLinearRegressor(mw, ref.PDMat.from_matrix(Λ))(ref.ColVecs(X), σ2)
    po, lpo = ref.posterior(fxo, y), ref.logpdf(fxo, y)
    assert abs(lp - lpo) / abs(lpo) < RTOL
    assert relerr(post.mw, po.mw) < RTOL
    assert relerr(post.Λw.dense(), ref.dense(po.Λw)) < RTOL
    T = post.Λw.U
    assert np.allclose(T, np.triu(T)) and np.all(np.diag(T) > 0)
    assert relerr(T, ref.posterior_factor_T(fxo, y)) < RTOL      # T = Λεy.U * Uw (:67), here as (Lw Lε)'
    assert relerr(T.T @ T, post.Λw.dense()) < 1e-11
    rng = np.random.default_rng(D)
    for Nt in (1, 31, 32, 33, 97):  # either side of the solver's 32-column slab
        Xt, σt = rng.standard_normal((D, Nt)), np.exp(rng.standard_normal(Nt))
        mo, vo = ref.mean_and_var(po(ref.ColVecs(Xt), σt))
        m, v = blr.mean_and_var(post(wrap(Xt), σt))
        assert relerr(m, mo) < RTOL and relerr(v, vo) < RTOL, (Nt, relerr(m, mo), relerr(v, vo))
        if Nt <= 33:
            assert relerr(blr.cov(post(wrap(Xt), σt)), ref.cov(po(ref.ColVecs(Xt), σt))) < RTOL
        for S in (1, 33):
            Zw, Zy = rng.standard_normal((D, S)), rng.standard_normal((Nt, S))
            assert relerr(blr.rand_with_draws(post(wrap(Xt), σt), Zw, Zy), ref.rand(po(ref.ColVecs(Xt), σt), Zw, Zy)) < RTOL


@pytest.mark.parametrize("D,N", [(5, 30), (64, 300), (130, 400), (512, 1000)])
def test_whitened_diagonal_prior_and_scalar_noise(whitened, D, N):
    X, _, y, mw, lam = problem(D, N, seed=D + 3, diagonal=True)
    post, lp = blr.posterior_and_logpdf(blr.BayesianLinearRegressor(mw, blr.Diagonal(lam))(blr.ColVecs(X), 0.3), y)
    fxo = ref.BayesianLinearRegressor(mw, ref.Diagonal(lam))(ref.ColVecs(X), 0.3)
    po = ref.posterior(fxo, y)
    assert abs(lp - ref.logpdf(fxo, y)) / abs(lp) < RTOL
    assert relerr(post.mw, po.mw) < RTOL and relerr(post.Λw.dense(), ref.dense(po.Λw)) < RTOL
    Xt = np.random.default_rng(D).standard_normal((D, 77))
    mo, vo = ref.mean_and_var(po(ref.ColVecs(Xt), 0.3))
    m, v = blr.mean_and_var(post(blr.ColVecs(Xt), 0.3))
    assert relerr(m, mo) < RTOL and relerr(v, vo) < RTOL


@pytest.mark.parametrize("D,N", [(48, 400), (256, 1500), (1024, 2100)])
def test_whitened_and_direct_forms_agree(D, N):
    """Same inputs through both forms of the same context: outputs agree far inside the parity bar on a well-conditioned problem."""
    X, σ2, y, mw, Λ = problem(D, N, seed=5 * D)
    ctx = blr.default_context()
    Xt = np.random.default_rng(D).standard_normal((D, 200))
    out = {}
    for form in ("direct", "whitened"):
        ctx.set_form(form)
        try:
            post, lp = blr.posterior_and_logpdf(blr.BayesianLinearRegressor(mw, blr.PDMat(Λ))(blr.ColVecs(X), σ2), y)
            m, v = blr.mean_and_var(post(blr.ColVecs(Xt), 0.1))
            out[form] = (lp, post.mw, post.Λw.U, m, v)
        finally:
            ctx.set_form("direct")
    a, b = out["direct"], out["whitened"]
    assert abs(a[0] - b[0]) / abs(a[0]) < 1e-12
    for i in (1, 2, 3, 4):
        assert relerr(a[i], b[i]) < 1e-11, (i, relerr(a[i], b[i]))


def test_whitened_logpdf_matrix_and_pos_def_error(whitened):
    D, N, k = 96, 500, 5
    X, σ2, y, mw, Λ = problem(D, N, seed=77)
    Y = np.stack([y + 0.1 * j for j in range(k)], axis=1)
    f = blr.BayesianLinearRegressor(mw, Λ)
    lps = blr.logpdf(f(blr.ColVecs(X), σ2), Y)
    fxo = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2)
    for j in range(k):
        lpo = ref.logpdf(fxo, Y[:, j])
        assert abs(lps[j] - lpo) / abs(lpo) < RTOL
    bad = Λ.copy()
    bad[5, 5] = -1.0  # the prior precision is not positive definite: the reference throws PosDefException from _cholesky(Λw) (:78)
    with pytest.raises(blr.PosDefException):
        blr.posterior(blr.BayesianLinearRegressor(mw, bad)(blr.ColVecs(X), σ2), y)


@pytest.mark.parametrize("tag,D,N,lam,noise,noise_t", ILL_CASES, ids=[f"{c[0]}-D{c[1]}" for c in ILL_CASES])
def test_whitened_ill_conditioned_against_extended_precision(whitened, tag, D, N, lam, noise, noise_t):
    """The stress cases of tests/test_gpu_illcond.py under the literal form: same truth, same criterion."""
    _ill_case(tag, D, N, lam, noise, noise_t)


@pytest.mark.parametrize("D,N", [(64, 600), (256, 2000)])
def test_whitened_sequential_conditioning(whitened, D, N):
    _seq_case(D, N)
