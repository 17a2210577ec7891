"""CPU: the committed golden fixture must still be what the oracle computes (guards both against drift)."""
import os

import numpy as np

from oracle import blr_oracle as ref

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "blr_golden.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files if "/" in k})


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_cases_present():
    assert len(CASES) == 6 and "toy_d2" in CASES
    assert np.array_equal(GOLD["doctest_var"], [2.0, 1.25, 1.0, 1.25, 2.0])


def test_oracle_reproduces_golden():
    for c in CASES:
        g = lambda k: GOLD[f"{c}/{k}"]  # noqa: E731
        Λw = g("Lambda") if bool(g("dense")) else ref.Diagonal(np.diag(g("Lambda")).copy())
        σ2 = g("sigma2")
        f = ref.BayesianLinearRegressor(g("mw"), Λw)
        fx = f(ref.ColVecs(g("X")), σ2 if σ2.ndim else float(σ2))
        assert relerr(ref.rand(fx, g("Zw"), g("Zy")), g("rand")) < 1e-13
        assert abs(ref.logpdf(fx, g("y")) - g("logpdf")) <= 1e-12 * abs(g("logpdf"))
        post = ref.posterior(fx, g("y"))
        assert relerr(post.mw, g("m_post")) < 1e-12
        assert relerr(ref.dense(post.Λw), g("Lambda_post")) < 1e-13
        m, v = ref.mean_and_var(post(ref.ColVecs(g("Xt")), np.finfo(np.float64).eps))
        assert relerr(m, g("mean_t")) < 1e-11 and relerr(v, g("var_t")) < 1e-10
