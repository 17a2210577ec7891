#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- CPU study behind a design decision (run: `python -m tests.illcond_study > profiles/r02/illcond_study.txt`).

Question (VERDICT r01 "missing" #2): the device factorises Λ' = Λw + G directly and applies an explicit inverse factor
W = L'^-1 in `var`; the reference whitens first (`Bt = Σy.U' \\ (Uw' \\ X)'`, `chol(Bt'Bt + I)`,
src/bayesian_linear_regression.jl:81,86) and uses a triangular solve in `var` (:41).  Is the device's form a latent 1e-9
violation once Λw / Λ' are ill-conditioned?

Method: both forms in Float64 on the CPU (the oracle's literal restatement = the whitened form; the direct form written out
below with numpy/scipy exactly as the kernels compute it: chol(Λw + G), substitution, W = L^-1 by substitution), each compared
with an extended-precision evaluation (tests/highprec.py, longdouble, eps 1e-19, itself pinned to 50-digit mpmath).

Reading the table: every entry is a relative error against the truth.  "whitened" is what the reference's op sequence gives.
The direct form is at least as accurate in every regime (often 10x better for m' and var); the explicit inverse in `var` is
indistinguishable from the triangular solve; where cond * eps exceeds 1e-9 NEITHER form reaches 1e-9 -- and the two forms
differ from each other by more than 1e-9 there, so "agree with the reference to 1e-9" is not a well-posed requirement in that
regime; "no further from the truth than the reference" is, and tests/test_gpu_illcond.py asserts exactly that on the GPU.
"""
import numpy as np
import scipy.linalg as sl

from oracle import blr_oracle as ref
from tests import highprec as hp

EPS = np.finfo(float).eps


def direct_form(mw, Λ, G, r, q, ell, N, Xt, noise_t):
    Lp, Lw = np.linalg.cholesky(Λ + G), np.linalg.cholesky(Λ)
    z = sl.solve_triangular(Lp, r, lower=True)
    u = sl.solve_triangular(Lp.T, z, lower=False)
    lp = -(N * np.log(2 * np.pi) + ell + q + 2 * np.log(np.diag(Lp)).sum() - 2 * np.log(np.diag(Lw)).sum() - z @ z) / 2
    W = sl.solve_triangular(Lp, np.eye(len(mw)), lower=True)
    return lp, mw + u, ((W @ Xt) ** 2).sum(0) + noise_t, (sl.solve_triangular(Lp, Xt, lower=True) ** 2).sum(0) + noise_t


def equivalent_orders(mw, Λ, G, r, q, ell, N):
    """Three more backward-stable Float64 evaluation orders of the same quantities on the CPU -- the device's direct form, and
    the reference's whitened form with the whitening applied by triangular solves and by an explicit inverse factor.  Returns
    {name: (logpdf, m')}.  Together with the oracle's literal statement order they give the SPREAD that rounding alone produces
    between correct implementations; tests/test_gpu_illcond.py uses the largest error of the set as the yardstick, because a
    single implementation's error on a single draw is a random variable (seed ensemble at the end of the study output: the
    oracle's own logpdf error at cond(Λw) = 1e13 ranges from 5e-10 to 3e-8 over five seeds, and so does every other order's)."""
    D = len(mw)
    Lw = np.linalg.cholesky(Λ)
    out = {}
    lp, m, _, _ = direct_form(mw, Λ, G, r, q, ell, N, np.zeros((D, 1)), 0.0)
    out["direct"] = (lp, m)
    eye = np.eye(D)
    for name, white in (("whitened-trsm", lambda B: sl.solve_triangular(Lw, B, lower=True)),
                        ("whitened-inverse", lambda B, Ww=sl.solve_triangular(Lw, eye, lower=True): Ww @ B)):
        A = white(white(G).T) + eye
        LA = np.linalg.cholesky((A + A.T) / 2)
        z = sl.solve_triangular(LA, white(r), lower=True)
        lp = -(N * np.log(2 * np.pi) + ell + q + 2 * np.log(np.diag(LA)).sum() - z @ z) / 2
        u = sl.solve_triangular(LA.T, z, lower=False)
        out[name] = (lp, mw + sl.solve_triangular(Lw.T, u, lower=False))
    return out


def ensemble(D, N, lam, seeds):
    """logpdf / m' error of each evaluation order over several seeds of the same regime."""
    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        X = rng.standard_normal((D, N))
        σ2 = np.exp(rng.standard_normal(N))
        y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
        Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        Λ = (Q * np.geomspace(lam[0], lam[1], D)) @ Q.T
        Λ = (Λ + Λ.T) / 2
        mw = rng.standard_normal(D)
        tr = hp.ld_truth(hp.ld_stats(X, y, σ2, mw), mw, Λ)
        G, r, q, ell = ref.gram_stats(X, y, σ2, mw)
        fx = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2)
        res = {"reference order": (ref.logpdf(fx, y), ref.posterior(fx, y).mw)}
        res.update(equivalent_orders(mw, Λ, G, r, q, ell, N))
        print(f"λ in [{lam[0]:g}, {lam[1]:g}] D={D} seed {seed}: " + " | ".join(
            f"{k}: logpdf {hp.rel(v[0], tr['logpdf']):7.1e} m' {hp.rel(v[1], tr['m_post']):7.1e}" for k, v in res.items()))


def run(tag, D, N, lam, noise, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((D, N))
    σ2 = np.exp(rng.standard_normal(N)) if noise is None else np.full(N, noise)
    noise_t = 0.1 if noise is None else noise
    y = X.T @ rng.standard_normal(D) + np.sqrt(σ2) * rng.standard_normal(N)
    Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    Λ = (Q * np.geomspace(lam[0], lam[1], D)) @ Q.T
    Λ = (Λ + Λ.T) / 2
    mw = rng.standard_normal(D)
    Xt = np.concatenate([X[:, : min(24, N)], rng.standard_normal((D, 24))], axis=1)
    tr = hp.ld_truth(hp.ld_stats(X, y, σ2, mw), mw, Λ, Xt, noise_t)
    G, r, q, ell = ref.gram_stats(X, y, σ2, mw)
    fmt = lambda v: v if isinstance(v, str) else f"{v:8.1e}"  # noqa: E731
    try:
        fx = ref.BayesianLinearRegressor(mw, Λ)(ref.ColVecs(X), σ2)
        lpo, po = ref.logpdf(fx, y), ref.posterior(fx, y)
        vo = ref.var(po(ref.ColVecs(Xt), noise_t))
        w = (hp.rel(lpo, tr["logpdf"]), hp.rel(po.mw, tr["m_post"]), hp.rel(vo, tr["var_t"]))
    except Exception as e:  # the reference itself fails (PosDefException) on some rank-deficient low-noise problems
        w = (type(e).__name__,) * 3
    try:
        lpd, md, v_inv, v_trsm = direct_form(mw, Λ, G, r, q, ell, N, Xt, noise_t)
        d = (hp.rel(lpd, tr["logpdf"]), hp.rel(md, tr["m_post"]), hp.rel(v_inv, tr["var_t"]), hp.rel(v_trsm, tr["var_t"]))
    except Exception as e:
        d = (type(e).__name__,) * 4
    print(f"{tag:34s} D={D:4d} N={N:5d} cond(Λ')={np.linalg.cond(Λ + G):7.1e} | logpdf whitened {fmt(w[0])} direct {fmt(d[0])} | "
          f"m' whitened {fmt(w[1])} direct {fmt(d[1])} | var whitened+trsm {fmt(w[2])} direct+inverse {fmt(d[2])} direct+trsm {fmt(d[3])}")


def main():
    print(__doc__)
    for D, N in ((64, 500), (256, 1500)):
        run("cond(Λw) = 1e6", D, N, (1.0, 1e6), None, 1)
        run("cond(Λw) = 1e10", D, N, (1.0, 1e10), None, 2)
        run("cond(Λw) = 1e13", D, N, (1.0, 1e13), None, 3)
        run("weak prior λ in [1e-6, 1]", D, N, (1e-6, 1.0), None, 4)
        run("weak prior λ in [1e-13, 1]", D, N, (1e-13, 1.0), None, 5)
        run("λ in [1e-5, 1e5]", D, N, (1e-5, 1e5), None, 6)
    run("N < D, weak ill prior", 64, 40, (1e-8, 1.0), 0.5, 7)
    run("N < D, prior 1..1e8", 64, 40, (1.0, 1e8), 0.5, 8)
    run("N < D, noise 1e-8", 64, 40, (1.0, 10.0), 1e-8, 9)
    run("N < D, noise eps() (rank deficient)", 64, 40, (1.0, 10.0), EPS, 10)
    run("noise eps(), N > D (interpolation)", 64, 100, (1.0, 10.0), EPS, 11)
    run("noise eps(), N > D (interpolation)", 256, 300, (1.0, 10.0), EPS, 12)
    run("N < D, λ in [1e-10, 1]", 256, 100, (1e-10, 1.0), 0.5, 13)
    run("N < D, noise 1e-6", 256, 100, (1.0, 10.0), 1e-6, 14)
    print("\nSeed ensemble: is the whitened form more accurate for logpdf once cond(Λw) is extreme?  No -- every evaluation order,\n"
          "the reference's included, draws its error from the same band (the problem's own conditioning: all of them start from\n"
          "chol(Λw), whose backward error already perturbs the small eigen-directions by cond * eps).")
    ensemble(64, 500, (1.0, 1e13), range(5))
    ensemble(64, 500, (1e-13, 1.0), range(5))
    ensemble(256, 1500, (1.0, 1e13), range(3))


if __name__ == "__main__":
    main()
