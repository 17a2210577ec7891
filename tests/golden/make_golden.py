"""Generates tests/golden/blr_golden.npz: seeded inputs and the CPU oracle's outputs for them.

Run from the repo root:  python tests/golden/make_golden.py
The oracle (oracle/blr_oracle.py) is the literal restatement of the reference, itself pinned against the
reference's doctest vector and property suite (tests/test_oracle_reference_suite.py).  The fixture lets the GPU
parity tests (and later rounds) detect drift in either side without recomputing.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import blr_oracle as ref  # noqa: E402

CASES = [  # name, D, N, dense prior?, zero prior mean?, scalar noise?
    ("toy_d2", 2, 10, False, True, False),
    ("d7_n13", 7, 13, True, False, False),
    ("d33_n257", 33, 257, True, False, True),
    ("d64_n500", 64, 500, True, False, False),
    ("d130_n515", 130, 515, False, True, False),
    ("d128_n777", 128, 777, True, False, False),
]


def make_case(name, D, N, dense, zero_mean, scalar_noise, seed):
    rng = np.random.default_rng(seed)
    if name == "toy_d2":  # README.md:44-50
        X = np.vstack([np.linspace(-5.0, 5.0, N), np.ones(N)])
    else:
        X = rng.standard_normal((D, N))
    mw = np.zeros(D) if zero_mean else rng.standard_normal(D)
    if dense:
        B = rng.standard_normal((D, D))
        Λ = B @ B.T + np.eye(D)
        Λw = Λ
    else:
        Λ = np.diag(np.ones(D))
        Λw = ref.Diagonal(np.ones(D))
    σ2 = np.float64(0.37) if scalar_noise else np.exp(rng.standard_normal(N))
    f = ref.BayesianLinearRegressor(mw, Λw)
    fx = f(ref.ColVecs(X), σ2)
    Zw, Zy = rng.standard_normal((D, 3)), rng.standard_normal((N, 3))
    Y = ref.rand(fx, Zw, Zy)
    y = Y[:, 0]
    post = ref.posterior(fx, y)
    Nt = 17
    Xt = rng.standard_normal((D, Nt))
    if name == "toy_d2":
        Xt = np.vstack([np.linspace(-6.0, 6.0, Nt), np.ones(Nt)])
    fpt = post(ref.ColVecs(Xt), np.finfo(np.float64).eps)
    m_t, v_t = ref.mean_and_var(fpt)
    out = {
        "X": X, "mw": mw, "Lambda": Λ, "dense": np.bool_(dense), "sigma2": np.asarray(σ2), "Zw": Zw, "Zy": Zy,
        "rand": Y, "y": y, "logpdf": np.float64(ref.logpdf(fx, y)), "m_post": post.mw, "Lambda_post": ref.dense(post.Λw),
        "T_post": ref.posterior_factor_T(fx, y), "Xt": Xt, "mean_t": m_t, "var_t": v_t, "cov_t": ref.cov(fpt),
        "prior_mean": ref.mean(fx), "prior_var": ref.var(fx),
    }
    return {f"{name}/{k}": v for k, v in out.items()}


def main():
    blob = {}
    for i, c in enumerate(CASES):
        blob.update(make_case(*c, seed=1000 + i))
    blob["doctest_var"] = np.array([2.0, 1.25, 1.0, 1.25, 2.0])  # src/basis_function_regression.jl:18-28
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "blr_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
