"""Fixtures shared by the oracle suite and the GPU parity suite.

Restates test/test_utils.jl:4-30 of the reference (generate_toy_problem and the toy ϕ) with
numpy's Generator in place of MersenneTwister (Julia's stream is not reproducible here).
"""
import numpy as np


def generate_toy_problem(rng, N, D, Tx, ns):
    """test/test_utils.jl:4-20: unstructured mean / precision / DENSE noise. ``ns`` is the
    namespace providing ColVecs / RowVecs / BayesianLinearRegressor (oracle or product)."""
    X = rng.standard_normal((D, N))
    B = rng.standard_normal((D, D))
    C = 0.1 * rng.standard_normal((N, N))
    mw, Λw, Σy = rng.standard_normal(D), B @ B.T + np.eye(D), C @ C.T + np.eye(N)
    f = ns.BayesianLinearRegressor(mw, Λw)
    if Tx == "Matrix":
        return X, f, Σy
    if Tx == "ColVecs":
        return ns.ColVecs(X), f, Σy
    return ns.RowVecs(np.ascontiguousarray(X.T)), f, Σy


def make_phi(ns):
    """test/test_utils.jl:28-30: ϕ(x) = [1, prod(x)] for RowVecs / ColVecs / Matrix."""

    def ϕ(x):
        if isinstance(x, ns.RowVecs):
            return ns.RowVecs(np.column_stack([np.ones(len(x)), np.prod(x.X, axis=1)]))
        if isinstance(x, ns.ColVecs):
            return ns.ColVecs(np.vstack([np.ones(len(x)), np.prod(x.X, axis=0)]))
        return ϕ(ns.ColVecs(np.asarray(x))).X

    return ϕ


def as_matrix(X, ns):
    """D x N dense matrix of whatever input kind."""
    if isinstance(X, ns.ColVecs):
        return X.X
    if isinstance(X, ns.RowVecs):
        return X.X.T
    return X


def take(X, idx, ns):
    """X[idx] for vectors-of-inputs, X[:, idx] for a bare matrix (test/bayesian_linear_regression.jl:61-62)."""
    if isinstance(X, (ns.ColVecs, ns.RowVecs)):
        return X[idx]
    return X[:, idx]
