"""Independent high-precision truth for the FiniteBLR path -- TEST INFRASTRUCTURE.

Two evaluators of the *mathematical* quantities the reference computes (src/bayesian_linear_regression.jl:33-93),
neither of which shares code or rounding behaviour with the oracle (oracle/blr_oracle.py) or the CUDA library:

  * ``mp_*``   : mpmath at 50 significant digits (small problems, D <= 64, N <= 200): the naive N x N Gaussian of
                 test/bayesian_linear_regression.jl:22-38 for logpdf, and the closed-form posterior
                 Λ' = Λw + X Σy⁻¹ X', m' = Λ'⁻¹ (Λw mw + X Σy⁻¹ y).
  * ``ld_*``   : numpy ``longdouble`` (x87 80-bit extended, eps = 1.08e-19 -- 3.3 more digits than Float64) with
                 hand-written Cholesky / substitution, for the sizes where the TMA kernels run (D = 64 ... 1024).

They answer "how far is each implementation from the truth", which is the only meaningful comparison once
cond(Λ) * eps approaches the 1e-9 bar: two backward-stable Float64 algorithms then differ from each other by more
than 1e-9 and neither is wrong.
"""
from __future__ import annotations

import math

import numpy as np

LD = np.longdouble


# ----------------------------------------------------------------------------------------------- longdouble kernels
def ld_cholesky(A: np.ndarray) -> np.ndarray:
    """Lower Cholesky factor in extended precision (right-looking, one column per step, vectorised over rows)."""
    A = np.array(A, dtype=LD)
    n = A.shape[0]
    for k in range(n):
        d = A[k, k]
        if not d > 0:
            raise np.linalg.LinAlgError(f"not positive definite at {k + 1}")
        d = np.sqrt(d)
        A[k, k] = d
        if k + 1 < n:
            A[k + 1 :, k] /= d
            c = A[k + 1 :, k]
            # rank-1 update of the trailing lower triangle (full block: simpler, still O(n^3/3 * 2))
            A[k + 1 :, k + 1 :] -= np.outer(c, c)
    return np.tril(A)


def ld_solve_lower(L: np.ndarray, B: np.ndarray) -> np.ndarray:
    """L \\ B by forward substitution (extended precision)."""
    L = np.asarray(L, dtype=LD)
    X = np.array(B, dtype=LD, copy=True)
    n = L.shape[0]
    for i in range(n):
        if i:
            X[i] -= L[i, :i] @ X[:i]
        X[i] /= L[i, i]
    return X


def ld_solve_upper(U: np.ndarray, B: np.ndarray) -> np.ndarray:
    """U \\ B by backward substitution (extended precision)."""
    U = np.asarray(U, dtype=LD)
    X = np.array(B, dtype=LD, copy=True)
    n = U.shape[0]
    for i in range(n - 1, -1, -1):
        if i + 1 < n:
            X[i] -= U[i, i + 1 :] @ X[i + 1 :]
        X[i] /= U[i, i]
    return X


def ld_matmul(A, B, blk: int = 256):
    """A @ B in extended precision (numpy has no BLAS for longdouble; blocked to bound temporaries)."""
    A, B = np.asarray(A, dtype=LD), np.asarray(B, dtype=LD)
    return A @ B


def ld_stats(X, y, σ2, mw):
    """Sufficient statistics G = X S X', r = X S δ, q = δ'Sδ, ℓ = Σ log σ² in extended precision (the expensive part:
    compute once per data set and reuse across priors)."""
    X = np.asarray(X, dtype=LD)
    y = np.asarray(y, dtype=LD)
    D, N = X.shape
    v = np.broadcast_to(np.asarray(σ2, dtype=LD), (N,))
    s = 1 / v
    mw = np.asarray(mw, dtype=LD)
    Xs = X * s
    G = Xs @ X.T
    G = (G + G.T) / 2
    δ = y - X.T @ mw
    return {"G": G, "r": Xs @ δ, "q": (s * δ) @ δ, "ell": np.sum(np.log(v)), "N": N}


def ld_truth(st, mw, Λw, Xt=None, σ2t=0.0, Zw=None, Zy=None):
    """Diagonal-noise BLR in extended precision from ld_stats(): returns a dict with logpdf, m_post, Λ_post, T (upper
    factor of Λ_post), and -- if Xt is given -- predictive mean / var of the posterior at Xt (noise σ2t) and rand with the
    supplied draws.  Direct (unwhitened) form: in extended precision the form does not matter at the 1e-9 level unless
    cond(Λ') exceeds ~1e9, and the fixtures stay below that."""
    G, r, q, ell, N = st["G"], st["r"], st["q"], st["ell"], st["N"]
    mw = np.asarray(mw, dtype=LD)
    Λw = np.asarray(Λw, dtype=LD)
    if Λw.ndim == 1:
        Λw = np.diag(Λw)
    Λp = Λw + G
    Lp = ld_cholesky(Λp)
    Lw = ld_cholesky(Λw)
    z = ld_solve_lower(Lp, r)
    u = ld_solve_upper(Lp.T, z)
    logdet_p = 2 * np.sum(np.log(np.diag(Lp)))
    logdet_w = 2 * np.sum(np.log(np.diag(Lw)))
    LOG2PI = np.log(LD(2)) + np.log(np.arccos(LD(-1)))
    lp = -(N * LOG2PI + ell + q + (logdet_p - logdet_w) - z @ z) / 2
    out = {"logpdf": lp, "m_post": mw + u, "Lambda_post": Λp, "T": Lp.T}
    if Xt is not None:
        Xt = np.asarray(Xt, dtype=LD)
        α = ld_solve_lower(Lp, Xt)
        out["mean_t"] = Xt.T @ out["m_post"]
        out["var_t"] = np.sum(α * α, axis=0) + np.broadcast_to(np.asarray(σ2t, dtype=LD), (Xt.shape[1],))
        if Zw is not None:
            W = out["m_post"][:, None] + ld_solve_upper(Lp.T, np.asarray(Zw, dtype=LD))
            sd = np.sqrt(np.broadcast_to(np.asarray(σ2t, dtype=LD), (Xt.shape[1],)))
            out["rand_t"] = Xt.T @ W + sd[:, None] * np.asarray(Zy, dtype=LD)
    return out


def rel(a, b) -> float:
    """norm-wise relative error of a (any precision) against the extended-precision b."""
    a, b = np.asarray(a, dtype=LD), np.asarray(b, dtype=LD)
    if a.ndim == 0:
        return float(abs(a - b) / max(abs(b), LD(1e-300)))
    return float(np.sqrt(np.sum((a - b) ** 2)) / max(np.sqrt(np.sum(b**2)), LD(1e-300)))


# ----------------------------------------------------------------------------------------------- mpmath (50 digits)
def mp_truth(X, y, Σy, mw, Λw, dps: int = 50):
    """The naive Gaussian of test/bayesian_linear_regression.jl:22-38 at `dps` digits:
        m = X'mw,  Σ = X' Λw⁻¹ X + Σy,  logpdf = -(N log 2π + logdet Σ + δ'Σ⁻¹δ)/2,
    and the closed-form posterior Λ' = Λw + X Σy⁻¹ X', m' = Λ'⁻¹ (Λw mw + X Σy⁻¹ y).  Σy: scalar, vector (diagonal) or
    dense N x N.  Returns Float64 roundings of the 50-digit results."""
    import mpmath as mp

    mp.mp.dps = dps
    X = np.asarray(X, dtype=np.float64)
    D, N = X.shape
    Xm = mp.matrix(X.tolist())
    ym = mp.matrix([mp.mpf(float(v)) for v in y])
    mwm = mp.matrix([mp.mpf(float(v)) for v in mw])
    Lw = np.asarray(Λw, dtype=np.float64)
    Λm = mp.matrix(Lw.tolist()) if Lw.ndim == 2 else mp.diag([mp.mpf(float(v)) for v in Lw])
    S = np.asarray(Σy, dtype=np.float64)
    if S.ndim == 0:
        Sm = mp.diag([mp.mpf(float(S))] * N)
    elif S.ndim == 1:
        Sm = mp.diag([mp.mpf(float(v)) for v in S])
    else:
        Sm = mp.matrix(S.tolist())
    # naive marginal
    m = Xm.T * mwm
    Σ = Xm.T * (mp.inverse(Λm) * Xm) + Sm   # 50 digits: an explicit inverse costs ~log10 cond digits, harmless
    Σ = (Σ + Σ.T) / 2
    δ = ym - m
    Lc = mp.cholesky(Σ)
    logdet = 2 * sum(mp.log(Lc[i, i]) for i in range(N))
    quad = (δ.T * mp.lu_solve(Σ, δ))[0]
    lp = -(N * mp.log(2 * mp.pi) + logdet + quad) / 2
    # posterior
    Si = mp.inverse(Sm)
    SiXt = Si * Xm.T                      # Σy⁻¹ X'   (N x D)
    Λp = Λm + Xm * SiXt
    Λp = (Λp + Λp.T) / 2
    rhs = Λm * mwm + Xm * (Si * ym)
    mp_post = mp.lu_solve(Λp, rhs)
    return {
        "logpdf": float(lp),
        "m_post": np.array([float(mp_post[i]) for i in range(D)]),
        "Lambda_post": np.array([[float(Λp[i, j]) for j in range(D)] for i in range(D)]),
    }


def mp_predict(Xt, m_post, Λ_post, σ2t, dps: int = 50):
    """mean / var at Xt for w ~ N(m_post, Λ_post⁻¹) at `dps` digits (src/bayesian_linear_regression.jl:33,40-43)."""
    import mpmath as mp

    mp.mp.dps = dps
    Xt = np.asarray(Xt, dtype=np.float64)
    D, Nt = Xt.shape
    Xm = mp.matrix(Xt.tolist())
    Λm = mp.matrix(np.asarray(Λ_post, dtype=np.float64).tolist())
    mm = mp.matrix([mp.mpf(float(v)) for v in m_post])
    mean = Xm.T * mm
    Z = mp.inverse(Λm) * Xm
    s2 = np.broadcast_to(np.asarray(σ2t, dtype=np.float64), (Nt,))
    var = [sum(Xm[d, n] * Z[d, n] for d in range(D)) + mp.mpf(float(s2[n])) for n in range(Nt)]
    return np.array([float(mean[i]) for i in range(Nt)]), np.array([float(v) for v in var])
