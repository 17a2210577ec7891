"""CPU oracle for the FiniteBLR inference path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, operation by operation and in the reference's own order, what
BayesianLinearRegressors.jl v0.3.9 computes on the path BASELINE.json names
(posterior / logpdf / mean / var / cov / rand / BasisFunctionRegressor forwarding /
weight-space function samples).  It runs on host cores in Float64 through scipy's
LAPACK/BLAS (OpenBLAS: dpotrf / dtrsm / dsyrk / dgemv / dgemm / dpotrs -- the same BLAS
family Julia's LinearAlgebra stdlib dispatches to).

Who may use it: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and there only as the checker / the reported
CPU baseline.  The product package never imports it and has no CPU fallback.

Where the arithmetic really lives: the reference is 234 lines of Julia whose every flop
is executed by UN-VENDORED dependencies that are not under /root/reference --
Julia's LinearAlgebra stdlib (julia compat "1.10", Project.toml:18), AbstractGPs "0.5"
(Project.toml:14: FiniteGP, _cholesky, marginals), PDMats "0.11" (Project.toml:16),
KernelFunctions (ColVecs / RowVecs).  There is no root Manifest, so exact versions are
unpinned.  Their published behaviour used here: ``cholesky(A)`` = LAPACK dpotrf upper
factor; ``cholesky(Diagonal)`` = element-wise sqrt; ``logdet(::Cholesky)`` = 2 sum log
diag; ``A \\ B`` on triangular = dtrsm/dtrsv; ``f(X, s::Real)`` = Diagonal(Fill(s, N));
``f(X, v::Vector)`` = Diagonal(v); ``f(X)`` = noise 1e-18; ``f(X::Matrix)`` = ColVecs.

Pinning status (see tests/test_oracle_reference_suite.py): Julia cannot run in this image
(no binary, no depot, no network), so the oracle is pinned against
  * the only literal golden vector in the reference, the doctest
    ``var(bfr(x)) == [2.0, 1.25, 1.0, 1.25, 2.0]`` (src/basis_function_regression.jl:11-28);
  * every property / known-answer-by-construction test the reference's own suite holds for
    this path (test/bayesian_linear_regression.jl:22-122, test/basis_function_regression.jl:13-41,
    test/sampling_functions.jl:3-47), restated with the same shapes and tolerances.
  * round 2: an independent 50-digit evaluation (mpmath) of the mathematical definitions -- the naive N x N
    Gaussian of test/bayesian_linear_regression.jl:22-38 and the closed-form posterior -- which this oracle
    matches to 1e-12 on well-conditioned problems, dense and diagonal noise (tests/highprec.py,
    tests/test_highprec_pins.py; the reference's own suite pins the same identity only to sqrt(eps)).
What stays unpinned is what cannot be had without a Julia binary: seed-specific MersenneTwister streams and the
last-bit rounding of Julia+OpenBLAS itself ("parity unpinned" at the bit level; pinned at 1e-12 against the truth).

Conventions: X is always the D x N matrix of ``ColVecs`` (observation = column), exactly as
``x_as_colvecs`` produces (src/bayesian_linear_regression.jl:20-31).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.linalg as sl

LOG2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# Input wrappers (KernelFunctions.ColVecs / RowVecs) and matrix kinds (LinearAlgebra / PDMats)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ColVecs:
    """D x N matrix, each column one input (README.md:24)."""

    X: np.ndarray

    def __len__(self) -> int:
        return self.X.shape[1]

    def __getitem__(self, idx) -> "ColVecs":
        return ColVecs(self.X[:, idx])


@dataclass(frozen=True)
class RowVecs:
    """N x D matrix, each row one input (README.md:25)."""

    X: np.ndarray

    def __len__(self) -> int:
        return self.X.shape[0]

    def __getitem__(self, idx) -> "RowVecs":
        return RowVecs(self.X[idx, :])


@dataclass(frozen=True)
class Diagonal:
    diag: np.ndarray

    def dense(self) -> np.ndarray:
        return np.diag(self.diag)


@dataclass(frozen=True)
class Symmetric:
    data: np.ndarray

    def dense(self) -> np.ndarray:
        return self.data


@dataclass(frozen=True)
class PDMat:
    """PDMats.PDMat: a positive definite matrix carrying its upper Cholesky factor."""

    mat: np.ndarray
    U: np.ndarray

    @staticmethod
    def from_upper(U: np.ndarray) -> "PDMat":
        U = np.triu(np.asarray(U, dtype=np.float64))
        return PDMat(U.T @ U, U)

    @staticmethod
    def from_matrix(A: np.ndarray) -> "PDMat":
        A = np.asarray(A, dtype=np.float64)
        return PDMat(A, chol_upper(A))

    def dense(self) -> np.ndarray:
        return self.mat


MatrixLike = Union[np.ndarray, Diagonal, Symmetric, PDMat]


class PosDefException(np.linalg.LinAlgError):
    """LinearAlgebra.PosDefException(info): cholesky hit a non-positive pivot."""

    def __init__(self, info: int):
        super().__init__(f"matrix is not positive definite; Cholesky factorization failed (info={info})")
        self.info = info


def dense(A: MatrixLike) -> np.ndarray:
    return A if isinstance(A, np.ndarray) else A.dense()


def chol_upper(A: np.ndarray) -> np.ndarray:
    """LAPACK dpotrf('U'); what ``cholesky(A).U`` returns for a dense / Symmetric A."""
    A = np.asarray(A, dtype=np.float64)
    potrf = sl.get_lapack_funcs("potrf", (A,))
    c, info = potrf(A, lower=False, clean=True)
    if info > 0:
        raise PosDefException(int(info))
    if info < 0:
        raise ValueError(f"illegal value in argument {-info} of dpotrf")
    return c


def _cholesky_U(A: MatrixLike) -> MatrixLike:
    """``AbstractGPs._cholesky(A).U`` for every matrix kind the reference is used with.

    Dense / Symmetric -> dpotrf; Diagonal -> element-wise sqrt (stays Diagonal);
    PDMat -> the stored factor (no work).
    """
    if isinstance(A, Diagonal):
        if np.any(A.diag <= 0.0):
            raise PosDefException(int(np.argmax(A.diag <= 0.0)) + 1)
        return Diagonal(np.sqrt(A.diag))
    if isinstance(A, PDMat):
        return A.U
    return chol_upper(dense(A))


def _ut_ldiv(U: MatrixLike, B: np.ndarray, trans: bool) -> np.ndarray:
    """``U \\ B`` (trans=False) or ``U' \\ B`` (trans=True) for upper-triangular or Diagonal U."""
    if isinstance(U, Diagonal):
        return B / (U.diag[:, None] if B.ndim == 2 else U.diag)
    return sl.solve_triangular(U, B, lower=False, trans="T" if trans else "N", check_finite=False)


def _logdet_from_U(U: MatrixLike) -> float:
    d = U.diag if isinstance(U, Diagonal) else np.diag(U)
    return 2.0 * float(np.sum(np.log(d)))


# --------------------------------------------------------------------------------------
# Model types
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class BayesianLinearRegressor:
    """w ~ N(mw, inv(Lw)); f(x) = dot(x, w)   (src/bayesian_linear_regression.jl:11-14)."""

    mw: np.ndarray
    Λw: MatrixLike

    def __call__(self, x, Σy=1e-18) -> "FiniteGP":
        return FiniteGP(self, _wrap_inputs(x), Σy)


@dataclass(frozen=True)
class BasisFunctionRegressor:
    """blr(ϕ(x))   (src/basis_function_regression.jl:34-37)."""

    blr: BayesianLinearRegressor
    ϕ: Callable

    def __call__(self, x, Σy=1e-18) -> "FiniteGP":
        return FiniteGP(self, _wrap_inputs(x), Σy)


def _wrap_inputs(x):
    # AbstractGPs: a bare Matrix is read as ColVecs (D x N)  (README.md:26).
    if isinstance(x, np.ndarray) and x.ndim == 2:
        return ColVecs(x)
    return x


@dataclass(frozen=True)
class FiniteGP:
    """AbstractGPs.FiniteGP{f,x,Σy}: fields .f .x .Σy (used at bayesian_linear_regression.jl:20,33,37,42)."""

    f: Union[BayesianLinearRegressor, BasisFunctionRegressor]
    x: object
    Σy: object

    def noise(self, N: int) -> MatrixLike:
        """FiniteGP's noise normalisation: Real -> Diagonal(Fill), Vector -> Diagonal, Matrix as is."""
        S = self.Σy
        if isinstance(S, (Diagonal, Symmetric, PDMat)):
            return S
        S = np.asarray(S, dtype=np.float64)
        if S.ndim == 0:
            return Diagonal(np.full(N, float(S)))
        if S.ndim == 1:
            return Diagonal(S)
        return S


def x_as_colvecs(x) -> np.ndarray:
    """src/bayesian_linear_regression.jl:20-31 -- returns the D x N matrix (RowVecs: lazy transpose view)."""
    if isinstance(x, ColVecs):
        return x.X
    if isinstance(x, RowVecs):
        return x.X.T
    raise RuntimeError(
        f"{type(x).__name__} is not a subtype of AbstractVector that is known. "
        "Please provide either aColVecs or RowVecs."
    )


def _to_finite_blr(fx: FiniteGP) -> FiniteGP:
    """src/basis_function_regression.jl:41 -- ϕ is re-evaluated on every call."""
    if isinstance(fx.f, BasisFunctionRegressor):
        return fx.f.blr(fx.f.ϕ(fx.x), fx.Σy)
    return fx


def _diag_of(S: MatrixLike) -> np.ndarray:
    return S.diag if isinstance(S, Diagonal) else np.diag(dense(S)).copy()


# --------------------------------------------------------------------------------------
# Predictive side (src/bayesian_linear_regression.jl:33-53)
# --------------------------------------------------------------------------------------
def mean(fx: FiniteGP) -> np.ndarray:
    fx = _to_finite_blr(fx)
    return x_as_colvecs(fx.x).T @ fx.f.mw  # :33  dgemv('T')


def cov(fx: FiniteGP) -> np.ndarray:
    fx = _to_finite_blr(fx)
    X = x_as_colvecs(fx.x)
    α = _ut_ldiv(_cholesky_U(fx.f.Λw), X, trans=True)  # :36  Uw' \ X
    return α.T @ α + dense(fx.noise(X.shape[1]))  # :37


def var(fx: FiniteGP) -> np.ndarray:
    fx = _to_finite_blr(fx)
    X = x_as_colvecs(fx.x)
    α = _ut_ldiv(_cholesky_U(fx.f.Λw), X, trans=True)  # :41
    return np.sum(α * α, axis=0) + _diag_of(fx.noise(X.shape[1]))  # :42


def mean_and_cov(fx: FiniteGP):
    return mean(fx), cov(fx)  # :45


def mean_and_var(fx: FiniteGP):
    return mean(fx), var(fx)  # :47


def marginals(fx: FiniteGP) -> Tuple[np.ndarray, np.ndarray]:
    """AbstractGPs.marginals: Normal.(m, sqrt.(v)); returned as (mean, std)."""
    m, v = mean_and_var(fx)
    return m, np.sqrt(v)


def rand(fx: FiniteGP, Zw: np.ndarray, Zy: np.ndarray) -> np.ndarray:
    """src/bayesian_linear_regression.jl:49-53 with the standard-normal draws supplied.

    The reference draws ``Zw = randn(rng, D, S)`` FIRST (:51) and ``Zy = randn(rng, N, S)``
    SECOND (:52); callers wanting parity feed both in that order.  Returns the N x S matrix.
    """
    fx = _to_finite_blr(fx)
    X = x_as_colvecs(fx.x)
    Zw = np.asarray(Zw, dtype=np.float64).reshape(X.shape[0], -1)
    Zy = np.asarray(Zy, dtype=np.float64).reshape(X.shape[1], -1)
    w = fx.f.mw[:, None] + _ut_ldiv(_cholesky_U(fx.f.Λw), Zw, trans=False)  # :51
    Uy = _cholesky_U(fx.noise(X.shape[1]))
    noise = Uy.diag[:, None] * Zy if isinstance(Uy, Diagonal) else Uy.T @ Zy
    return X.T @ w + noise  # :52


# --------------------------------------------------------------------------------------
# Inference side (src/bayesian_linear_regression.jl:55-93)
# --------------------------------------------------------------------------------------
def _compute_inference_quantities(fx: FiniteGP, y: np.ndarray):
    """src/bayesian_linear_regression.jl:72-89, literal op order."""
    X = x_as_colvecs(fx.x)  # :73
    y = np.asarray(y, dtype=np.float64)
    if y.shape[0] != X.shape[1]:  # :74
        raise RuntimeError("length(y) != size(fx.x.X, 2)")
    blr = fx.f
    N = y.shape[0]

    Uw = _cholesky_U(blr.Λw)  # :78
    Uy = _cholesky_U(fx.noise(N))  # :79

    Bt = _ut_ldiv(Uy, _ut_ldiv(Uw, X, trans=True).T, trans=True)  # :81  N x D
    δy = _ut_ldiv(Uy, y - mean(fx), trans=True)  # :82

    logpdf_δy = -(N * LOG2PI + _logdet_from_U(Uy) + float(δy @ δy)) / 2  # :84

    BtB = sl.blas.dsyrk(1.0, Bt, trans=1, lower=0)  # :86  Bt'Bt (upper triangle)
    BtB = np.triu(BtB) + np.triu(BtB, 1).T
    Λεy_U = chol_upper(BtB + np.eye(X.shape[0]))  # :86

    return Uw, Bt, δy, logpdf_δy, Λεy_U  # :88


def logpdf(fx: FiniteGP, y: np.ndarray) -> float:
    """src/bayesian_linear_regression.jl:55-58."""
    fx = _to_finite_blr(fx)
    _, Bt, δy, logpdf_δy, Λεy_U = _compute_inference_quantities(fx, y)
    v = sl.solve_triangular(Λεy_U, Bt.T @ δy, lower=False, trans="T", check_finite=False)
    return -(_logdet_from_U(Λεy_U) - float(v @ v)) / 2 + logpdf_δy  # :57


def _dense_U(U: MatrixLike) -> np.ndarray:
    return np.diag(U.diag) if isinstance(U, Diagonal) else U


def _build_Λ(prior_Λw: MatrixLike, T: np.ndarray) -> MatrixLike:
    """src/bayesian_linear_regression.jl:92-93: PDMat prior -> PDMat(Cholesky(UpperTriangular(T)));
    anything else -> Symmetric(T'T)."""
    if isinstance(prior_Λw, PDMat):
        return PDMat.from_upper(T)
    return Symmetric(T.T @ T)


def posterior(fx: FiniteGP, y: np.ndarray):
    """src/bayesian_linear_regression.jl:60-69 (and basis_function_regression.jl:62-65)."""
    if isinstance(fx.f, BasisFunctionRegressor):
        return BasisFunctionRegressor(posterior(_to_finite_blr(fx), y), fx.f.ϕ)
    Uw, Bt, δy, _, Λεy_U = _compute_inference_quantities(fx, y)
    mεy = sl.cho_solve((Λεy_U, False), Bt.T @ δy, check_finite=False)  # :64  dpotrs
    T = Λεy_U @ _dense_U(Uw)  # :67  (upper x upper = upper)
    m_post = fx.f.mw + _ut_ldiv(Uw, mεy, trans=False)  # :68
    return BayesianLinearRegressor(m_post, _build_Λ(fx.f.Λw, T))


def posterior_factor_T(fx: FiniteGP, y: np.ndarray) -> np.ndarray:
    """The upper-triangular T of :67 (T'T = posterior precision); exposed for parity checks."""
    fx = _to_finite_blr(fx)
    Uw, _, _, _, Λεy_U = _compute_inference_quantities(fx, y)
    return Λεy_U @ _dense_U(Uw)


# --------------------------------------------------------------------------------------
# Weight-space function samples (src/sampling_functions.jl)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class BLRFunctionSample:
    """src/sampling_functions.jl:12-19."""

    w: np.ndarray
    ϕ: Callable

    def __call__(self, X):
        Z = self.ϕ(X)
        if isinstance(Z, ColVecs):
            return Z.X.T @ self.w  # :18
        if isinstance(Z, RowVecs):
            return Z.X @ self.w  # :19
        return np.asarray(Z).T @ self.w  # :17  Matrix (D x N)


def _identity(x):
    return x


def _blr_and_mapping(b):
    """src/sampling_functions.jl:51-52."""
    if isinstance(b, BasisFunctionRegressor):
        return b.blr, b.ϕ
    return b, _identity


def rand_weights(b, Z: np.ndarray) -> np.ndarray:
    """``blr.mw .+ _cholesky(blr.Λw).U \\ Z`` (src/sampling_functions.jl:29,35,44); Z is D or D x S."""
    blr, _ = _blr_and_mapping(b)
    Z = np.asarray(Z, dtype=np.float64)
    sol = _ut_ldiv(_cholesky_U(blr.Λw), Z, trans=False)
    return blr.mw + sol if Z.ndim == 1 else blr.mw[:, None] + sol


def rand_function(b, Z: np.ndarray) -> BLRFunctionSample:
    """src/sampling_functions.jl:27-31 with the draw supplied (Z of length D)."""
    _, ϕ = _blr_and_mapping(b)
    return BLRFunctionSample(rand_weights(b, np.asarray(Z).reshape(-1)), ϕ)


def rand_functions(b, Z: np.ndarray, dims: Sequence[int]):
    """src/sampling_functions.jl:33-38: Z is D x prod(dims); returns an object array of shape dims
    filled in column-major order (Julia's reshape)."""
    _, ϕ = _blr_and_mapping(b)
    ws = rand_weights(b, Z)
    flat = np.empty(ws.shape[1], dtype=object)
    for i in range(ws.shape[1]):
        flat[i] = BLRFunctionSample(ws[:, i].copy(), ϕ)
    return flat.reshape(tuple(dims), order="F")


# --------------------------------------------------------------------------------------
# Streaming Gram-form oracle: the same outputs without the two N x D temporaries.
# --------------------------------------------------------------------------------------
def gram_stats(X: np.ndarray, y: np.ndarray, σ2: np.ndarray, mw: np.ndarray, chunk: int = 1 << 15):
    """Sufficient statistics of the path in closed form:

        G = X S X',  r = X S δ,  q = δ' S δ,  ℓ = sum log σ²,   S = diag(1/σ²),  δ = y - X'mw.

    Accumulated chunk by chunk with dsyrk so N can be far larger than host RAM allows for the
    literal form.  Validated against the literal form in tests/test_oracle_streaming.py.
    """
    D, N = X.shape
    σ2 = np.broadcast_to(np.asarray(σ2, dtype=np.float64), (N,))
    G = np.zeros((D, D), order="F")
    r = np.zeros(D)
    q = 0.0
    ℓ = 0.0
    for a in range(0, N, chunk):
        b = min(N, a + chunk)
        Xc = X[:, a:b]
        isd = 1.0 / np.sqrt(σ2[a:b])
        Bc = np.asfortranarray(Xc * isd)  # D x n, column scaled by 1/σ
        δ = (y[a:b] - Xc.T @ mw) * isd
        G = sl.blas.dsyrk(1.0, Bc, beta=1.0, c=G, trans=0, lower=0, overwrite_c=1)
        r += Bc @ δ
        q += float(δ @ δ)
        ℓ += float(np.sum(np.log(σ2[a:b])))
    G = np.triu(G) + np.triu(G, 1).T
    return G, r, q, ℓ


def infer_from_stats(mw: np.ndarray, Λw: MatrixLike, G, r, q, ℓ, N: int):
    """posterior + logpdf from the reduced statistics, following :78-:86 on D x D objects:
    Bt'Bt + I = Uw^-T G Uw^-1 + I, Bt'δy = Uw^-T r.  Returns (logpdf, m_post, T)."""
    Uw = _cholesky_U(Λw)
    D = mw.shape[0]
    Wt = _ut_ldiv(Uw, _ut_ldiv(Uw, G, trans=True).T, trans=True)  # Uw^-T G Uw^-1
    Wt = (Wt + Wt.T) / 2
    Λεy_U = chol_upper(Wt + np.eye(D))
    b = _ut_ldiv(Uw, r, trans=True)
    v = sl.solve_triangular(Λεy_U, b, lower=False, trans="T", check_finite=False)
    lp = -(_logdet_from_U(Λεy_U) - float(v @ v)) / 2 - (N * LOG2PI + ℓ + q) / 2
    mεy = sl.cho_solve((Λεy_U, False), b, check_finite=False)
    T = Λεy_U @ _dense_U(Uw)
    return lp, mw + _ut_ldiv(Uw, mεy, trans=False), T


def infer_streaming(mw, Λw, X, y, σ2, chunk: int = 1 << 15):
    """posterior + logpdf for diagonal noise without N x D temporaries."""
    G, r, q, ℓ = gram_stats(X, y, σ2, mw, chunk)
    return infer_from_stats(mw, Λw, G, r, q, ℓ, X.shape[1])
