// The reference's LITERAL numerical form, as an opt-in of the context (BLR_FORM_WHITENED, blr_ctx_set_form / BLR_FORM=whitened).
//
// The default device path factorises Λ' = Λw + G directly and applies an explicit inverse factor in `var` / `rand`
// (chol.cu, predict*.cu).  The reference (src/bayesian_linear_regression.jl) instead
//   * whitens with the prior factor before it factorises:  Bt = Σy.U' \ (Uw' \ X)'  (:81),  Λεy = chol(Bt'Bt + I)  (:86),
//     mεy = Λεy \ (Bt'δy)  (:64),  T = Λεy.U * Uw  (:67),  m' = mw + Uw \ mεy  (:68),
//     logpdf = -(logdet(Λεy) - |Λεy.U' \ (Bt'δy)|²) / 2 + logpdf_δy  (:57);
//   * uses triangular SOLVES against the factor in `var` / `cov` (`Uw' \ X`, :36, :41) and `rand` (`Uw \ randn`, :51).
// On the reduced statistics (G = X Σy⁻¹ X', r = X Σy⁻¹ δ) the whitened quantities are
//     Bt'Bt = Lw⁻¹ G Lw⁻ᵀ,   Bt'δy = Lw⁻¹ r        (Lw = Uw' lower, Λw = Lw Lw'),
// so this file runs, with substitution-based triangular solves and no inverse anywhere:
//     M = Lw⁻¹ G Lw⁻ᵀ + I,  Lε = chol(M),  z = Lε⁻¹ Lw⁻¹ r,  u = Lε⁻ᵀ z,  m' = mw + Lw⁻ᵀ u,  L' = Lw Lε  (T = L'ᵀ),
//     logpdf = -1/2 [ n log 2π + ℓ + q + logdet M − z'z ].
// It costs about 8x the D x D flops of the direct form (two D-column triangular solves and a triangular product on plain DFMA) and
// was measured to buy no accuracy (profiles/r02/illcond_study.txt) -- it exists so that a user who wants the
// reference's evaluation order, e.g. to reproduce its rounding behaviour on an ill-conditioned prior, can have it on the device.
#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace blr {

// ---------------------------------------------------------------------------------------------------------------------
// B (D x K, column-major, ldb) <- L⁻¹ B (TRANS = false) or L⁻ᵀ B (TRANS = true); L lower triangular D x D (ldl), only its lower
// triangle is read.  One CTA per slab of 32 right-hand sides (columns are independent).  Block rows of 64 are visited in
// dependency order: the contribution of the block rows already solved (read back from B -- this CTA is their only writer) is
// accumulated by a register-tiled DFMA GEMM, then the 64 x 64 diagonal block is applied by SUBSTITUTION: 8 threads per column,
// each holding 8 of its 64 entries in registers, the pivot entry broadcast by a shuffle.
constexpr int TS_NB = 64, TS_KC = 32, TS_COLS = 32, TS_THREADS = 256;

template <bool TRANS>
__global__ void __launch_bounds__(TS_THREADS) trsm_lower_kernel(const double* __restrict__ L, int64_t ldl, int D, double* B,
                                                                int64_t ldb, int64_t K) {
    __shared__ __align__(16) double sm[TS_NB * TS_NB + TS_NB * TS_COLS];
    double* As = sm;                      // accumulation: 64 x 32 coefficients, As[k * 64 + m]
    double* Bs = sm + TS_NB * TS_KC;      //               32 x 32 solved entries, Bs[k * 33 + c]  (33 * 32 <= 2048)
    double* Ld = sm;                      // substitution: diagonal block, Ld[c * 64 + r]  (overlays As / Bs)
    double* T = sm + TS_NB * TS_NB;       //               right-hand sides of the block row, T[c * 64 + r]
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int64_t c0 = (int64_t)blockIdx.x * TS_COLS;
    const int nblk = (D + TS_NB - 1) / TS_NB;
    for (int step = 0; step < nblk; ++step) {
        const int bi = TRANS ? nblk - 1 - step : step;
        const int r0 = bi * TS_NB;
        const int kbeg = TRANS ? r0 + TS_NB : 0, kend = TRANS ? D : r0;
        double acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.0;
        for (int k0 = kbeg; k0 < kend; k0 += TS_KC) {
            __syncthreads();
            for (int e = t; e < TS_NB * TS_KC; e += TS_THREADS) {
                if (!TRANS) {  // A(m, k) = L[r0 + m, k0 + k]: m is the unit-stride direction
                    const int m = e % TS_NB, k = e / TS_NB;
                    As[k * TS_NB + m] = (r0 + m < D) ? L[(int64_t)(k0 + k) * ldl + r0 + m] : 0.0;
                } else {  // A(m, k) = L[k0 + k, r0 + m]: k is the unit-stride direction
                    const int k = e % TS_KC, m = e / TS_KC;
                    As[k * TS_NB + m] = (k0 + k < D && r0 + m < D) ? L[(int64_t)(r0 + m) * ldl + k0 + k] : 0.0;
                }
            }
            for (int e = t; e < TS_KC * TS_COLS; e += TS_THREADS) {
                const int k = e % TS_KC, c = e / TS_KC;
                Bs[k * 33 + c] = (k0 + k < D && c0 + c < K) ? B[(c0 + c) * ldb + k0 + k] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < TS_KC; ++k) {
                const double2 a01 = *reinterpret_cast<const double2*>(As + k * TS_NB + tx * 4);
                const double2 a23 = *reinterpret_cast<const double2*>(As + k * TS_NB + tx * 4 + 2);
                const double b0 = Bs[k * 33 + ty * 2], b1 = Bs[k * 33 + ty * 2 + 1];
                acc[0][0] = fma(a01.x, b0, acc[0][0]);
                acc[0][1] = fma(a01.x, b1, acc[0][1]);
                acc[1][0] = fma(a01.y, b0, acc[1][0]);
                acc[1][1] = fma(a01.y, b1, acc[1][1]);
                acc[2][0] = fma(a23.x, b0, acc[2][0]);
                acc[2][1] = fma(a23.x, b1, acc[2][1]);
                acc[3][0] = fma(a23.y, b0, acc[3][0]);
                acc[3][1] = fma(a23.y, b1, acc[3][1]);
            }
        }
        __syncthreads();  // As / Bs are dead: Ld overlays them
        for (int e = t; e < TS_NB * TS_NB; e += TS_THREADS) {
            const int r = e % TS_NB, c = e / TS_NB;
            Ld[e] = (r0 + r < D && r0 + c < D && r >= c) ? L[(int64_t)(r0 + c) * ldl + r0 + r] : (r == c ? 1.0 : 0.0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = tx * 4 + i, c = ty * 2 + j;
                const double b = (r0 + r < D && c0 + c < K) ? B[(c0 + c) * ldb + r0 + r] : 0.0;
                T[c * TS_NB + r] = b - acc[i][j];
            }
        __syncthreads();
        // substitution on the diagonal block: column c of the slab, rows s, s + 8, ..., s + 56 in registers
        const int c = t >> 3, s = t & 7, lane = t & 31;
        double x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = T[c * TS_NB + s + 8 * q];
#pragma unroll
        for (int jj = 0; jj < TS_NB; ++jj) {
            const int j = TRANS ? TS_NB - 1 - jj : jj;
            double xj = x[j >> 3] / Ld[j * TS_NB + j];  // meaningful on the owner (s == j % 8) only
            xj = __shfl_sync(0xffffffffu, xj, (lane & ~7) | (j & 7));
            if (s == (j & 7)) x[j >> 3] = xj;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = s + 8 * q;
                // forward: rows below the pivot, coefficient L[i, j];  transposed: rows above it, coefficient L'[i, j] = L[j, i]
                const bool live = TRANS ? (8 * q < j) : (8 * q + 7 > j);  // static after unrolling: skips dead register rows
                if (live) {
                    const double l = TRANS ? Ld[i * TS_NB + j] : Ld[j * TS_NB + i];
                    if (TRANS ? (i < j) : (i > j)) x[q] = fma(-l, xj, x[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = s + 8 * q;
            if (r0 + r < D && c0 + c < K) B[(c0 + c) * ldb + r0 + r] = x[q];
        }
        // the next block row reads these entries back from global memory: the barrier at the top of its accumulation loop (or
        // of its diagonal phase, when it has no accumulation) orders the accesses within the CTA
    }
}

int trsm_lower(blr_ctx* ctx, const double* L, int64_t ldl, int64_t D, double* B, int64_t ldb, int64_t K, bool trans) {
    if (D == 0 || K == 0) return 0;
    const int64_t grid = (K + TS_COLS - 1) / TS_COLS;
    if (grid > 0x7fffffff) return set_err(ctx, BLR_E_INVALID, "trsm: too many right-hand sides for one launch");
    if (trans)
        trsm_lower_kernel<true><<<(int)grid, TS_THREADS, 0, ctx->stream>>>(L, ldl, (int)D, B, ldb, K);
    else
        trsm_lower_kernel<false><<<(int)grid, TS_THREADS, 0, ctx->stream>>>(L, ldl, (int)D, B, ldb, K);
    BLR_CHECK_LAUNCH(ctx, "trsm_lower_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
__global__ void diag_sqrt_matrix_kernel(double* __restrict__ A, const double* __restrict__ d, int D) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) A[(int64_t)i * D + i] = sqrt(d[i]);
}
__global__ void add_identity_kernel(double* __restrict__ A, int D) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) A[(int64_t)i * D + i] += 1.0;
}
__global__ void vec_sum_kernel(const double* __restrict__ a, const double* __restrict__ b, int D, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) out[i] = a[i] + b[i];
}
__global__ void transpose_ld_kernel(const double* __restrict__ A, double* __restrict__ At, int D) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = bx + threadIdx.x, c = by + j;
        tile[j][threadIdx.x] = (r < D && c < D) ? A[(int64_t)c * D + r] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + threadIdx.x, c = bx + j;
        if (r < D && c < D) At[(int64_t)c * D + r] = tile[threadIdx.x][j];
    }
}

// The whitened D x D phase.  On entry: p->Lam = Λw + G (kept: the posterior precision), rhs = r, Lw = lower factor of the prior
// precision (strict upper part zero), mwd = prior mean, sc / info as in infer_solve.  On exit: p->L = L' = Lw Lε, p->mw = m',
// sc[3] = logpdf, info_post[0] = LAPACK-style info of chol(M).
int dxd_whitened(blr_ctx* ctx, blr_post* p, const double* Lw, const blr_stats* st, double* rhs, double* usol, const double* mwd,
                 double* sc, int* info_post) {
    const int64_t D = p->D, n2 = D * D;
    cudaStream_t sm = ctx->stream;
    if (!p->W) BLR_CUDA_OK(ctx, dev_alloc(ctx, &p->W, (size_t)n2 * sizeof(double)));
    p->has_W = false;
    const dim3 tgrid((unsigned)((D + 31) / 32), (unsigned)((D + 31) / 32)), tblock(32, 8);
    const int vgrid = (int)std::min<int64_t>((D + 255) / 256, 64);
    // M - I = Lw⁻¹ (Lw⁻¹ G)'   (G is symmetric; the second solve acts on the transpose of the first)
    BLR_CUDA_OK(ctx, cudaMemcpyAsync(p->W, st->G(), (size_t)n2 * sizeof(double), cudaMemcpyDeviceToDevice, sm));
    BLR_TRY(trsm_lower(ctx, Lw, D, D, p->W, D, D, false));
    transpose_ld_kernel<<<tgrid, tblock, 0, sm>>>(p->W, p->L, (int)D);
    BLR_CHECK_LAUNCH(ctx, "transpose_ld_kernel");
    BLR_TRY(trsm_lower(ctx, Lw, D, D, p->L, D, D, false));
    add_identity_kernel<<<vgrid, 256, 0, sm>>>(p->L, (int)D);
    BLR_CHECK_LAUNCH(ctx, "add_identity_kernel");
    // Bt'δy = Lw⁻¹ r
    BLR_TRY(trsm_lower(ctx, Lw, D, D, rhs, D, 1, false));
    // Lε = chol(M), z = Lε⁻¹ (Bt'δy), u = mεy = Lε⁻ᵀ z, logdet M, z'z, logpdf (logdet Λw does not enter: sc[4] = 0)
    BLR_CUDA_OK(ctx, cudaMemsetAsync(sc + 4, 0, sizeof(double), sm));
    DxdFinalize fin;
    fin.stat_scal = st->scal();
    fin.mw = mwd;
    fin.m_post = p->mw;  // receives mw + u here; overwritten with mw + Lw⁻ᵀ u below
    fin.logdet_w = sc + 4;
    fin.sc = sc;
    BLR_TRY(dxd_fused(ctx, p->L, D, info_post, rhs, usol, &fin));
    // m' = mw + Uw \ mεy
    BLR_TRY(trsm_lower(ctx, Lw, D, D, usol, D, 1, true));
    vec_sum_kernel<<<vgrid, 256, 0, sm>>>(mwd, usol, (int)D, p->mw);
    BLR_CHECK_LAUNCH(ctx, "vec_sum_kernel");
    // T = Λεy.U * Uw, stored lower: L' = Lw Lε (both factors have exact zeros above the diagonal, so has the product)
    BLR_TRY(gemm_generic(ctx, D, D, D, Lw, 1, D, p->L, 1, D, p->W, 1, D, 0.0));
    std::swap(p->L, p->W);
    return 0;
}

// Lower factor of a Diagonal prior precision as a dense matrix (the whitened phase treats both prior kinds alike).
int diag_factor_dense(blr_ctx* ctx, const double* diag_dev, int64_t D, double* Lw) {
    BLR_CUDA_OK(ctx, cudaMemsetAsync(Lw, 0, (size_t)D * D * sizeof(double), ctx->stream));
    diag_sqrt_matrix_kernel<<<(int)std::min<int64_t>((D + 255) / 256, 64), 256, 0, ctx->stream>>>(Lw, diag_dev, (int)D);
    BLR_CHECK_LAUNCH(ctx, "diag_sqrt_matrix_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// var (:40-43) in the reference's form: α = Uw' \ X by triangular solve, var_n = Σ_d α[d, n]² + σ²_n.
__global__ void __launch_bounds__(256) colnorm_noise_kernel(const double* __restrict__ A, int64_t D, int64_t ld, int64_t N,
                                                            const double* __restrict__ sigma2, double sigma2_scalar,
                                                            double* __restrict__ var) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); n < N; n += warps) {
        const double* col = A + n * ld;
        double acc = 0.0;
        for (int64_t d = lane; d < D; d += 32) acc = fma(col[d], col[d], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) var[n] = acc + (sigma2 ? sigma2[n] : sigma2_scalar);
    }
}

// α (D x n, column-major, ld = D) = L⁻¹ X[:, n0 : n0 + n] for either input layout
int solve_alpha(blr_ctx* ctx, const blr_post* p, const blr_x* x, int64_t n0, int64_t n, double* alpha) {
    const int64_t D = p->D;
    if (x->layout == BLR_COLVECS) {
        BLR_CUDA_OK(ctx, cudaMemcpy2DAsync(alpha, (size_t)D * sizeof(double), x->p + n0 * x->ld, (size_t)x->ld * sizeof(double),
                                           (size_t)D * sizeof(double), (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        blr_x view = *x;
        view.p = x->p + n0;
        view.N = n;
        view.owned = false;
        BLR_TRY(transpose_to_colvecs(ctx, &view, alpha, D));
    }
    return trsm_lower(ctx, p->L, D, D, alpha, D, n, false);
}

int predict_mean_var_literal(blr_ctx* ctx, blr_post* p, const blr_x* x, const double* sigma2, double sigma2_scalar,
                             double* mean_dev, double* var_dev) {
    const int64_t D = p->D, N = x->N;
    if (N == 0) return 0;
    if (mean_dev) BLR_TRY(apply_weights(ctx, x, p->mw, mean_dev));
    if (!var_dev) return 0;
    const int64_t chunk = std::max<int64_t>(TS_COLS, std::min<int64_t>(N, ((int64_t)1 << 25) / D));  // <= 256 MiB of α at a time
    double* alpha = nullptr;
    BLR_CUDA_OK(ctx, dev_alloc(ctx, &alpha, (size_t)D * chunk * sizeof(double)));
    int rc = 0;
    for (int64_t n0 = 0; n0 < N && rc == 0; n0 += chunk) {
        const int64_t n = std::min(chunk, N - n0);
        rc = solve_alpha(ctx, p, x, n0, n, alpha);
        if (rc != 0) break;
        colnorm_noise_kernel<<<(int)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
            alpha, D, D, n, sigma2 ? sigma2 + n0 : nullptr, sigma2_scalar, var_dev + n0);
        ctx->launches++;
    }
    dev_free(ctx->stream, alpha);
    return rc;
}

}  // namespace blr
