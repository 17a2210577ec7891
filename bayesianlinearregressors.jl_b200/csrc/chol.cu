// K3/K4: the replicated D x D phase -- blocked Cholesky, triangular solves, logdet, posterior assembly.
//
// Reference work replaced (src/bayesian_linear_regression.jl): `_cholesky(Λw)` :78, `_cholesky(Symmetric(Bt'Bt + I))`
// :86, `Λεy.U' \ (Bt'δy)` + `logdet(Λεy)` :57, `Λεy \ (Bt'δy)` :64, `Λεy.U * Uw` :67, `Uw \ mεy` :68, `U'U` :92.
// In closed form on the reduced statistics:
//     Λ' = Λw + G,  L' = chol(Λ') (lower),  z = L'^-1 r,  m' = mw + L'^-T z,
//     logpdf = -1/2 [ n log 2π + ℓ + q + logdet Λ' - logdet Λw - z'z ],   T = L'^T.
// All matrices column-major; factors are stored LOWER (T of the reference is the transpose).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "blockgemm.cuh"
#include "cholblock.cuh"
#include "common.cuh"
#include "internal.h"

namespace blr {

// Panel step j of the right-looking Cholesky: every CTA factors the diagonal block A[j0:j0+64, j0:j0+64]
// redundantly in shared memory (saves a launch + a dependency), CTA 0 writes it back, and each CTA solves
// 256 rows of the panel below:  L21 = A21 * L11^-T  (one thread per row, unrolled substitution in registers).
__global__ void __launch_bounds__(PANEL_THREADS, 1) potrf_panel_kernel(double* __restrict__ A, int64_t ld, int D, int j0,
                                                                       int* __restrict__ info) {
    __shared__ __align__(16) double Ls[NB * LDL];
    __shared__ double rdiag[NB];
    const int nbj = min(NB, D - j0);
    for (int e = threadIdx.x; e < NB * NB; e += PANEL_THREADS) {
        const int r = e % NB, c = e / NB;
        double v = (r == c) ? 1.0 : 0.0;
        if (r < nbj && c < nbj && r >= c) v = A[(int64_t)(j0 + c) * ld + j0 + r];
        Ls[c * LDL + r] = v;
    }
    __syncthreads();
    const int fail = factor_block_smem(Ls, rdiag);
    if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < NB * NB; e += PANEL_THREADS) {
            const int r = e % NB, c = e / NB;
            if (r < nbj && c < nbj) A[(int64_t)(j0 + c) * ld + j0 + r] = (r >= c) ? Ls[c * LDL + r] : 0.0;
        }
        if (threadIdx.x == 0 && fail != 0 && *info == 0) *info = j0 + fail;
    }
    const int row = j0 + NB + blockIdx.x * PANEL_THREADS + threadIdx.x;
    if (row < D) {
        double a[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) a[c] = A[(int64_t)(j0 + c) * ld + row];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const double xk = a[k] * rdiag[k];
            a[k] = xk;
#pragma unroll
            for (int c = k + 1; c < NB; ++c) a[c] = fma(-xk, Ls[k * LDL + c], a[c]);  // L[c][k], broadcast read
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) A[(int64_t)(j0 + c) * ld + row] = a[c];
    }
}

// Trailing update of step j:  A22 -= L21 L21'  on the 64 x 64 blocks of the lower triangle (DMMA).
__global__ void __launch_bounds__(bg::THREADS) potrf_update_kernel(double* __restrict__ A, int64_t ld, int D, int j0) {
    __shared__ double smA[bg::SMEM_A], smB[bg::SMEM_B];
    int bi, bj;
    {
        int i = 0;
        const int idx = blockIdx.x;
        while ((i + 1) * (i + 2) / 2 <= idx) ++i;
        bi = i;
        bj = idx - i * (i + 1) / 2;
    }
    const int r0 = j0 + NB + bi * NB, c0 = j0 + NB + bj * NB;
    double acc[4][4][2];
    acc_zero(acc);
    // A operand: L[r0 + m, j0 + k] (m contiguous); B operand (k, n) = L[c0 + n, j0 + k] (n contiguous)
    cta_gemm64<false>(acc, A + (int64_t)j0 * ld + r0, ld, D - r0, A + (int64_t)j0 * ld + c0, ld, D - c0, NB, smA, smB);
    acc_foreach(acc, [&](int row, int col, double& v) {
        const int gr = r0 + row, gc = c0 + col;
        if (gr < D && gc < D && gr >= gc) A[(int64_t)gc * ld + gr] -= v;
    });
}

__global__ void zero_strict_upper_kernel(double* __restrict__ A, int64_t ld, int D) {
    const int64_t total = (int64_t)D * D;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e % D), c = (int)(e / D);
        if (r < c) A[(int64_t)c * ld + r] = 0.0;
    }
}

// info_dev: 4 device ints ([0] = LAPACK-style info; [3] = watchdog flag of the fused kernel).
int potrf_lower(blr_ctx* ctx, double* A, int64_t D64, int* info_dev) {
    if (!ctx->dxd_legacy) return dxd_fused(ctx, A, D64, info_dev, nullptr, nullptr, nullptr);  // one cooperative launch
    const int D = (int)D64;
    cudaStream_t sm = ctx->stream;
    BLR_CUDA_OK(ctx, cudaMemsetAsync(info_dev, 0, 4 * sizeof(int), sm));
    for (int j0 = 0; j0 < D; j0 += NB) {
        const int below = D - j0 - NB;
        const int pgrid = below > 0 ? (below + PANEL_THREADS - 1) / PANEL_THREADS : 1;
        potrf_panel_kernel<<<pgrid, PANEL_THREADS, 0, sm>>>(A, D64, D, j0, info_dev);
        BLR_CHECK_LAUNCH(ctx, "potrf_panel_kernel");
        if (below > 0) {
            const int nblk = (below + NB - 1) / NB;
            potrf_update_kernel<<<nblk * (nblk + 1) / 2, bg::THREADS, 0, sm>>>(A, D64, D, j0);
            BLR_CHECK_LAUNCH(ctx, "potrf_update_kernel");
        }
    }
    zero_strict_upper_kernel<<<std::min(ctx->sm_count * 4, (int)((D64 * D64 + 255) / 256)), 256, 0, sm>>>(A, D64, D);
    BLR_CHECK_LAUNCH(ctx, "zero_strict_upper_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Triangular solves with a single right-hand side, as a wavefront over 64-row blocks: one CTA per block row,
//     x_j = inv(L_jj) (b_j - Σ_{k<j} L_jk x_k)            (forward;  backward is the transposed mirror image)
// CTA j consumes x_k as soon as CTA k publishes it (flag = launch epoch, release/acquire through global memory),
// so the D x D factor is streamed by D/64 SMs at once and the critical path is D/64 short steps instead of one
// SM reading the whole matrix.  inv(L_jj) comes from trtri_diag_packed_kernel.
// Deadlock freedom rests on CO-RESIDENCY, not on dispatch order (which CUDA does not guarantee): a CTA spins on flags of other
// CTAs, so every CTA of the grid must be resident at once.  The grid has D / 64 <= 256 CTAs (D <= 16384) of 256 threads and
// <= 35 KB of shared memory, of which an SM holds at least two; the launcher checks nblk <= 2 x #SMs and refuses otherwise.
constexpr int TRSV_THREADS = 256;

__device__ __forceinline__ void wait_flag(const int* flag, int epoch) {
    if (threadIdx.x == 0) {
        while (*reinterpret_cast<const volatile int*>(flag) < epoch) {
        }
        __threadfence();
    }
    __syncthreads();
}

// Dinv[b][c * 64 + r] = (inv(L_bb))[r][c]; identity padding for a partial last block
__global__ void __launch_bounds__(NB) trtri_diag_packed_kernel(const double* __restrict__ L, int64_t ld, int D,
                                                               double* __restrict__ Dinv) {
    __shared__ double Ls[NB * 65];
    const int j0 = blockIdx.x * NB, nbj = min(NB, D - j0), c = threadIdx.x;
    for (int e = threadIdx.x; e < NB * NB; e += NB) {
        const int r = e % NB, cc = e / NB;
        Ls[cc * 65 + r] = (r < nbj && cc < nbj && r >= cc) ? L[(int64_t)(j0 + cc) * ld + j0 + r] : (r == cc ? 1.0 : 0.0);
    }
    __syncthreads();
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) acc = fma(-Ls[k * 65 + i], x[k], acc);
        x[i] = (i >= c) ? acc / Ls[i * 65 + i] : 0.0;
    }
    double* out = Dinv + (int64_t)blockIdx.x * NB * NB + c * NB;
#pragma unroll
    for (int i = 0; i < NB; ++i) out[i] = x[i];
}

// TRANS == false: L z = b.  TRANS == true: L' u = b (block rows visited from the bottom; blockIdx 0 = last block).
template <bool TRANS>
__global__ void __launch_bounds__(TRSV_THREADS) trsv_wavefront_kernel(const double* __restrict__ L, int64_t ld, int D,
                                                                      const double* __restrict__ Dinv, double* b,
                                                                      int* flags, int epoch) {
    __shared__ double xs[NB];
    __shared__ double part[4][NB];
    __shared__ double Ts[TRANS ? NB * 65 : 1];
    const int nblk = (D + NB - 1) / NB;
    const int j = TRANS ? nblk - 1 - (int)blockIdx.x : (int)blockIdx.x;
    const int j0 = j * NB;
    const int r = threadIdx.x & 63, q = threadIdx.x >> 6;
    double acc = 0.0;
    const int nprev = TRANS ? nblk - 1 - j : j;
    for (int s = 0; s < nprev; ++s) {
        const int k = TRANS ? nblk - 1 - s : s;  // blocks become ready in this order
        const int k0 = k * NB;
        if (TRANS) {
            // stage L[k0 + rr, j0 + cc] (coalesced over rr) before waiting: the factor itself is already final
            for (int e = threadIdx.x; e < NB * NB; e += TRSV_THREADS) {
                const int rr = e % NB, cc = e / NB;
                Ts[cc * 65 + rr] = (k0 + rr < D && j0 + cc < D) ? L[(int64_t)(j0 + cc) * ld + k0 + rr] : 0.0;
            }
        }
        double lv[16];
        if (!TRANS) {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int c = q * 16 + t;
                lv[t] = (j0 + r < D && k0 + c < D) ? L[(int64_t)(k0 + c) * ld + j0 + r] : 0.0;
            }
        }
        wait_flag(flags + k, epoch);
        if (threadIdx.x < NB) xs[threadIdx.x] = (k0 + threadIdx.x < D) ? __ldcg(b + k0 + threadIdx.x) : 0.0;
        __syncthreads();
        if (TRANS) {
            // column r of the staged block: Σ_rr L[k0 + rr, j0 + r] x_k[rr], rows rr split over the 4 phases q
#pragma unroll
            for (int t = 0; t < 16; ++t) acc = fma(Ts[r * 65 + q * 16 + t], xs[q * 16 + t], acc);
        } else {
#pragma unroll
            for (int t = 0; t < 16; ++t) acc = fma(lv[t], xs[q * 16 + t], acc);
        }
        __syncthreads();
    }
    part[q][r] = acc;
    __syncthreads();
    if (threadIdx.x < NB) {
        const double rhs = ((j0 + r < D) ? b[j0 + r] : 0.0) - (part[0][r] + part[1][r] + part[2][r] + part[3][r]);
        xs[r] = rhs;
    }
    __syncthreads();
    if (threadIdx.x < NB) {
        // x_j = inv(L_jj) rhs   or   inv(L_jj)' rhs
        const double* Dj = Dinv + (int64_t)j * NB * NB;
        double v = 0.0;
        if (TRANS) {
            for (int c = r; c < NB; ++c) v = fma(Dj[r * NB + c], xs[c], v);  // (inv L)'[r][c] = (inv L)[c][r]
        } else {
            for (int c = 0; c <= r; ++c) v = fma(Dj[c * NB + r], xs[c], v);
        }
        if (j0 + r < D) b[j0 + r] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicExch(flags + j, epoch);
    }
}

static int trsv_wavefront(blr_ctx* ctx, const double* L, int64_t D, const double* Dinv, double* b, bool trans) {
    const int nblk = (int)((D + NB - 1) / NB);
    if (nblk > 2 * ctx->sm_count) return set_err(ctx, BLR_E_INVALID, "trsv wavefront: grid would not be co-resident");
    const int epoch = ++ctx->flag_epoch;
    if (trans)
        trsv_wavefront_kernel<true><<<nblk, TRSV_THREADS, 0, ctx->stream>>>(L, D, (int)D, Dinv, b, ctx->d_flags, epoch);
    else
        trsv_wavefront_kernel<false><<<nblk, TRSV_THREADS, 0, ctx->stream>>>(L, D, (int)D, Dinv, b, ctx->d_flags, epoch);
    BLR_CHECK_LAUNCH(ctx, "trsv_wavefront_kernel");
    return 0;
}
int trsv_lower_forward(blr_ctx* ctx, const double* L, int64_t D, const double* Dinv, double* b) {
    return trsv_wavefront(ctx, L, D, Dinv, b, false);
}
int trsv_lower_backward(blr_ctx* ctx, const double* L, int64_t D, const double* Dinv, double* b) {
    return trsv_wavefront(ctx, L, D, Dinv, b, true);
}
int trtri_diag_packed(blr_ctx* ctx, const double* L, int64_t D, double* Dinv) {
    const int nblk = (int)((D + NB - 1) / NB);
    trtri_diag_packed_kernel<<<nblk, NB, 0, ctx->stream>>>(L, D, (int)D, Dinv);
    BLR_CHECK_LAUNCH(ctx, "trtri_diag_packed_kernel");
    return 0;
}

__global__ void __launch_bounds__(256) logdet_kernel(const double* __restrict__ L, int64_t ld, int D,
                                                     double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < D; i += 256) acc += log(L[(int64_t)i * ld + i]);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = 2.0 * acc;
}
int logdet_from_chol(blr_ctx* ctx, const double* L, int64_t D, double* out_dev) {
    logdet_kernel<<<1, 256, 0, ctx->stream>>>(L, D, (int)D, out_dev);
    BLR_CHECK_LAUNCH(ctx, "logdet_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// W = inv(L), lower triangular.  Diagonal 64 x 64 blocks are inverted by substitution; off-diagonal
// blocks follow by block distance d = i - j:  W_ij = -W_ii Σ_{k=j}^{i-1} L_ik W_kj   (DMMA block GEMMs).
__global__ void __launch_bounds__(NB) trtri_diag_kernel(const double* __restrict__ L, int64_t ld, int D,
                                                        double* __restrict__ W) {
    __shared__ double Ls[NB * 65];
    const int j0 = blockIdx.x * NB, nbj = min(NB, D - j0), c = threadIdx.x;
    for (int e = threadIdx.x; e < NB * NB; e += NB) {
        const int r = e % NB, cc = e / NB;
        Ls[cc * 65 + r] = (r < nbj && cc < nbj && r >= cc) ? L[(int64_t)(j0 + cc) * ld + j0 + r] : (r == cc ? 1.0 : 0.0);
    }
    __syncthreads();
    // thread c solves L x = e_c by forward substitution; x lives in registers (loops fully unrolled)
    double x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) acc = fma(-Ls[k * 65 + i], x[k], acc);  // x[k] == 0 for k < c
        x[i] = (i >= c) ? acc / Ls[i * 65 + i] : 0.0;
    }
    __syncthreads();
    // stage through shared memory (reusing Ls) so the global store is coalesced over rows
#pragma unroll
    for (int i = 0; i < NB; ++i) Ls[c * 65 + i] = x[i];
    __syncthreads();
    for (int e = threadIdx.x; e < NB * NB; e += NB) {
        const int r = e % NB, cc = e / NB;
        if (r < nbj && cc < nbj) W[(int64_t)(j0 + cc) * ld + j0 + r] = Ls[cc * 65 + r];
    }
}

__global__ void __launch_bounds__(bg::THREADS) trtri_offdiag_kernel(const double* __restrict__ L, int64_t ld, int D,
                                                                    double* W, int dist) {
    __shared__ double smA[bg::SMEM_A], smB[bg::SMEM_B];
    const int bj = blockIdx.x, bi = bj + dist;
    const int r0 = bi * NB, c0 = bj * NB;
    double acc[4][4][2];
    acc_zero(acc);
    for (int bk = bj; bk < bi; ++bk) {
        const int k0 = bk * NB;
        // A (m, k) = L[r0 + m, k0 + k];  B (k, n) = W[k0 + k, c0 + n]  (k contiguous)
        cta_gemm64<true>(acc, L + (int64_t)k0 * ld + r0, ld, D - r0, W + (int64_t)c0 * ld + k0, ld, D - c0, NB, smA, smB);
    }
    // park the intermediate sum C in the destination block (this CTA is its only reader and writer)
    acc_foreach(acc, [&](int row, int col, double& v) {
        const int gr = r0 + row, gc = c0 + col;
        if (gr < D && gc < D) W[(int64_t)gc * ld + gr] = v;
    });
    __threadfence_block();
    __syncthreads();
    // W_ij = -W_ii * C.   A (m, k) = W[r0 + m, r0 + k];  B (k, n) = C[k, n] = W[r0 + k, c0 + n] (k contiguous)
    acc_zero(acc);
    cta_gemm64<true>(acc, W + (int64_t)r0 * ld + r0, ld, D - r0, W + (int64_t)c0 * ld + r0, ld, D - c0, min(NB, D - r0),
                     smA, smB);
    __syncthreads();
    acc_foreach(acc, [&](int row, int col, double& v) {
        const int gr = r0 + row, gc = c0 + col;
        if (gr < D && gc < D) W[(int64_t)gc * ld + gr] = -v;
    });
}

int trtri_lower(blr_ctx* ctx, const double* L, double* W, int64_t D64) {
    const int D = (int)D64, nblk = (D + NB - 1) / NB;
    cudaStream_t sm = ctx->stream;
    BLR_CUDA_OK(ctx, cudaMemsetAsync(W, 0, (size_t)D64 * D64 * sizeof(double), sm));
    trtri_diag_kernel<<<nblk, NB, 0, sm>>>(L, D64, D, W);
    BLR_CHECK_LAUNCH(ctx, "trtri_diag_kernel");
    for (int dist = 1; dist < nblk; ++dist) {
        trtri_offdiag_kernel<<<nblk - dist, bg::THREADS, 0, sm>>>(L, D64, D, W, dist);
        BLR_CHECK_LAUNCH(ctx, "trtri_offdiag_kernel");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Posterior assembly
__global__ void add_inplace_kernel(double* __restrict__ A, const double* __restrict__ G, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        A[e] += G[e];
}
__global__ void set_diag_kernel(double* __restrict__ A, const double* __restrict__ d, int D) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) A[(int64_t)i * D + i] = d[i];
}
__global__ void transpose_square_kernel(const double* __restrict__ A, double* __restrict__ At, int D) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = bx + threadIdx.x, c = by + j;
        tile[j][threadIdx.x] = (r < D && c < D) ? A[(int64_t)c * D + r] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + threadIdx.x, c = bx + j;  // At[r, c] = A[c, r]
        if (r < D && c < D) At[(int64_t)c * D + r] = tile[threadIdx.x][j];
    }
}
// z'z (block reduce), m' = mw + u, logpdf
__global__ void __launch_bounds__(256) dot_self_kernel(const double* __restrict__ z, int D, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < D; i += 256) acc = fma(z[i], z[i], acc);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = acc;
}
__global__ void finalize_kernel(const double* __restrict__ scal /* q, ℓ, n */, const double* __restrict__ sc,
                                const double* __restrict__ logdet_w, const double* __restrict__ mw, const double* __restrict__ u, int D,
                                double* __restrict__ m_post, double* __restrict__ logpdf, int* __restrict__ noise_info) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) m_post[i] = mw[i] + u[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double LOG2PI = 1.8378770664093454835606594728112;
        const double q = scal[0], l = scal[1], n = scal[2];
        // sc[1] = logdet Λ', sc[2] = z'z
        logpdf[0] = -0.5 * (n * LOG2PI + l + q + (sc[1] - logdet_w[0]) - sc[2]);
        if (!isfinite(l)) *noise_info = 1;  // Σ log σ²: some variance is <= 0 or not finite
    }
}

// Z_j = L^-1 R_j in place for the K columns of R (column j = D contiguous doubles), zz_dev[j] = Z_j'Z_j.  Substitution against
// the factor itself (no inverse factor): the diagonal 64 x 64 blocks are inverted once, each column is one wavefront launch.
int forward_solve_multi(blr_ctx* ctx, const blr_post* p, double* R, int64_t K, double* zz_dev) {
    const int64_t D = p->D, nblk = (D + NB - 1) / NB;
    BLR_TRY(ensure_dinv(ctx, (size_t)nblk * NB * NB * sizeof(double)));
    BLR_TRY(trtri_diag_packed(ctx, p->L, D, ctx->dinv));
    for (int64_t j = 0; j < K; ++j) {
        BLR_TRY(trsv_lower_forward(ctx, p->L, D, ctx->dinv, R + j * D));
        dot_self_kernel<<<1, 256, 0, ctx->stream>>>(R + j * D, (int)D, zz_dev + j);
        BLR_CHECK_LAUNCH(ctx, "dot_self_kernel");
    }
    return 0;
}

void post_release(blr_post* p) {
    if (!p) return;
    dev_free(p->stream, p->mw);
    dev_free(p->stream, p->L);
    dev_free(p->stream, p->W);
    dev_free(p->stream, p->Lam);
    dev_free(p->stream, p->Wp);
    delete p;
}

static int alloc_post(blr_ctx* ctx, int64_t D, blr_post** out) {
    blr_post* p = new blr_post();
    p->D = D;
    p->stream = ctx->stream;
    cudaError_t e = dev_alloc(ctx, &p->mw, (size_t)D * sizeof(double));
    if (e == cudaSuccess) e = dev_alloc(ctx, &p->L, (size_t)D * D * sizeof(double));
    if (e == cudaSuccess) e = dev_alloc(ctx, &p->Lam, (size_t)D * D * sizeof(double));
    if (e != cudaSuccess) {
        post_release(p);
        return cuda_fail(ctx, e, "cudaMalloc(post)");
    }
    *out = p;
    return 0;
}

// Upload the prior precision as a dense column-major D x D device matrix.  For a Diagonal prior the
// log-determinant and positivity check are done on the host (D numbers); returns info > 0 on failure.
static int upload_prior_precision(blr_ctx* ctx, const blr_prior* prior, int64_t D, double* Lam_dev, double* logdet_host,
                                  bool* have_logdet) {
    cudaStream_t sm = ctx->stream;
    *have_logdet = false;
    if (prior->lambda_kind == BLR_LAMBDA_DIAGONAL) {
        double ld = 0.0;
        for (int64_t i = 0; i < D; ++i) {
            if (!(prior->lambda[i] > 0.0)) return (int)(i + 1);
            ld += log(prior->lambda[i]);
        }
        *logdet_host = ld;
        *have_logdet = true;
        double* dtmp = ctx->small + SMALL_DTMP;
        BLR_CUDA_OK(ctx, cudaMemsetAsync(Lam_dev, 0, (size_t)D * D * sizeof(double), sm));
        BLR_CUDA_OK(ctx, cudaMemcpyAsync(dtmp, prior->lambda, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, sm));
        set_diag_kernel<<<(int)((D + 255) / 256), 256, 0, sm>>>(Lam_dev, dtmp, (int)D);
        BLR_CHECK_LAUNCH(ctx, "set_diag_kernel");
        return 0;  // no host sync: dtmp reuse is stream-ordered and pageable H2D copies are staged before returning
    }
    if (prior->lambda_kind == BLR_LAMBDA_DENSE) {
        const int64_t ld = prior->ld > 0 ? prior->ld : D;
        if (ld < D) return set_err(ctx, BLR_E_INVALID, "prior.ld < D");
        BLR_CUDA_OK(ctx, cudaMemcpy2DAsync(Lam_dev, (size_t)D * sizeof(double), prior->lambda, (size_t)ld * sizeof(double),
                                           (size_t)D * sizeof(double), (size_t)D, cudaMemcpyHostToDevice, sm));
        return 0;
    }
    return set_err(ctx, BLR_E_INVALID, "unknown lambda_kind");
}

int read_info(blr_ctx* ctx, int* info_host) {
    int h[4] = {0, 0, 0, 0};
    BLR_CUDA_OK(ctx, cudaMemcpyAsync(h, ctx->d_info, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BLR_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[3] != 0) return set_err(ctx, BLR_E_CUDA, "D x D phase: dependency wait timed out (internal error)");
    *info_host = h[0];
    return 0;
}

int post_from_prior(blr_ctx* ctx, const blr_prior* prior, int64_t D, blr_post** out) {
    blr_post* p = nullptr;
    BLR_TRY(alloc_post(ctx, D, &p));
    cudaStream_t sm = ctx->stream;
    double ldh;
    bool have;
    int r = upload_prior_precision(ctx, prior, D, p->Lam, &ldh, &have);
    if (r != 0) {
        post_release(p);
        return r;
    }
    cudaError_t e = cudaMemcpyAsync(p->mw, prior->mw, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, sm);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->L, p->Lam, (size_t)D * D * sizeof(double), cudaMemcpyDeviceToDevice, sm);
    if (e != cudaSuccess) {
        post_release(p);
        return cuda_fail(ctx, e, "post upload");
    }
    r = potrf_lower(ctx, p->L, D, ctx->d_info);
    int info = 0;
    if (r == 0) r = read_info(ctx, &info);
    if (r == 0 && info != 0) r = info;
    if (r != 0) {
        post_release(p);
        return r;
    }
    *out = p;
    return 0;
}

int post_ensure_W(blr_ctx* ctx, blr_post* p) {
    if (p->has_W) return 0;
    if (!p->W) BLR_CUDA_OK(ctx, dev_alloc(ctx, &p->W, (size_t)p->D * p->D * sizeof(double)));
    BLR_TRY(trtri_lower(ctx, p->L, p->W, p->D));
    p->has_W = true;
    return 0;
}

// Λ' = Λw + G into both Lam (kept: the posterior precision) and L (factored in place), rhs = r -- one pass instead of
// memset + set_diag + add + two device copies.  diag != nullptr: Diagonal prior (Lam holds nothing yet);
// else Lam already holds the dense prior precision.
__global__ void __launch_bounds__(256) assemble_posterior_kernel(double* __restrict__ Lam, double* __restrict__ L,
                                                                 const double* __restrict__ G, const double* __restrict__ diag,
                                                                 int D, const double* __restrict__ r, double* __restrict__ rhs) {
    const int64_t n2 = (int64_t)D * D;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += (int64_t)gridDim.x * blockDim.x) {
        double v = G[e];
        if (diag) {
            const int row = (int)(e % D), col = (int)(e / D);
            if (row == col) v += diag[row];
        } else {
            v += Lam[e];
        }
        Lam[e] = v;
        L[e] = v;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) rhs[i] = r[i];
}

int infer_solve(blr_ctx* ctx, const blr_prior* prior, const blr_stats* st, double* logpdf_out, double* m_post,
                double* T_post, double* L_post, blr_post** post_out) {
    const int64_t D = st->D;
    cudaStream_t sm = ctx->stream;
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[4], sm));
    blr_post* p = nullptr;
    BLR_TRY(alloc_post(ctx, D, &p));
    // scratch scalars live behind the prep partials in ctx->small
    double* sc = ctx->small + SMALL_SC;    // [0] logdet Λw (dense prior)  [1] logdet Λ'  [2] z'z  [3] logpdf
    double* rhs = ctx->small + SMALL_RHS;  // D doubles: r in, z = L^-1 r out
    double* usol = ctx->small + SMALL_U;   // D doubles: u = L^-T z
    double* in = ctx->small + SMALL_MW;    // host inputs, ONE page-locked H2D copy: [mw (D) | diag Λw (D) | logdet Λw]
    double* mwd = in;
    double* diagd = in + D;
    const double* logdet_w = in + 2 * D;
    int* info_post = ctx->d_info;          // [0] info of chol(Λ'), [1] noise flag, [3] watchdog
    int* info_prior = ctx->d_info + 4;     // same for chol(Λw) of a dense prior
    int rc = 0;
    bool lam_on_copy_stream = false;  // the precision download is in flight on the copy stream
    auto fail = [&](int code) {
        if (lam_on_copy_stream) cudaEventSynchronize(ctx->ev[7]);  // it reads p->Lam and writes the caller's buffer
        post_release(p);
        return code;
    };
    if (D > SMALL_VEC) return fail(set_err(ctx, BLR_E_INVALID, "D > 16384 not supported"));
    const bool diagonal = prior->lambda_kind == BLR_LAMBDA_DIAGONAL;
    if (!diagonal && prior->lambda_kind != BLR_LAMBDA_DENSE) return fail(set_err(ctx, BLR_E_INVALID, "unknown lambda_kind"));

    // ---- host inputs.  Diagonal prior: positivity and logdet on the host (D numbers).  The previous inference on this
    // context ended with a stream synchronisation, so the staging buffer is free.
    double* hin = ctx->h_in;
    memcpy(hin, prior->mw, (size_t)D * sizeof(double));
    double ldh = 0.0;
    if (diagonal) {
        for (int64_t i = 0; i < D; ++i) {
            if (!(prior->lambda[i] > 0.0)) return fail((int)(i + 1));
            ldh += log(prior->lambda[i]);
        }
        memcpy(hin + D, prior->lambda, (size_t)D * sizeof(double));
    }
    hin[2 * D] = ldh;
    BLR_CUDA_OK(ctx, cudaMemcpyAsync(in, hin, (size_t)(2 * D + 1) * sizeof(double), cudaMemcpyHostToDevice, sm));
    BLR_CUDA_OK(ctx, cudaMemsetAsync(info_prior, 0, 4 * sizeof(int), sm));
    if (!diagonal) {
        // dense prior: logdet Λw (and its PosDef check) from a device Cholesky of a copy in p->L.  No host synchronisation
        // here: a failed prior factorisation is reported with the results (one sync per inference).
        const int64_t ld = prior->ld > 0 ? prior->ld : D;
        if (ld < D) return fail(set_err(ctx, BLR_E_INVALID, "prior.ld < D"));
        BLR_CUDA_OK(ctx, cudaMemcpy2DAsync(p->Lam, (size_t)D * sizeof(double), prior->lambda, (size_t)ld * sizeof(double),
                                           (size_t)D * sizeof(double), (size_t)D, cudaMemcpyHostToDevice, sm));
        BLR_CUDA_OK(ctx, cudaMemcpyAsync(p->L, p->Lam, (size_t)D * D * sizeof(double), cudaMemcpyDeviceToDevice, sm));
        rc = potrf_lower(ctx, p->L, D, info_prior);
        if (rc == 0) rc = logdet_from_chol(ctx, p->L, D, sc + 0);
        if (rc != 0) return fail(rc);
        logdet_w = sc + 0;
    }

    // ---- the reference's whitened form (opt-in): keep the prior factor Lw -- the assembly below overwrites p->L
    const int64_t n2 = D * D;
    const bool whitened = ctx->form == BLR_FORM_WHITENED;
    double* Lw = nullptr;
    auto fail_w = [&](int code) {
        dev_free(sm, Lw);
        return fail(code);
    };
    if (whitened) {
        cudaError_t e = dev_alloc(ctx, &Lw, (size_t)n2 * sizeof(double));
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "cudaMalloc(Lw)"));
        if (diagonal) {
            rc = diag_factor_dense(ctx, diagd, D, Lw);
            if (rc != 0) return fail_w(rc);
        } else {
            e = cudaMemcpyAsync(Lw, p->L, (size_t)n2 * sizeof(double), cudaMemcpyDeviceToDevice, sm);
            if (e != cudaSuccess) return fail_w(cuda_fail(ctx, e, "copy Lw"));
        }
    }

    // ---- Λ' = Λw + G  (kept in p->Lam), a copy to factor in p->L, rhs = r
    assemble_posterior_kernel<<<(int)std::min<int64_t>((n2 + 255) / 256, ctx->sm_count * 8), 256, 0, sm>>>(
        p->Lam, p->L, st->G(), diagonal ? diagd : nullptr, (int)D, st->r(), rhs);
    BLR_CHECK_LAUNCH(ctx, "assemble_posterior_kernel");
    // The posterior precision is final here: its download (8 MiB at D = 1024, 128 MiB at D = 4096) runs on the copy stream
    // underneath the factorisation and the solves instead of after them (small matrices keep the plain in-order copy:
    // measured, the extra stream hand-over costs more than a sub-100 us copy saves).
    if (L_post && D >= 1024) {
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[6], sm));
        BLR_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[6], 0));
        BLR_CUDA_OK(ctx, cudaMemcpyAsync(L_post, p->Lam, (size_t)n2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[7], ctx->copy_stream));
        lam_on_copy_stream = true;
    }
    if (whitened) {
        // Lε = chol(Lw^-1 G Lw^-T + I), ..., L' = Lw Lε: the reference's own evaluation order (whitened.cu)
        rc = dxd_whitened(ctx, p, Lw, st, rhs, usol, mwd, sc, info_post);
        dev_free(sm, Lw);
        Lw = nullptr;
        if (rc != 0) return fail(rc);
    } else if (!ctx->dxd_legacy) {
        // L = chol(Λ'), z = L^-1 r, u = L^-T z, logdet, z'z, m' = mw + u, logpdf: one cooperative launch (chol_tiled.cu)
        DxdFinalize fin;
        fin.stat_scal = st->scal();
        fin.mw = mwd;
        fin.m_post = p->mw;
        fin.logdet_w = logdet_w;
        fin.sc = sc;
        rc = dxd_fused(ctx, p->L, D, info_post, rhs, usol, &fin);
        if (rc != 0) return fail(rc);
    } else {
        rc = potrf_lower(ctx, p->L, D, info_post);
        if (rc != 0) return fail(rc);
        // z = L'^-1 r ; z'z ; u = L'^-T z ; m' = mw + u  (inverse diagonal blocks for the wavefront solves)
        const int64_t nblk = (D + NB - 1) / NB;
        rc = ensure_dinv(ctx, (size_t)nblk * NB * NB * sizeof(double));
        if (rc != 0) return fail(rc);
        double* Dinv = ctx->dinv;
        rc = trtri_diag_packed(ctx, p->L, D, Dinv);
        if (rc == 0) rc = trsv_lower_forward(ctx, p->L, D, Dinv, rhs);
        if (rc != 0) return fail(rc);
        dot_self_kernel<<<1, 256, 0, sm>>>(rhs, (int)D, sc + 2);
        BLR_CHECK_LAUNCH(ctx, "dot_self_kernel");
        rc = trsv_lower_backward(ctx, p->L, D, Dinv, rhs);
        if (rc == 0) rc = logdet_from_chol(ctx, p->L, D, sc + 1);
        if (rc != 0) return fail(rc);
        finalize_kernel<<<(int)std::min<int64_t>((D + 255) / 256, 64), 256, 0, sm>>>(st->scal(), sc, logdet_w, mwd, rhs, (int)D,
                                                                                  p->mw, sc + 3, info_post + 1);
        BLR_CHECK_LAUNCH(ctx, "finalize_kernel");
    }
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[5], sm));
    ctx->ev_valid[3] = true;

    // Results: the small ones (pivot status, logpdf, m') go through page-locked staging so that every copy is truly
    // asynchronous and the whole inference costs ONE host synchronisation (a cudaMemcpyAsync into pageable memory
    // blocks the host, i.e. one GPU-idle round trip per output).
    double* hr = ctx->h_res;  // [8 ints | logpdf | m']
    BLR_CUDA_OK(ctx, cudaMemcpyAsync(hr, ctx->d_info, 8 * sizeof(int), cudaMemcpyDeviceToHost, sm));
    if (logpdf_out) BLR_CUDA_OK(ctx, cudaMemcpyAsync(hr + 4, sc + 3, sizeof(double), cudaMemcpyDeviceToHost, sm));
    if (m_post) BLR_CUDA_OK(ctx, cudaMemcpyAsync(hr + 5, p->mw, (size_t)D * sizeof(double), cudaMemcpyDeviceToHost, sm));
    if (L_post && !lam_on_copy_stream)
        BLR_CUDA_OK(ctx, cudaMemcpyAsync(L_post, p->Lam, (size_t)n2 * sizeof(double), cudaMemcpyDeviceToHost, sm));
    if (T_post) {
        // T = L'^T (upper).  Transpose into the (not yet built) W buffer, then download.
        if (!p->W) BLR_CUDA_OK(ctx, dev_alloc(ctx, &p->W, (size_t)n2 * sizeof(double)));
        dim3 grid((unsigned)((D + 31) / 32), (unsigned)((D + 31) / 32)), block(32, 8);
        transpose_square_kernel<<<grid, block, 0, sm>>>(p->L, p->W, (int)D);
        BLR_CHECK_LAUNCH(ctx, "transpose_square_kernel");
        BLR_CUDA_OK(ctx, cudaMemcpyAsync(T_post, p->W, (size_t)n2 * sizeof(double), cudaMemcpyDeviceToHost, sm));
    }
    BLR_CUDA_OK(ctx, cudaStreamSynchronize(sm));
    if (lam_on_copy_stream) BLR_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev[7]));
    int hi[8];
    memcpy(hi, hr, sizeof(hi));
    if (hi[3] != 0 || hi[7] != 0) return fail(set_err(ctx, BLR_E_CUDA, "D x D phase: dependency wait timed out (internal error)"));
    if (hi[4] != 0) return fail(hi[4]);          // the prior precision is not positive definite (:78)
    if (hi[1] != 0) return fail(BLR_INFO_NOISE);  // a non-positive observation-noise variance (:79)
    if (hi[0] != 0) return fail(hi[0]);          // non-positive pivot: the outputs above are not meaningful
    if (logpdf_out) *logpdf_out = hr[4];
    if (m_post) memcpy(m_post, hr + 5, (size_t)D * sizeof(double));
    if (post_out)
        *post_out = p;
    else
        post_release(p);
    return 0;
}

}  // namespace blr
