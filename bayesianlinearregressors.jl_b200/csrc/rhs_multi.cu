// Right-hand sides of several observation vectors under ONE design matrix (AbstractGPs' `logpdf(fx, Y::AbstractMatrix)`, the
// form the reference's conformance test calls, test/bayesian_linear_regression.jl:7-9).
//
// In `__compute_inference_quantities` (src/bayesian_linear_regression.jl:72-89) only δy = y - X'mw depends on y: the Gram
// matrix, its factorisation and every log-determinant are shared by all columns of Y.  So the O(N D²) Gram pass runs once (for
// column 0, the ordinary path) and each further column needs
//     r_j = X Σy⁻¹ (y_j - X'mw)   (D)        q_j = (y_j - X'mw)' Σy⁻¹ (y_j - X'mw)
// which this file forms for all columns in ONE pass over X:  R = X T,  T[n, j] = s_n (Y[n, j] - pm_n),  a skinny
// (D x N) x (N x K) product that is HBM-bound for small K (8 FMA per 8 bytes of X at the K-block of 8 used here).
//
// Grid = (row blocks of 256) x (column blocks of 8) x (observation chunks).  A CTA walks its chunk in tiles of 32 observations:
// the tile of T is built in shared memory (Y is read coalesced along n), then thread `d` streams its row of X and updates its
// 8 accumulators.  ColVecs data is read straight from global memory (consecutive threads = consecutive features of one
// observation); RowVecs data is staged through a padded shared tile so that the global reads run along n.  Chunk partials are
// written to a workspace and summed in FIXED order by a second kernel -- no atomics, bit-reproducible like the Gram path.
#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace blr {
namespace rm {
constexpr int THREADS = 256;  // = rows of X per CTA
constexpr int KB = 8;         // columns of Y per CTA
constexpr int OT = 32;        // observations per tile
constexpr int XPAD = OT + 1;  // RowVecs staging: xs[row][o], odd stride -> conflict-free both ways
}  // namespace rm

template <bool COLV>
__global__ void __launch_bounds__(rm::THREADS) rhs_multi_kernel(const double* __restrict__ X, int64_t ld, int D, int64_t N,
                                                               const double* __restrict__ Y, int64_t ldy, int K,
                                                               const double* __restrict__ sigma2, double sigma2_scalar,
                                                               const double* __restrict__ pm, int64_t chunk_obs,
                                                               double* __restrict__ part_R, double* __restrict__ part_q) {
    using namespace rm;
    __shared__ __align__(16) double ts[OT][KB];
    extern __shared__ double xs[];  // RowVecs only: [THREADS][XPAD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d0 = blockIdx.x * THREADS, d = d0 + tid;
    const int j0 = blockIdx.y * KB;
    const int64_t n_begin = (int64_t)blockIdx.z * chunk_obs, n_end = min(N, n_begin + chunk_obs);
    double acc[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc[j] = 0.0;
    double qacc = 0.0;                  // this thread's share of q_{j0 + warp} (its T entry has column index `warp`)
    const bool col_ok = j0 + warp < K;  // THREADS / 32 == KB: warp w builds column j0 + w of the T tile

    for (int64_t n0 = n_begin; n0 < n_end; n0 += OT) {
        const int valid = (int)min((int64_t)OT, n_end - n0);
        __syncthreads();  // the previous tile has been consumed
        {
            const int64_t n = n0 + lane;
            double t = 0.0;
            if (lane < valid && col_ok) {
                const double s = 1.0 / (sigma2 ? sigma2[n] : sigma2_scalar);
                const double dy = Y[n + (int64_t)(j0 + warp) * ldy] - (pm ? pm[n] : 0.0);
                t = s * dy;
                qacc = fma(t, dy, qacc);
            }
            ts[lane][warp] = t;
        }
        if (!COLV) {  // stage X[d0 .. d0+255, n0 .. n0+31]: lanes run along n (the contiguous direction of RowVecs)
#pragma unroll 4
            for (int rr = warp; rr < THREADS; rr += THREADS / 32) {
                const int row = d0 + rr;
                xs[rr * XPAD + lane] = (row < D && lane < valid) ? X[(int64_t)row * ld + n0 + lane] : 0.0;
            }
        }
        __syncthreads();
        if (d < D) {
            const double* xp = COLV ? X + n0 * ld + d : nullptr;
#pragma unroll 8
            for (int o = 0; o < OT; ++o) {
                if (o >= valid) break;
                const double x = COLV ? xp[(int64_t)o * ld] : xs[tid * XPAD + o];
                const double2* t2 = reinterpret_cast<const double2*>(&ts[o][0]);
#pragma unroll
                for (int j = 0; j < KB / 2; ++j) {
                    const double2 tv = t2[j];
                    acc[2 * j] = fma(x, tv.x, acc[2 * j]);
                    acc[2 * j + 1] = fma(x, tv.y, acc[2 * j + 1]);
                }
            }
        }
    }
    // partial R: [chunk][column][D]
    if (d < D) {
#pragma unroll
        for (int j = 0; j < KB; ++j)
            if (j0 + j < K) part_R[((int64_t)blockIdx.z * K + j0 + j) * D + d] = acc[j];
    }
    if (blockIdx.x == 0) {  // every row block forms the same T: one of them reports q
        qacc = warp_sum(qacc);
        if (lane == 0 && col_ok) part_q[(int64_t)blockIdx.z * K + j0 + warp] = qacc;
    }
}

// out[e] = Σ_c part[c * len + e], c ascending (fixed order)
__global__ void __launch_bounds__(256) rhs_reduce_kernel(const double* __restrict__ part, int nchunk, int64_t len,
                                                         double* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (int64_t)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int c = 0; c < nchunk; ++c) a += part[(int64_t)c * len + e];
        out[e] = a;
    }
}

// R_out (K x D, column j contiguous) and q_out (K) for the K columns of Y (device, ld = ldy); pm = X'mw (device, N) or nullptr.
int rhs_multi(blr_ctx* ctx, const blr_x* x, const double* Y, int64_t ldy, int64_t K, const double* sigma2,
              double sigma2_scalar, const double* pm, double* R_out, double* q_out) {
    using namespace rm;
    const int64_t D = x->D, N = x->N;
    cudaStream_t sm = ctx->stream;
    if (K == 0) return 0;
    if (N == 0) {
        BLR_CUDA_OK(ctx, cudaMemsetAsync(R_out, 0, (size_t)(K * D) * sizeof(double), sm));
        BLR_CUDA_OK(ctx, cudaMemsetAsync(q_out, 0, (size_t)K * sizeof(double), sm));
        return 0;
    }
    const int64_t gx = (D + THREADS - 1) / THREADS, gy = (K + KB - 1) / KB;
    if (gy > 65535) return set_err(ctx, BLR_E_INVALID, "too many columns in Y");
    const int64_t tiles = (N + OT - 1) / OT;
    int64_t nchunk = std::max<int64_t>(1, (4 * (int64_t)ctx->sm_count + gx * gy - 1) / (gx * gy));
    nchunk = std::min<int64_t>(std::min<int64_t>(nchunk, tiles), 65535);
    const int64_t chunk_obs = (tiles + nchunk - 1) / nchunk * OT;
    nchunk = (N + chunk_obs - 1) / chunk_obs;
    double* part = nullptr;
    BLR_CUDA_OK(ctx, dev_alloc(ctx, &part, (size_t)(nchunk * K * (D + 1)) * sizeof(double)));
    double* part_q = part + nchunk * K * D;
    const dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)nchunk);
    if (x->layout == BLR_COLVECS) {
        rhs_multi_kernel<true><<<grid, THREADS, 0, sm>>>(x->p, x->ld, (int)D, N, Y, ldy, (int)K, sigma2, sigma2_scalar, pm,
                                                       chunk_obs, part, part_q);
    } else {
        const int smem = THREADS * XPAD * (int)sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(rhs_multi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            dev_free(sm, part);
            return cuda_fail(ctx, e, "cudaFuncSetAttribute(rhs_multi_kernel)");
        }
        rhs_multi_kernel<false><<<grid, THREADS, smem, sm>>>(x->p, x->ld, (int)D, N, Y, ldy, (int)K, sigma2, sigma2_scalar, pm,
                                                             chunk_obs, part, part_q);
    }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        rhs_reduce_kernel<<<(int)std::min<int64_t>((K * D + 255) / 256, 1024), 256, 0, sm>>>(part, (int)nchunk, K * D, R_out);
        rhs_reduce_kernel<<<1, 256, 0, sm>>>(part_q, (int)nchunk, K, q_out);
        ctx->launches += 2;
        e = cudaGetLastError();
    }
    dev_free(sm, part);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch rhs_multi_kernel");
    return 0;
}

}  // namespace blr
