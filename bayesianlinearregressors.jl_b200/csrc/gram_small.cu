// K1 small-D path: the Gram statistics for D <= 64 features (README toy D = 2, the reference's D = 3..7 fixtures, and the
// "few features, very many observations" regime) as an HBM-bound streaming kernel.
//
// With D this small the whole D x D Gram matrix fits in the registers of ONE warp (lower-triangular 8 x 8 sub-tiles of a
// DMMA.8x8x4 accumulator grid), so every warp streams its own contiguous range of observations and no tile is shared:
// the m8n8k4 operand fragments are loaded straight from global memory -- lane (g, k) reads element (feature 8*mi + g,
// observation n + k), which for ColVecs is 8 consecutive doubles per observation and for RowVecs 4 consecutive doubles
// per feature, i.e. whole 32-byte sectors either way -- and the B fragment is the A fragment times s_k, so X is read
// exactly once and nothing is staged in shared memory.  Works for any D <= 64, any leading dimension / alignment and
// both layouts (plain 8-byte loads).  Per-CTA partial matrices are summed in a fixed order by gram_small_reduce_kernel.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace blr {
namespace gs {
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
}  // namespace gs

template <int MI>
__global__ void __launch_bounds__(gs::THREADS, (MI == 8 ? 1 : (MI == 4 ? 2 : 4)))
    gram_small_kernel(const double* __restrict__ X, int64_t sd, int64_t sn, int D, int64_t N, const double* __restrict__ s,
                      const double* __restrict__ t, double* __restrict__ P, double* __restrict__ Pr, int64_t obs_per_warp) {
    using namespace gs;
    constexpr int DP = MI * 8;
    __shared__ double tile[DP * DP];
    __shared__ double rsum[DP];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, kq = lane & 3;
    const int64_t wid = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t n0 = wid * obs_per_warp, n1 = min(N, n0 + obs_per_warp);

    double acc[MI][MI][2];
    double racc[MI];
    bool rowok[MI];
    const double* xrow[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        racc[mi] = 0.0;
        rowok[mi] = (mi * 8 + g) < D;
        xrow[mi] = X + (int64_t)(mi * 8 + g) * sd;
#pragma unroll
        for (int ni = 0; ni < MI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }
#pragma unroll 2
    for (int64_t n = n0; n < n1; n += 4) {
        const int64_t nk = n + kq;
        const bool ok = nk < n1;
        const double sk = ok ? s[nk] : 0.0, tk = ok ? t[nk] : 0.0;
        double a[MI], b[MI];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) a[mi] = (ok && rowok[mi]) ? xrow[mi][nk * sn] : 0.0;
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            racc[mi] = fma(a[mi], tk, racc[mi]);
            b[mi] = a[mi] * sk;
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni <= mi; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
    }
    // r: fold the four observation lanes of every feature
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        racc[mi] += __shfl_xor_sync(0xffffffffu, racc[mi], 1);
        racc[mi] += __shfl_xor_sync(0xffffffffu, racc[mi], 2);
    }
    // CTA reduction in warp order (fixed => bit-reproducible)
    for (int e = tid; e < DP * DP; e += THREADS) tile[e] = 0.0;
    if (tid < DP) rsum[tid] = 0.0;
    for (int w = 0; w < WARPS; ++w) {
        __syncthreads();
        if (warp == w) {
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
                for (int ni = 0; ni <= mi; ++ni) {
                    double* dst = &tile[(mi * 8 + g) * DP + ni * 8 + kq * 2];
                    dst[0] += acc[mi][ni][0];
                    dst[1] += acc[mi][ni][1];
                }
                if (kq == 0) rsum[mi * 8 + g] += racc[mi];
            }
        }
    }
    __syncthreads();
    double* Pt = P + (int64_t)blockIdx.x * (DP * DP);
    for (int e = tid; e < DP * DP; e += THREADS) Pt[e] = tile[e];
    if (tid < DP) Pr[(int64_t)blockIdx.x * DP + tid] = rsum[tid];
}

// stats.G += Σ_blocks P (lower part mirrored), stats.r += Σ_blocks Pr, scalars from the prep partials.
__global__ void __launch_bounds__(256) gram_small_reduce_kernel(const double* __restrict__ P, const double* __restrict__ Pr,
                                                                int DP, int nblocks, int D, double* __restrict__ G,
                                                                double* __restrict__ r, double* __restrict__ scal,
                                                                const double* __restrict__ prep_partial, int prep_blocks,
                                                                double n_obs) {
    __shared__ double red[32];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < DP * DP) {
        const int row = e / DP, col = e % DP;
        if (row < D && col <= row) {
            double v = 0.0;
            for (int b = 0; b < nblocks; ++b) v += P[(int64_t)b * DP * DP + e];
            const double nv = G[(int64_t)col * D + row] + v;
            G[(int64_t)col * D + row] = nv;
            if (row != col) G[(int64_t)row * D + col] = nv;
        }
    }
    if (blockIdx.x == 0) {
        for (int m = threadIdx.x; m < D; m += blockDim.x) {
            double v = 0.0;
            for (int b = 0; b < nblocks; ++b) v += Pr[(int64_t)b * DP + m];
            r[m] += v;
        }
        double q = 0.0, l = 0.0;
        for (int b = threadIdx.x; b < prep_blocks; b += blockDim.x) {
            q += prep_partial[2 * b];
            l += prep_partial[2 * b + 1];
        }
        q = block_sum(q, red);
        l = block_sum(l, red);
        if (threadIdx.x == 0) {
            scal[0] += q;
            scal[1] += l;
            scal[2] += n_obs;
        }
    }
}

template <int MI>
static int launch_small(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* s, const double* t,
                        const double* prep_partial, int prep_blocks) {
    constexpr int DP = MI * 8;
    const int D = (int)x->D;
    const int64_t N = x->N;
    const int per_sm = (MI == 8 ? 1 : (MI == 4 ? 2 : 4));
    int nblocks = ctx->sm_count * per_sm;
    const int64_t groups = (N + 3) / 4;  // k4 steps
    nblocks = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks, (groups + gs::WARPS * 8 - 1) / (gs::WARPS * 8)));
    const int64_t total_warps = (int64_t)nblocks * gs::WARPS;
    const int64_t obs_per_warp = ((groups + total_warps - 1) / total_warps) * 4;
    BLR_TRY(ensure_ws(ctx, (size_t)nblocks * (DP * DP + DP) * sizeof(double)));
    double* P = ctx->ws;
    double* Pr = ctx->ws + (size_t)nblocks * DP * DP;
    const bool colv = x->layout == BLR_COLVECS;
    const int64_t sd = colv ? 1 : x->ld, sn = colv ? x->ld : 1;
    gram_small_kernel<MI><<<nblocks, gs::THREADS, 0, ctx->stream>>>(x->p, sd, sn, D, N, s, t, P, Pr, obs_per_warp);
    BLR_CHECK_LAUNCH(ctx, "gram_small_kernel");
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    gram_small_reduce_kernel<<<(DP * DP + 255) / 256, 256, 0, ctx->stream>>>(P, Pr, DP, nblocks, D, st->G(), st->r(), st->scal(),
                                                                            prep_partial, prep_blocks, (double)N);
    BLR_CHECK_LAUNCH(ctx, "gram_small_reduce_kernel");
    return 0;
}

int gram_small(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* s, const double* t, const double* prep_partial,
               int prep_blocks) {
    const int D = (int)x->D;
    if (D <= 8) return launch_small<1>(ctx, st, x, s, t, prep_partial, prep_blocks);
    if (D <= 16) return launch_small<2>(ctx, st, x, s, t, prep_partial, prep_blocks);
    if (D <= 32) return launch_small<4>(ctx, st, x, s, t, prep_partial, prep_blocks);
    return launch_small<8>(ctx, st, x, s, t, prep_partial, prep_blocks);
}

}  // namespace blr
