// Bench-only synthetic inputs generated in place on the device (cfg 3-5 cannot be staged
// through host RAM), the device-resident random-Fourier feature map for BasisFunctionRegressor
// (replaces ϕ(x) of src/basis_function_regression.jl:41 for ϕ(x) = sqrt(2/D) cos(W x + b)), and the
// RowVecs -> ColVecs transposition used by the Gram fast path.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "internal.h"
#include "philox.cuh"

namespace blr {

// out[r + c * ld] = N(0,1) for element id (col_offset + c) * rows + r of stream `stream_id`
__global__ void __launch_bounds__(256) synth_normal_kernel(double* __restrict__ out, int64_t rows, int64_t cols,
                                                           int64_t ld, uint64_t seed, uint32_t stream_id,
                                                           int64_t col_offset) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    if ((rows & 1) == 0) {
        const int64_t hr = rows / 2, npairs = hr * cols;
        for (int64_t pidx = (int64_t)blockIdx.x * 256 + threadIdx.x; pidx < npairs; pidx += stride) {
            const int64_t c = pidx / hr, r2 = pidx % hr;
            double z0, z1;
            philox_normal_pair(seed, stream_id, (uint64_t)((col_offset + c) * hr + r2), z0, z1);
            *reinterpret_cast<double2*>(out + c * ld + 2 * r2) = make_double2(z0, z1);
        }
    } else {
        const int64_t total = rows * cols;
        for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += stride) {
            const int64_t c = e / rows, r = e % rows;
            const uint64_t id = (uint64_t)((col_offset + c) * rows + r);
            double z0, z1;
            philox_normal_pair(seed, stream_id, id >> 1, z0, z1);
            out[c * ld + r] = (id & 1) ? z1 : z0;
        }
    }
}

int synth_normal(blr_ctx* ctx, double* out, int64_t rows, int64_t cols, int64_t ld, uint64_t seed, uint64_t stream_id,
                 int64_t col_offset) {
    if (rows == 0 || cols == 0) return 0;
    const bool vec_ok = ((rows & 1) == 0) && ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int64_t work = vec_ok ? rows / 2 * cols : rows * cols;
    const int grid = (int)std::min<int64_t>((work + 255) / 256, (int64_t)ctx->sm_count * 32);
    // the paired path needs 16-byte aligned columns; otherwise present an odd row count to take the scalar path
    if (vec_ok || (rows & 1)) {
        synth_normal_kernel<<<grid, 256, 0, ctx->stream>>>(out, rows, cols, ld, seed, (uint32_t)stream_id, col_offset);
    } else {
        return set_err(ctx, BLR_E_INVALID, "synth_normal: even row count needs an even leading dimension");
    }
    BLR_CHECK_LAUNCH(ctx, "synth_normal_kernel");
    return 0;
}

// σ²_n = exp(z_n)  (heteroscedastic diagonal noise as README.md:50)
__global__ void __launch_bounds__(256) synth_noise_kernel(double* __restrict__ sigma2, int64_t n, uint64_t seed,
                                                          int64_t n_offset) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const uint64_t id = (uint64_t)(n_offset + i);
        double z0, z1;
        philox_normal_pair(seed, 1, id >> 1, z0, z1);
        sigma2[i] = exp((id & 1) ? z1 : z0);
    }
}
int synth_noise(blr_ctx* ctx, double* sigma2, int64_t n, uint64_t seed, int64_t n_offset) {
    if (n == 0) return 0;
    synth_noise_kernel<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        sigma2, n, seed, n_offset);
    BLR_CHECK_LAUNCH(ctx, "synth_noise_kernel");
    return 0;
}

// y_n = x_n' w* + σ_n ε_n   (w* = stream 2, ε = stream 3)
template <int LAYOUT>
__global__ void __launch_bounds__(256) synth_targets_kernel(const double* __restrict__ X, int64_t ld, int D, int64_t N,
                                                            const double* __restrict__ wstar,
                                                            const double* __restrict__ sigma2, uint64_t seed,
                                                            int64_t n_offset, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); n < N; n += warps) {
        double dot = 0.0;
        if (LAYOUT == BLR_COLVECS) {
            const double* col = X + n * ld;
            for (int d = lane; d < D; d += 32) dot = fma(col[d], __ldg(wstar + d), dot);
        } else {
            for (int d = lane; d < D; d += 32) dot = fma(X[(int64_t)d * ld + n], __ldg(wstar + d), dot);
        }
        dot = warp_sum(dot);
        if (lane == 0) {
            const uint64_t id = (uint64_t)(n_offset + n);
            double z0, z1;
            philox_normal_pair(seed, 3, id >> 1, z0, z1);
            y[n] = dot + sqrt(sigma2[n]) * ((id & 1) ? z1 : z0);
        }
    }
}
int synth_targets(blr_ctx* ctx, const blr_x* x, const double* sigma2, uint64_t seed, int64_t n_offset, double* y) {
    if (x->N == 0) return 0;
    if (x->D > SMALL_VEC) return set_err(ctx, BLR_E_INVALID, "D too large");
    double* wstar = ctx->small + SMALL_DTMP;
    // w* as a D x 1 "matrix" with an odd-safe path
    BLR_TRY(synth_normal(ctx, wstar, x->D, 1, x->D + (x->D & 1), seed, 2, 0));
    const int grid = (int)std::min<int64_t>((x->N + 7) / 8, (int64_t)ctx->sm_count * 16);
    if (x->layout == BLR_COLVECS)
        synth_targets_kernel<BLR_COLVECS><<<grid, 256, 0, ctx->stream>>>(x->p, x->ld, (int)x->D, x->N, wstar, sigma2, seed,
                                                                        n_offset, y);
    else
        synth_targets_kernel<BLR_ROWVECS><<<grid, 256, 0, ctx->stream>>>(x->p, x->ld, (int)x->D, x->N, wstar, sigma2, seed,
                                                                        n_offset, y);
    BLR_CHECK_LAUNCH(ctx, "synth_targets_kernel");
    return 0;
}

// out[d + n * ldo] = scale * act(Σ_i W[d, i] x[i, n] + b[d]);  one CTA = 16 observations x all D features.
// ACT: 0 cos (random Fourier features with scale = sqrt(2 / D)), 1 tanh, 2 relu, 3 identity, 4 sin.
// A thread owns one feature at a time and 16 observations: per input dimension one coalesced load of W (re-read once per CTA
// from L2), 8 broadcast LDS.128 of the observations' inputs ([i][observation] in shared memory) and 16 DFMA.  Measured at
// cfg5's shape (d_in = 32, D = 4096, 2^19 inputs, tools/bench_features.py): 16 observations x 2 CTAs / SM 16.7 ms, 32 x 2 19.7,
// 32 x 1 21.5, 8 x 2 20.2 -- about 60 % of the kernel's fp64 instruction bound (32 DFMA + ~50 for cos per output).
#ifndef BLR_RFF_OBS
#define BLR_RFF_OBS 16
#endif
#ifndef BLR_RFF_MINB
#define BLR_RFF_MINB 2
#endif
constexpr int RFF_OBS = BLR_RFF_OBS;
template <int ACT>
__device__ __forceinline__ double feature_act(double z) {
    if (ACT == 0) return cos(z);
    if (ACT == 1) return tanh(z);
    if (ACT == 2) return z > 0.0 ? z : 0.0;
    if (ACT == 4) return sin(z);
    return z;
}
template <int ACT>
__global__ void __launch_bounds__(256, BLR_RFF_MINB) features_kernel(const double* __restrict__ Xin, int64_t ldx, int din, int64_t N,
                                                       const double* __restrict__ W, const double* __restrict__ b, int D,
                                                       double scale, double* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(16) double xs[];  // [din][RFF_OBS]
    const int64_t n0 = (int64_t)blockIdx.x * RFF_OBS;
    for (int e = threadIdx.x; e < RFF_OBS * din; e += 256) {
        const int o = e / din, i = e % din;  // i fastest: coalesced over an observation's inputs
        xs[i * RFF_OBS + o] = (n0 + o < N) ? Xin[(n0 + o) * ldx + i] : 0.0;
    }
    __syncthreads();
    const int nobs = (int)min((int64_t)RFF_OBS, N - n0);
    for (int d = threadIdx.x; d < D; d += 256) {
        double acc[RFF_OBS];
        const double bd = b[d];
#pragma unroll
        for (int o = 0; o < RFF_OBS; ++o) acc[o] = bd;
#pragma unroll 4
        for (int i = 0; i < din; ++i) {
            const double w = W[(int64_t)i * D + d];
            const double2* xi = reinterpret_cast<const double2*>(xs + i * RFF_OBS);
#pragma unroll
            for (int o = 0; o < RFF_OBS / 2; ++o) {
                const double2 x2 = xi[o];
                acc[2 * o] = fma(w, x2.x, acc[2 * o]);
                acc[2 * o + 1] = fma(w, x2.y, acc[2 * o + 1]);
            }
        }
#pragma unroll
        for (int o = 0; o < RFF_OBS; ++o)
            if (o < nobs) out[(n0 + o) * ldo + d] = scale * feature_act<ACT>(acc[o]);
    }
}
int affine_features(blr_ctx* ctx, const blr_x* xin, const double* W_dev, const double* b_dev, int64_t D, int act, double scale,
                    double* out, int64_t ldo) {
    if (xin->N == 0) return 0;
    if (xin->layout != BLR_COLVECS) return set_err(ctx, BLR_E_INVALID, "device feature maps expect ColVecs inputs");
    const size_t smem = (size_t)RFF_OBS * xin->D * sizeof(double);
    if (smem > 48 * 1024) return set_err(ctx, BLR_E_INVALID, "device feature map: d_in too large");
    const int64_t grid = (xin->N + RFF_OBS - 1) / RFF_OBS;
#define BLR_FEATURES_LAUNCH(A)                                                                                              \
    features_kernel<A><<<(unsigned)grid, 256, smem, ctx->stream>>>(xin->p, xin->ld, (int)xin->D, xin->N, W_dev, b_dev, (int)D, \
                                                                  scale, out, ldo)
    switch (act) {
        case 0: BLR_FEATURES_LAUNCH(0); break;
        case 1: BLR_FEATURES_LAUNCH(1); break;
        case 2: BLR_FEATURES_LAUNCH(2); break;
        case 3: BLR_FEATURES_LAUNCH(3); break;
        case 4: BLR_FEATURES_LAUNCH(4); break;
        default: return set_err(ctx, BLR_E_INVALID, "unknown activation");
    }
#undef BLR_FEATURES_LAUNCH
    BLR_CHECK_LAUNCH(ctx, "features_kernel");
    return 0;
}
int rff_features(blr_ctx* ctx, const blr_x* xin, const double* W_dev, const double* b_dev, int64_t D, double* out,
                 int64_t ldo) {
    return affine_features(ctx, xin, W_dev, b_dev, D, 0, sqrt(2.0 / (double)D), out, ldo);
}

// RowVecs (N x D column-major, element (n, d) at d * ld + n)  ->  ColVecs (D x N, element (d, n) at n * ldo + d)
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, int64_t ldi, int64_t N, int D,
                                                        double* __restrict__ out, int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t n0 = (int64_t)blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int64_t n = n0 + tx;
        const int d = d0 + j;
        tile[j][tx] = (n < N && d < D) ? in[(int64_t)d * ldi + n] : 0.0;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int64_t n = n0 + j;
        const int d = d0 + tx;
        if (n < N && d < D) out[n * ldo + d] = tile[tx][j];
    }
}
int transpose_to_colvecs(blr_ctx* ctx, const blr_x* x, double* out, int64_t ldo) {
    const int64_t gx = (x->N + 31) / 32;
    const int gy = (int)((x->D + 31) / 32);
    transpose_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ctx->stream>>>(x->p, x->ld, x->N, (int)x->D, out, ldo);
    BLR_CHECK_LAUNCH(ctx, "transpose_kernel");
    return 0;
}

}  // namespace blr
