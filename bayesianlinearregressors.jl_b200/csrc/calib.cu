// On-box calibration of the fp64 roofline denominators ("fp64 peak must be calibrated on-box"):
// a pure DMMA.8x8x4 issue loop, a pure DFMA loop, and a device-to-device streaming copy.
#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace blr {

constexpr int CAL_THREADS = 512;  // 4 warps per SM sub-partition
constexpr int CAL_ACC = 8;

__global__ void __launch_bounds__(CAL_THREADS) calib_dmma_kernel(double* __restrict__ out, int iters, double seed) {
    double c[CAL_ACC][2];
#pragma unroll
    for (int i = 0; i < CAL_ACC; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CAL_ACC; ++i) dmma884(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CAL_ACC; ++i) s += c[i][0] + c[i][1];
    out[(int64_t)blockIdx.x * CAL_THREADS + threadIdx.x] = s;
}

__global__ void __launch_bounds__(CAL_THREADS) calib_dfma_kernel(double* __restrict__ out, int iters, double seed) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[(int64_t)blockIdx.x * CAL_THREADS + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) calib_copy_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                                         int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) out[i] = in[i];
}

template <typename Launch>
static int time_best(blr_ctx* ctx, int reps, Launch launch, float* best_ms) {
    cudaEvent_t e0, e1;
    BLR_CUDA_OK(ctx, cudaEventCreate(&e0));
    BLR_CUDA_OK(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps + 2; ++r) {
        BLR_CUDA_OK(ctx, cudaEventRecord(e0, ctx->stream));
        launch();
        ctx->launches++;
        BLR_CUDA_OK(ctx, cudaEventRecord(e1, ctx->stream));
        BLR_CUDA_OK(ctx, cudaEventSynchronize(e1));
        float ms;
        BLR_CUDA_OK(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    BLR_CUDA_OK(ctx, cudaGetLastError());
    *best_ms = best;
    return 0;
}

int calib_dmma(blr_ctx* ctx, double* tflops) {
    const int blocks = ctx->sm_count * 2, iters = 20000;
    BLR_TRY(ensure_ws(ctx, (size_t)blocks * CAL_THREADS * sizeof(double)));
    float ms;
    BLR_TRY(time_best(ctx, 5, [&] { calib_dmma_kernel<<<blocks, CAL_THREADS, 0, ctx->stream>>>(ctx->ws, iters, 1.0); }, &ms));
    const double flop = (double)blocks * (CAL_THREADS / 32) * (double)iters * CAL_ACC * 512.0;
    *tflops = flop / (ms * 1e-3) / 1e12;
    return 0;
}

// DMMA issue-rate probe: `warps` warps per SM (one CTA per SM), NACC independent accumulators per warp.
template <int NACC>
__global__ void calib_dmma_cfg_kernel(double* __restrict__ out, int iters, double seed) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int calib_dmma_cfg(blr_ctx* ctx, int warps, int nacc, double* tflops) {
    if (warps < 1 || warps > 32) return set_err(ctx, BLR_E_INVALID, "warps per SM must be 1..32");
    const int blocks = ctx->sm_count, threads = warps * 32;
    const int iters = 400000 / nacc;
    BLR_TRY(ensure_ws(ctx, (size_t)blocks * threads * sizeof(double)));
    float ms;
    auto launch = [&] {
        switch (nacc) {
            case 1: calib_dmma_cfg_kernel<1><<<blocks, threads, 0, ctx->stream>>>(ctx->ws, iters, 1.0); break;
            case 2: calib_dmma_cfg_kernel<2><<<blocks, threads, 0, ctx->stream>>>(ctx->ws, iters, 1.0); break;
            case 4: calib_dmma_cfg_kernel<4><<<blocks, threads, 0, ctx->stream>>>(ctx->ws, iters, 1.0); break;
            case 8: calib_dmma_cfg_kernel<8><<<blocks, threads, 0, ctx->stream>>>(ctx->ws, iters, 1.0); break;
            case 16: calib_dmma_cfg_kernel<16><<<blocks, threads, 0, ctx->stream>>>(ctx->ws, iters, 1.0); break;
            default: calib_dmma_cfg_kernel<32><<<blocks, threads, 0, ctx->stream>>>(ctx->ws, iters, 1.0); break;
        }
    };
    if (nacc != 1 && nacc != 2 && nacc != 4 && nacc != 8 && nacc != 16 && nacc != 32)
        return set_err(ctx, BLR_E_INVALID, "nacc must be 1, 2, 4, 8, 16 or 32");
    BLR_TRY(time_best(ctx, 3, launch, &ms));
    const double flop = (double)blocks * warps * (double)iters * nacc * 512.0;
    *tflops = flop / (ms * 1e-3) / 1e12;
    return 0;
}

// Are DMMA and DFMA separate pipes?  Half of the warps of every SM run the DMMA loop, the other half a DFMA loop.
__global__ void __launch_bounds__(CAL_THREADS) calib_mixed_kernel(double* __restrict__ out, int iters, double seed) {
    const int warp = threadIdx.x >> 5;
    const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    double s = 0.0;
    if ((warp >> 2) & 1) {  // warps 4-7, 12-15: one per SM sub-partition each
        double c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(a, c[i], b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c[i];
    } else {
        double c[CAL_ACC][2];
#pragma unroll
        for (int i = 0; i < CAL_ACC; ++i) c[i][0] = c[i][1] = 0.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < CAL_ACC; ++i) dmma884(c[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < CAL_ACC; ++i) s += c[i][0] + c[i][1];
    }
    out[(int64_t)blockIdx.x * CAL_THREADS + threadIdx.x] = s;
}

// returns the time of the mixed kernel in ms via tflops[0] = DMMA-only-equivalent TF/s, tflops[1] = DFMA TF/s
int calib_mixed(blr_ctx* ctx, double* tflops2) {
    const int blocks = ctx->sm_count, iters = 20000;
    BLR_TRY(ensure_ws(ctx, (size_t)blocks * CAL_THREADS * sizeof(double)));
    float ms;
    BLR_TRY(time_best(ctx, 3, [&] { calib_mixed_kernel<<<blocks, CAL_THREADS, 0, ctx->stream>>>(ctx->ws, iters, 1.0); }, &ms));
    const double warps_each = (double)blocks * (CAL_THREADS / 64);
    tflops2[0] = warps_each * iters * CAL_ACC * 512.0 / (ms * 1e-3) / 1e12;
    tflops2[1] = warps_each * 32.0 * iters * 16 * 2.0 / (ms * 1e-3) / 1e12;
    return 0;
}

int calib_dfma(blr_ctx* ctx, double* tflops) {
    const int blocks = ctx->sm_count * 2, iters = 20000;
    BLR_TRY(ensure_ws(ctx, (size_t)blocks * CAL_THREADS * sizeof(double)));
    float ms;
    BLR_TRY(time_best(ctx, 5, [&] { calib_dfma_kernel<<<blocks, CAL_THREADS, 0, ctx->stream>>>(ctx->ws, iters, 1.0); }, &ms));
    const double flop = (double)blocks * CAL_THREADS * (double)iters * 16 * 2.0;
    *tflops = flop / (ms * 1e-3) / 1e12;
    return 0;
}

int calib_hbm(blr_ctx* ctx, double* gbs) {
    const size_t bytes = (size_t)2 << 30;  // 2 GiB each way: far larger than the 126 MB L2
    double *a = nullptr, *b = nullptr;
    BLR_CUDA_OK(ctx, cudaMalloc(&a, bytes));
    cudaError_t e = cudaMalloc(&b, bytes);
    if (e != cudaSuccess) {
        cudaFree(a);
        return cuda_fail(ctx, e, "cudaMalloc(calib)");
    }
    cudaMemsetAsync(a, 0, bytes, ctx->stream);
    float ms = 0;
    const int64_t n = (int64_t)(bytes / sizeof(double2));
    int rc = time_best(ctx, 5, [&] {
        calib_copy_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(reinterpret_cast<const double2*>(a),
                                                                       reinterpret_cast<double2*>(b), n);
    }, &ms);
    cudaFree(a);
    cudaFree(b);
    if (rc != 0) return rc;
    *gbs = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
    return 0;
}

}  // namespace blr
