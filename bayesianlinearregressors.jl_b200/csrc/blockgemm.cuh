// CTA-level 64 x 64 fp64 block GEMM on DMMA.8x8x4, used by the D x D phases (Cholesky trailing update,
// triangular inverse, weight sampling) and by the first version of the predictive-variance kernel.
//
//   C(64 x 64) += A(64 x kc) * B(kc x 64)      kc <= 64
//
// CTA = 128 threads = 2 x 2 warps, each warp a 32 x 32 register tile.  Operands are staged through
// shared memory in K halves of 32 with padded rows (stride 68 == 4 mod 16 doubles: conflict-free
// LDS.64 for the m8n8k4 fragment pattern).
#pragma once
#include "common.cuh"

namespace blr {
namespace bg {
constexpr int BS = 64;        // block edge
constexpr int KH = 32;        // k extent staged per pass
constexpr int LDP = BS + 4;   // padded stride of a k-major tile row  ([k][m])
constexpr int LDQ = KH + 4;   // padded stride of an n-major tile row ([n][k])
constexpr int THREADS = 128;
constexpr int SMEM_A = KH * LDP;                                        // doubles
constexpr int SMEM_B = (KH * LDP > BS * LDQ) ? KH * LDP : BS * LDQ;     // doubles
}  // namespace bg

// A element (m, k) at Ag[m + k * a_ks]            (m contiguous: column-major block)
// B element (k, n):  B_NMAJOR == false -> Bg[n + k * b_ks]   (n contiguous)
//                    B_NMAJOR == true  -> Bg[k + n * b_ns]   (k contiguous: column-major block)
// Rows/cols beyond *_valid and k beyond kc read as zero.
template <bool B_NMAJOR>
__device__ __forceinline__ void cta_gemm64(double (&acc)[4][4][2], const double* __restrict__ Ag, int64_t a_ks,
                                           int a_valid, const double* __restrict__ Bg, int64_t b_stride, int b_valid,
                                           int kc, double* smA, double* smB) {
    using namespace bg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    for (int kh = 0; kh < kc; kh += KH) {
        __syncthreads();
        for (int e = tid; e < KH * BS; e += THREADS) {
            const int m = e % BS, k = e / BS;
            smA[k * LDP + m] = (m < a_valid && kh + k < kc) ? Ag[m + (int64_t)(kh + k) * a_ks] : 0.0;
        }
        if (B_NMAJOR) {
            for (int e = tid; e < KH * BS; e += THREADS) {
                const int k = e % KH, n = e / KH;
                smB[n * LDQ + k] = (n < b_valid && kh + k < kc) ? Bg[(kh + k) + (int64_t)n * b_stride] : 0.0;
            }
        } else {
            for (int e = tid; e < KH * BS; e += THREADS) {
                const int n = e % BS, k = e / BS;
                smB[k * LDP + n] = (n < b_valid && kh + k < kc) ? Bg[n + (int64_t)(kh + k) * b_stride] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k4 = 0; k4 < KH; k4 += 4) {
            if (B_NMAJOR)
                warp_mma_k4<4, 4>(acc, smA + k4 * LDP + wm * 32, 1, LDP, smB + (wn * 32) * LDQ + k4, LDQ, 1, lane);
            else
                warp_mma_k4<4, 4>(acc, smA + k4 * LDP + wm * 32, 1, LDP, smB + k4 * LDP + wn * 32, 1, LDP, lane);
        }
    }
}

__device__ __forceinline__ void acc_zero(double (&acc)[4][4][2]) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}

// Visit every accumulator element owned by this thread: f(row, col, value&) with row, col in [0, 64).
template <typename F>
__device__ __forceinline__ void acc_foreach(double (&acc)[4][4][2], F f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, kq = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int c = 0; c < 2; ++c) f(wm * 32 + mi * 8 + g, wn * 32 + ni * 8 + kq * 2 + c, acc[mi][ni][c]);
}

}  // namespace blr
