// Diagonal-block Cholesky shared by the D x D kernels (chol.cu: legacy panel kernel; chol_tiled.cu: fused tiled kernel).
#pragma once
#include "common.cuh"

namespace blr {

constexpr int NB = 64;  // block size of the D x D phase

// ---------------------------------------------------------------------------------------------
// Factor a 64 x 64 diagonal block held in shared memory, in place: Ls[c * LDL + r] = element (r, c) for r >= c (the
// strict upper part is never read or written).  The sequential pivot chain is what bounds this phase (D dependent
// column steps for the whole matrix), so it is kept in registers: the block is processed as four 16-column strips;
//   1. warp 0 factors the strip's 16 x 16 diagonal sub-block with lane r holding row r (16 registers); pivots and
//      scaled columns travel by warp shuffle, 1/sqrt comes from rsqrt + one Newton step for the diagonal itself --
//      one column step is a shuffle, an rsqrt and a multiply deep, no barrier, no shared-memory round trip;
//   2. one thread per row below solves its 16 entries of the strip against that sub-block (registers);
//   3. all 256 threads apply the rank-16 update to the rest of the block.
// On return L[r][c] = Ls[c * LDL + r] for r >= c and rdiag[k] = 1 / L[k][k].  Blocks smaller than 64 are padded
// with the identity by the caller.  Returns (to all threads) the 1-based index of the first non-positive pivot, or 0.
constexpr int PANEL_THREADS = 256;
constexpr int LDL = NB + 2;  // even stride: 16-byte aligned column starts, conflict-free for consecutive rows
constexpr int SB = 16;       // strip width
static __device__ int factor_block_smem(double* Ls, double* rdiag) {
    __shared__ int fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) fail = 0;
    __syncthreads();
#pragma unroll 1
    for (int k0 = 0; k0 < NB; k0 += SB) {
        if (warp == 0) {
            const int r = lane & (SB - 1);  // lanes 16..31 mirror lanes 0..15 so that every shuffle is full-warp
            double a[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) a[c] = (c <= r) ? Ls[(k0 + c) * LDL + k0 + r] : 0.0;
            int bad = 0;
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                const double akk = __shfl_sync(0xffffffffu, a[k], k);
                const double rd = rsqrt(akk);
                double d = akk * rd;
                d = fma(fma(-d, d, akk), 0.5 * rd, d);  // sqrt(akk) to the last bit or two
                if (!(akk > 0.0) && bad == 0) bad = k0 + k + 1;
                a[k] = (r == k) ? d : a[k] * rd;        // rows r > k: L[r][k]; rows r < k hold zeros
                if (lane == k) rdiag[k0 + k] = rd;
#pragma unroll
                for (int c = k + 1; c < SB; ++c) {
                    const double lck = __shfl_sync(0xffffffffu, a[k], c);  // L[c][k]
                    a[c] = fma(-a[k], lck, a[c]);
                }
            }
            __syncwarp();  // the mirror lanes' loads above precede the stores below (explicit for racecheck; free here)
            if (lane < SB) {
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (c <= r) Ls[(k0 + c) * LDL + k0 + r] = a[c];
            }
            if (lane == 0 && bad != 0 && fail == 0) fail = bad;
        }
        __syncthreads();
        const int below = NB - k0 - SB;  // rows of the block under this strip's diagonal sub-block
        if (below > 0) {
            if (tid < below) {
                const int row = k0 + SB + tid;
                double a[SB];
#pragma unroll
                for (int c = 0; c < SB; ++c) a[c] = Ls[(k0 + c) * LDL + row];
#pragma unroll
                for (int k = 0; k < SB; ++k) {
                    const double xk = a[k] * rdiag[k0 + k];
                    a[k] = xk;
#pragma unroll
                    for (int c = k + 1; c < SB; ++c) a[c] = fma(-xk, Ls[(k0 + k) * LDL + k0 + c], a[c]);
                }
#pragma unroll
                for (int c = 0; c < SB; ++c) Ls[(k0 + c) * LDL + row] = a[c];
            }
            __syncthreads();
            const int i = tid & (NB - 1), ty = tid >> 6;
            if (i >= k0 + SB) {
                double li[SB];
#pragma unroll
                for (int k = 0; k < SB; ++k) li[k] = Ls[(k0 + k) * LDL + i];
                for (int j = k0 + SB + ty; j <= i; j += PANEL_THREADS / NB) {
                    double dot = 0.0;
#pragma unroll
                    for (int k = 0; k < SB; ++k) dot = fma(li[k], Ls[(k0 + k) * LDL + j], dot);
                    Ls[j * LDL + i] -= dot;
                }
            }
            __syncthreads();
        }
    }
    return fail;
}

}  // namespace blr
