// Shared device helpers for libblr_cuda (sm_100a only).
//
//  * mbarrier + cp.async.bulk (TMA 1-D bulk copy, SASS: UBLKCP / SYNCS.*) wrappers
//  * fp64 tensor-core MMA: mma.sync.m8n8k4.f64 (SASS: DMMA.8x8x4 -- the native fp64 tensor
//    op on sm_100a; tcgen05 has no f64 kind, SURVEY.md section 7 "Hard parts")
//  * a warp-level register-tiled MMA step over shared-memory operand tiles
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace blr {

// ------------------------------------------------------------------ PTX: mbarrier / bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Ring hand-over.  Production: ONE elected lane arrives for its warp after __syncwarp() -- the other lanes' shared-memory accesses
// are ordered before that arrive by bar.warp.sync (PTX memory model: causality order is transitive), the idiom of every
// warp-specialised TMA pipeline.  -DBLR_STRICT_ARRIVE builds the variant used to cross-check compute-sanitizer racecheck
// reports (profiles/r02/sanitizer/): every lane arrives itself and the barrier counts are 32x larger, so no ordering relies on
// composing bar.warp.sync with mbarrier.arrive.
#ifdef BLR_STRICT_ARRIVE
constexpr int RING_LANES = 32;
#else
constexpr int RING_LANES = 1;
#endif
// consumer warp: finished reading the slot
__device__ __forceinline__ void ring_release(uint32_t bar, int lane) {
#ifdef BLR_STRICT_ARRIVE
    mbar_arrive(bar);
#else
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
#endif
}
// producer warp: announce the bytes its copies will deliver (all lanes' earlier generic stores to the slot included)
__device__ __forceinline__ void ring_expect(uint32_t bar, uint32_t bytes, int lane) {
    if (lane == 0) mbar_arrive_expect_tx(bar, bytes);
#ifdef BLR_STRICT_ARRIVE
    else mbar_arrive(bar);
#endif
    __syncwarp();
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (TMA engine).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// 2-D tiled TMA load (tensor map built on the host by make_tmap_2d_f64, predict_tma.cu): box -> shared memory, completion on
// an mbarrier.  c0 = coordinate along the contiguous (inner) dimension, c1 = along the outer one; elements outside the tensor
// are filled with zeros and still count towards the transaction bytes (always the full box).
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// Offset (in doubles) of element (row p, column k) of a box with 16 doubles (128 bytes) per row that TMA stored with
// CU_TENSOR_MAP_SWIZZLE_128B: the 16-byte chunk index (k >> 1) is XORed with the row index modulo 8.  The destination must be
// 1024-byte aligned.  With p = 8 j + g the eight rows g = 0..7 of an m8n8k4 fragment land in eight different chunks: the
// fragment load is conflict-free (two wavefronts for 32 x 8 bytes) although the rows are a dense 128 bytes apart.
__device__ __forceinline__ int swz128(int p, int k) { return p * 16 + ((((k >> 1) ^ (p & 7)) << 1) | (k & 1)); }

// ------------------------------------------------------------------ fp64 tensor core
// D(8x8) += A(8x4, row) * B(4x8, col).  Fragment ownership (PTX ISA, mma.m8n8k4 .f64):
//   a  = A[lane>>2][lane&3]     b = B[lane&3][lane>>2]     c[0..1] = C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// One k4 step of a warp tile of (8*MI) x (8*NI):
//   acc[mi][ni] += A[m, k] * B[k, n],  k = 0..3
// A operand element (m, k) lives at  A[m * a_sm + k * a_sk];  B element (k, n) at  B[n * b_sn + k * b_sk].
// Bank-conflict-free whenever the non-unit stride of an operand is == 4 (mod 16) doubles.
template <int MI, int NI>
__device__ __forceinline__ void warp_mma_k4(double (&acc)[MI][NI][2], const double* __restrict__ A, int a_sm,
                                            int a_sk, const double* __restrict__ B, int b_sn, int b_sk, int lane) {
    const int g = lane >> 2, k = lane & 3;
    double a[MI], b[NI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) a[mi] = A[(mi * 8 + g) * a_sm + k * a_sk];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) b[ni] = B[(ni * 8 + g) * b_sn + k * b_sk];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide sum, deterministic (fixed tree).  `red` holds >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    t = red[0];
    return t;
}

}  // namespace blr
