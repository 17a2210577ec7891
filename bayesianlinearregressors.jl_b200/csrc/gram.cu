// K0/K1: fused noise-scale + SYRK Gram accumulation
//
//     G = X S X'   r = X S δ   q = δ'Sδ   ℓ = Σ log σ²_n        S = diag(1/σ²),  δ = y - X'mw
//
// These are the O(N D²) statistics behind `__compute_inference_quantities`
// (reference src/bayesian_linear_regression.jl:72-89: Bt = Uy^-T X' Uw^-1 at :81, δy at :82,
// logdet Σy + |δy|² at :84, Bt'Bt at :86, Bt'δy at :57/:64) written in closed form so that the two
// N x D temporaries of the reference never exist.
//
// Data layout in HBM: X is the D x N column-major ColVecs matrix (one observation = D contiguous
// doubles); s_n = 1/σ²_n and t_n = s_n δ_n are N-vectors produced by the prep kernel.
//
// Fast path (gram_tma_kernel): the lower triangle of G is cut into 128 x 128 tiles.  One persistent CTA per SM
// owns an equal share of the (tile, observation) work space (stream-K: a few segments = tile x observation
// range), keeps the current tile in registers as fp64 tensor-core accumulators (DMMA.8x8x4), and streams the
// tile's two 128-row panels of X through a 4-stage shared-memory ring filled by TMA bulk copies
// (cp.async.bulk + mbarrier) issued by a dedicated producer warp.  Diagonal tiles skip the 8 x 8 sub-tiles above
// the diagonal.  Each segment's partial tile goes to its own workspace slot; gram_reduce_kernel sums the slots of
// a tile in a FIXED order, so results are bit-reproducible run to run.
#include <math.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "schedule.h"

namespace blr {

// ====================================================================== K0: per-observation prep
// s_n = 1/σ²_n, t_n = s_n δ_n, block partials of q = Σ s δ² and ℓ = Σ log σ².
// LAYOUT 0 = ColVecs, 1 = RowVecs.  HAS_MEAN = prior mean is non-zero (δ needs a pass over X).
constexpr int PREP_THREADS = 256;

template <int LAYOUT, bool HAS_MEAN>
__global__ void __launch_bounds__(PREP_THREADS) prep_kernel(const double* __restrict__ X, int64_t ld, int D, int64_t N,
                                                            int64_t npad, const double* __restrict__ y,
                                                            const double* __restrict__ sigma2, double sigma2_scalar,
                                                            const double* __restrict__ mw, double* __restrict__ s,
                                                            double* __restrict__ t, double* __restrict__ partial) {
    __shared__ double red[32];
    double q = 0.0, l = 0.0;
    if (HAS_MEAN && LAYOUT == BLR_COLVECS && D > 64) {
        // one warp per observation: the column is contiguous, lanes stride over features.  (Only for D > 64: with fewer features
        // most lanes idle and a warp per observation is an order of magnitude too many warps -- 13.7 ms for 2^25 observations of
        // 24 features; those take the thread-per-observation branch below, whose strided loads are served sector by sector from L1.)
        const int lane = threadIdx.x & 31;
        const int64_t warps = (int64_t)gridDim.x * (PREP_THREADS / 32);
        // 16-byte streaming loads with four independent accumulator chains when the columns are 16-byte aligned (D and ld even):
        // this pass is pure HBM streaming (8 D bytes per observation) and runs serially in front of the Gram kernel
        const bool vec2 = ((D | ld) & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(mw) & 15) == 0;
        for (int64_t n = (int64_t)blockIdx.x * (PREP_THREADS / 32) + (threadIdx.x >> 5); n < npad; n += warps) {
            if (n < N) {
                const double* col = X + n * ld;
                double dot = 0.0;
                if (vec2) {
                    const double2* c2 = reinterpret_cast<const double2*>(col);
                    const double2* m2 = reinterpret_cast<const double2*>(mw);
                    const int D2 = D >> 1;
                    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
                    int i = lane;
                    for (; i + 32 < D2; i += 64) {
                        const double2 xa = __ldcs(c2 + i), xb = __ldcs(c2 + i + 32);
                        const double2 ma = __ldg(m2 + i), mb = __ldg(m2 + i + 32);
                        d0 = fma(xa.x, ma.x, d0);
                        d1 = fma(xa.y, ma.y, d1);
                        d2 = fma(xb.x, mb.x, d2);
                        d3 = fma(xb.y, mb.y, d3);
                    }
                    if (i < D2) {
                        const double2 xa = __ldcs(c2 + i), ma = __ldg(m2 + i);
                        d0 = fma(xa.x, ma.x, d0);
                        d1 = fma(xa.y, ma.y, d1);
                    }
                    dot = (d0 + d1) + (d2 + d3);
                } else {
                    for (int d = lane; d < D; d += 32) dot = fma(col[d], __ldg(mw + d), dot);
                }
                dot = warp_sum(dot);
                if (lane == 0) {
                    const double v = sigma2 ? sigma2[n] : sigma2_scalar;
                    const double sn = 1.0 / v, dl = y[n] - dot;
                    s[n] = sn;
                    t[n] = sn * dl;
                    q += sn * dl * dl;
                    l += log(v);
                }
            } else if (lane == 0) {
                s[n] = 0.0;
                t[n] = 0.0;
            }
        }
    } else {
        const int64_t stride = (int64_t)gridDim.x * PREP_THREADS;
        for (int64_t n = (int64_t)blockIdx.x * PREP_THREADS + threadIdx.x; n < npad; n += stride) {
            if (n < N) {
                double dot = 0.0;
                if (HAS_MEAN) {
                    if (LAYOUT == BLR_COLVECS) {  // small D: each thread walks its own (contiguous) observation
                        const double* col = X + n * ld;
                        for (int d = 0; d < D; ++d) dot = fma(col[d], __ldg(mw + d), dot);
                    } else {  // RowVecs: feature-contiguous, threads of a warp read consecutive n
                        for (int d = 0; d < D; ++d) dot = fma(X[(int64_t)d * ld + n], __ldg(mw + d), dot);
                    }
                }
                const double v = sigma2 ? sigma2[n] : sigma2_scalar;
                const double sn = 1.0 / v, dl = y[n] - dot;
                s[n] = sn;
                t[n] = sn * dl;
                q += sn * dl * dl;
                l += log(v);
            } else {
                s[n] = 0.0;
                t[n] = 0.0;
            }
        }
    }
    q = block_sum(q, red);
    l = block_sum(l, red);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = q;
        partial[2 * blockIdx.x + 1] = l;
    }
}

// ====================================================================== K1 fast path
namespace gk {
constexpr int TM = 128;                   // tile edge (rows of G per panel)
constexpr int KT_MAX = 32;                // observations per pipeline stage: 16 (4-stage ring) or 32 (3-stage ring)
constexpr int LDT = TM + 4;               // padded smem row: stride == 4 (mod 16) doubles -> conflict-free LDS.64
constexpr int CONSUMER_WARPS = 8;         // 2 (m) x 4 (n) warps, warp tile 64 x 32
// bulk copies are uniform-datapath ops (one lane at a time), so they are spread over four producer warps: two per panel
// (ColVecs: each takes half of the stage's observations; feature-major: half of the panel's 128 feature rows)
constexpr int PRODUCER_WARPS = 4;
constexpr int PRODUCER_WARPS_ROW = 4;
constexpr int THREADS = (CONSUMER_WARPS + PRODUCER_WARPS) * 32;
constexpr int THREADS_ROW = (CONSUMER_WARPS + PRODUCER_WARPS_ROW) * 32;
constexpr int W_OFF = 64;                 // schedule weight of one stage of an off-diagonal tile
// ROW == false: ColVecs input, a panel is staged [k][m] (one bulk copy per observation: its 128-row slice);
// ROW == true : RowVecs input (feature-major), staged [m][k] (one bulk copy per feature: its KT observations).
// Either way the non-unit stride is == 4 (mod 16) doubles, so the m8n8k4 fragment reads are bank-conflict free.
template <int KT, bool ROW = false>
struct __align__(16) Stage {
    static constexpr int SK = ROW ? 1 : LDT;        // element (k, m) of a panel lives at k * SK + m * SM
    static constexpr int SM = ROW ? KT + 4 : 1;
    static constexpr int PANEL = ROW ? TM * (KT + 4) : KT * LDT;
    double a[PANEL];     // panel I
    double b[PANEL];     // panel J (unused on diagonal tiles)
    double s[KT];        // 1/σ²
    double t[KT];        // δ/σ²
};
template <int KT, int STAGES, bool ROW = false>
struct Smem {
    Stage<KT, ROW> st[STAGES];
    double rred[TM];
    unsigned long long full[STAGES];
    unsigned long long empty[STAGES];
};
}  // namespace gk

// Stream-K schedule (built on the host, see build_schedule): the (tile, stage) work space is linearised
// tile-major with per-tile weights (diagonal tiles skip their strictly-upper 8 x 8 sub-tiles and are cheaper), cut
// into one equal share per CTA; a CTA therefore owns a few SEGMENTS (tile, stage range), each flushed to its own
// partial slot.  Segments of one tile are contiguous in the global segment list, so the reduction order is fixed.
struct GramParams {
    const double* X;
    int64_t ld;
    int D;
    int64_t N;
    const double* s;
    const double* t;
    double* P;    // [nseg][TM*TM]
    double* Pr;   // [nseg][TM]
    const int* cta_seg_begin;  // [grid + 1]
    const int* seg_tile;       // [nseg]
    const int* seg_g0;         // [nseg] first stage WITHIN a period, 16.16 fixed point (see schedule.h)
    const int* seg_g1;         // [nseg] one past the last stage within a period, 16.16 fixed point
    int PS;                    // stages per period: the same cut of the (tile, stage) space is repeated every period,
    int NP;                    //   so all CTAs sweep the observations together and a period's panels stay L2-resident
    int n_stages;              // total stages = ceil(N / KT)
    int fix_bits;              // fractional bits of seg_g0 / seg_g1
    unsigned int* period_counter;  // soft barrier: number of (CTA, period) pairs whose loads have all been issued
};


// integer stage of a fixed-point boundary in period `per` (dither θ_per shared by all CTAs, θ_0 = 0)
__device__ __forceinline__ int sched_stage(int fix, int per, int bits) {
    const unsigned int theta = per == 0 ? 0u : (((unsigned int)per * 40503u) & 0xffffu) >> (16 - bits);
    return (int)(((unsigned int)fix + theta) >> bits);
}

__device__ __forceinline__ void tile_from_index(int idx, int& ti, int& tj) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= idx) ++i;
    ti = i;
    tj = idx - i * (i + 1) / 2;
}

// One pipeline stage of a consumer warp.  THR is a compile-time threshold: sub-tile (mi, ni) is computed only when
// mi - ni >= THR (diagonal tiles skip the 8 x 8 sub-tiles above the diagonal; (wm*8 + mi) >= (wn*4 + ni) <=>
// mi - ni >= wn*4 - wm*8).  Compile-time, because predicating an mma.sync at run time makes ptxas wrap every DMMA in
// WARPSYNC.ALL, which serialises the tensor pipe (measured: no speed-up at all from run-time skipping).
//   THR = -8: all 32 sub-tiles;  THR = 0: 26;  THR = 4: 10.
template <bool DIAG, int THR, int KT, bool ROW = false>
__device__ __forceinline__ void consume_stage(double (&acc)[8][4][2], const gk::Stage<KT, ROW>& S, int wm, int wn, int g,
                                              int kq) {
    using namespace gk;
    constexpr int SK = Stage<KT, ROW>::SK, SM = Stage<KT, ROW>::SM;
    const double* Ap = S.a + (wm * 64 + g) * SM;
    const double* Bp = (DIAG ? S.a : S.b) + (wn * 32 + g) * SM;
#pragma unroll
    for (int kk = 0; kk < KT / 4; ++kk) {
        const int kl = kk * 4 + kq;
        const double sk = S.s[kl];
        double a[8], b[4];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
            if (mi - 0 >= THR) a[mi] = Ap[kl * SK + mi * 8 * SM];  // row block mi is used by some ni iff mi >= THR
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            if (7 - ni >= THR) b[ni] = Bp[kl * SK + ni * 8 * SM] * sk;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                if ((mi - ni) >= THR) dmma884(acc[mi][ni], a[mi], b[ni]);
            }
    }
}

// All stages of one segment for one consumer warp.  MODE: 0 off-diagonal; 1..3 diagonal with THR -8 / 0 / 4;
// 4 diagonal, warp entirely above the diagonal (only helps with the r block).
template <int MODE, int KT, int STAGES, bool ROW>
__device__ __forceinline__ void run_segment(double (&acc)[8][4][2], double& racc, gk::Smem<KT, STAGES, ROW>& sm, int& it,
                                            int nst, int wm, int wn, int g, int kq, int rm, int rhalf, int lane) {
    using namespace gk;
    for (int i = 0; i < nst; ++i, ++it) {
        const int stg = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(smem_u32(&sm.full[stg]), ph);
        const Stage<KT, ROW>& S = sm.st[stg];
        if (MODE == 0) consume_stage<false, -8, KT, ROW>(acc, S, wm, wn, g, kq);
        if (MODE == 1) consume_stage<true, -8, KT, ROW>(acc, S, wm, wn, g, kq);
        if (MODE == 2) consume_stage<true, 0, KT, ROW>(acc, S, wm, wn, g, kq);
        if (MODE == 3) consume_stage<true, 4, KT, ROW>(acc, S, wm, wn, g, kq);
        if (MODE != 0) {
#pragma unroll
            for (int k = 0; k < KT / 2; ++k) {  // r block of this row panel: r[m] += Σ_k X[m,k] t_k
                const int kl = rhalf * (KT / 2) + k;
                racc = fma(S.a[kl * Stage<KT, ROW>::SK + rm * Stage<KT, ROW>::SM], S.t[kl], racc);
            }
        }
        ring_release(smem_u32(&sm.empty[stg]), lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// "Column-strip" consumer tiling (ColVecs ring only): the 8 consumer warps sit 1 (m) x 8 (n), warp tile 128 x 16.
// Each B fragment is then scaled by exactly ONE warp -- 2 DMUL per 32 DMMA instead of 4 (the DMULs share the fp64 pipe
// with the DMMAs: the 2 x 4 mix tops out at 95.7 % of the pure-DMMA rate, this one at 97.9 %, `blr_calibrate_gram_inner`
// with BLR_PROBE_VARIANT) -- and fragments are fetched with 16-byte loads: lane g of an A fragment pair holds rows
// 16 p + 2 g and 16 p + 2 g + 1 (one LDS.128), lane g of the B pair holds columns 16 wn + 2 g, + 1.  The accumulator of
// (pair p, e; b; c) is element (row 16 p + 2 g + e, col 16 wn + 4 kq + 2 c + b): a thread owns 2 x 4 blocks, flushed as
// two 16-byte stores per row.  Diagonal tiles: the 16-row band p is needed by column strip wn iff p >= wn (compile-time
// P0 = wn; 36 live 8 x 8 sub-tiles per SM sub-partition with the warp -> strip map {0,1,2,3,7,6,5,4}).
template <int P0, int KT, bool UNIT>
__device__ __forceinline__ void consume_stage_cs(double (&acc)[16][2][2], const double* __restrict__ Asrc,
                                                 const double* __restrict__ Bsrc, const double* __restrict__ Ssrc, int wn,
                                                 int g, int kq) {
    using namespace gk;
    const double* Ap = Asrc + 2 * g;
    const double* Bp = Bsrc + wn * 16 + 2 * g;
#pragma unroll
    for (int kk = 0; kk < KT / 4; ++kk) {
        const int kl = kk * 4 + kq;
        const double sk = Ssrc[kl];
        double a[16], b[2];
#pragma unroll
        for (int pr = P0; pr < 8; ++pr) {
            const double2 v = *reinterpret_cast<const double2*>(Ap + kl * LDT + pr * 16);
            a[2 * pr] = v.x;
            a[2 * pr + 1] = v.y;
        }
        const double2 bv = *reinterpret_cast<const double2*>(Bp + kl * LDT);
        b[0] = UNIT ? bv.x : bv.x * sk;  // UNIT: homoscedastic noise, the common factor 1/σ² is applied by the reduction
        b[1] = UNIT ? bv.y : bv.y * sk;
#pragma unroll
        for (int mi = 2 * P0; mi < 16; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
    }
}

template <int P0, int KT, int STAGES, bool UNIT>
__device__ __forceinline__ void run_segment_cs(double (&acc)[16][2][2], double& racc, gk::Smem<KT, STAGES, false>& sm, int& it,
                                               int nst, bool diag, int wn, int g, int kq, int rm, int rhalf, int lane) {
    using namespace gk;
    for (int i = 0; i < nst; ++i, ++it) {
        const int stg = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(smem_u32(&sm.full[stg]), ph);
        const Stage<KT, false>& S = sm.st[stg];
        consume_stage_cs<P0, KT, UNIT>(acc, S.a, diag ? S.a : S.b, S.s, wn, g, kq);
        if (diag) {
#pragma unroll
            for (int k = 0; k < KT / 2; ++k) {  // r block of this row panel: r[m] += Σ_k X[m,k] t_k
                const int kl = rhalf * (KT / 2) + k;
                racc = fma(S.a[kl * LDT + rm], S.t[kl], racc);
            }
        }
        ring_release(smem_u32(&sm.empty[stg]), lane);
    }
}

// Diagonal tiles of the hybrid kernel: warp w owns the 8-row blocks w and 15 - w of the 16 x 16 grid of 8 x 8 sub-tiles.
// Block row R has R + 1 sub-tiles on or below the diagonal, so every warp carries exactly 17 of the 136 live sub-tiles and
// the two warps of an SM sub-partition always have equal work to overlap.  Measured on the alternatives: maps that leave a
// sub-partition one heavy and one light (or idle) warp run the heavy warp's DMMA stream at 1 per 21-32 cycles instead of 16
// -- a diagonal tile then costs as much as a full one.  Only the two A fragments are scaled by s_k (2 DMUL per 17 DMMA).
// acc[c] (c <= R0) is sub-tile (R0, c); acc[R0 + 1 + c] (c <= R1) is sub-tile (R1, c).
template <int W, int KT, int DU, bool UNIT, bool ROW>
__device__ __forceinline__ void consume_stage_rp(double (&acc)[17][2], const gk::Stage<KT, ROW>& S, int g, int kq) {
    using namespace gk;
    constexpr int R0 = W, R1 = 15 - W;          // block rows of this warp (R0 < R1)
    constexpr int N0 = R0 + 1, N1 = R1 + 1;     // live sub-tiles in each (columns 0 .. R)
    constexpr int SK = Stage<KT, ROW>::SK, SM = Stage<KT, ROW>::SM;  // element (k, m) of the panel at k * SK + m * SM
    const double* Ap = S.a + g * SM;
    // DU = k4 steps per loop trip.  Eight warps run eight different instruction streams here; fully unrolled (DU = 8:
    // 56 KB for the eight variants) they miss the instruction cache once most SMs of the chip run diagonal tiles at the
    // same time (D <= 512: 37 % of the stall samples were `no_instructions`, a diagonal tile cost as much as a full one),
    // while at larger D, where few CTAs are on diagonal tiles, the full unroll is the faster one (35.2 vs 34.8 TF at D = 1024).
#pragma unroll(DU)
    for (int kk = 0; kk < KT / 4; ++kk) {
        const int kl = kk * 4 + kq;
        const double sk = S.s[kl];
        const double a0 = UNIT ? Ap[kl * SK + R0 * 8 * SM] : Ap[kl * SK + R0 * 8 * SM] * sk;
        const double a1 = UNIT ? Ap[kl * SK + R1 * 8 * SM] : Ap[kl * SK + R1 * 8 * SM] * sk;
        double b[N1];
#pragma unroll
        for (int c = 0; c < N1; ++c) b[c] = Ap[kl * SK + c * 8 * SM];
#pragma unroll
        for (int c = 0; c < N1; ++c) {
            if (c < N0) dmma884(acc[c], a0, b[c]);
            dmma884(acc[N0 + c], a1, b[c]);
        }
    }
}

template <int W, int KT, int STAGES, int DU, bool UNIT, bool ROW = false>
__device__ __forceinline__ void run_segment_rp(double (&acc)[17][2], double& racc, gk::Smem<KT, STAGES, ROW>& sm, int& it,
                                               int nst, int g, int kq, int rm, int rhalf, int lane) {
    using namespace gk;
    for (int i = 0; i < nst; ++i, ++it) {
        const int stg = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(smem_u32(&sm.full[stg]), ph);
        const Stage<KT, ROW>& S = sm.st[stg];
        consume_stage_rp<W, KT, DU, UNIT, ROW>(acc, S, g, kq);
#pragma unroll
        for (int k = 0; k < KT / 2; ++k) {  // r block of this row panel: r[m] += Σ_k X[m,k] t_k
            const int kl = rhalf * (KT / 2) + k;
            racc = fma(S.a[kl * Stage<KT, ROW>::SK + rm * Stage<KT, ROW>::SM], S.t[kl], racc);
        }
        ring_release(smem_u32(&sm.empty[stg]), lane);
    }
}

template <int W>
__device__ __forceinline__ void flush_rp(const double (&acc)[17][2], double* __restrict__ Pt, int g, int kq, bool add) {
    using namespace gk;
    constexpr int R0 = W, R1 = 15 - W, N0 = R0 + 1, N1 = R1 + 1;
#pragma unroll
    for (int c = 0; c < N1; ++c) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == 0 && c >= N0) continue;
            const int R = h == 0 ? R0 : R1;
            const double(&v)[2] = acc[h == 0 ? c : N0 + c];
            double2* dst = reinterpret_cast<double2*>(Pt + (R * 8 + g) * TM + c * 8 + kq * 2);
            double2 o = make_double2(v[0], v[1]);
            if (add) {
                const double2 old = *dst;
                o.x += old.x;
                o.y += old.y;
            }
            *dst = o;
        }
    }
}

template <int KT, int STAGES, bool ROW = false, bool CS = false, int DU = 8, bool UNIT = false>
__global__ void __launch_bounds__(ROW ? gk::THREADS_ROW : gk::THREADS, 1) gram_tma_kernel(const GramParams p) {
    using namespace gk;
    using Stage = gk::Stage<KT, ROW>;
    using Smem = gk::Smem<KT, STAGES, ROW>;
    constexpr int NTHREADS = ROW ? THREADS_ROW : THREADS;
    constexpr int NPRODUCERS = ROW ? PRODUCER_WARPS_ROW : PRODUCER_WARPS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int seg_begin = p.cta_seg_begin[blockIdx.x], seg_end = p.cta_seg_begin[blockIdx.x + 1];

    // zero the ring once: rows past D (tail tile) and observations past N (tail stage) must read as 0.
    {
        double* z = reinterpret_cast<double*>(sm.st);
        const int nz = (int)(sizeof(Stage) * STAGES / sizeof(double));
        for (int i = tid; i < nz; i += NTHREADS) z[i] = 0.0;
    }
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&sm.full[i]), NPRODUCERS * RING_LANES);
            mbar_init(smem_u32(&sm.empty[i]), CONSUMER_WARPS * RING_LANES);
        }
        mbar_fence_init();
    }
    fence_proxy_async();  // order the generic-proxy zero fill before the async-proxy (TMA) writes
    __syncthreads();

    if (warp >= CONSUMER_WARPS) {
        if constexpr (CS || ROW) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");  // hand registers to the consumer warpgroups
        // ------------------------------------------------------------ producer warps (TMA): 0 -> panel I + s, 1 -> panel J + t
        // (feature-major: warps 0, 1 -> halves of panel I, warps 2, 3 -> halves of panel J)
        const int pwr = warp - CONSUMER_WARPS;
        const int pw = pwr >> 1;   // which panel
        const int half = pwr & 1;  // which half of its rows (feature-major) / of the stage's observations (ColVecs)
        int it = 0;  // running stage counter of this CTA: ring slot and phase continue across segments and periods
        for (int per = 0; per < p.NP; ++per) {
            // Soft barrier: nobody starts period `per` before every CTA has issued all loads of period per - 2, which
            // bounds the skew between CTAs to < 2 periods (the working set the L2 has to hold).  Needs all CTAs
            // co-resident: the kernel is launched cooperatively with one CTA per SM.
            if (per >= 2) {
                if (lane == 0) {
                    const unsigned int need = (unsigned int)(per - 1) * gridDim.x;
                    while (*reinterpret_cast<volatile unsigned int*>(p.period_counter) < need) {
                    }
                }
                __syncwarp();
            }
            const int base = per * p.PS;
            for (int sg = seg_begin; sg < seg_end; ++sg) {
                int ti, tj;
                tile_from_index(p.seg_tile[sg], ti, tj);
                const bool diag = (ti == tj);
                const int i0 = ti * TM, j0 = tj * TM;
                const int rowsA = min(TM, p.D - i0), rowsB = min(TM, p.D - j0);
                const int g0 = base + sched_stage(p.seg_g0[sg], per, p.fix_bits);
                const int g1 = min(base + sched_stage(p.seg_g1[sg], per, p.fix_bits), p.n_stages);
                // a tail tile has fewer rows than the previous tenant of the ring slot: stale rows would be read as data
                const bool narrow = (rowsA < TM) || (!diag && rowsB < TM);
                for (int gi = g0; gi < g1; ++gi, ++it) {
                    const int stg = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    mbar_wait(smem_u32(&sm.empty[stg]), ph ^ 1u);
                    const int64_t k0 = (int64_t)gi * KT;
                    const int kc = (int)min((int64_t)KT, p.N - k0);
                    Stage& S = sm.st[stg];
                    const uint32_t bar = smem_u32(&sm.full[stg]);
                    // UNIT: observations beyond N are not neutralised by s = 0, so the one partial stage clears its slot too
                    if (narrow || (UNIT && kc < KT)) {  // rare (tiles in the last block row when D % 128 != 0): clear my part first
                        // (the part this warp is about to fill: a quarter of panel I on a ColVecs diagonal tile, see below)
                        const bool quarter = !ROW && diag;
                        double* z = quarter ? S.a + pwr * (Stage::PANEL / 4) : (pw == 0 ? S.a : S.b) + half * (Stage::PANEL / 2);
                        const int nz = quarter ? Stage::PANEL / 4 : Stage::PANEL / 2;
                        for (int i = lane; i < nz; i += 32) z[i] = 0.0;
                        fence_proxy_async();
                        __syncwarp();
                    }
                    const bool have_panel = (pw == 0) || !diag;
                    const int rows = pw == 0 ? rowsA : rowsB;
                    if (ROW) {
                        // feature-major input: one copy per feature row of the panel (its kc observations); this warp
                        // owns rows [64 half, 64 half + 64) of its panel, the first warp of a panel also brings s / t
                        const int my_rows = have_panel ? max(0, min(TM / 2, rows - half * (TM / 2))) : 0;
                        ring_expect(bar, (uint32_t)kc * (uint32_t)my_rows * 8u + (half == 0 ? KT * 8u : 0u), lane);
                        double* panel = pw == 0 ? S.a : S.b;
                        const int r0 = pw == 0 ? i0 : j0;
#pragma unroll
                        for (int q = 0; q < TM / 64; ++q) {
                            const int m = half * (TM / 2) + q * 32 + lane;
                            if (m - half * (TM / 2) < my_rows)
                                bulk_g2s(smem_u32(panel + m * Stage::SM), p.X + (int64_t)(r0 + m) * p.ld + k0, (uint32_t)kc * 8u,
                                         bar);
                        }
                        if (lane == 0 && half == 0) bulk_g2s(smem_u32(pw == 0 ? S.s : S.t), (pw == 0 ? p.s : p.t) + k0, KT * 8u, bar);
                    } else {
                        // one copy per observation (its rows-long slice).  Off-diagonal tile: this warp owns observations
                        // [KT/2 half, KT/2 half + KT/2) of its panel.  Diagonal tile (panel I only): all four warps share
                        // it, KT/4 observations each -- the per-lane copies of a warp issue one after another (~100 cycles
                        // apiece), and with two warps a diagonal stage took as long to ISSUE as to consume.
                        const int kbeg = diag ? pwr * (KT / 4) : half * (KT / 2);
                        const int kcnt = diag ? KT / 4 : KT / 2;
                        const int my_k = (diag || have_panel) ? max(0, min(kcnt, kc - kbeg)) : 0;
                        const int rws = diag ? rowsA : rows;
                        ring_expect(bar, (uint32_t)my_k * (uint32_t)rws * 8u + (half == 0 ? KT * 8u : 0u), lane);
                        const int kl = kbeg + lane;
                        if (lane < my_k)
                            bulk_g2s(smem_u32(((diag || pw == 0) ? S.a : S.b) + kl * LDT),
                                     p.X + (k0 + kl) * p.ld + ((diag || pw == 0) ? i0 : j0), (uint32_t)rws * 8u, bar);
                        if (lane == 0 && half == 0) bulk_g2s(smem_u32(pw == 0 ? S.s : S.t), (pw == 0 ? p.s : p.t) + k0, KT * 8u, bar);
                    }
                }
            }
            if (pwr == 0 && lane == 0) atomicAdd(p.period_counter, 1u);
        }
        return;
    }

    // ---------------------------------------------------------------- consumer warps (DMMA)
    const int g = lane >> 2, kq = lane & 3;
    const int rm = tid & (TM - 1), rhalf = tid >> 7;
    if constexpr (CS) {
        // Hybrid tiling: off-diagonal tiles (86 % of the work at D = 1024, 97 % at D = 4096) run the column-strip consumer
        // (consume_stage_cs: half the DMULs); diagonal tiles keep the 2 x 4 tiling below, whose compile-time sub-tile
        // skipping leaves every SM sub-partition a second warp to overlap with (measured: a column-strip diagonal tile,
        // where the heavy strips run nearly alone on their sub-partition, costs more than a full tile).  Consumer
        // warpgroups take the producers' registers: 128 x 40 + 256 x 232 = 384 x 168.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int wn_cs = warp < 4 ? warp : 11 - warp;

        int it = 0;
        const bool single = (seg_end - seg_begin) == 1;
        double acc[16][2][2];  // column-strip view; the diagonal path uses the same 64 registers as [8][4][2]
        auto& acc_dg = reinterpret_cast<double (&)[17][2]>(acc);  // diagonal tiles: see consume_stage_rp
#pragma unroll
        for (int mi = 0; mi < 16; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        double racc = 0.0;
        auto flush = [&](int sg, bool diag, bool add) {
            double* Pt = p.P + (int64_t)sg * (TM * TM);
            if (!diag) {
#pragma unroll
                for (int mi = 0; mi < 16; ++mi) {
                    const int row = (mi >> 1) * 16 + 2 * g + (mi & 1);
                    double2* dst = reinterpret_cast<double2*>(Pt + row * TM + wn_cs * 16 + 4 * kq);
                    double2 v0 = make_double2(acc[mi][0][0], acc[mi][1][0]);  // cols 4 kq + 0, + 1
                    double2 v1 = make_double2(acc[mi][0][1], acc[mi][1][1]);  // cols 4 kq + 2, + 3
                    if (add) {
                        const double2 o0 = dst[0], o1 = dst[1];
                        v0.x += o0.x;
                        v0.y += o0.y;
                        v1.x += o1.x;
                        v1.y += o1.y;
                    }
                    dst[0] = v0;
                    dst[1] = v1;
                }
                return;
            }
            if (warp == 0) flush_rp<0>(acc_dg, Pt, g, kq, add);
            else if (warp == 1) flush_rp<1>(acc_dg, Pt, g, kq, add);
            else if (warp == 2) flush_rp<2>(acc_dg, Pt, g, kq, add);
            else if (warp == 3) flush_rp<3>(acc_dg, Pt, g, kq, add);
            else if (warp == 4) flush_rp<4>(acc_dg, Pt, g, kq, add);
            else if (warp == 5) flush_rp<5>(acc_dg, Pt, g, kq, add);
            else if (warp == 6) flush_rp<6>(acc_dg, Pt, g, kq, add);
            else flush_rp<7>(acc_dg, Pt, g, kq, add);
            asm volatile("bar.sync 1, 256;" ::: "memory");  // rred may still be read by the previous flush
            if (rhalf == 1) sm.rred[rm] = racc;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (rhalf == 0) {
                double* dst = p.Pr + (int64_t)sg * TM + rm;
                *dst = (add ? *dst : 0.0) + (racc + sm.rred[rm]);
            }
        };
        for (int per = 0; per < p.NP; ++per) {
            const int base = per * p.PS;
            for (int sg = seg_begin; sg < seg_end; ++sg) {
                int ti, tj;
                tile_from_index(p.seg_tile[sg], ti, tj);
                const bool diag = (ti == tj);
                const int nst = max(0, min(base + sched_stage(p.seg_g1[sg], per, p.fix_bits), p.n_stages) - (base + sched_stage(p.seg_g0[sg], per, p.fix_bits)));
                // warp-uniform dispatch to a fully unrolled, unpredicated instruction stream
                if (!diag) run_segment_cs<0, KT, STAGES, UNIT>(acc, racc, sm, it, nst, false, wn_cs, g, kq, rm, rhalf, lane);
                else if (warp == 0) run_segment_rp<0, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 1) run_segment_rp<1, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 2) run_segment_rp<2, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 3) run_segment_rp<3, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 4) run_segment_rp<4, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 5) run_segment_rp<5, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 6) run_segment_rp<6, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else run_segment_rp<7, KT, STAGES, DU, UNIT>(acc_dg, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                if (!single) {
                    flush(sg, diag, per > 0);
#pragma unroll
                    for (int mi = 0; mi < 16; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                    racc = 0.0;
                }
            }
        }
        if (single) {
            int ti, tj;
            tile_from_index(p.seg_tile[seg_begin], ti, tj);
            flush(seg_begin, ti == tj, false);
        }
        return;
    }
    // off-diagonal tiles: warp -> (wm, wn) row-major.  Diagonal tiles: remapped so that the surviving sub-tile counts
    // {26, 10, 0, 0, 32, 32, 26, 10} pair up evenly over the four SM sub-partitions (warp % 4): 32, 32, 36, 36.
    // The feature-major instantiation (ROW) takes the producers' registers and runs diagonal tiles as row pairs
    // (consume_stage_rp) like the hybrid kernel; its off-diagonal tiles stay 2 x 4.
    constexpr bool RPD = ROW;
    if constexpr (RPD) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int wm_off = warp >> 2, wn_off = warp & 3;
    const int dmap = (0x7132'6054 >> (4 * warp)) & 0xf;  // warp 0..7 -> (1,0) (1,1) (0,0) (1,2) (0,2) (0,3) (0,1) (1,3)
    const int wm_dg = dmap >> 2, wn_dg = dmap & 3;
    int it = 0;
    // A CTA with a single segment keeps its tile in registers across all periods and flushes once; a CTA that
    // switches tiles parks the partial tile in its workspace slot at every switch (first period: store, later: add).
    const bool single = (seg_end - seg_begin) == 1;
    double acc[8][4][2];
    auto& acc_rp = reinterpret_cast<double (&)[17][2]>(acc);  // RPD: diagonal tiles use the first 34 of the 64 accumulators
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    double racc = 0.0;

    auto flush = [&](int sg, bool diag, int wm, int wn, int thr, bool add) {
        double* Pt = p.P + (int64_t)sg * (TM * TM);
        if (RPD && diag) {
            if (warp == 0) flush_rp<0>(acc_rp, Pt, g, kq, add);
            else if (warp == 1) flush_rp<1>(acc_rp, Pt, g, kq, add);
            else if (warp == 2) flush_rp<2>(acc_rp, Pt, g, kq, add);
            else if (warp == 3) flush_rp<3>(acc_rp, Pt, g, kq, add);
            else if (warp == 4) flush_rp<4>(acc_rp, Pt, g, kq, add);
            else if (warp == 5) flush_rp<5>(acc_rp, Pt, g, kq, add);
            else if (warp == 6) flush_rp<6>(acc_rp, Pt, g, kq, add);
            else flush_rp<7>(acc_rp, Pt, g, kq, add);
        } else
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            const int row = wm * 64 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int col = wn * 32 + ni * 8 + kq * 2;
                if (!diag || (mi - ni) >= thr) {
                    double2* dst = reinterpret_cast<double2*>(Pt + row * TM + col);
                    double2 v = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                    if (add) {
                        const double2 o = *dst;
                        v.x += o.x;
                        v.y += o.y;
                    }
                    *dst = v;
                }
            }
        }
        if (diag) {
            asm volatile("bar.sync 1, 256;" ::: "memory");  // rred may still be read by the previous flush
            if (rhalf == 1) sm.rred[rm] = racc;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (rhalf == 0) {
                double* dst = p.Pr + (int64_t)sg * TM + rm;
                *dst = (add ? *dst : 0.0) + (racc + sm.rred[rm]);
            }
        }
    };

    for (int per = 0; per < p.NP; ++per) {
        const int base = per * p.PS;
        for (int sg = seg_begin; sg < seg_end; ++sg) {
            int ti, tj;
            tile_from_index(p.seg_tile[sg], ti, tj);
            const bool diag = (ti == tj);
            const int nst = max(0, min(base + sched_stage(p.seg_g1[sg], per, p.fix_bits), p.n_stages) - (base + sched_stage(p.seg_g0[sg], per, p.fix_bits)));
            const int wm = diag ? wm_dg : wm_off, wn = diag ? wn_dg : wn_off;
            const int thr = wn * 4 - wm * 8;  // sub-tile (mi, ni) needed iff mi - ni >= thr

            // warp-uniform dispatch to a fully unrolled, unpredicated instruction stream
            if (!diag) run_segment<0, KT, STAGES, ROW>(acc, racc, sm, it, nst, wm, wn, g, kq, rm, rhalf, lane);
            else if (RPD) {
                if (warp == 0) run_segment_rp<0, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 1) run_segment_rp<1, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 2) run_segment_rp<2, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 3) run_segment_rp<3, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 4) run_segment_rp<4, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 5) run_segment_rp<5, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else if (warp == 6) run_segment_rp<6, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
                else run_segment_rp<7, KT, STAGES, 4, false, ROW>(acc_rp, racc, sm, it, nst, g, kq, rm, rhalf, lane);
            } else if (thr <= -3) run_segment<1, KT, STAGES, ROW>(acc, racc, sm, it, nst, wm, wn, g, kq, rm, rhalf, lane);
            else if (thr == 0) run_segment<2, KT, STAGES, ROW>(acc, racc, sm, it, nst, wm, wn, g, kq, rm, rhalf, lane);
            else if (thr == 4) run_segment<3, KT, STAGES, ROW>(acc, racc, sm, it, nst, wm, wn, g, kq, rm, rhalf, lane);
            else run_segment<4, KT, STAGES, ROW>(acc, racc, sm, it, nst, wm, wn, g, kq, rm, rhalf, lane);

            if (!single) {
                flush(sg, diag, wm, wn, thr, per > 0);
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
                racc = 0.0;
            }
        }
    }
    if (single) {
        int ti, tj;
        tile_from_index(p.seg_tile[seg_begin], ti, tj);
        const bool diag = (ti == tj);
        const int wm = diag ? wm_dg : wm_off, wn = diag ? wn_dg : wn_off;
        flush(seg_begin, diag, wm, wn, wn * 4 - wm * 8, false);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Ceiling probe: the consumer instruction mix alone (12 LDS.64 + LDS + 4 DMUL + 32 DMMA per k4 step, 8 warps per SM) on a
// resident shared-memory stage -- no TMA, no mbarriers, no epilogue.  The gap between this and the real kernel is what
// the pipeline synchronisation costs; the gap between this and the pure-DMMA calibration is the cost of the mix itself.
template <int KT>
__global__ void __launch_bounds__(256, 1) gram_inner_probe_kernel(double* __restrict__ out, int iters) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    gk::Stage<KT>& S = *reinterpret_cast<gk::Stage<KT>*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(gk::Stage<KT>) / sizeof(double)); i += 256)
        reinterpret_cast<double*>(&S)[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    for (int it = 0; it < iters; ++it) consume_stage<false, -8, KT>(acc, S, warp >> 2, warp & 3, lane >> 2, lane & 3);
    double sum = 0.0;
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) sum += acc[mi][ni][0] + acc[mi][ni][1];
    out[(int64_t)blockIdx.x * 256 + threadIdx.x] = sum;
}

// Alternative consumer mixes for the same probe (BLR_PROBE_VARIANT): what would a different warp tiling buy?
//   1: warps 1 (m) x 8 (n), warp tile 128 x 16: 16 A fragments fetched as 8 LDS.128 (row pairs 16 p + 2 g, + 1),
//      2 B fragments as one LDS.128, 2 DMUL per 32 DMMA (each B fragment is scaled by exactly one warp)
//   2: the current 2 x 4 tiling without the DMULs (not a valid kernel: isolates what the scaling costs)
//   3: variant 1 without the DMULs
template <int KT, int VARIANT>
__global__ void __launch_bounds__(256, 1) gram_inner_probe_alt_kernel(double* __restrict__ out, int iters) {
    using namespace gk;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage<KT>& S = *reinterpret_cast<Stage<KT>*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(Stage<KT>) / sizeof(double)); i += 256)
        reinterpret_cast<double*>(&S)[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, kq = lane & 3;
    double sum = 0.0;
    if (VARIANT == 2) {
        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const int wm = warp >> 2, wn = warp & 3;
        const double* Ap = S.a + wm * 64 + g;
        const double* Bp = S.b + wn * 32 + g;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < KT / 4; ++kk) {
                const int kl = kk * 4 + kq;
                double a[8], b[4];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) a[mi] = Ap[kl * LDT + mi * 8];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = Bp[kl * LDT + ni * 8];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
            }
        }
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) sum += acc[mi][ni][0] + acc[mi][ni][1];
    } else {
        double acc[16][2][2];
#pragma unroll
        for (int mi = 0; mi < 16; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const double* Ap = S.a + 2 * g;
        const double* Bp = S.b + warp * 16 + 2 * g;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < KT / 4; ++kk) {
                const int kl = kk * 4 + kq;
                const double sk = S.s[kl];
                double a[16], b[2];
#pragma unroll
                for (int pr = 0; pr < 8; ++pr) {
                    const double2 v = *reinterpret_cast<const double2*>(Ap + kl * LDT + pr * 16);
                    a[2 * pr] = v.x;
                    a[2 * pr + 1] = v.y;
                }
                const double2 bv = *reinterpret_cast<const double2*>(Bp + kl * LDT);
                if (VARIANT == 1) {
                    b[0] = bv.x * sk;
                    b[1] = bv.y * sk;
                } else {
                    b[0] = bv.x;
                    b[1] = bv.y;
                }
#pragma unroll
                for (int mi = 0; mi < 16; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
            }
        }
#pragma unroll
        for (int mi = 0; mi < 16; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) sum += acc[mi][ni][0] + acc[mi][ni][1];
    }
    out[(int64_t)blockIdx.x * 256 + threadIdx.x] = sum;
}

int calib_gram_inner(blr_ctx* ctx, double* tflops) {
    constexpr int KT = 32;
    const int blocks = ctx->sm_count, iters = 2000;
    const int variant = getenv("BLR_PROBE_VARIANT") ? atoi(getenv("BLR_PROBE_VARIANT")) : 0;
    BLR_TRY(ensure_ws(ctx, (size_t)blocks * 256 * sizeof(double)));
    const int smem = (int)sizeof(gk::Stage<KT>);
    BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_inner_probe_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_inner_probe_alt_kernel<KT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_inner_probe_alt_kernel<KT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_inner_probe_alt_kernel<KT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    BLR_CUDA_OK(ctx, cudaEventCreate(&e0));
    BLR_CUDA_OK(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        BLR_CUDA_OK(ctx, cudaEventRecord(e0, ctx->stream));
        if (variant == 1) gram_inner_probe_alt_kernel<KT, 1><<<blocks, 256, smem, ctx->stream>>>(ctx->ws, iters);
        else if (variant == 2) gram_inner_probe_alt_kernel<KT, 2><<<blocks, 256, smem, ctx->stream>>>(ctx->ws, iters);
        else if (variant == 3) gram_inner_probe_alt_kernel<KT, 3><<<blocks, 256, smem, ctx->stream>>>(ctx->ws, iters);
        else gram_inner_probe_kernel<KT><<<blocks, 256, smem, ctx->stream>>>(ctx->ws, iters);
        ctx->launches++;
        BLR_CUDA_OK(ctx, cudaEventRecord(e1, ctx->stream));
        BLR_CUDA_OK(ctx, cudaEventSynchronize(e1));
        float ms;
        BLR_CUDA_OK(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    BLR_CUDA_OK(ctx, cudaGetLastError());
    // hardware flops of the DMMAs only: per block per iteration 128 x 128 x KT FMAs
    *tflops = (double)blocks * iters * 128.0 * 128.0 * KT * 2.0 / (best * 1e-3) / 1e12;
    return 0;
}

// ====================================================================== K1 generic path
// Any D, any leading dimension, either layout (element (d, n) at d*sd + n*sn).  32 x 32 tiles of the
// lower triangle, 2 x 2 outputs per thread, plain DFMA.  Used for small / odd shapes (README toy D = 2,
// the reference's D = 3..7 test problems); not a throughput path.
namespace gg {
constexpr int TS = 32, KC = 32, LDS_ = 33, THREADS = 256;
}

__global__ void __launch_bounds__(gg::THREADS) gram_generic_kernel(const double* __restrict__ X, int64_t sd, int64_t sn,
                                                                   int coalesce_n, int D, int64_t N,
                                                                   const double* __restrict__ s,
                                                                   const double* __restrict__ t, double* __restrict__ P,
                                                                   double* __restrict__ Pr, int nsplit, int64_t chunk) {
    using namespace gg;
    __shared__ double Xi[KC * LDS_], Xj[KC * LDS_], ss[KC], tt[KC];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int ti, tj;
    tile_from_index(blockIdx.x, ti, tj);
    const bool diag = (ti == tj);
    const int split = blockIdx.y;
    const int64_t n0 = (int64_t)split * chunk, n1 = min(N, n0 + chunk);
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0, racc = 0;
    for (int64_t k0 = n0; k0 < n1; k0 += KC) {
        __syncthreads();
        for (int e = tid; e < TS * KC; e += THREADS) {
            const int dl = coalesce_n ? e / KC : e % TS;
            const int k = coalesce_n ? e % KC : e / TS;
            const int64_t n = k0 + k;
            const int di = ti * TS + dl, dj = tj * TS + dl;
            Xi[k * LDS_ + dl] = (di < D && n < n1) ? X[(int64_t)di * sd + n * sn] : 0.0;
            Xj[k * LDS_ + dl] = (dj < D && n < n1) ? X[(int64_t)dj * sd + n * sn] : 0.0;
        }
        if (tid < KC) {
            const int64_t n = k0 + tid;
            ss[tid] = (n < n1) ? s[n] : 0.0;
            tt[tid] = (n < n1) ? t[n] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < KC; ++k) {
            const double a0 = Xi[k * LDS_ + ty * 2], a1 = Xi[k * LDS_ + ty * 2 + 1];
            const double sk = ss[k];
            const double b0 = Xj[k * LDS_ + tx * 2] * sk, b1 = Xj[k * LDS_ + tx * 2 + 1] * sk;
            c00 = fma(a0, b0, c00);
            c01 = fma(a0, b1, c01);
            c10 = fma(a1, b0, c10);
            c11 = fma(a1, b1, c11);
        }
        if (diag && tid < TS) {
            for (int k = 0; k < KC; ++k) racc = fma(Xi[k * LDS_ + tid], tt[k], racc);
        }
    }
    const int64_t slot = (int64_t)blockIdx.x * nsplit + split;
    double* Pt = P + slot * (TS * TS);
    Pt[(ty * 2) * TS + tx * 2] = c00;
    Pt[(ty * 2) * TS + tx * 2 + 1] = c01;
    Pt[(ty * 2 + 1) * TS + tx * 2] = c10;
    Pt[(ty * 2 + 1) * TS + tx * 2 + 1] = c11;
    if (diag && tid < TS) Pr[slot * TS + tid] = racc;
}

// ====================================================================== fixed-order partial reduction
// stats.G += Σ_slots P (lower tiles mirrored to both triangles), stats.r += Σ_slots Pr,
// stats.{q, ℓ, n} += prep partials.  One CTA per tile; the slots of a tile are visited in index order, so the
// result does not depend on timing.  tile_slot_begin == nullptr: uniform layout, slots [t * nsplit, (t+1) * nsplit).
__global__ void __launch_bounds__(256) gram_reduce_kernel(const double* __restrict__ P, const double* __restrict__ Pr,
                                                          int TS, const int* __restrict__ tile_slot_begin, int nsplit,
                                                          int D, double* __restrict__ G, double* __restrict__ r,
                                                          double* __restrict__ scal,
                                                          const double* __restrict__ prep_partial, int prep_blocks,
                                                          double n_obs, double gscale) {
    __shared__ double red[32];
    int ti, tj;
    tile_from_index(blockIdx.x, ti, tj);
    const int tsz = TS * TS;
    const int s0 = tile_slot_begin ? tile_slot_begin[blockIdx.x] : blockIdx.x * nsplit;
    const int s1 = tile_slot_begin ? tile_slot_begin[blockIdx.x + 1] : (blockIdx.x + 1) * nsplit;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < tsz; e += blockDim.x * gridDim.y) {
        const int row = e / TS, col = e % TS;
        const int gi = ti * TS + row, gj = tj * TS + col;
        // diagonal tiles of the fast path only hold sub-tiles whose 8-row block is not above their 8-col block
        if (gi < D && gj < D && gi >= gj) {
            double v = 0.0;
            for (int sl = s0; sl < s1; ++sl) v += P[(int64_t)sl * tsz + e];
            const double nv = G[(int64_t)gj * D + gi] + v * gscale;  // gscale = 1/σ² when the kernel ran unscaled, else 1
            G[(int64_t)gj * D + gi] = nv;
            if (gi != gj) G[(int64_t)gi * D + gj] = nv;
        }
    }
    if (ti == tj && blockIdx.y == 0) {
        for (int m = threadIdx.x; m < TS; m += blockDim.x) {
            const int gi = ti * TS + m;
            if (gi < D) {
                double v = 0.0;
                for (int sl = s0; sl < s1; ++sl) v += Pr[(int64_t)sl * TS + m];
                r[gi] += v;
            }
        }
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
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

// ====================================================================== host orchestration
static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// out (ldo x n, ldo >= D) = X[:, 0:n] with rows D .. ldo - 1 zeroed: one warp per observation, coalesced both ways
__global__ void __launch_bounds__(256) repack_colvecs_kernel(const double* __restrict__ X, int64_t ld, int D, int64_t n,
                                                             double* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); c < n; c += warps) {
        const double* src = X + c * ld;
        double* dst = out + c * ldo;
        for (int d = lane; d < ldo; d += 32) dst[d] = d < D ? src[d] : 0.0;
    }
}
int repack_colvecs(blr_ctx* ctx, const double* X, int64_t ld, int64_t D, int64_t n, double* out, int64_t ldo) {
    repack_colvecs_kernel<<<(int)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        X, ld, (int)D, n, out, ldo);
    BLR_CHECK_LAUNCH(ctx, "repack_colvecs_kernel");
    return 0;
}

int gram_accumulate(blr_ctx* ctx, blr_stats* st, const double* mw_dev, bool mw_is_zero, const blr_x* x,
                    const double* y, const double* sigma2, double sigma2_scalar, bool padded_odd) {
    const int64_t N = x->N;
    const int D = (int)x->D;
    if (N == 0) return 0;
    cudaStream_t sm = ctx->stream;
    // A ColVecs matrix that breaks the alignment rules of the bulk copies -- an odd number of features (a bias feature next to
    // 2^k learned ones), an odd leading dimension or a base that is not 16-byte aligned -- used to fall to the generic DFMA-fed
    // kernel (D = 127: 4.4 TF against 29 TF at D = 128).  Instead, blocks of observations are repacked into an aligned staging
    // buffer with an even leading dimension (one extra read + write of X at HBM speed: 93 / D of the Gram pass) and run through
    // the tiled TMA kernel; for odd D the kernel sees D + 1 features, the last one all zeros, and the reduction drops its row
    // and column.  The statistics are additive, so the staging buffer is bounded (1 GiB) whatever N.
    if (x->layout == BLR_COLVECS && D > 64 && !padded_odd &&
        ((D % 2) != 0 || (x->ld % 2) != 0 || (reinterpret_cast<uintptr_t>(x->p) & 15) != 0)) {
        const int64_t ldt = D + (D % 2);
        const int64_t block = std::min<int64_t>(N, std::max<int64_t>((int64_t)1 << 18, ((int64_t)1 << 27) / ldt));
        double* stage = nullptr;
        BLR_CUDA_OK(ctx, cudaMallocAsync(&stage, (size_t)ldt * block * sizeof(double), sm));
        int rc = 0;
        for (int64_t a = 0; a < N && rc == 0; a += block) {
            blr_x sub = *x;
            sub.p = stage;
            sub.N = std::min(block, N - a);
            sub.ld = ldt;
            sub.owned = false;
            rc = repack_colvecs(ctx, x->p + a * x->ld, x->ld, D, sub.N, stage, ldt);
            if (rc == 0)
                rc = gram_accumulate(ctx, st, mw_dev, mw_is_zero, &sub, y + a, sigma2 ? sigma2 + a : nullptr, sigma2_scalar,
                                     (D % 2) != 0);
        }
        cudaFreeAsync(stage, sm);
        return rc;
    }
    // A large RowVecs (feature-major) matrix is consumed in blocks of observations: each block is transposed into a
    // bounded ColVecs staging buffer and accumulated like any other chunk (the statistics are additive), so the extra
    // memory is 8 * D * ROWVECS_BLOCK bytes instead of a second copy of X.
    constexpr int64_t ROWVECS_BLOCK = 1 << 18;
    // (feature-major input that meets the TMA alignment rules needs no staging at all: see row_native below)
    const bool row_native = x->layout == BLR_ROWVECS && D > 64 && (N % 2) == 0 && (x->ld % 2) == 0 &&
                            (reinterpret_cast<uintptr_t>(x->p) & 15) == 0;
    if (x->layout == BLR_ROWVECS && !row_native && D >= 64 && (D % 2) == 0 && N > ROWVECS_BLOCK) {
        for (int64_t a = 0; a < N; a += ROWVECS_BLOCK) {
            blr_x sub = *x;
            sub.p = x->p + a;
            sub.N = std::min(ROWVECS_BLOCK, N - a);
            sub.owned = false;
            BLR_TRY(gram_accumulate(ctx, st, mw_dev, mw_is_zero, &sub, y + a, sigma2 ? sigma2 + a : nullptr, sigma2_scalar));
        }
        return 0;
    }
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], sm));

    // ---- 64 < D <= 96: two warps hold the Gram matrix together (gram_mid.cu), K0 fused in
    if (gram_mid_eligible(ctx, x, padded_odd)) {
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], sm));
        BLR_TRY(gram_mid(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, ctx->small + SMALL_PREP, padded_odd));
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], sm));
        ctx->ev_valid[0] = ctx->ev_valid[1] = ctx->ev_valid[2] = true;
        return 0;
    }

    // ---- smallest D: one warp holds the whole Gram matrix, observations stream straight from HBM, K0 fused in
    if (gram_small_fused(ctx, x)) {
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], sm));
        BLR_TRY(gram_small(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, nullptr, nullptr, ctx->small + SMALL_PREP, 0));
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], sm));
        ctx->ev_valid[0] = ctx->ev_valid[1] = ctx->ev_valid[2] = true;
        return 0;
    }

    // ---- K0: prep
    const int64_t npad = round_up(N, gk::KT_MAX) + gk::KT_MAX;
    BLR_TRY(ensure_nbuf(ctx, (size_t)(2 * npad) * sizeof(double)));
    double* s = ctx->nbuf;
    double* t = ctx->nbuf + npad;
    const int max_blocks = ctx->sm_count * 8;
    int prep_blocks;
    const bool warp_per_obs = (!mw_is_zero && x->layout == BLR_COLVECS && D > 64);  // must match prep_kernel's branch
    {
        const int64_t per_block = warp_per_obs ? PREP_THREADS / 32 : PREP_THREADS;
        prep_blocks = (int)std::min<int64_t>((npad + per_block - 1) / per_block, max_blocks);
    }
    double* prep_partial = ctx->small + SMALL_PREP;
    if (mw_is_zero) {
        prep_kernel<BLR_COLVECS, false><<<prep_blocks, PREP_THREADS, 0, sm>>>(x->p, x->ld, D, N, npad, y, sigma2,
                                                                              sigma2_scalar, mw_dev, s, t, prep_partial);
    } else if (x->layout == BLR_COLVECS) {
        prep_kernel<BLR_COLVECS, true><<<prep_blocks, PREP_THREADS, 0, sm>>>(x->p, x->ld, D, N, npad, y, sigma2,
                                                                             sigma2_scalar, mw_dev, s, t, prep_partial);
    } else {
        prep_kernel<BLR_ROWVECS, true><<<prep_blocks, PREP_THREADS, 0, sm>>>(x->p, x->ld, D, N, npad, y, sigma2,
                                                                             sigma2_scalar, mw_dev, s, t, prep_partial);
    }
    BLR_CHECK_LAUNCH(ctx, "prep_kernel");
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], sm));

    // ---- K1, small D: same streaming kernel, s / t from K0
    if (D <= 64) {
        BLR_TRY(gram_small(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, s, t, prep_partial, prep_blocks));
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], sm));
        ctx->ev_valid[0] = ctx->ev_valid[1] = ctx->ev_valid[2] = true;
        return 0;
    }

    // ---- K1
    const double* Xc = x->p;
    int64_t ldc = x->ld;
    double* xt = nullptr;  // transposed copy for a large RowVecs input
    bool fast = row_native || ((D >= 64) && (D % 2 == 0 || padded_odd));
    if (fast && x->layout == BLR_ROWVECS && !row_native) {
        const int64_t ldt = round_up(D, 2);
        BLR_CUDA_OK(ctx, cudaMallocAsync(&xt, (size_t)ldt * N * sizeof(double), sm));
        BLR_TRY(transpose_to_colvecs(ctx, x, xt, ldt));
        Xc = xt;
        ldc = ldt;
    }
    if (fast && ((ldc % 2) != 0 || (reinterpret_cast<uintptr_t>(Xc) & 15) != 0)) fast = false;
    const int KT = row_native ? 32 : ctx->gram_kt;  // the feature-major ring is built for 32-observation stages

    if (fast) {
        const int TS = gk::TM;
        const int nt = (D + TS - 1) / TS;
        const int64_t n_stages = (N + KT - 1) / KT;
        const int G = ctx->sm_count;
        // Periods are OPT-IN (BLR_GRAM_PERIOD_OBS): repeating the cut every ~32 MB of X keeps a period's panels
        // L2-resident (DRAM traffic ~1x instead of ~8x at D = 1024) but the lockstep costs 8-18 % throughput on B200
        // (measured), and HBM is nowhere near its limit -- so the default is one period.
        const int64_t period_obs = ctx->gram_period_obs > 0 ? ctx->gram_period_obs : INT64_MAX / 4;
        int64_t PS = std::max<int64_t>(1, period_obs / KT);
        if (n_stages <= PS + PS / 2) PS = n_stages;  // short inputs: one period (plain stream-K)
        const int64_t NP = (n_stages + PS - 1) / PS;
        const int64_t flush_cost = NP > 1 ? gk::W_OFF / 2 : 0;  // ~half a stage per tile switch
        const bool hybrid = KT == 32 && ctx->gram_cs && !row_native;
        const bool unit = hybrid && sigma2 == nullptr && ctx->gram_unit;  // Σy = σ² I
        const int diag_weight = ctx->diag_weight > 0 ? ctx->diag_weight : ((hybrid || row_native) ? 38 : 40);
        if (ctx->sched_key[0] != nt || ctx->sched_key[1] != PS || ctx->sched_key[2] != G ||
            ctx->sched_key[3] != diag_weight * 1000 + flush_cost) {
            Schedule sc;
            build_schedule(sc, nt, PS, G, diag_weight, flush_cost);
            const size_t bytes = sc.table.size() * sizeof(int);
            if (ctx->sched_bytes < bytes) {
                BLR_CUDA_OK(ctx, cudaStreamSynchronize(sm));
                if (ctx->sched) BLR_CUDA_OK(ctx, cudaFree(ctx->sched));
                ctx->sched = nullptr;
                ctx->sched_bytes = 0;
                BLR_CUDA_OK(ctx, cudaMalloc(&ctx->sched, bytes * 2));
                ctx->sched_bytes = bytes * 2;
            }
            BLR_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sched, sc.table.data(), bytes, cudaMemcpyHostToDevice, sm));
            BLR_CUDA_OK(ctx, cudaStreamSynchronize(sm));  // the host vector dies at the end of this scope
            ctx->sched_key[0] = nt;
            ctx->sched_key[1] = PS;
            ctx->sched_key[2] = G;
            ctx->sched_key[3] = diag_weight * 1000 + flush_cost;
            ctx->sched_T = sc.T;
            ctx->sched_nseg = sc.nseg;
        }
        const int T = ctx->sched_T, nseg = ctx->sched_nseg;
        const size_t p_elems = (size_t)nseg * TS * TS, pr_elems = (size_t)nseg * TS;
        BLR_TRY(ensure_ws(ctx, (p_elems + pr_elems) * sizeof(double)));
        GramParams gp;
        gp.X = Xc;
        gp.ld = ldc;
        gp.D = padded_odd ? D + 1 : D;  // staging buffer of an odd-D input: feature D is a row of zeros (same tile count)
        gp.N = N;
        gp.s = s;
        gp.t = t;
        gp.P = ctx->ws;
        gp.Pr = ctx->ws + p_elems;
        gp.cta_seg_begin = ctx->sched;
        const int* tile_slot_begin = ctx->sched + (G + 1);
        gp.seg_tile = tile_slot_begin + (T + 1);
        gp.seg_g0 = gp.seg_tile + nseg;
        gp.seg_g1 = gp.seg_g0 + nseg;
        gp.PS = (int)PS;
        gp.NP = (int)NP;
        gp.n_stages = (int)n_stages;
        gp.fix_bits = sched_fix_bits(PS);
        gp.period_counter = reinterpret_cast<unsigned int*>(ctx->d_flags + PERIOD_COUNTER_SLOT);
        BLR_CUDA_OK(ctx, cudaMemsetAsync(gp.period_counter, 0, sizeof(unsigned int), sm));
        void* args[] = {(void*)&gp};
        // cooperative launch: the soft barrier between periods needs every CTA resident (grid = #SMs, 1 CTA / SM)
        if (row_native) {
            using SM = gk::Smem<32, 3, true>;
            BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_tma_kernel<32, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)sizeof(SM)));
            BLR_CUDA_OK(ctx, cudaLaunchCooperativeKernel((void*)gram_tma_kernel<32, 3, true>, dim3(G), dim3(gk::THREADS_ROW),
                                                         args, sizeof(SM), sm));
        } else if (hybrid) {
            using SM = gk::Smem<32, 3>;
            // diagonal-tile code: partially unrolled when a large share of the SMs runs diagonal tiles (see consume_stage_rp);
            // homoscedastic noise: no per-observation scaling in the kernel at all, 1/σ² is applied by the reduction
            void* kern = unit ? (nt <= 4 ? (void*)gram_tma_kernel<32, 3, false, true, 4, true>
                                         : (void*)gram_tma_kernel<32, 3, false, true, 8, true>)
                              : (nt <= 4 ? (void*)gram_tma_kernel<32, 3, false, true, 4, false>
                                         : (void*)gram_tma_kernel<32, 3, false, true, 8, false>);
            BLR_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)));
            BLR_CUDA_OK(ctx, cudaLaunchCooperativeKernel(kern, dim3(G), dim3(gk::THREADS), args, sizeof(SM), sm));
        } else if (KT == 32) {
            using SM = gk::Smem<32, 3>;
            BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_tma_kernel<32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)sizeof(SM)));
            BLR_CUDA_OK(ctx, cudaLaunchCooperativeKernel((void*)gram_tma_kernel<32, 3>, dim3(G), dim3(gk::THREADS), args,
                                                         sizeof(SM), sm));
        } else if (ctx->gram_stages == 6) {
            using SM = gk::Smem<16, 6>;
            BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_tma_kernel<16, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)sizeof(SM)));
            BLR_CUDA_OK(ctx, cudaLaunchCooperativeKernel((void*)gram_tma_kernel<16, 6>, dim3(G), dim3(gk::THREADS), args,
                                                         sizeof(SM), sm));
        } else {
            using SM = gk::Smem<16, 4>;
            BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_tma_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)sizeof(SM)));
            BLR_CUDA_OK(ctx, cudaLaunchCooperativeKernel((void*)gram_tma_kernel<16, 4>, dim3(G), dim3(gk::THREADS), args,
                                                         sizeof(SM), sm));
        }
        BLR_CHECK_LAUNCH(ctx, "gram_tma_kernel");
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], sm));
        // the slot sum of one tile is spread over several CTAs (few tiles at small D, many slots per tile)
        const int ysplit = std::max(1, std::min(16, (2 * G) / T));
        gram_reduce_kernel<<<dim3(T, ysplit), 256, 0, sm>>>(gp.P, gp.Pr, TS, tile_slot_begin, 0, D, st->G(), st->r(), st->scal(),
                                              prep_partial, prep_blocks, (double)N, unit ? 1.0 / sigma2_scalar : 1.0);
        BLR_CHECK_LAUNCH(ctx, "gram_reduce_kernel");
    } else {
        const int TS = gg::TS;
        const int nt = (D + TS - 1) / TS, ntiles = nt * (nt + 1) / 2;
        int nsplit = std::max(1, (ctx->sm_count * 4) / ntiles);
        const int64_t chunk = round_up((N + nsplit - 1) / nsplit, gg::KC);
        nsplit = (int)((N + chunk - 1) / chunk);
        const size_t p_elems = (size_t)nsplit * ntiles * TS * TS, pr_elems = (size_t)nsplit * ntiles * TS;
        BLR_TRY(ensure_ws(ctx, (p_elems + pr_elems) * sizeof(double)));
        double* P = ctx->ws;
        double* Pr = ctx->ws + p_elems;
        const bool colv = (x->layout == BLR_COLVECS);
        const int64_t sd = colv ? 1 : x->ld, sn = colv ? x->ld : 1;
        gram_generic_kernel<<<dim3(ntiles, nsplit), gg::THREADS, 0, sm>>>(x->p, sd, sn, colv ? 0 : 1, D, N, s, t, P, Pr,
                                                                          nsplit, chunk);
        BLR_CHECK_LAUNCH(ctx, "gram_generic_kernel");
        BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], sm));
        gram_reduce_kernel<<<ntiles, 256, 0, sm>>>(P, Pr, TS, nullptr, nsplit, D, st->G(), st->r(), st->scal(),
                                                   prep_partial, prep_blocks, (double)N, 1.0);
        BLR_CHECK_LAUNCH(ctx, "gram_reduce_kernel");
    }
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], sm));
    ctx->ev_valid[0] = ctx->ev_valid[1] = ctx->ev_valid[2] = true;
    if (xt) BLR_CUDA_OK(ctx, cudaFreeAsync(xt, sm));
    return 0;
}

}  // namespace blr
