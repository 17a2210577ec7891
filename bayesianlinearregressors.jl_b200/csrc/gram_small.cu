// K1 small-D path: the Gram statistics for D <= 64 features (README toy D = 2, the reference's D = 3..7 fixtures, and the
// "few features, very many observations" regime) as an HBM-bound streaming kernel.
//
// With D this small the whole D x D Gram matrix fits in the registers of ONE warp (lower-triangular 8 x 8 sub-tiles of a
// DMMA.8x8x4 accumulator grid), so every warp streams its own contiguous range of observations and no tile is shared:
// the m8n8k4 operand fragments are loaded straight from global memory -- lane (g, k) reads element (feature 8*mi + g,
// observation n + k), which for ColVecs is 8 consecutive doubles per observation and for RowVecs 4 consecutive doubles
// per feature, i.e. whole 32-byte sectors either way -- and the B fragment is the A fragment times s_k, so X is read
// exactly once and nothing is staged in shared memory.  The preparation pass (s = 1/σ², t = s δ, q, ℓ) is fused in.  Works for any D <= 64, any leading dimension / alignment and
// both layouts (plain 8-byte loads).  Per-CTA partial matrices are summed in a fixed order by gram_small_reduce_kernel.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "internal.h"

namespace blr {
namespace gs {
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int ctas_per_sm(int MI) { return MI == 1 ? 3 : (MI == 8 ? 1 : 2); }  // register budget: 85 / 128 / 255 per thread
}  // namespace gs

// K0 is fused in: the kernel reads y and σ² itself.  Per trip a warp takes 32 observations; lane L first handles
// observation n + L on its own (coalesced loads of y and σ², s = 1/σ², log σ² -- one divide and one log per observation,
// not per fragment lane), then the eight k4 steps fetch s and y for their fragment lanes by shuffle.  δ = y - x'mw is
// formed from the fragments already in registers (three xor-shuffles over the feature lanes) when the prior mean is
// non-zero.  So the whole path reads 8 (D + 2) bytes per observation -- the algorithmic minimum -- and writes nothing
// but the per-CTA partials.
// FUSED == false (D > 16, where the extra shuffles cost more than the 32 bytes per observation they save): s and t come
// from the separate preparation kernel in gram.cu, as do q and ℓ.
template <int MI, bool HAS_MEAN, bool FUSED>
__global__ void __launch_bounds__(gs::THREADS, gs::ctas_per_sm(MI))
    gram_small_kernel(const double* __restrict__ X, int64_t sd, int64_t sn, int D, int64_t N, const double* __restrict__ y,
                      const double* __restrict__ sigma2, double sigma2_scalar, const double* __restrict__ mw,
                      const double* __restrict__ s, const double* __restrict__ t, double* __restrict__ P,
                      double* __restrict__ Pr, double* __restrict__ Pq, int64_t obs_per_warp) {
    using namespace gs;
    constexpr int DP = MI * 8;
    __shared__ double tile[DP * DP];
    __shared__ double rsum[DP];
    __shared__ double qsum[WARPS][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, kq = lane & 3;
    const int64_t wid = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t n0 = wid * obs_per_warp, n1 = min(N, n0 + obs_per_warp);

    double acc[MI][MI][2];
    double racc[MI], mwr[MI];
    bool rowok[MI];
    const double* xbase = X + (int64_t)g * sd;
    const int64_t sd8 = 8 * sd;
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        racc[mi] = 0.0;
        rowok[mi] = (mi * 8 + g) < D;
        mwr[mi] = (HAS_MEAN && rowok[mi]) ? mw[mi * 8 + g] : 0.0;
#pragma unroll
        for (int ni = 0; ni < MI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }
    double qacc = 0.0, lacc = 0.0;
    if (FUSED) {
        // Software-pipelined over trips of 32 observations: the loads of trip i + 1 (lane-own y, σ² and the eight fragment
        // steps) are issued before trip i is reduced, so two trips per warp are in flight -- with only a few doubles per
        // observation the kernel is bound by bytes in flight (Little: ~6.5 MB for HBM at ~1 us), not by the tensor pipe.
        double a[8][MI], vo, yo;
        auto load_trip = [&](int64_t n, double (&aa)[8][MI], double& v, double& yy) {
            const int64_t no = n + lane;
            const bool oko = no < n1;
            v = oko ? (sigma2 ? sigma2[no] : sigma2_scalar) : 1.0;
            yy = oko ? y[no] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t nk = n + 4 * u + kq;
                const bool ok = nk < n1;
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) aa[u][mi] = (ok && rowok[mi]) ? xbase[mi * sd8 + nk * sn] : 0.0;
            }
        };
        load_trip(n0, a, vo, yo);
        for (int64_t n = n0; n < n1; n += 32) {
            double an[8][MI], vn, yn;
            load_trip(n + 32, an, vn, yn);  // fully predicated off past n1
            const double so = (n + lane < n1) ? 1.0 / vo : 0.0;
            lacc += log(vo);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int src = 4 * u + kq;
                const double sk = __shfl_sync(0xffffffffu, so, src);
                double dk = __shfl_sync(0xffffffffu, yo, src);
                if (HAS_MEAN) {
                    double dot = 0.0;
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) dot = fma(a[u][mi], mwr[mi], dot);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 8);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 16);
                    dk -= dot;
                }
                const double tk = sk * dk;
                qacc = fma(tk, dk, qacc);  // identical on the eight feature lanes of an observation; lanes g == 0 count
                double b[MI];
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) {
                    racc[mi] = fma(a[u][mi], tk, racc[mi]);
                    b[mi] = a[u][mi] * sk;
                }
#pragma unroll
                for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                    for (int ni = 0; ni <= mi; ++ni) dmma884(acc[mi][ni], a[u][mi], b[ni]);
            }
            vo = vn;
            yo = yn;
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) a[u][mi] = an[u][mi];
        }
    } else {
        // U k4-steps per batch, all loads issued before the first DMMA
        constexpr int U = 2;
        for (int64_t n = n0; n < n1; n += 32) {
#pragma unroll(MI == 8 ? 1 : 2)
            for (int u0 = 0; u0 < 8; u0 += U) {
                double a[U][MI], sk[U], tk[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t nk = n + 4 * (u0 + u) + kq;
                    const bool ok = nk < n1;
                    sk[u] = ok ? s[nk] : 0.0;
                    tk[u] = ok ? t[nk] : 0.0;
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) a[u][mi] = (ok && rowok[mi]) ? xbase[mi * sd8 + nk * sn] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double b[MI];
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) {
                        racc[mi] = fma(a[u][mi], tk[u], racc[mi]);
                        b[mi] = a[u][mi] * sk[u];
                    }
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                        for (int ni = 0; ni <= mi; ++ni) dmma884(acc[mi][ni], a[u][mi], b[ni]);
                }
            }
        }
    }
    // q: the four observation lanes with g == 0; ℓ: all 32 lanes (xor tree: fixed order)
    if (g != 0) qacc = 0.0;
    qacc = warp_sum(qacc);
    lacc = warp_sum(lacc);
    if (lane == 0) {
        qsum[warp][0] = qacc;
        qsum[warp][1] = lacc;
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
    if (FUSED && tid == 0) {
        double q = 0.0, l = 0.0;
        for (int w = 0; w < WARPS; ++w) {
            q += qsum[w][0];
            l += qsum[w][1];
        }
        Pq[2 * blockIdx.x] = q;
        Pq[2 * blockIdx.x + 1] = l;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 16 < D <= 64, D even, ColVecs with 16-byte aligned observations: the same warp-owns-the-Gram-matrix scheme, but the
// observations reach the operand fragments through a PER-WARP ring in shared memory fed by TMA bulk copies, so that every warp
// keeps three stages of observations in flight.  The register-fed kernel above has one batch of two k4-steps in flight per
// warp and only 8 warps per SM at D = 64 (the accumulators take the register file): it is latency-bound between the HBM and
// the tensor roofline (D = 64: 44 % of the DMMA peak, D = 32: 70 % of HBM).  The block grid is exactly ceil(D / 8) wide (the kernel
// above rounds 33..64 features up to 8 x 8 blocks: 36 block-MMAs per step at D = 40 instead of 15).
// One stage = 8 observations = two k4-steps; with ld == D a stage is ONE contiguous bulk copy (8 D doubles), otherwise one
// copy per observation.  The warp that consumes a slot is the warp that refills it (program order + __syncwarp), so the ring
// needs full barriers only.  Dense rows of D doubles give a 4-way conflict on the fragment loads (the four observations of a
// k4-step sit D doubles apart): 32 wavefronts per 36 block-MMAs at D = 64, far from binding.
namespace gr {
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int KO = 8;      // observations per stage
constexpr int STAGES = 4;  // per warp
constexpr int ctas_per_sm(int MI) { return MI <= 5 ? 2 : 1; }  // registers: <= 128 per thread up to D = 40; ring: 10 KB x D / 8 x ... per CTA
}  // namespace gr

// The preparation pass is fused in exactly as in the D <= 16 kernel: per trip of 32 observations (four stages) lane L handles
// observation n + L on its own (coalesced loads of y and σ², one divide and one log per observation, next trip prefetched),
// the fragment lanes of a k4-step fetch s and y by shuffle and, with a non-zero prior mean, form x'mw from the fragments they
// already hold (three xor-shuffles over the feature lanes).  No s / t arrays, no second pass over X for mw != 0.
template <int MI, bool HAS_MEAN>
__global__ void __launch_bounds__(gr::THREADS, gr::ctas_per_sm(MI))
    gram_small_ring_kernel(const double* __restrict__ X, int64_t ld, int D, int64_t N, const double* __restrict__ y,
                           const double* __restrict__ sigma2, double sigma2_scalar, const double* __restrict__ mw,
                           double* __restrict__ P, double* __restrict__ Pr, double* __restrict__ Pq, int64_t obs_per_warp) {
    using namespace gr;
    // D (even) features in DP = 8 MI >= D block rows.  A stage slot holds KO observations D doubles apart (dense, as they arrive) in
    // a buffer of KO * DP doubles: fragment rows >= D read the head of the next observation (finite values inside the slot) and
    // only ever feed accumulator rows / columns >= D, which are never written out.
    constexpr int DP = MI * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring_all = reinterpret_cast<double*>(smem_raw);                                    // [WARPS][STAGES][KO * DP]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ring_all + WARPS * STAGES * KO * DP);  // [WARPS][STAGES]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, kq = lane & 3;
    double* ring = ring_all + warp * (STAGES * KO * DP);
    const int64_t wid = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t n0 = min(N, wid * obs_per_warp), n1 = min(N, n0 + obs_per_warp);
    const int nstages = (int)((n1 - n0 + KO - 1) / KO);

    for (int e = tid; e < WARPS * STAGES * KO * DP; e += THREADS) ring_all[e] = 0.0;  // a partial last stage reads finite values
    if (lane == 0)
        for (int i = 0; i < STAGES; ++i) mbar_init(smem_u32(&bars[warp * STAGES + i]), 1);
    mbar_fence_init();
    fence_proxy_async();
    __syncthreads();

    const bool dense = (ld == D);
    auto issue = [&](int j) {  // stage j of this warp -> slot j % STAGES
        if (j >= nstages) return;
        const int64_t nb = n0 + (int64_t)j * KO;
        const int cnt = (int)min((int64_t)KO, n1 - nb);
        const uint32_t bar = smem_u32(&bars[warp * STAGES + j % STAGES]);
        double* dst = ring + (j % STAGES) * (KO * DP);
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)cnt * (uint32_t)D * 8u);
        __syncwarp();
        if (dense) {
            if (lane == 0) bulk_g2s(smem_u32(dst), X + nb * ld, (uint32_t)cnt * (uint32_t)D * 8u, bar);
        } else if (lane < cnt) {
            bulk_g2s(smem_u32(dst + lane * D), X + (nb + lane) * ld, (uint32_t)D * 8u, bar);
        }
    };

    double acc[MI][MI][2];
    double racc[MI], mwr[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        racc[mi] = 0.0;
        mwr[mi] = (HAS_MEAN && mi * 8 + g < D) ? mw[mi * 8 + g] : 0.0;  // rows >= D must not enter x'mw
#pragma unroll
        for (int ni = 0; ni < MI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }
    double qacc = 0.0, lacc = 0.0;
    // Σ log σ² without a log per observation (the scalar fp64 pipe is the tensor pipe): σ² = m 2^e with m in [1, 2); the lane keeps
    // the running product of the m's and the integer sum of the e's and takes ONE log per 512 observations (the product of 512
    // mantissas is < 2^512 and carries a relative error of ~512 eps = 6e-14, i.e. an absolute error of 6e-14 in its log).
    // Anything that is not a positive normal number (<= 0, NaN, inf, denormal) goes through log() itself, which also keeps the
    // "ℓ not finite => a noise variance is not positive" contract.
    double lmant = 1.0;
    int lexp = 0, lcnt = 0;
    auto log_acc = [&](double v) {
        const long long bits = __double_as_longlong(v);
        const int ef = (int)((bits >> 52) & 0x7ff);
        if (bits > 0 && ef != 0 && ef != 0x7ff) {
            lmant *= __longlong_as_double((bits & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
            lexp += ef - 1023;
            if (++lcnt == 512) {
                lacc += fma((double)lexp, 0.69314718055994530942, log(lmant));
                lmant = 1.0;
                lexp = 0;
                lcnt = 0;
            }
        } else {
            lacc += log(v);
        }
    };
#pragma unroll 1
    for (int j = 0; j < STAGES - 1; ++j) issue(j);
    // lane-own observation of the current trip (vo, yo -> so) and of the next one (vn, yn)
    double vo = 1.0, yo = 0.0, so = 0.0, vn, yn;
    auto load_vy = [&](int64_t nb, double& v, double& yy) {
        const int64_t no = nb + lane;
        const bool ok = no < n1;
        v = ok ? (sigma2 ? sigma2[no] : sigma2_scalar) : 1.0;
        yy = ok ? y[no] : 0.0;
    };
    load_vy(n0, vn, yn);
#pragma unroll 1
    for (int i = 0; i < nstages; ++i) {
        issue(i + STAGES - 1);  // refills the slot this warp finished reading in iteration i - 1
        if ((i & 3) == 0) {     // a new trip of 32 observations
            const int64_t nb32 = n0 + (int64_t)i * KO;
            vo = vn;
            yo = yn;
            so = (nb32 + lane < n1) ? __drcp_rn(vo) : 0.0;  // correctly rounded: == 1.0 / vo
            log_acc(vo);                                        // vo == 1 beyond the range
            load_vy(nb32 + 32, vn, yn);
        }
        // s and y of this stage's fragment lanes, fetched from the owning lanes BEFORE waiting for the stage (they do not depend
        // on its data; inside the k4-steps the shuffles sat on the critical path in front of the first block-MMA)
        double skv[2], ykv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int src = 8 * (i & 3) + 4 * u + kq;
            skv[u] = __shfl_sync(0xffffffffu, so, src);
            ykv[u] = __shfl_sync(0xffffffffu, yo, src);
        }
        mbar_wait(smem_u32(&bars[warp * STAGES + i % STAGES]), (uint32_t)(i / STAGES) & 1u);
        const double* st = ring + (i % STAGES) * (KO * DP) + g;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double a[MI], b[MI];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) a[mi] = st[(4 * u + kq) * D + mi * 8];
            const double sku = skv[u];
            double dk = ykv[u];
            if (HAS_MEAN) {
                double dot = 0.0;
#pragma unroll
                for (int mi = 0; mi < MI; ++mi) dot = fma(a[mi], mwr[mi], dot);
                dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                dot += __shfl_xor_sync(0xffffffffu, dot, 8);
                dot += __shfl_xor_sync(0xffffffffu, dot, 16);
                dk -= dot;
            }
            const double tku = sku * dk;
            qacc = fma(tku, dk, qacc);  // identical on the eight feature lanes of an observation; lanes g == 0 count
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                racc[mi] = fma(a[mi], tku, racc[mi]);
                b[mi] = a[mi] * sku;
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni <= mi; ++ni) dmma884(acc[mi][ni], a[mi], b[ni]);
        }
        __syncwarp();  // every lane is done with the slot before lane 0 refills it in the next iteration
    }
    // r: fold the four observation lanes of every feature
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        racc[mi] += __shfl_xor_sync(0xffffffffu, racc[mi], 1);
        racc[mi] += __shfl_xor_sync(0xffffffffu, racc[mi], 2);
    }
    // q: the four observation lanes with g == 0; ℓ: all 32 lanes (xor tree: fixed order)
    if (g != 0) qacc = 0.0;
    qacc = warp_sum(qacc);
    lacc += fma((double)lexp, 0.69314718055994530942, log(lmant));
    lacc = warp_sum(lacc);
    // CTA reduction in warp order (fixed => bit-reproducible); the tile aliases the (now idle) rings
    __syncthreads();
    double* tile = ring_all;
    double* rsum = ring_all + DP * DP;
    double* qsum = rsum + DP;  // [WARPS][2]
    for (int e = tid; e < DP * DP + DP; e += THREADS) tile[e] = 0.0;
    if (lane == 0) {
        qsum[2 * warp] = qacc;
        qsum[2 * warp + 1] = lacc;
    }
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
    if (tid == 0) {
        double q = 0.0, l = 0.0;
        for (int w = 0; w < WARPS; ++w) {
            q += qsum[2 * w];
            l += qsum[2 * w + 1];
        }
        Pq[2 * blockIdx.x] = q;
        Pq[2 * blockIdx.x + 1] = l;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// D <= 8: one THREAD per observation.  With so few features an m8n8k4 tile is mostly padding and the warp-per-4-
// observations schedule above is bound by instruction issue (~11 warp instructions per observation); here a warp
// instruction advances 32 observations: D (D + 1) / 2 + 2 D + 3 DFMAs, one divide and one log per observation per lane,
// statistics in registers, X / y / σ² read exactly once with fully coalesced loads (ColVecs with ld == D: 16-byte vector
// loads over a contiguous stream; any other layout: 8-byte loads, coalesced over observations for RowVecs).  K0 is fused.
// Per-CTA partials use the 8 x 8 layout of the kernel above, so gram_small_reduce_kernel finishes both.
namespace gt {
__host__ __device__ constexpr int threads(int DT) { return DT == 8 ? 128 : 256; }
constexpr int ctas_per_sm(int DT) { return DT == 2 ? 4 : 3; }  // register budget per thread: 64 (DT 2), 85 (4), 168 (8)
}  // namespace gt

template <int DT, bool HAS_MEAN, bool VEC>
__global__ void __launch_bounds__(gt::threads(DT), gt::ctas_per_sm(DT))
    gram_tiny_kernel(const double* __restrict__ X, int64_t sd, int64_t sn, int D, int64_t N, const double* __restrict__ y,
                     const double* __restrict__ sigma2, double sigma2_scalar, const double* __restrict__ mw,
                     double* __restrict__ P, double* __restrict__ Pr, double* __restrict__ Pq) {
    constexpr int THREADS = gt::threads(DT);
    constexpr int NG = DT * (DT + 1) / 2;
    __shared__ double red[THREADS / 32][NG + DT + 2];
    double G[NG], r[DT], mwr[DT], q = 0.0, l = 0.0;
#pragma unroll
    for (int e = 0; e < NG; ++e) G[e] = 0.0;
#pragma unroll
    for (int d = 0; d < DT; ++d) {
        r[d] = 0.0;
        mwr[d] = (HAS_MEAN && d < D) ? mw[d] : 0.0;
    }
    const int64_t stride = (int64_t)gridDim.x * THREADS;
#pragma unroll(DT == 8 ? 1 : 2)
    for (int64_t n = (int64_t)blockIdx.x * THREADS + threadIdx.x; n < N; n += stride) {
        double x[DT];
        if (VEC) {  // ColVecs, ld == D == DT, 16-byte aligned base: the observation is DT / 2 aligned double2
            const double2* col = reinterpret_cast<const double2*>(X + n * DT);
#pragma unroll
            for (int h = 0; h < DT / 2; ++h) {
                const double2 v2 = col[h];
                x[2 * h] = v2.x;
                x[2 * h + 1] = v2.y;
            }
        } else {
#pragma unroll
            for (int d = 0; d < DT; ++d) x[d] = (d < D) ? X[(int64_t)d * sd + n * sn] : 0.0;
        }
        const double v = sigma2 ? sigma2[n] : sigma2_scalar;
        double dl = y[n];
        const double sk = 1.0 / v;
        l += log(v);
        if (HAS_MEAN) {
#pragma unroll
            for (int d = 0; d < DT; ++d) dl = fma(-x[d], mwr[d], dl);
        }
        const double tk = sk * dl;
        q = fma(tk, dl, q);
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            r[i] = fma(x[i], tk, r[i]);
            const double xs = x[i] * sk;
#pragma unroll
            for (int j = 0; j <= i; ++j) G[i * (i + 1) / 2 + j] = fma(xs, x[j], G[i * (i + 1) / 2 + j]);
        }
    }
    // CTA reduction: xor tree inside the warp, then warps in index order (fixed => bit-reproducible)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < NG; ++e) {
        const double t = warp_sum(G[e]);
        if (lane == 0) red[warp][e] = t;
    }
#pragma unroll
    for (int d = 0; d < DT; ++d) {
        const double t = warp_sum(r[d]);
        if (lane == 0) red[warp][NG + d] = t;
    }
    q = warp_sum(q);
    l = warp_sum(l);
    if (lane == 0) {
        red[warp][NG + DT] = q;
        red[warp][NG + DT + 1] = l;
    }
    __syncthreads();
    if (threadIdx.x < NG + DT + 2) {
        double t = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) t += red[w][threadIdx.x];
        const int e = threadIdx.x;
        if (e < NG) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            const int j = e - i * (i + 1) / 2;
            P[(int64_t)blockIdx.x * 64 + i * 8 + j] = t;  // 8 x 8 tile, element (row i, col j <= i)
        } else if (e < NG + DT) {
            Pr[(int64_t)blockIdx.x * 8 + (e - NG)] = t;
        } else {
            Pq[2 * blockIdx.x + (e - NG - DT)] = t;
        }
    }
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

// host wrapper for the other translation units that produce per-CTA partials in this layout (gram_mid.cu)
int gram_small_reduce(blr_ctx* ctx, blr_stats* st, const double* P, const double* Pr, int DP, int nblocks, int D, const double* partial,
                      int partial_blocks, double n_obs) {
    gram_small_reduce_kernel<<<(DP * DP + 255) / 256, 256, 0, ctx->stream>>>(P, Pr, DP, nblocks, D, st->G(), st->r(), st->scal(), partial,
                                                                            partial_blocks, n_obs);
    BLR_CHECK_LAUNCH(ctx, "gram_small_reduce_kernel");
    return 0;
}

// FUSED: `partial` receives this launch's per-CTA (q, ℓ) partials; otherwise it holds the preparation kernel's
// `partial_blocks` partials and s, t are its outputs.
template <int MI, bool FUSED>
static int launch_small(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* y, const double* sigma2,
                        double sigma2_scalar, const double* mw_dev, bool mw_is_zero, const double* s, const double* t,
                        double* partial, int partial_blocks) {
    constexpr int DP = MI * 8;
    const int D = (int)x->D;
    const int64_t N = x->N;
    const int per_sm = gs::ctas_per_sm(MI);
    int nblocks = ctx->sm_count * per_sm;
    const int64_t groups = (N + 31) / 32;  // warp trips of 32 observations
    nblocks = (int)std::max<int64_t>(1, std::min<int64_t>(nblocks, (groups + gs::WARPS - 1) / gs::WARPS));
    const int64_t total_warps = (int64_t)nblocks * gs::WARPS;
    const int64_t obs_per_warp = ((groups + total_warps - 1) / total_warps) * 32;
    BLR_TRY(ensure_ws(ctx, (size_t)nblocks * (DP * DP + DP) * sizeof(double)));
    double* P = ctx->ws;
    double* Pr = ctx->ws + (size_t)nblocks * DP * DP;
    const bool colv = x->layout == BLR_COLVECS;
    const int64_t sd = colv ? 1 : x->ld, sn = colv ? x->ld : 1;
    if (FUSED && !mw_is_zero)
        gram_small_kernel<MI, true, FUSED><<<nblocks, gs::THREADS, 0, ctx->stream>>>(
            x->p, sd, sn, D, N, y, sigma2, sigma2_scalar, mw_dev, s, t, P, Pr, partial, obs_per_warp);
    else
        gram_small_kernel<MI, false, FUSED><<<nblocks, gs::THREADS, 0, ctx->stream>>>(
            x->p, sd, sn, D, N, y, sigma2, sigma2_scalar, mw_dev, s, t, P, Pr, partial, obs_per_warp);
    BLR_CHECK_LAUNCH(ctx, "gram_small_kernel");
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    gram_small_reduce_kernel<<<(DP * DP + 255) / 256, 256, 0, ctx->stream>>>(P, Pr, DP, nblocks, D, st->G(), st->r(), st->scal(),
                                                                            partial, FUSED ? nblocks : partial_blocks, (double)N);
    BLR_CHECK_LAUNCH(ctx, "gram_small_reduce_kernel");
    return 0;
}

template <int MI>
static int launch_ring(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* y, const double* sigma2, double sigma2_scalar,
                       const double* mw_dev, bool mw_is_zero, double* partial) {
    constexpr int DP = MI * 8;
    const int64_t N = x->N;
    const int smem = gr::WARPS * gr::STAGES * gr::KO * DP * (int)sizeof(double) + gr::WARPS * gr::STAGES * (int)sizeof(unsigned long long);
    const int64_t groups = (N + 31) / 32;
    const int nblocks = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * gr::ctas_per_sm(MI), (groups + gr::WARPS - 1) / gr::WARPS));
    const int64_t total_warps = (int64_t)nblocks * gr::WARPS;
    const int64_t obs_per_warp = ((groups + total_warps - 1) / total_warps) * 32;
    BLR_TRY(ensure_ws(ctx, (size_t)nblocks * (DP * DP + DP) * sizeof(double)));
    double* P = ctx->ws;
    double* Pr = ctx->ws + (size_t)nblocks * DP * DP;
    const int D = (int)x->D;
    if (mw_is_zero) {
        BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_small_ring_kernel<MI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        gram_small_ring_kernel<MI, false><<<nblocks, gr::THREADS, smem, ctx->stream>>>(x->p, x->ld, D, N, y, sigma2, sigma2_scalar, mw_dev, P,
                                                                                      Pr, partial, obs_per_warp);
    } else {
        BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_small_ring_kernel<MI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        gram_small_ring_kernel<MI, true><<<nblocks, gr::THREADS, smem, ctx->stream>>>(x->p, x->ld, D, N, y, sigma2, sigma2_scalar, mw_dev, P,
                                                                                     Pr, partial, obs_per_warp);
    }
    BLR_CHECK_LAUNCH(ctx, "gram_small_ring_kernel");
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    gram_small_reduce_kernel<<<(DP * DP + 255) / 256, 256, 0, ctx->stream>>>(P, Pr, DP, nblocks, D, st->G(), st->r(), st->scal(),
                                                                            partial, nblocks, (double)N);
    BLR_CHECK_LAUNCH(ctx, "gram_small_reduce_kernel");
    return 0;
}

// the ring kernel applies to: ColVecs, 16 < D <= 64 with D even, observations 16-byte aligned (BLR_SMALL_RING=0: off)
static bool ring_eligible(const blr_ctx* ctx, const blr_x* x) {
    return ctx->small_ring && x->layout == BLR_COLVECS && x->D > 16 && x->D <= 64 && (x->D % 2) == 0 && (x->ld % 2) == 0 &&
           (reinterpret_cast<uintptr_t>(x->p) & 15) == 0 && x->N >= 64;
}

// true: the small-D kernel forms s, t, q, ℓ itself (no separate preparation pass, `partial` receives 2 doubles per CTA)
bool gram_small_fused(const blr_ctx* ctx, const blr_x* x) { return x->D <= 16 || ring_eligible(ctx, x); }

template <int DT>
static int launch_tiny(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* y, const double* sigma2,
                       double sigma2_scalar, const double* mw_dev, bool mw_is_zero, double* partial) {
    const int D = (int)x->D;
    const int64_t N = x->N;
    constexpr int THREADS = gt::threads(DT);
    const int nblocks =
        (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * gt::ctas_per_sm(DT), (N + THREADS - 1) / THREADS));
    BLR_TRY(ensure_ws(ctx, (size_t)nblocks * (64 + 8) * sizeof(double)));
    double* P = ctx->ws;
    double* Pr = ctx->ws + (size_t)nblocks * 64;
    const bool colv = x->layout == BLR_COLVECS;
    const int64_t sd = colv ? 1 : x->ld, sn = colv ? x->ld : 1;
    const bool vec = colv && D == DT && x->ld == DT && (reinterpret_cast<uintptr_t>(x->p) & 15) == 0;
#define BLR_TINY(HM, VC)                                                                                              \
    gram_tiny_kernel<DT, HM, VC><<<nblocks, THREADS, 0, ctx->stream>>>(x->p, sd, sn, D, N, y, sigma2, sigma2_scalar, \
                                                                          mw_dev, P, Pr, partial)
    if (mw_is_zero) {
        if (vec) BLR_TINY(false, true);
        else BLR_TINY(false, false);
    } else {
        if (vec) BLR_TINY(true, true);
        else BLR_TINY(true, false);
    }
#undef BLR_TINY
    BLR_CHECK_LAUNCH(ctx, "gram_tiny_kernel");
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    gram_small_reduce_kernel<<<1, 256, 0, ctx->stream>>>(P, Pr, 8, nblocks, D, st->G(), st->r(), st->scal(), partial, nblocks,
                                                         (double)N);
    BLR_CHECK_LAUNCH(ctx, "gram_small_reduce_kernel");
    return 0;
}

int gram_small(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* y, const double* sigma2, double sigma2_scalar,
               const double* mw_dev, bool mw_is_zero, const double* s, const double* t, double* partial, int partial_blocks) {
    const int D = (int)x->D;
    if (D <= 2) return launch_tiny<2>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
    if (D <= 4) return launch_tiny<4>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
    if (D <= 8) return launch_tiny<8>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
    if (D <= 16)
        return launch_small<2, true>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, s, t, partial, partial_blocks);
    if (ring_eligible(ctx, x)) {
        switch ((D + 7) / 8) {
            case 3: return launch_ring<3>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
            case 4: return launch_ring<4>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
            case 5: return launch_ring<5>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
            case 6: return launch_ring<6>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
            case 7: return launch_ring<7>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
            default: return launch_ring<8>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial);
        }
    }
    if (D <= 32)
        return launch_small<4, false>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, s, t, partial, partial_blocks);
    return launch_small<8, false>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, s, t, partial, partial_blocks);
}

}  // namespace blr
