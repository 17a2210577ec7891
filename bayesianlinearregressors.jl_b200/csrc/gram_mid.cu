// K1m: the Gram statistics for 64 < D <= 96 (D even, ColVecs with 16-byte aligned observations).
//
// Between the warp-owns-the-matrix kernels (D <= 64, gram_small.cu) and the 128 x 128 tiles of K1 (gram.cu) lies a trough: K1
// pads such a problem to one 128-row diagonal tile and spends (128 / D)² of the useful tensor work (D = 66: 8 TF, D = 96: 16 TF),
// while the lower triangle no longer fits one warp's registers (78 sub-tiles of 8 x 8 at D = 96).  Here TWO warps -- a team --
// own the matrix together: the lower-triangular sub-tiles are dealt out so that both members carry the same number (rows in
// the boustrophedon order 0 | 1 2 | 3 0 ..., a leftover odd row split by column parity), and both consume the SAME ring of
// stages (8 observations each, dense as they arrive, TMA bulk copies) -- so X is still read exactly once.  Member 0 issues the
// copies; a slot is refilled once both members have released it (an `empty` mbarrier per slot, one arrive per member), three
// stages ahead of the consumers in a ring of five.  Four teams per CTA, one CTA per SM, every team streams its own contiguous
// range of observations.  The preparation (s = 1/σ², δ = y - x'mw, q, ℓ) is fused in exactly as in gram_small_ring_kernel: per
// trip of 32 observations lane L of each warp owns observation n + L, the fragment lanes fetch s and y by shuffle; q and ℓ are
// counted by member 0 only.  Per-CTA partial matrices are summed in a fixed order by gram_small_reduce_kernel.
// (reference: src/bayesian_linear_regression.jl:81-86 -- Bt'Bt and Bt'δy on the reduced statistics, see gram.cu.)
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "internal.h"
#include "mid_deal.h"

namespace blr {
namespace gm {
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int TEAM = 2;
constexpr int TEAMS = WARPS / TEAM;
constexpr int KO = 8;      // observations per stage
constexpr int STAGES = 5;  // per team
constexpr int AHEAD = 3;   // stages in flight: stage i + AHEAD is issued in iteration i, into the slot released in iteration i - 2

// the deal of the sub-tiles (block_owner, row_owner): mid_deal.h
}  // namespace gm

struct MidArgs {
    const double* X;
    int64_t ld;
    int D;   // features per observation as stored (even; D_logical + 1 with a zero feature for a staged odd-D input)
    int Dm;  // features the prior mean has
    const double* y;
    const double* sigma2;
    double sigma2_scalar;
    int64_t n0, n1;  // the team's observations
};

// the whole stream of one team member, then its share of the CTA's partial matrix (teams in fixed order)
template <int MI, bool HAS_MEAN, int MEMBER>
__device__ __forceinline__ void mid_member(const MidArgs& p, double* ring, unsigned long long* full, unsigned long long* empty,
                                           const double* mw, int lane, int team, double* tile, unsigned long long* done) {
    using namespace gm;
    constexpr int DP = MI * 8;
    const int g = lane >> 2, kq = lane & 3;
    // accumulators live HERE, per member: only the sub-tiles this member owns are ever touched, the others are dead code
    double acc[MI][MI][2];
    double racc[MI];
    const int D = p.D;
    const int64_t n0 = p.n0, n1 = p.n1;
    const int nstages = (int)((n1 - n0 + KO - 1) / KO);
    const bool dense = (p.ld == D);
    auto issue = [&](int j) {  // member 0 only: stage j of this team -> slot j % STAGES
        if (j >= nstages) return;
        const int slot = j % STAGES;
        if (j >= STAGES) mbar_wait(smem_u32(&empty[slot]), (uint32_t)((j - STAGES) / STAGES) & 1u);  // both members released it
        const int64_t nb = n0 + (int64_t)j * KO;
        const int cnt = (int)min((int64_t)KO, n1 - nb);
        const uint32_t bar = smem_u32(&full[slot]);
        double* dst = ring + slot * (KO * DP);
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)cnt * (uint32_t)D * 8u);
        __syncwarp();
        if (dense) {
            if (lane == 0) bulk_g2s(smem_u32(dst), p.X + nb * p.ld, (uint32_t)cnt * (uint32_t)D * 8u, bar);
        } else if (lane < cnt) {
            bulk_g2s(smem_u32(dst + lane * D), p.X + (nb + lane) * p.ld, (uint32_t)D * 8u, bar);
        }
    };
    double mwr[MI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
        racc[mi] = 0.0;
        mwr[mi] = (HAS_MEAN && mi * 8 + g < p.Dm) ? mw[mi * 8 + g] : 0.0;  // rows >= D must not enter x'mw
#pragma unroll
        for (int ni = 0; ni <= mi; ++ni)
            if (block_owner(mi, ni, MI) == MEMBER) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    }
    double qacc = 0.0, lacc = 0.0;
    // Σ log σ² as a running mantissa product and an integer exponent sum, one log per 512 observations (see gram_small_ring_kernel)
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
    if (MEMBER == 0) {
#pragma unroll 1
        for (int j = 0; j < AHEAD; ++j) issue(j);
    }
    double vo = 1.0, yo = 0.0, so = 0.0, vn, yn;
    auto load_vy = [&](int64_t nb, double& v, double& yy) {
        const int64_t no = nb + lane;
        const bool ok = no < n1;
        v = ok ? (p.sigma2 ? p.sigma2[no] : p.sigma2_scalar) : 1.0;
        yy = ok ? p.y[no] : 0.0;
    };
    load_vy(n0, vn, yn);
#pragma unroll 1
    for (int i = 0; i < nstages; ++i) {
        if (MEMBER == 0) issue(i + AHEAD);
        if ((i & 3) == 0) {  // a new trip of 32 observations
            const int64_t nb32 = n0 + (int64_t)i * KO;
            vo = vn;
            yo = yn;
            so = (nb32 + lane < n1) ? __drcp_rn(vo) : 0.0;  // correctly rounded: == 1.0 / vo
            if (MEMBER == 0) log_acc(vo);                   // vo == 1 beyond the range
            load_vy(nb32 + 32, vn, yn);
        }
        double skv[2], ykv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int src = 8 * (i & 3) + 4 * u + kq;
            skv[u] = __shfl_sync(0xffffffffu, so, src);
            ykv[u] = __shfl_sync(0xffffffffu, yo, src);
        }
        const int slot = i % STAGES;
        mbar_wait(smem_u32(&full[slot]), (uint32_t)(i / STAGES) & 1u);
        const double* st = ring + slot * (KO * DP) + g;
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
            if (MEMBER == 0) qacc = fma(tku, dk, qacc);  // identical on the eight feature lanes of an observation; lanes g == 0 count
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                if (row_owner(mi, MI) == MEMBER) racc[mi] = fma(a[mi], tku, racc[mi]);
                b[mi] = a[mi] * sku;
            }
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni <= mi; ++ni)
                    if (block_owner(mi, ni, MI) == MEMBER) dmma884(acc[mi][ni], a[mi], b[ni]);
        }
        ring_release(smem_u32(&empty[slot]), lane);  // this member is done with the slot
    }
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {  // r: fold the four observation lanes of every feature
        racc[mi] += __shfl_xor_sync(0xffffffffu, racc[mi], 1);
        racc[mi] += __shfl_xor_sync(0xffffffffu, racc[mi], 2);
    }
    if (g != 0) qacc = 0.0;
    qacc = warp_sum(qacc);
    lacc += fma((double)lexp, 0.69314718055994530942, log(lmant));
    lacc = warp_sum(lacc);

    // CTA reduction, teams in order (fixed => bit-reproducible); the members of a team own disjoint sub-tiles.  The tile has its
    // own shared memory (zeroed before the stream started), so no CTA-wide barrier is needed here -- this code runs in the
    // member-specific branch, where __syncthreads() would be divergent: team t waits for team t - 1 on an mbarrier that all
    // 64 threads of that team arrive on after their additions (release / acquire), then adds its own sub-tiles and arrives on its own.
    double* rsum = tile + DP * DP;
    double* qsum = rsum + DP;  // [TEAMS][2]
    if (team > 0) mbar_wait(smem_u32(&done[team - 1]), 0u);
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
#pragma unroll
        for (int ni = 0; ni <= mi; ++ni) {
            if (block_owner(mi, ni, MI) == MEMBER) {
                double* dst = &tile[(mi * 8 + g) * DP + ni * 8 + kq * 2];
                dst[0] += acc[mi][ni][0];
                dst[1] += acc[mi][ni][1];
            }
        }
        if (row_owner(mi, MI) == MEMBER && kq == 0) rsum[mi * 8 + g] += racc[mi];
    }
    if (MEMBER == 0 && lane == 0) {
        qsum[2 * team] = qacc;
        qsum[2 * team + 1] = lacc;
    }
    mbar_arrive(smem_u32(&done[team]));  // EVERY lane arrives (once per kernel: free) -- each thread's additions are ordered by its own arrive
}

template <int MI, bool HAS_MEAN>
__global__ void __launch_bounds__(gm::THREADS, 1)
    gram_mid_ring_kernel(const double* __restrict__ X, int64_t ld, int D, int Dm, int64_t N, const double* __restrict__ y,
                         const double* __restrict__ sigma2, double sigma2_scalar, const double* __restrict__ mw,
                         double* __restrict__ P, double* __restrict__ Pr, double* __restrict__ Pq, int64_t obs_per_team) {
    using namespace gm;
    constexpr int DP = MI * 8;
    constexpr int RING = TEAMS * STAGES * KO * DP;   // doubles
    constexpr int TILE = DP * DP + DP + 2 * TEAMS;   // doubles: the CTA's partial matrix, r and (q, ℓ) per team
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* buf = reinterpret_cast<double*>(smem_raw);
    double* tile = buf + RING;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(buf + RING + TILE);  // [TEAMS][2][STAGES] full, empty; [TEAMS] done
    unsigned long long* done = bars + TEAMS * 2 * STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int team = warp / TEAM, member = warp % TEAM;
    double* ring = buf + team * (STAGES * KO * DP);
    unsigned long long* full = bars + team * 2 * STAGES;
    unsigned long long* empty = full + STAGES;

    for (int e = tid; e < RING + TILE; e += THREADS) buf[e] = 0.0;  // rings: a partial last stage / rows >= D read finite values
    if (member == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&full[i]), 1);
            mbar_init(smem_u32(&empty[i]), TEAM * RING_LANES);
        }
        mbar_init(smem_u32(&done[team]), TEAM * 32);
    }
    mbar_fence_init();
    fence_proxy_async();
    __syncthreads();

    MidArgs p;
    p.X = X;
    p.ld = ld;
    p.D = D;
    p.Dm = Dm;
    p.y = y;
    p.sigma2 = sigma2;
    p.sigma2_scalar = sigma2_scalar;
    const int64_t tg = (int64_t)blockIdx.x * TEAMS + team;
    p.n0 = min(N, tg * obs_per_team);
    p.n1 = min(N, p.n0 + obs_per_team);

    if (member == 0) mid_member<MI, HAS_MEAN, 0>(p, ring, full, empty, mw, lane, team, tile, done);
    else mid_member<MI, HAS_MEAN, 1>(p, ring, full, empty, mw, lane, team, tile, done);
    __syncthreads();  // convergent again: every warp has added its sub-tiles

    const double* rsum = tile + DP * DP;
    const double* qsum = rsum + DP;
    double* Pt = P + (int64_t)blockIdx.x * (DP * DP);
    for (int e = tid; e < DP * DP; e += THREADS) Pt[e] = tile[e];
    if (tid < DP) Pr[(int64_t)blockIdx.x * DP + tid] = rsum[tid];
    if (tid == 0) {
        double q = 0.0, l = 0.0;
        for (int t = 0; t < TEAMS; ++t) {
            q += qsum[2 * t];
            l += qsum[2 * t + 1];
        }
        Pq[2 * blockIdx.x] = q;
        Pq[2 * blockIdx.x + 1] = l;
    }
}

template <int MI>
static int launch_mid(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* y, const double* sigma2, double sigma2_scalar,
                      const double* mw_dev, bool mw_is_zero, double* partial, bool padded_odd) {
    using namespace gm;
    constexpr int DP = MI * 8;
    constexpr int RING = TEAMS * STAGES * KO * DP, TILE = DP * DP + DP + 2 * TEAMS;
    const int smem = (RING + TILE) * (int)sizeof(double) + (TEAMS * 2 * STAGES + TEAMS) * (int)sizeof(unsigned long long);
    const int64_t N = x->N;
    const int64_t groups = (N + 31) / 32;  // trips of 32 observations
    const int nblocks = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, (groups + TEAMS - 1) / TEAMS));
    const int64_t total_teams = (int64_t)nblocks * TEAMS;
    const int64_t obs_per_team = ((groups + total_teams - 1) / total_teams) * 32;
    BLR_TRY(ensure_ws(ctx, (size_t)nblocks * (DP * DP + DP) * sizeof(double)));
    double* P = ctx->ws;
    double* Pr = ctx->ws + (size_t)nblocks * DP * DP;
    const int Dm = (int)x->D;                 // logical features (the statistics' dimension)
    const int D = Dm + (padded_odd ? 1 : 0);  // as stored: a staged odd-D input carries one zero feature
    if (mw_is_zero) {
        BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_mid_ring_kernel<MI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        gram_mid_ring_kernel<MI, false><<<nblocks, THREADS, smem, ctx->stream>>>(x->p, x->ld, D, Dm, N, y, sigma2, sigma2_scalar, mw_dev, P,
                                                                                Pr, partial, obs_per_team);
    } else {
        BLR_CUDA_OK(ctx, cudaFuncSetAttribute(gram_mid_ring_kernel<MI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        gram_mid_ring_kernel<MI, true><<<nblocks, THREADS, smem, ctx->stream>>>(x->p, x->ld, D, Dm, N, y, sigma2, sigma2_scalar, mw_dev, P,
                                                                               Pr, partial, obs_per_team);
    }
    BLR_CHECK_LAUNCH(ctx, "gram_mid_ring_kernel");
    BLR_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    return gram_small_reduce(ctx, st, P, Pr, DP, nblocks, Dm, partial, nblocks, (double)N);  // drops the zero feature's row / column
}

// ColVecs, 64 < D <= 96 with D even, observations 16-byte aligned (BLR_MID_RING=0: off -> one padded tile of K1).
// padded_odd: x is the aligned staging buffer of an odd-D input (gram.cu, repack_colvecs): x->D is the logical (odd) dimension,
// every observation carries D + 1 doubles, the last one zero.
bool gram_mid_eligible(const blr_ctx* ctx, const blr_x* x, bool padded_odd) {
    const int64_t Ds = x->D + (padded_odd ? 1 : 0);
    return ctx->mid_ring && x->layout == BLR_COLVECS && Ds > 64 && Ds <= 96 && (Ds % 2) == 0 && (x->ld % 2) == 0 &&
           (reinterpret_cast<uintptr_t>(x->p) & 15) == 0 && x->N >= 64;
}

int gram_mid(blr_ctx* ctx, blr_stats* st, const blr_x* x, const double* y, const double* sigma2, double sigma2_scalar,
             const double* mw_dev, bool mw_is_zero, double* partial, bool padded_odd) {
    switch ((x->D + (padded_odd ? 1 : 0) + 7) / 8) {
        case 9: return launch_mid<9>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial, padded_odd);
        case 10: return launch_mid<10>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial, padded_odd);
        case 11: return launch_mid<11>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial, padded_odd);
        default: return launch_mid<12>(ctx, st, x, y, sigma2, sigma2_scalar, mw_dev, mw_is_zero, partial, padded_odd);
    }
}

}  // namespace blr
