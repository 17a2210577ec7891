// K3/K4 fused: the replicated D x D phase as ONE cooperative kernel -- tiled left-looking Cholesky, the forward solve as a
// border row, the backward solve, log-determinant, quadratic term and the posterior mean / log marginal likelihood.
//
// Reference work replaced (src/bayesian_linear_regression.jl): `_cholesky(Symmetric(Bt'Bt + I))` :86, `Λεy.U' \ (Bt'δy)` +
// `logdet(Λεy)` :57, `Λεy \ (Bt'δy)` :64, `Uw \ mεy` :68 -- in closed form on the reduced statistics (chol.cu header):
//     L = chol(Λ'),  z = L^-1 r,  u = L^-T z,  m' = mw + u,  logpdf = -1/2 [n log 2π + ℓ + q + logdet Λ' - logdet Λw - z'z].
//
// Why one kernel.  Round 1 ran a right-looking blocked Cholesky as two launches per 64-column panel plus six more launches
// for the solves (~40 dependent launches at D = 1024, 1.06 ms; it capped the 8-GPU scaling curve and cfg2).  The phase is
// bound by its dependency chain (D sequential pivots), not by flops, so the launches are replaced by a task graph executed
// by a persistent grid:
//   * the matrix is cut into 64 x 64 tiles; tile (i, j) of the lower triangle is ONE task:  accumulate
//     S = Σ_{k<j} L_ik L_jk' in fp64 tensor-core accumulators (DMMA.8x8x4, registers -- the tile is read once and written
//     once, no read-modify-write of global memory as in the right-looking form), then finish:  i == j: factor A_jj - S in
//     shared memory (register strips, cholblock.cuh);  i > j: (A_ij - S) L_jj^-T by substitution in registers;
//   * every task publishes a ready flag (release store); consumers spin on the flags of exactly the tiles they read
//     (acquire loads), so a tile's update starts the moment its inputs exist: panel look-ahead of any depth falls out of the
//     schedule, and the only work on the critical path per block column is one K = 64 tile product, one diagonal-block
//     factorisation and one 64-row substitution;
//   * the right-hand side r rides along as a border ROW of the factor (z_j = L_jj^-1 (r_j - Σ_k L_jk z_k): task "border j"),
//     the backward solve runs as tasks "bsolve j" (j descending) behind the last column, and the CTA that finishes bsolve 0
//     forms logdet, z'z, m' and the log marginal likelihood -- nothing of the phase is left for a second launch.
// Tasks are ordered so that every dependency has a smaller index (columns left to right, the backward solves last) and are
// dealt round-robin to a grid that is launched cooperatively (all CTAs co-resident), which makes the spin waits deadlock-free:
// the unfinished task with the smallest index always has its inputs complete and its CTA has nothing older to do.  A watchdog
// (~1 s of SM clocks) turns any violated assumption into an error code instead of a hung device.
#include <math.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "cholblock.cuh"
#include "common.cuh"
#include "internal.h"

namespace blr {
namespace tc {
constexpr int THREADS = 256;     // 8 warps: 4 (m) x 2 (n), warp tile 16 x 32
constexpr int KH = 32;           // k extent staged per pipeline step
constexpr int LDP = NB + 4;      // [k][m] staging rows: stride == 4 (mod 16) doubles, conflict-free m8n8k4 fragment reads
constexpr long long WATCHDOG_CLOCKS = 1ll << 31;

struct Smem {
    union {
        struct {
            double a[KH * LDP];  // rows of L_ik, [k][m]
            double b[KH * LDP];  // rows of L_jk, [k][n]
        } g;
        double Ls[NB * LDL];     // L_jj (substitution / solves), staged tiles of the backward solve
    } u;
    double T[NB * LDL];          // the task's own tile, T[c * LDL + r]
    double rdiag[NB];
    double xs[NB];
    double part[4][NB];
    double red[32];
    unsigned long long t_ready;  // trace: when the last wait of the current task returned
    int ready;
    int ok;
};
}  // namespace tc

struct TiledParams {
    double* A;          // D x D column-major, lower triangle in / factor out (strict upper zeroed)
    int64_t ld;
    int D, nb;
    int* flags;         // [nb (nb + 1) / 2 tiles | nb border | nb bsolve | abort]
    int epoch;
    int* info;          // [0] first non-positive pivot (1-based, 0 = none)  [3] watchdog / abort
    double* z;          // D: r in, z = L^-1 r out (nullptr: factor only)
    double* u;          // D: u = L^-T z out (nullptr: no backward solve)
    // finalize (needs z and u)
    const double* stat_scal;  // q, ℓ, n
    const double* mw;         // prior mean
    double* m_post;           // mw + u
    const double* logdet_w;   // logdet Λw (device scalar)
    double* sc;               // out: [1] logdet Λ'  [2] z'z  [3] logpdf
    int* noise_info;          // set to 1 when ℓ is not finite (a non-positive noise variance)
    unsigned long long* trace;  // debugging (BLR_DXD_TRACE): per task [kind/i/j, t_start, t_inputs_ready, t_end] in ns (globaltimer)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int tile_flag(int i, int j) { return i * (i + 1) / 2 + j; }

// Block until flags f0 and f1 (f1 < 0: none) carry this launch's epoch.  Returns false to every thread if the launch was
// aborted (watchdog expired here or in another CTA).
__device__ bool wait_flags(const TiledParams& p, tc::Smem& sm, int f0, int f1) {
    if (threadIdx.x == 0) {
        const int abort_slot = p.nb * (p.nb + 1) / 2 + 2 * p.nb;
        int ok = 1;
        const long long t0 = clock64();
#pragma unroll 1
        for (int pass = 0; pass < 2 && ok; ++pass) {
            const int f = pass ? f1 : f0;
            if (f < 0) continue;
            int spins = 0;
            while (ld_acquire_gpu(p.flags + f) != p.epoch) {
                if ((++spins & 63) == 0 && (ld_acquire_gpu(p.flags + abort_slot) == p.epoch || clock64() - t0 > tc::WATCHDOG_CLOCKS)) {
                    ok = 0;
                    break;
                }
            }
        }
        if (!ok) {
            st_release_gpu(p.flags + abort_slot, p.epoch);
            atomicExch(p.info + 3, 1);
        }
        if (p.trace) sm.t_ready = globaltimer_ns();
        sm.ok = ok;
    }
    __syncthreads();
    const bool ok = sm.ok != 0;
    __syncthreads();  // sm.ok may be rewritten by the next wait
    return ok;
}
__device__ __forceinline__ bool flags_ready(const TiledParams& p, int f0, int f1) {
    return ld_acquire_gpu(p.flags + f0) == p.epoch && (f1 < 0 || ld_acquire_gpu(p.flags + f1) == p.epoch);
}
// all stores of the CTA before this call become visible to whoever acquires the flag
__device__ __forceinline__ void publish(const TiledParams& p, int f) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(p.flags + f, p.epoch);
}

// L_jj (lower, identity-padded beyond `rows`) -> Ls[c * LDL + r], rdiag[k] = 1 / L[k][k]
__device__ void load_diag_block(const TiledParams& p, int j0, int rows, double* Ls, double* rdiag) {
    for (int e = threadIdx.x; e < NB * NB; e += tc::THREADS) {
        const int r = e & (NB - 1), c = e >> 6;
        double v = (r == c) ? 1.0 : 0.0;
        if (r < rows && c < rows && r >= c) v = __ldcg(p.A + (int64_t)(j0 + c) * p.ld + j0 + r);
        Ls[c * LDL + r] = v;
    }
    __syncthreads();
    if (threadIdx.x < NB) rdiag[threadIdx.x] = 1.0 / Ls[threadIdx.x * LDL + threadIdx.x];
    __syncthreads();
}

// Warp 0 solves the 64 x 64 triangular system held in Ls against xs (in place).  TRANS == false: L x = b (forward);
// TRANS == true: L' x = b (backward).  Lane l holds rows l and l + 32; the pivot value travels by shuffle.
template <bool TRANS>
__device__ void tri_solve64_warp0(const double* Ls, const double* rdiag, double* xs) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    double lo = xs[lane], hi = xs[lane + 32];
    if (!TRANS) {
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const double xc = __shfl_sync(0xffffffffu, c < 32 ? lo : hi, c & 31) * rdiag[c];
            if (c < 32) {
                if (lane > c) lo = fma(-xc, Ls[c * LDL + lane], lo);
                else if (lane == c) lo = xc;
                hi = fma(-xc, Ls[c * LDL + lane + 32], hi);
            } else {
                if (lane + 32 > c) hi = fma(-xc, Ls[c * LDL + lane + 32], hi);
                else if (lane + 32 == c) hi = xc;
            }
        }
    } else {
#pragma unroll
        for (int c = NB - 1; c >= 0; --c) {
            const double xc = __shfl_sync(0xffffffffu, c < 32 ? lo : hi, c & 31) * rdiag[c];
            // (L')[r][c] = L[c][r] = Ls[r * LDL + c], rows r < c
            if (c >= 32) {
                if (lane + 32 < c) hi = fma(-xc, Ls[(lane + 32) * LDL + c], hi);
                else if (lane + 32 == c) hi = xc;
                lo = fma(-xc, Ls[lane * LDL + c], lo);
            } else {
                if (lane < c) lo = fma(-xc, Ls[lane * LDL + c], lo);
                else if (lane == c) lo = xc;
            }
        }
    }
    xs[lane] = lo;
    xs[lane + 32] = hi;
}

// ---------------------------------------------------------------------------------------------------------------------
// Tile task (i, j), i >= j.
__device__ bool tile_task(const TiledParams& p, tc::Smem& sm, int i, int j) {
    using namespace tc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp & 3, wn = warp >> 2, g = lane >> 2, kq = lane & 3;
    const int i0 = i * NB, j0 = j * NB;
    const int rows_i = min(NB, p.D - i0), rows_j = min(NB, p.D - j0);
    const bool diag = (i == j);

    // the tile's own entries (still the input matrix: nobody else writes them) -> T; identity padding on a partial diagonal block
    for (int e = tid; e < NB * NB; e += THREADS) {
        const int r = e & (NB - 1), c = e >> 6;
        double v = (diag && r == c) ? 1.0 : 0.0;
        if (r < rows_i && c < rows_j) v = __ldcg(p.A + (int64_t)(j0 + c) * p.ld + i0 + r);
        sm.T[c * LDL + r] = v;
    }

    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // S = Σ_{k<j} L_ik L_jk'  -- software pipeline over (k, half): the next stage's global loads are in flight while the
    // current one is multiplied; a new k is fetched early only if its flags are already up (peek), otherwise after the math
    const int m = tid & (NB - 1), kb = tid >> 6;
    double ra[8], rb[8];
    auto fetch = [&](int step) {
        const int k0 = (step >> 1) * NB + (step & 1) * KH;
        const double* Ai = p.A + (int64_t)k0 * p.ld + i0 + m;
        const double* Aj = p.A + (int64_t)k0 * p.ld + j0 + m;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int64_t off = (int64_t)(kb + 4 * t) * p.ld;
            ra[t] = (m < rows_i) ? __ldcg(Ai + off) : 0.0;
            rb[t] = diag ? ra[t] : ((m < rows_j) ? __ldcg(Aj + off) : 0.0);
        }
    };
    const int nsteps = 2 * j;
    if (nsteps > 0) {
        if (!wait_flags(p, sm, tile_flag(i, 0), diag ? -1 : tile_flag(j, 0))) return false;
        fetch(0);
    }
    for (int step = 0; step < nsteps; ++step) {
        __syncthreads();  // the previous stage has been consumed
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            sm.u.g.a[(kb + 4 * t) * LDP + m] = ra[t];
            sm.u.g.b[(kb + 4 * t) * LDP + m] = rb[t];
        }
        const int nxt = step + 1;
        const bool new_k = nxt < nsteps && (nxt & 1) == 0;
        if (new_k && tid == 0) sm.ready = flags_ready(p, tile_flag(i, nxt >> 1), diag ? -1 : tile_flag(j, nxt >> 1)) ? 1 : 0;
        __syncthreads();
        bool fetched = false;
        if (nxt < nsteps && (!new_k || sm.ready)) {
            fetch(nxt);
            fetched = true;
        }
#pragma unroll
        for (int k4 = 0; k4 < KH; k4 += 4)
            warp_mma_k4<2, 4>(acc, sm.u.g.a + k4 * LDP + wm * 16, 1, LDP, sm.u.g.b + k4 * LDP + wn * 32, 1, LDP, lane);
        if (nxt < nsteps && !fetched) {
            if (!wait_flags(p, sm, tile_flag(i, nxt >> 1), diag ? -1 : tile_flag(j, nxt >> 1))) return false;
            fetch(nxt);
        }
    }
    __syncthreads();  // T is loaded, the staging buffers are free

    // T -= S
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int row = wm * 16 + mi * 8 + g, col = wn * 32 + ni * 8 + kq * 2 + c;
                sm.T[col * LDL + row] -= acc[mi][ni][c];
            }
    __syncthreads();

    if (diag) {
        const int fail = factor_block_smem(sm.T, sm.rdiag);
        for (int e = tid; e < NB * NB; e += THREADS) {
            const int r = e & (NB - 1), c = e >> 6;
            if (r < rows_j && c < rows_j) p.A[(int64_t)(j0 + c) * p.ld + j0 + r] = (r >= c) ? sm.T[c * LDL + r] : 0.0;
        }
        if (tid == 0 && fail != 0) {  // keep the smallest failing order (LAPACK's info), whatever order the tiles finish in
            const int val = j0 + fail;
            int old = atomicCAS(p.info, 0, val);
            while (old != 0 && old > val) {
                const int prev = atomicCAS(p.info, old, val);
                if (prev == old) break;
                old = prev;
            }
        }
        publish(p, tile_flag(j, j));
        return true;
    }

    // i > j:  X L_jj' = T  ->  row r of X by forward substitution.  Four threads share a row (columns c = 4 t + q, t = 0..15),
    // the solved entry of column k travels to the other three by shuffle: 16 registers per thread instead of 64.
    if (!wait_flags(p, sm, tile_flag(j, j), -1)) return false;
    load_diag_block(p, j0, NB, sm.u.Ls, sm.rdiag);  // an off-diagonal tile exists only under FULL diagonal blocks
    {
        const int r = tid >> 2, q = tid & 3;
        double a[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) a[t] = sm.T[(4 * t + q) * LDL + r];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const int tk = k >> 2, qk = k & 3;
            double xk = a[tk] * sm.rdiag[k];
            xk = __shfl_sync(0xffffffffu, xk, (lane & ~3) | qk);
            if (q == qk) a[tk] = xk;
            const double* Lk = sm.u.Ls + k * LDL + q;  // L_jj[c][k], c = 4 t + q
            if (q > qk) a[tk] = fma(-xk, Lk[4 * tk], a[tk]);
#pragma unroll
            for (int t = tk + 1; t < 16; ++t) a[t] = fma(-xk, Lk[4 * t], a[t]);
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) sm.T[(4 * t + q) * LDL + r] = a[t];
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += THREADS) {
        const int r = e & (NB - 1), c = e >> 6;
        if (r < rows_i) p.A[(int64_t)(j0 + c) * p.ld + i0 + r] = sm.T[c * LDL + r];
        if (c < rows_i) p.A[(int64_t)(i0 + c) * p.ld + j0 + r] = 0.0;  // mirror tile (j, i): the factor's strict upper part
    }
    publish(p, tile_flag(i, j));
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Border task j:  z_j = L_jj^-1 (r_j - Σ_{k<j} L_jk z_k)
__device__ bool border_task(const TiledParams& p, tc::Smem& sm, int j) {
    using namespace tc;
    const int tid = threadIdx.x, m = tid & (NB - 1), qd = tid >> 6;
    const int j0 = j * NB, rows_j = min(NB, p.D - j0);
    const int ntile = p.nb * (p.nb + 1) / 2;
    double acc = 0.0;
    for (int k = 0; k < j; ++k) {
        if (!wait_flags(p, sm, tile_flag(j, k), ntile + k)) return false;
        if (tid < NB) sm.xs[tid] = __ldcg(p.z + k * NB + tid);
        __syncthreads();
        if (m < rows_j) {
            const double* Lp = p.A + (int64_t)(k * NB + qd * 16) * p.ld + j0 + m;
#pragma unroll
            for (int t = 0; t < 16; ++t) acc = fma(__ldcg(Lp + (int64_t)t * p.ld), sm.xs[qd * 16 + t], acc);
        }
        __syncthreads();
    }
    sm.part[qd][m] = acc;
    if (!wait_flags(p, sm, tile_flag(j, j), -1)) return false;
    load_diag_block(p, j0, rows_j, sm.u.Ls, sm.rdiag);
    if (tid < NB) {
        const double rj = (tid < rows_j) ? __ldcg(p.z + j0 + tid) : 0.0;
        sm.xs[tid] = rj - (sm.part[0][tid] + sm.part[1][tid] + sm.part[2][tid] + sm.part[3][tid]);
    }
    __syncthreads();
    tri_solve64_warp0<false>(sm.u.Ls, sm.rdiag, sm.xs);
    __syncthreads();
    if (tid < rows_j) p.z[j0 + tid] = sm.xs[tid];
    publish(p, ntile + j);
    return true;
}

// Backward-solve task j:  u_j = L_jj^-T (z_j - Σ_{i>j} L_ij' u_i).  Everything that does not depend on the u's is staged
// before waiting for them (L_jj and its reciprocal diagonal in T; each tile L_ij before its u_i), so that the chain
// u_{nb-1} -> ... -> u_0 pays per block one 64-vector load, 16 FMAs per thread and one 64 x 64 substitution.
__device__ bool bsolve_task(const TiledParams& p, tc::Smem& sm, int j) {
    using namespace tc;
    const int tid = threadIdx.x, c = tid & (NB - 1), rq = tid >> 6;
    const int j0 = j * NB, rows_j = min(NB, p.D - j0);
    const int ntile = p.nb * (p.nb + 1) / 2;
    if (!wait_flags(p, sm, tile_flag(j, j), ntile + j)) return false;  // L_jj and z_j (both final long before u_{j+1})
    load_diag_block(p, j0, rows_j, sm.T, sm.rdiag);
    const double zj = (tid < rows_j) ? __ldcg(p.z + j0 + tid) : 0.0;
    double acc = 0.0;
    for (int i = p.nb - 1; i > j; --i) {
        const int i0 = i * NB, rows_i = min(NB, p.D - i0);
        if (!wait_flags(p, sm, tile_flag(i, j), -1)) return false;
        for (int e = tid; e < NB * NB; e += THREADS) {  // stage L_ij [cc][rr], coalesced over rr
            const int rr = e & (NB - 1), cc = e >> 6;
            sm.u.Ls[cc * LDL + rr] = (rr < rows_i) ? __ldcg(p.A + (int64_t)(j0 + cc) * p.ld + i0 + rr) : 0.0;
        }
        if (!wait_flags(p, sm, ntile + p.nb + i, -1)) return false;
        if (tid < NB) sm.xs[tid] = (tid < rows_i) ? __ldcg(p.u + i0 + tid) : 0.0;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 16; ++t) acc = fma(sm.u.Ls[c * LDL + rq * 16 + t], sm.xs[rq * 16 + t], acc);
        __syncthreads();
    }
    sm.part[rq][c] = acc;
    __syncthreads();
    if (tid < NB) sm.xs[tid] = zj - (sm.part[0][tid] + sm.part[1][tid] + sm.part[2][tid] + sm.part[3][tid]);
    __syncthreads();
    tri_solve64_warp0<true>(sm.T, sm.rdiag, sm.xs);
    __syncthreads();
    if (tid < rows_j) p.u[j0 + tid] = sm.xs[tid];
    publish(p, ntile + p.nb + j);
    return true;
}

// The CTA that finished bsolve 0: logdet Λ' = 2 Σ log L_ii, z'z, m' = mw + u, logpdf.
__device__ void finalize_task(const TiledParams& p, tc::Smem& sm) {
    using namespace tc;
    const int tid = threadIdx.x;
    const int ntile = p.nb * (p.nb + 1) / 2;
    // bsolve 0 is the end of every dependency chain; acquire the remaining flags explicitly so that all of z, u, diag(L)
    // is visible to this CTA by the letter of the memory model
    for (int k = 0; k < p.nb; ++k)
        if (!wait_flags(p, sm, ntile + k, ntile + p.nb + k)) return;
    double ld = 0.0, zz = 0.0;
    for (int i = tid; i < p.D; i += THREADS) {
        ld += log(__ldcg(p.A + (int64_t)i * p.ld + i));
        const double zi = __ldcg(p.z + i);
        zz = fma(zi, zi, zz);
        p.m_post[i] = p.mw[i] + __ldcg(p.u + i);
    }
    ld = block_sum(ld, sm.red);
    zz = block_sum(zz, sm.red);
    if (tid == 0) {
        const double LOG2PI = 1.8378770664093454835606594728112;
        const double q = p.stat_scal[0], l = p.stat_scal[1], n = p.stat_scal[2];
        p.sc[1] = 2.0 * ld;
        p.sc[2] = zz;
        p.sc[3] = -0.5 * (n * LOG2PI + l + q + (2.0 * ld - p.logdet_w[0]) - zz);
        if (!isfinite(l)) *p.noise_info = 1;  // Σ log σ²: some variance is <= 0 (-inf / NaN) or not finite
    }
}

__global__ void __launch_bounds__(tc::THREADS, 2) dxd_fused_kernel(const TiledParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tc::Smem& sm = *reinterpret_cast<tc::Smem*>(smem_raw);
    const int nb = p.nb;
    const int per_col_extra = p.z ? 1 : 0;
    const int ncol_tasks = nb * (nb + 1) / 2 + per_col_extra * nb;
    const int ntasks = ncol_tasks + (p.u ? nb : 0);
    for (int t = blockIdx.x; t < ntasks; t += gridDim.x) {
        bool ok;
        unsigned long long t_start = 0;
        int kind, ti, tj;
        if (p.trace && threadIdx.x == 0) {
            t_start = globaltimer_ns();
            sm.t_ready = t_start;
        }
        if (t < ncol_tasks) {
            // column j holds (nb - j) tile tasks followed by its border task; off(j) = j nb - j (j - 1) / 2 + e j
            int j = 0;
            while (j + 1 < nb && (j + 1) * nb - (j + 1) * j / 2 + per_col_extra * (j + 1) <= t) ++j;
            const int within = t - (j * nb - j * (j - 1) / 2 + per_col_extra * j);
            tj = j;
            if (within < nb - j) {
                ti = j + within;
                kind = (ti == j) ? 0 : 1;
                ok = tile_task(p, sm, ti, j);
            } else {
                ti = nb;
                kind = 2;
                ok = border_task(p, sm, j);
            }
        } else {
            const int j = nb - 1 - (t - ncol_tasks);
            ti = tj = j;
            kind = 3;
            ok = bsolve_task(p, sm, j);
            if (ok && j == 0 && p.m_post) finalize_task(p, sm);
        }
        if (!ok) return;
        __syncthreads();
        if (p.trace && threadIdx.x == 0) {
            unsigned long long* tr = p.trace + 4 * (size_t)t;
            tr[0] = ((unsigned long long)kind << 32) | ((unsigned long long)ti << 16) | (unsigned long long)tj;
            tr[1] = t_start;
            tr[2] = sm.t_ready;
            tr[3] = globaltimer_ns();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side
int dxd_fused(blr_ctx* ctx, double* A, int64_t D64, int* info_dev, double* z, double* u, const DxdFinalize* fin) {
    const int D = (int)D64, nb = (D + NB - 1) / NB;
    cudaStream_t sm = ctx->stream;
    const int ntile = nb * (nb + 1) / 2;
    const size_t nflags = (size_t)ntile + 2 * nb + 1;
    if (ctx->tflags_n < nflags) {
        BLR_CUDA_OK(ctx, cudaStreamSynchronize(sm));
        if (ctx->tflags) BLR_CUDA_OK(ctx, cudaFree(ctx->tflags));
        ctx->tflags = nullptr;
        ctx->tflags_n = 0;
        const size_t cap = std::max<size_t>(nflags, 4096);
        BLR_CUDA_OK(ctx, cudaMalloc(&ctx->tflags, cap * sizeof(int)));
        BLR_CUDA_OK(ctx, cudaMemset(ctx->tflags, 0, cap * sizeof(int)));
        ctx->tflags_n = cap;
    }
    const int smem = (int)sizeof(tc::Smem);
    if (ctx->dxd_occ == 0) {
        BLR_CUDA_OK(ctx, cudaFuncSetAttribute(dxd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int occ = 0;
        BLR_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dxd_fused_kernel, tc::THREADS, smem));
        if (occ < 1) return set_err(ctx, BLR_E_CUDA, "dxd_fused_kernel does not fit on an SM");
        ctx->dxd_occ = occ;
    }
    TiledParams p;
    p.A = A;
    p.ld = D64;
    p.D = D;
    p.nb = nb;
    p.flags = ctx->tflags;
    p.epoch = ++ctx->flag_epoch;
    p.info = info_dev;
    p.z = z;
    p.u = (z != nullptr) ? u : nullptr;
    p.stat_scal = fin ? fin->stat_scal : nullptr;
    p.mw = fin ? fin->mw : nullptr;
    p.m_post = (fin && p.u) ? fin->m_post : nullptr;
    p.sc = fin ? fin->sc : nullptr;
    p.logdet_w = fin ? fin->logdet_w : nullptr;
    p.noise_info = info_dev + 1;
    p.trace = nullptr;
    const int ntasks = ntile + (p.z ? nb : 0) + (p.u ? nb : 0);
    const char* trace_path = getenv("BLR_DXD_TRACE");
    if (trace_path) {
        BLR_CUDA_OK(ctx, cudaMalloc(&p.trace, (size_t)ntasks * 4 * sizeof(unsigned long long)));
        BLR_CUDA_OK(ctx, cudaMemsetAsync(p.trace, 0, (size_t)ntasks * 4 * sizeof(unsigned long long), sm));
    }
    BLR_CUDA_OK(ctx, cudaMemsetAsync(info_dev, 0, 4 * sizeof(int), sm));
    const int grid = std::min(ntasks, ctx->dxd_occ * ctx->sm_count);
    void* args[] = {(void*)&p};
    BLR_CUDA_OK(ctx, cudaLaunchCooperativeKernel((void*)dxd_fused_kernel, dim3(grid), dim3(tc::THREADS), args, smem, sm));
    BLR_CHECK_LAUNCH(ctx, "dxd_fused_kernel");
    if (trace_path) {  // debugging aid: dump the task timeline (kind: 0 potrf, 1 trsm, 2 border, 3 bsolve)
        std::vector<unsigned long long> h((size_t)ntasks * 4);
        BLR_CUDA_OK(ctx, cudaStreamSynchronize(sm));
        BLR_CUDA_OK(ctx, cudaMemcpy(h.data(), p.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(p.trace);
        if (FILE* fp = fopen(trace_path, "a")) {
            unsigned long long t0 = ~0ull;
            for (int t = 0; t < ntasks; ++t)
                if (h[4 * t + 1] && h[4 * t + 1] < t0) t0 = h[4 * t + 1];
            fprintf(fp, "# D=%d nb=%d grid=%d ntasks=%d: task kind i j start_us ready_us end_us\n", D, nb, grid, ntasks);
            for (int t = 0; t < ntasks; ++t)
                fprintf(fp, "%d %d %d %d %.2f %.2f %.2f\n", t, (int)(h[4 * t] >> 32), (int)((h[4 * t] >> 16) & 0xffff),
                        (int)(h[4 * t] & 0xffff), (h[4 * t + 1] - t0) * 1e-3, (h[4 * t + 2] - t0) * 1e-3, (h[4 * t + 3] - t0) * 1e-3);
            fclose(fp);
        }
    }
    return 0;
}

}  // namespace blr
