// Host-side stream-K schedule of the Gram fast path (pure C++, no CUDA): unit-tested on CPU by
// tests/test_schedule.py.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace blr {

constexpr int SCHED_W_OFF = 64;  // weight of one stage of an off-diagonal tile (== gk::W_OFF)

// Stream-K schedule for the fast path.  Layout of the int table: cta_seg_begin[G+1] | tile_slot_begin[T+1] |
// seg_tile[nseg] | seg_g0[nseg] | seg_g1[nseg].
// Segment boundaries are stage positions within one period in 16.16 FIXED POINT: in period p the kernel uses
// floor(boundary + θ_p) with a per-period dither θ_p in [0, 1) shared by all CTAs (θ_0 = 0), so neighbouring CTAs
// always agree on the cut, every stage is covered exactly once, and each CTA's stage count averages to its exact
// (fractional) share instead of being rounded the same way in every period.
inline int sched_fix_bits(int64_t n_stages) {  // fractional bits such that n_stages << bits fits an int
    int bits = 16;
    while (bits > 0 && (n_stages << bits) >= ((int64_t)1 << 30)) --bits;
    return bits;
}
struct Schedule {
    std::vector<int> table;
    int G = 0, T = 0, nseg = 0, fix_bits = 0;
};
// n_stages = stages of ONE period (the kernel repeats the same cut in every period of observations);
// flush_cost = weight charged per segment to CTAs that own more than one segment (they park their partial tile in
// the workspace at every segment switch); 0 = plain equal cut.
inline void build_schedule(Schedule& sc, int nt, int64_t n_stages, int G, int w_diag, int64_t flush_cost = 0) {
    const int T = nt * (nt + 1) / 2;
    std::vector<int64_t> start(T + 1, 0);
    std::vector<int> wt(T);
    for (int ti = 0, t = 0; ti < nt; ++ti)
        for (int tj = 0; tj <= ti; ++tj, ++t) {
            wt[t] = (ti == tj) ? w_diag : SCHED_W_OFF;
            start[t + 1] = start[t] + (int64_t)wt[t] * n_stages;
        }
    const int64_t W = start[T];
    const int SCHED_FIX = sched_fix_bits(n_stages);
    std::vector<int> cta_begin(G + 1, 0), tile_begin(T + 1, 0), seg_tile, seg_g0, seg_g1;
    auto bound = [&](int t, int64_t pos) -> int64_t {  // stage position (16.16) of weighted position pos inside tile t
        if (pos <= start[t]) return 0;
        if (pos >= start[t + 1]) return n_stages << SCHED_FIX;
        return ((pos - start[t]) << SCHED_FIX) / wt[t];
    };
    // CTA boundaries in weighted units
    std::vector<int64_t> cut(G + 1, 0);
    for (int k = 0; k <= G; ++k) cut[k] = (int64_t)((__int128)W * k / G);
    if (flush_cost > 0 && W > 0) {
        // smallest per-CTA budget tau such that G CTAs cover W when multi-segment CTAs pay flush_cost per segment
        auto walk = [&](int64_t tau, std::vector<int64_t>* out) -> bool {
            int64_t pos = 0;
            int u = 0;
            for (int k = 0; k < G; ++k) {
                if (out) (*out)[k] = pos;
                while (u < T && start[u + 1] <= pos) ++u;
                if (pos >= W) continue;
                if (tau <= start[u + 1] - pos) {
                    pos += tau;  // stays inside one tile: single segment, no flush
                } else {
                    int64_t budget = tau;
                    int v = u;
                    while (pos < W && budget > flush_cost) {
                        budget -= flush_cost;
                        const int64_t take = std::min(budget, start[v + 1] - pos);
                        pos += take;
                        budget -= take;
                        if (pos >= start[v + 1]) ++v;
                    }
                }
            }
            if (out) (*out)[G] = W;
            return pos >= W;
        };
        int64_t lo_t = W / G, hi_t = 2 * (W / G) + 2 * flush_cost * (T + 1) + 2;
        while (lo_t < hi_t) {
            const int64_t mid = lo_t + (hi_t - lo_t) / 2;
            if (walk(mid, nullptr)) hi_t = mid; else lo_t = mid + 1;
        }
        walk(lo_t, &cut);
        for (int k = 1; k <= G; ++k) cut[k] = std::min(std::max(cut[k], cut[k - 1]), W);
        cut[G] = W;
    }
    int t = 0;
    for (int k = 0; k < G; ++k) {
        const int64_t lo = cut[k], hi = cut[k + 1];
        cta_begin[k] = (int)seg_tile.size();
        while (t < T && start[t + 1] <= lo) ++t;
        for (int u = t; u < T && start[u] < hi; ++u) {
            const int64_t g0 = bound(u, lo), g1 = (k == G - 1) ? (n_stages << SCHED_FIX) : bound(u, hi);
            if ((g1 >> SCHED_FIX) > (g0 >> SCHED_FIX) || (g1 > g0 && flush_cost > 0)) {
                seg_tile.push_back(u);
                seg_g0.push_back((int)g0);
                seg_g1.push_back((int)g1);
            }
        }
    }
    cta_begin[G] = (int)seg_tile.size();
    const int nseg = (int)seg_tile.size();
    {
        int sidx = 0;
        for (int u = 0; u < T; ++u) {
            tile_begin[u] = sidx;
            while (sidx < nseg && seg_tile[sidx] == u) ++sidx;
        }
        tile_begin[T] = sidx;
    }
    sc.G = G;
    sc.T = T;
    sc.nseg = nseg;
    sc.fix_bits = SCHED_FIX;
    sc.table.clear();
    sc.table.insert(sc.table.end(), cta_begin.begin(), cta_begin.end());
    sc.table.insert(sc.table.end(), tile_begin.begin(), tile_begin.end());
    sc.table.insert(sc.table.end(), seg_tile.begin(), seg_tile.end());
    sc.table.insert(sc.table.end(), seg_g0.begin(), seg_g0.end());
    sc.table.insert(sc.table.end(), seg_g1.begin(), seg_g1.end());
}


}  // namespace blr
