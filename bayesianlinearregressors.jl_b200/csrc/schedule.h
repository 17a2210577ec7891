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
struct Schedule {
    std::vector<int> table;
    int G = 0, T = 0, nseg = 0;
};
inline void build_schedule(Schedule& sc, int nt, int64_t n_stages, int G, int w_diag) {
    const int T = nt * (nt + 1) / 2;
    std::vector<int64_t> start(T + 1, 0);
    std::vector<int> wt(T);
    for (int ti = 0, t = 0; ti < nt; ++ti)
        for (int tj = 0; tj <= ti; ++tj, ++t) {
            wt[t] = (ti == tj) ? w_diag : SCHED_W_OFF;
            start[t + 1] = start[t] + (int64_t)wt[t] * n_stages;
        }
    const int64_t W = start[T];
    std::vector<int> cta_begin(G + 1, 0), tile_begin(T + 1, 0), seg_tile, seg_g0, seg_g1;
    auto bound = [&](int t, int64_t pos) -> int64_t {  // first stage of tile t at or after weighted position pos
        if (pos <= start[t]) return 0;
        const int64_t g = (pos - start[t] + wt[t] - 1) / wt[t];
        return std::min<int64_t>(g, n_stages);
    };
    int t = 0;
    for (int k = 0; k < G; ++k) {
        const int64_t lo = (int64_t)((__int128)W * k / G), hi = (int64_t)((__int128)W * (k + 1) / G);
        cta_begin[k] = (int)seg_tile.size();
        while (t < T && start[t + 1] <= lo) ++t;
        for (int u = t; u < T && start[u] < hi; ++u) {
            const int64_t g0 = bound(u, lo), g1 = (k == G - 1) ? n_stages : bound(u, hi);
            if (g1 > g0) {
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
    sc.table.clear();
    sc.table.insert(sc.table.end(), cta_begin.begin(), cta_begin.end());
    sc.table.insert(sc.table.end(), tile_begin.begin(), tile_begin.end());
    sc.table.insert(sc.table.end(), seg_tile.begin(), seg_tile.end());
    sc.table.insert(sc.table.end(), seg_g0.begin(), seg_g0.end());
    sc.table.insert(sc.table.end(), seg_g1.begin(), seg_g1.end());
}


}  // namespace blr
