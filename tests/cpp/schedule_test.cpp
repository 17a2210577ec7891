// CPU unit test of the stream-K schedule (bayesianlinearregressors.jl_b200/csrc/schedule.h).
// usage: schedule_test nt n_stages G w_diag  -> prints "ok <nseg> <max_w> <min_w>" or "FAIL <why>"
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "schedule.h"

int main(int argc, char** argv) {
    if (argc != 5 && argc != 6) return 2;
    const long long flush = argc == 6 ? atoll(argv[5]) : 0;
    const int nt = atoi(argv[1]);
    const long long n_stages = atoll(argv[2]);
    const int G = atoi(argv[3]), wd = atoi(argv[4]);
    blr::Schedule sc;
    blr::build_schedule(sc, nt, n_stages, G, wd, flush);
    const int T = nt * (nt + 1) / 2;
    if (sc.T != T || sc.G != G) { printf("FAIL meta\n"); return 1; }
    const int* cta = sc.table.data();
    const int* tb = cta + G + 1;
    const int* st = tb + T + 1;
    const int* g0f = st + sc.nseg;
    const int* g1f = g0f + sc.nseg;
    // boundaries are 16.16 fixed point; check coverage for the un-dithered period (theta = 0) and a dithered one
    const unsigned theta = argc == 6 ? (0x9e37u >> (16 - sc.fix_bits)) : 0u;
    std::vector<int> g0(sc.nseg), g1(sc.nseg);
    for (int s = 0; s < sc.nseg; ++s) {
        g0[s] = (int)(((long long)g0f[s] + theta) >> sc.fix_bits);
        g1[s] = (int)(((long long)g1f[s] + theta) >> sc.fix_bits);
        if (g1[s] > n_stages) g1[s] = (int)n_stages;
    }
    if ((int)sc.table.size() != G + 1 + T + 1 + 3 * sc.nseg) { printf("FAIL size\n"); return 1; }
    if (cta[0] != 0 || cta[G] != sc.nseg || tb[0] != 0 || tb[T] != sc.nseg) { printf("FAIL ends\n"); return 1; }
    // every (tile, stage) covered exactly once, segments of a tile contiguous + ordered
    std::vector<long long> next(T, 0);
    for (int s = 0; s < sc.nseg; ++s) {
        const int t = st[s];
        if (t < 0 || t >= T || g1[s] < g0[s]) { printf("FAIL seg %d\n", s); return 1; }
        if (s < tb[t] || s >= tb[t + 1]) { printf("FAIL tile range %d\n", s); return 1; }
        if (g0[s] != next[t]) { printf("FAIL gap tile %d at seg %d\n", t, s); return 1; }
        next[t] = g1[s];
        if (s > 0 && st[s] < st[s - 1]) { printf("FAIL order\n"); return 1; }
    }
    for (int t = 0; t < T; ++t)
        if (next[t] != n_stages) { printf("FAIL coverage tile %d\n", t); return 1; }
    // balance: weighted work per CTA
    long long mx = 0, mn = 1LL << 62;
    for (int k = 0; k < G; ++k) {
        if (cta[k + 1] < cta[k]) { printf("FAIL cta order\n"); return 1; }
        long long w = (cta[k + 1] - cta[k] > 1) ? flush * (cta[k + 1] - cta[k]) : 0;
        for (int s = cta[k]; s < cta[k + 1]; ++s) {
            int ti = 0;
            while ((ti + 1) * (ti + 2) / 2 <= st[s]) ++ti;
            const bool diag = (st[s] - ti * (ti + 1) / 2) == ti;
            w += (long long)(g1[s] - g0[s]) * (diag ? wd : blr::SCHED_W_OFF);
        }
        mx = w > mx ? w : mx;
        mn = w < mn ? w : mn;
    }
    printf("ok %d %lld %lld\n", sc.nseg, mx, mn);
    return 0;
}
