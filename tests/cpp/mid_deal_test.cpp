// CPU check of the sub-tile deal of the K1m team kernel (csrc/mid_deal.h): every sub-tile of the lower triangle has exactly one
// owner in {0, 1}, the two members' counts differ by at most one, every block row has one r-owner who also owns a sub-tile of it.
#include <cstdio>
#include <cstdlib>

#include "mid_deal.h"

int main(int argc, char** argv) {
    const int MI = argc > 1 ? atoi(argv[1]) : 12;
    int cnt[2] = {0, 0};
    for (int mi = 0; mi < MI; ++mi) {
        bool row_has[2] = {false, false};
        for (int ni = 0; ni <= mi; ++ni) {
            const int o = blr::gm::block_owner(mi, ni, MI);
            if (o != 0 && o != 1) { printf("bad owner %d at (%d, %d)\n", o, mi, ni); return 1; }
            ++cnt[o];
            row_has[o] = true;
        }
        const int ro = blr::gm::row_owner(mi, MI);
        if ((ro != 0 && ro != 1) || !row_has[ro]) { printf("row owner %d of row %d owns none of its sub-tiles\n", ro, mi); return 1; }
    }
    if (cnt[0] + cnt[1] != MI * (MI + 1) / 2) { printf("coverage %d + %d\n", cnt[0], cnt[1]); return 1; }
    printf("ok %d %d\n", cnt[0], cnt[1]);
    return 0;
}
