// How the two warps of a K1m team (gram_mid.cu) share the lower triangle of an MI x MI grid of 8 x 8 sub-tiles.
// Host/device constexpr, no CUDA types: unit-tested on the CPU (tests/cpp/mid_deal_test.cpp).
#pragma once

#ifdef __CUDACC__
#define BLR_HD __host__ __device__
#else
#define BLR_HD
#endif

namespace blr {
namespace gm {
// which member (0 / 1) computes sub-tile (mi, ni), ni <= mi
BLR_HD constexpr int block_owner(int mi, int ni, int MI) {
    const int full = (MI / 4) * 4, rem = MI - full;
    if (mi < full) return ((mi & 3) == 0 || (mi & 3) == 3) ? 0 : 1;  // rows r have r + 1 sub-tiles: 0 + 3 == 1 + 2 per group of four
    if ((rem & 1) && mi == MI - 1) return ni & 1;                    // a leftover odd row is split by column parity
    return (mi - full) & 1;
}
// which member accumulates r for block row mi
BLR_HD constexpr int row_owner(int mi, int MI) { return block_owner(mi, 0, MI); }
}  // namespace gm
}  // namespace blr
