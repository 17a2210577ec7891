// Counter-based Philox4x32-10 and a Box-Muller pair of fp64 standard normals.
// Element e of stream `stream_id` under `seed` is a pure function of (seed, stream_id, e): synthetic data is
// identical however the observations are partitioned over GPUs.
#pragma once
#include <stdint.h>

namespace blr {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k0 += W0;
        k1 += W1;
    }
}

// two independent N(0,1) draws for pair index `idx`
__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint32_t stream_id, uint64_t idx, double& z0,
                                                   double& z1) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), stream_id, 0x5eedu};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c[1] << 32) | c[0], b = ((uint64_t)c[3] << 32) | c[2];
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0, 1)
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    z0 = r * co;
    z1 = r * s;
}

}  // namespace blr
