// Stateless Philox4x32-10 (Salmon et al., SC'11) and fp64 normal variates for the samplers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppcseq {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t y0 = hi1 ^ c1 ^ k0, y1 = lo1, y2 = hi0 ^ c3 ^ k1, y3 = lo0;
        c0 = y0; c1 = y1; c2 = y2; c3 = y3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two independent N(0,1) variates from one Philox block addressed by (idx, counter, stream) under key `seed`
// (Box-Muller in fp64 on 2 x 53-bit... here 2 x (32+21)-bit uniforms)
__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t idx, uint64_t counter, uint32_t stream, double *z0,
                                            double *z1) {
    uint32_t r[4];
    philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32) ^ (stream << 8), (uint32_t)counter, (uint32_t)(counter >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    // u1 in (0,1], u2 in [0,1)
    const double u1 = ((double)(((uint64_t)r[0] << 21) | (r[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = (double)(((uint64_t)r[2] << 21) | (r[3] >> 11)) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    *z0 = rad * c;
    *z1 = rad * s;
}

}  // namespace ppcseq
