// Counter-based Philox4x32-10 (Salmon et al., SC'11): reproducible random bits keyed on (seed, index, block),
// independent of how a batch is split over launches or ranks.  Used by the SpecAugment sampler and the
// training-mode GRU dropout.
#pragma once

#include <cstdint>

namespace sir {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
}

__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t index, uint32_t block, uint32_t (&c)[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    c[0] = (uint32_t)index;
    c[1] = (uint32_t)(index >> 32);
    c[2] = block;
    c[3] = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

}  // namespace sir
