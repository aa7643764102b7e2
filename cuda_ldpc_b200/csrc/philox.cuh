// Philox4x32-10 (Salmon et al., SC'11) and the Box-Muller step shared by the channel kernels and the fused
// load phase of the layered decoder: one call = the four normals of bits 4*nb .. 4*nb+3 of one frame.
#pragma once
#include <stdint.h>

namespace ldpcb {

__host__ __device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

#ifdef __CUDACC__
// uniform in (0, 1]: never 0, so log() is finite (24 bits: the angle of the Box-Muller pair)
__device__ __forceinline__ float philox_u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// the RADIUS uniform keeps all 32 bits near 0, where the tail of the normal lives: (x + 0.5) / 2^32 in (0, 1] — a float
// resolves 2^-33 there.  Largest |n| = sqrt(-2 ln 2^-33) = 6.76 sigma (P = 1.4e-11 per sample) instead of the 5.77 sigma
// (P = 8e-9) of a 24-bit uniform: with 38 400 samples per frame a 24-bit tail would clip one sample in every 3 000 frames.
__device__ __forceinline__ float philox_u01_tail(uint32_t x) { return __fmaf_rn((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// g[0..3] = N(0,1) samples of bits 4*nb .. 4*nb+3 of global frame gf
__device__ __forceinline__ void awgn_normals4(unsigned long long gf, int nb, uint32_t k0, uint32_t k1, float g[4])
{
    uint32_t c[4] = {(uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)nb, 0x4C445043u /* "LDPC" */};
    philox4x32_10(c, k0, k1);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float r = sqrtf(-2.0f * __logf(philox_u01_tail(c[2 * h])));
        float sn, cs;
        __sincosf(6.283185307179586f * philox_u01(c[2 * h + 1]), &sn, &cs);
        g[2 * h] = r * cs;
        g[2 * h + 1] = r * sn;
    }
}
// the channel value itself: BPSK point + sigma * normal (one rounding each, no contraction)
__device__ __forceinline__ float awgn_bpsk_sample(int bit, float sigma, float g)
{
    return __fadd_rn(1.0f - 2.0f * (float)bit, __fmul_rn(sigma, g));
}
#endif

}  // namespace ldpcb
