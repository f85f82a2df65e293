// Philox4x32-10 counter-based RNG (SPEC.md §3), device + host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CT_HD __host__ __device__ __forceinline__
#else
#define CT_HD inline
#endif

namespace ct {

struct Philox4 {
    uint32_t w[4];
};

CT_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

// counter = (c0, c1, c2, c3), key = (k0, k1)
CT_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 out;
    out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
    return out;
}

// Same generator with the ten round keys precomputed (rk[2r] = k0 + r*W0, rk[2r+1] = k1 + r*W1):
// when rk sits in constant memory the key schedule costs no instructions.
CT_HD Philox4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
        uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    Philox4 out;
    out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
    return out;
}

// 23-bit uniform strictly inside (0,1): ((w >> 9) + 0.5) * 2^-23, built without an int->float
// conversion: the mantissa trick gives [1,2), then subtract (1 - 2^-24).
CT_HD float u01(uint32_t w)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(0x3f800000u | (w >> 9)) - 0.99999994039535522461f;
#else
    return (float)(((double)(w >> 9) + 0.5) * (1.0 / 8388608.0));
#endif
}

}  // namespace ct
