// philox.cuh -- the project's counter-based RNG contract on the device.
//
// The reference draws from CPython's global Mersenne Twister (random.*, SC:10,43-92,153,346,372,
// 406,425,576,890,902-903,912).  The B200 build defines one Philox4x32-10 stream per env instead:
//
//   block(k) = Philox4x32-10(counter = (k, 0, env_lo, env_hi), key = (seed_lo, seed_hi))
//   one block per random.* call (k += 1); shuffle/sample consume one block per swap
//   u = ((w0 >> 5) * 2^26 + (w1 >> 6)) / 2^53
//   uniform(a,b) = a + (b-a)*u ; randint(a,b) = a + min(floor(u*(b-a+1)), b-a)
//   normalvariate(m,s) = m + s*sqrt(-2 ln(1-u))*cos(2 pi u2), u2 from (w2,w3) of the same block
//
// The normative statement (and the shim that drives the unmodified reference with the same stream)
// is oracle/refshim/philox.py; DESIGN.md "RNG" repeats it.
#pragma once
#include <stdint.h>

#ifndef MH_HD
#ifdef __CUDACC__
#define MH_HD __host__ __device__ __forceinline__
#else
#define MH_HD inline
#endif
#endif

// one out-of-line copy on the device: ~80 integer instructions per block, called from many places
#if defined(__CUDACC__) && defined(MH_INLINE_PHILOX)
#define MH_PHILOX static __host__ __device__ __forceinline__
#elif defined(__CUDACC__)
#define MH_PHILOX static __host__ __device__ __noinline__
#else
#define MH_PHILOX static inline
#endif

namespace mhppo {

struct PhiloxBlock { uint32_t w0, w1, w2, w3; };

MH_PHILOX PhiloxBlock philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    PhiloxBlock b; b.w0 = c0; b.w1 = c1; b.w2 = c2; b.w3 = c3;
    return b;
}

MH_HD double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// Per-env stream cursor kept in registers; `ctr` is persisted in the env word of the state.
struct Rng {
    uint32_t ctr, env_lo, env_hi, k0, k1;
    MH_HD PhiloxBlock next() { return philox4x32_10(ctr++, 0u, env_lo, env_hi, k0, k1); }
    MH_HD double random() { PhiloxBlock b = next(); return u53(b.w0, b.w1); }
    MH_HD double uniform(double a, double b) { return a + (b - a) * random(); }
    MH_HD int randint(int a, int b) {
        const int n = b - a + 1;
        int k = (int)floor(random() * (double)n);
        if (k > n - 1) k = n - 1;
        return a + k;
    }
    MH_HD double normal(double mu, double sigma) {
        PhiloxBlock b = next();
        const double u1 = u53(b.w0, b.w1), u2 = u53(b.w2, b.w3);
        return mu + sigma * (sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2));
    }
    MH_HD void skip(int n) { ctr += (uint32_t)n; }
};

}  // namespace mhppo
