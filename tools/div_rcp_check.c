/* dev tool: x / d against the correctly-rounded-reciprocal form of env_core.cuh div_rcp, 1.3e9 operands.  gcc -O2 -mfma tools/div_rcp_check.c -lm */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t nxt(void) { uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b; }
static inline double u01(void) { return (double)(nxt() >> 11) * (1.0 / 9007199254740992.0); }
static inline double div_rcp(double x, double d, double rd) { double q0 = x * rd; double r = fma(-d, q0, x); return fma(r, rd, q0); }
int main(void) {
    const double consts[] = {0.3, 10.0, 1.0 - (2.0 / 3.141592653589793), 8.0, 0.6, 2.0 * sqrt(2.0 * 4.0), 0.25, 3.0, 7.0, 1.1};
    long bad = 0, n = 0;
    for (int c = 0; c < 10; ++c) {
        const double d = consts[c], rd = 1.0 / d;
        for (long i = 0; i < 100000000; ++i) {
            double x;
            const int k = (int)(nxt() & 7);
            if (k < 4) x = (u01() - 0.5) * 200.0;                      /* kinematic magnitudes */
            else if (k < 6) x = ldexp(u01() + 0.5, (int)(nxt() % 120) - 60);   /* wide exponents */
            else { uint64_t b = (nxt() & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull; memcpy(&x, &b, 8); x *= d; }   /* near-exact quotients */
            const double a = x / d, b2 = div_rcp(x, d, rd);
            ++n;
            if (memcmp(&a, &b2, 8)) { if (bad < 5) printf("MISMATCH d=%a x=%a %a %a\n", d, x, a, b2); ++bad; }
        }
    }
    /* variable divisor with a correctly rounded reciprocal (the per-env crosswalk width) */
    for (long i = 0; i < 300000000; ++i) {
        const double d = 2.5 + 0.5 * u01() + (double)(nxt() & 3), rd = 1.0 / d;
        const double x = (u01() - 0.2) * 4.0 * d;
        const double a = x / d, b2 = div_rcp(x, d, rd);
        ++n;
        if (memcmp(&a, &b2, 8)) { if (bad < 10) printf("MISMATCH d=%a x=%a %a %a\n", d, x, a, b2); ++bad; }
    }
    printf("trials %ld mismatches %ld\n", n, bad);
    return bad != 0;
}
