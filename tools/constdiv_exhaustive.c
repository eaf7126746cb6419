/*
 * constdiv_exhaustive.c -- proof tool (CPU only, not part of the product).  Proves, by enumerating all 2^32 fp32 inputs, that
 *     q = RN(x * rc);  r = fma(-q, c, x);  q' = fma(r, rc, q)          (rc = RN(1/c))
 * returns the correctly rounded x / c for the three constant divisors of the reference's collision
 * (SerialCode/d2q9-bgk.c:367-393): c_sq, 2*c_sq and 2*c_sq*c_sq.  The strict CUDA kernel uses this
 * sequence instead of div.rn.f32 for those divisors; GPU fma.rn.f32 and x86 fmaf are both IEEE-754
 * fused multiply-adds (denormals included), so a proof here carries over.
 * Inputs whose quotient is NaN are compared as "both NaN".  Prints the number of mismatches per
 * constant, over all inputs and inside the range 2^-100 <= |x| < 2^120 in which the kernel's div_const()
 * uses the sequence (outside it the kernel falls back to div.rn.f32); exit status 0 iff the guarded
 * range has no mismatch.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline float cdiv(float x, float c, float rc)
{
    float q = x * rc;
    float r = fmaf(-q, c, x);
    return fmaf(r, rc, q);
}
static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(void)
{
    const float c_sq = 1.f / 3.f;
    const float cs[3] = { c_sq, 2.f * c_sq, 2.f * c_sq * c_sq };
    const char* names[3] = { "c_sq", "2*c_sq", "2*c_sq*c_sq" };
    int rc_all = 0;
    for (int k = 0; k < 3; k++) {
        const float c = cs[k];
        const float rc = 1.f / c;
        unsigned long long bad = 0, bad_guarded = 0;
        uint32_t lo_bad = 0xffffffffu, hi_bad = 0; /* magnitude range (abs bits) of offenders */
#pragma omp parallel for reduction(+ : bad, bad_guarded) reduction(min : lo_bad) reduction(max : hi_bad) schedule(static)
        for (long long i = 0; i < (1LL << 32); i++) {
            uint32_t u = (uint32_t)i;
            float x; memcpy(&x, &u, 4);
            volatile float ref = x / c;
            float got = cdiv(x, c, rc);
            float refv = ref;
            int same = (bits(refv) == bits(got)) || (isnan(refv) && isnan(got));
            if (!same) {
                bad++;
                /* the range in which div_const() (lbm_kernels.cuh) trusts the sequence: 2^-100 <= |x| < 2^120 */
                const float ax = fabsf(x);
                if (ax >= 0x1p-100f && ax < 0x1p120f) bad_guarded++;
                uint32_t a = u & 0x7fffffffu;
                if (a < lo_bad) lo_bad = a;
                if (a > hi_bad) hi_bad = a;
            }
        }
        printf("%-12s c=%.9g (0x%08x) rc=%.9g (0x%08x): mismatches=%llu", names[k], c, bits(c), rc, bits(rc), bad);
        if (bad) printf("  |x| bits in [0x%08x, 0x%08x]", lo_bad, hi_bad);
        printf("  mismatches with 2^-100 <= |x| < 2^120: %llu\n", bad_guarded);
        if (bad_guarded) rc_all = 1;
    }
    return rc_all;
}
