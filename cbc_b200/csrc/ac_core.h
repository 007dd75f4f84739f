/*
 * ac_core.h -- the reference's 26-bit arithmetic coder (src/Arithmetic_stream.c:274-454) as
 * closed-form integer steps.
 *
 * The reference renormalises one bit per loop iteration (E1/E2: equal MSBs -> emit; E3: l = 01..,
 * u = 10.. -> count a pending bit). After an interval update that loop always runs as
 *     k x (E1/E2)   followed by   m x (E3)
 * because an E3 step leaves l's MSB 0 and u's MSB 1, so no E1/E2 step can follow it. Here k is the
 * number of leading bits l and u share and m the length of the run, below the MSB, where l has ones
 * and u zeros; both come from one count-leading-zeros each, so a symbol costs a fixed handful of
 * integer instructions instead of a data-dependent loop. Compiled for host (unit tests against a
 * bit-serial restatement) and device.
 */
#pragma once
#include <stdint.h>
#include "cbcg_format.h"

#ifdef __CUDACC__
#define AC_HD __host__ __device__ __forceinline__
#else
#define AC_HD static inline
#endif

AC_HD uint32_t ac_clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

/* Exact floor((a * b - sub) / d) for a, b, d < 2^27 whose quotient is < 2^27 (every division of this coder:
 * range <= 2^26, counts and totals < 2^21). The product is exact in a double (< 2^54 ... here < 2^48), the
 * reciprocal comes from rcp.approx (relative error <= 2^-23) plus one Newton step (<= 2^-45), so the estimate
 * trunc(p * r) is off by at most one and the integer remainder check puts it right. One reciprocal serves both
 * quotients of an interval update: ~15 instructions each instead of a full 64-bit division. */
AC_HD double ac_rcp(uint32_t d) {
#ifdef __CUDA_ARCH__
    const double x = (double)d;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = __fma_rn(-x, r, 1.0);
    return __fma_rn(r, e, r);
#else
    return 1.0 / (double)d;
#endif
}
AC_HD uint32_t ac_muldiv(uint32_t a, uint32_t b, uint32_t sub, uint32_t d, double r) {
#ifdef __CUDA_ARCH__
    const double pd = __dmul_rn(__uint2double_rn(a), __uint2double_rn(b)) - (double)sub;
    uint32_t q = __double2uint_rz(__dmul_rn(pd, r));
#else
    const double pd = (double)a * (double)b - (double)sub;
    uint32_t q = (uint32_t)(pd * r);
#endif
    /* the estimate is off by at most one, so the true remainder lies in (-d, 2d): |.| < 2^28, and the low 32 bits
       of the product carry it exactly (no 64-bit arithmetic on the dependent chain) */
    const int32_t rem = (int32_t)(a * b - sub - q * d);
    if (rem < 0) q--; else if (rem >= (int32_t)d) q++;
    return q;
}

struct AcInterval { uint32_t l, u; };

/* Interval update of arithmetic_encoder_step / arithmetic_decoder_step (:295-296, :402-403):
 * u is computed from the old l, both products in 64 bits, truncated to 32. */
/* r = ac_rcp(n): a caller that knows n before it knows the symbol (every decoder step) forms it off the dependent chain */
AC_HD void ac_narrow_r(AcInterval &a, uint32_t lo, uint32_t hi, uint32_t n, double r) {
    const uint32_t range = a.u - a.l + 1u;                 /* <= 2^26 */
    const uint32_t nu = a.l + ac_muldiv(range, hi, 0u, n, r) - 1u;
    const uint32_t nl = a.l + ac_muldiv(range, lo, 0u, n, r);
    a.u = nu; a.l = nl;
}
AC_HD void ac_narrow(AcInterval &a, uint32_t lo, uint32_t hi, uint32_t n) { ac_narrow_r(a, lo, hi, n, ac_rcp(n)); }

/* Renormalisation shape: k E1/E2 shifts (the top k bits of l are the emitted bits), then m E3 shifts. */
AC_HD void ac_renorm_shape(const AcInterval &a, uint32_t &k, uint32_t &bits, uint32_t &m, AcInterval &out) {
    const uint32_t x = (a.l ^ a.u) & CBCG_AC_TOP;
    k = x ? (ac_clz32(x) - (32u - CBCG_AC_BITS)) : CBCG_AC_BITS;
    bits = k ? (a.l >> (CBCG_AC_BITS - k)) : 0u;
    uint32_t l = (a.l << k) & CBCG_AC_TOP;                            /* k <= 26: the bits shifted past 32 are masked anyway */
    uint32_t u = ((a.u << k) & CBCG_AC_TOP) | ((1u << k) - 1u);
    /* run of (l bit = 1, u bit = 0) from bit 24 downwards */
    const uint32_t z = (l & ~u) << (32u - (CBCG_AC_BITS - 1u));      /* bit 24 -> bit 31 */
    m = ac_clz32(~z);
    if (m > CBCG_AC_BITS - 1u) m = CBCG_AC_BITS - 1u;
    if (m) {
        l = (l << m) & CBCG_AC_LOWMASK;
        u = ((u << m) & CBCG_AC_LOWMASK) | CBCG_AC_MSB | ((1u << m) - 1u);
    }
    out.l = l; out.u = u;
}

/* Decoder tag update for the same shape (:431-453): k plain shifts, then m shifts whose last one
 * leaves the MSB flipped. `in` holds the next k + m stream bits, first bit most significant. */
AC_HD uint32_t ac_tag_shift(uint32_t t, uint32_t k, uint32_t m, uint32_t in) {
    const uint32_t s = k + m;
    if (s == 0) return t;
    uint32_t r = (uint32_t)(((((uint64_t)t << s) | in) & CBCG_AC_TOP));
    if (m) r ^= CBCG_AC_MSB;
    return r;
}

/* arithmetic_get_symbol_range (:373-381) */
AC_HD uint32_t ac_target(const AcInterval &a, uint32_t t, uint32_t n) {
    const uint32_t range = a.u - a.l + 1u;
    const uint32_t gap = t - a.l + 1u;
    return ac_muldiv(gap, n, 1u, range, ac_rcp(range));
}
