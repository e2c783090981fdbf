/*
 * dsim_math.h — scalar rules shared by host set-up code and device kernels.
 *
 * Every function here must evaluate identically on the host and on the device: only IEEE
 * +,-,*,/ in a fixed order (compile with -fmad=false / -ffp-contract=off), no libm calls.
 */
#ifndef DSIM_MATH_H
#define DSIM_MATH_H

#include "dsim_rng.h"

/* OneYearResources comparisons (src/ecology.rs:13-34): `partial_cmp` with incomparable = Equal, so
 * std::cmp::min(a,b) = b iff a > b, and Ord::max(a,b) = a iff a > b. */
DSIM_HD double dsim_oyr_min(double a, double b) { return (a > b) ? b : a; }
DSIM_HD double dsim_oyr_max(double a, double b) { return (a > b) ? a : b; }

/* Rust `as u32` on f64 (lib.rs:535,1271,1878,1916): truncate toward zero, saturate, NaN -> 0. */
DSIM_HD uint32_t dsim_f64_as_u32(double x) {
    if (x != x) return 0u;
    if (x <= 0.0) return 0u;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}

DSIM_HD double dsim_bits_to_f64(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    __builtin_memcpy(&d, &b, 8);
    return d;
#endif
}
DSIM_HD uint64_t dsim_f64_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    __builtin_memcpy(&b, &d, 8);
    return b;
#endif
}

/* Natural logarithm for the mutation clock `fastrand::f64().log(1. - rate) as u32` (lib.rs:1271).
 * The quotient of two logarithms is truncated to an integer, so host and device must agree to the
 * last bit: x = 2^e m with m in (sqrt(1/2), sqrt 2], ln m = 2 atanh(s), s = (m-1)/(m+1), evaluated
 * as a fixed Horner chain.  Specification in DESIGN.md §3; the oracle carries its own copy. */
DSIM_HD double dsim_log(double x) {
    if (x != x) return x;
    if (x < 0.0) return dsim_bits_to_f64(0x7FF8000000000000ull);
    if (x == 0.0) return dsim_bits_to_f64(0xFFF0000000000000ull);
    uint64_t bits = dsim_f64_to_bits(x);
    if (bits == 0x7FF0000000000000ull) return x;
    int e = 0;
    if ((bits >> 52) == 0) {
        x *= 18014398509481984.0;
        bits = dsim_f64_to_bits(x);
        e = -54;
    }
    e += (int)(bits >> 52) - 1023;
    double m = dsim_bits_to_f64((bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    const double s = (m - 1.0) / (m + 1.0);
    const double z = s * s;
    double p = 1.0 / 27.0;
    p = p * z + 1.0 / 25.0;
    p = p * z + 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z + 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z + 1.0;
    const double r = (2.0 * s) * p;
    const double de = (double)e;
    return de * 6.93147180369123816490e-01 + (de * 1.90821492927058770002e-10 + r);
}

/* time_till_next_mutation, lib.rs:1270-1272: ln(U) / ln(1 - rate), `as u32`.  log1m = dsim_log(1 - rate). */
DSIM_HD uint32_t dsim_time_till_next_mutation(double u, double log1m) { return dsim_f64_as_u32(dsim_log(u) / log1m); }

/* one application of Ecovector * 0.95 on one entry (src/ecology.rs:147-161): f64::max(minimum, v * 0.95) */
DSIM_HD double dsim_adapt_decay(double v, double minimum) {
    const double d = v * 0.95;
    /* f64::max ignores a NaN operand */
    if (d != d) return minimum;
    if (minimum != minimum) return d;
    return (d > minimum) ? d : minimum;
}

DSIM_HD bool dsim_will_cooperate(uint64_t a, uint64_t b, uint32_t threshold, uint32_t inclusive) {
#if defined(__CUDA_ARCH__)
    const uint32_t d = (uint32_t)__popcll(a ^ b);
#else
    const uint32_t d = (uint32_t)__builtin_popcountll(a ^ b);
#endif
    return inclusive ? (d <= threshold) : (d < threshold);
}

#endif
