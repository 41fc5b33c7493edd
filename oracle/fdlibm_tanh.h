// =============================================================================
// fdlibm_tanh.h -- ORACLE side (test infrastructure) of the exactly specified tanh of the mixing taper.
//
// mix_imp.f:837-857 (tprstb) calls the Fortran intrinsic tanh = the platform libm.  glibc's __tanh is the fdlibm algorithm
// (tanh through expm1); its last bit depends on the glibc build (FMA multiarch variants on x86-64), so "the reference's tanh" is
// not one bit pattern.  vmix_jac divides differences of it by 1e-8, so a bit-exact comparison of the device path needs ONE
// specified tanh on both sides.  This is the oracle's own restatement of the published fdlibm algorithm (s_tanh.c / s_expm1.c,
// Sun Microsystems 1993, "freely granted"), written independently of the product's copy (i-emic_b200/csrc/thcm_tanh.h);
// tests/test_oracle_pins.py checks both against the platform libm (differences in < 0.1 % of arguments, <= 3 ulp) and against
// each other (bit-equal).  oracle_set_tanh(0) switches the oracle back to std::tanh.
// =============================================================================
#pragma once
#include <stdint.h>
#include <string.h>

static inline uint32_t or_hi(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32); }
static inline uint32_t or_lo(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u & 0xffffffffu); }
static inline double or_make(uint32_t hi, uint32_t lo) { uint64_t u = ((uint64_t)hi << 32) | (uint64_t)lo; double x; memcpy(&x, &u, 8); return x; }
static inline double or_add_exponent(double y, int k) { return or_make(or_hi(y) + ((uint32_t)k << 20), or_lo(y)); }

static inline double oracle_expm1(double x) {
    static const double huge = 1.0e+300, tiny = 1.0e-300, o_threshold = 7.09782712893383973096e+02,
                        ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        invln2 = 1.44269504088896338700e+00;
    static const double Q[6] = {0.0, -3.33333333333331316428e-02, 1.58730158725481460165e-03, -7.93650757867487942473e-05,
                                4.00821782732936239552e-06, -2.01099218183624371326e-07};
    uint32_t hx = or_hi(x);
    const int negative = (hx >> 31) != 0;
    hx &= 0x7fffffffu;
    if (hx >= 0x4043687Au) {                                    /* |x| >= 56 ln2 */
        if (hx >= 0x40862E42u) {                                /* |x| >= 709.78 */
            if (hx >= 0x7ff00000u) {
                if (((hx & 0xfffffu) | or_lo(x)) != 0) return x + x;
                return negative ? -1.0 : x;
            }
            if (x > o_threshold) return huge * huge;
        }
        if (negative && x + tiny < 0.0) return tiny - 1.0;
    }
    int k = 0;
    double c = 0.0;
    if (hx > 0x3fd62e42u) {                                     /* |x| > 0.5 ln2: reduce */
        double hi, lo;
        if (hx < 0x3FF0A2B2u) {                                 /* |x| < 1.5 ln2 */
            if (!negative) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
            else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
        } else {
            k = (int)(invln2 * x + (negative ? -0.5 : 0.5));
            const double t = k;
            hi = x - t * ln2_hi;
            lo = t * ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x3c900000u) {                              /* |x| < 2^-54 */
        const double t = huge + x;
        return x - (t - (huge + x));
    }
    const double hfx = 0.5 * x, hxs = x * hfx;
    const double r1 = 1.0 + hxs * (Q[1] + hxs * (Q[2] + hxs * (Q[3] + hxs * (Q[4] + hxs * Q[5]))));
    double t = 3.0 - r1 * hfx;
    double e = hxs * ((r1 - t) / (6.0 - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = (x * (e - c) - c);
    e -= hxs;
    if (k == -1) return 0.5 * (x - e) - 0.5;
    if (k == 1) return x < -0.25 ? -2.0 * (e - (x + 0.5)) : 1.0 + 2.0 * (x - e);
    double y;
    if (k <= -2 || k > 56) return or_add_exponent(1.0 - (e - x), k) - 1.0;
    if (k < 20) {
        t = or_make(0x3ff00000u - (0x200000u >> k), 0u);        /* 1 - 2^-k */
        y = t - (e - x);
    } else {
        t = or_make((uint32_t)(0x3ff - k) << 20, 0u);           /* 2^-k */
        y = x - (e + t);
        y += 1.0;
    }
    return or_add_exponent(y, k);
}

static inline double oracle_tanh(double x) {
    const uint32_t jx = or_hi(x), ix = jx & 0x7fffffffu;
    const int negative = (jx >> 31) != 0;
    if (ix >= 0x7ff00000u) return negative ? 1.0 / x - 1.0 : 1.0 / x + 1.0;
    double z;
    if (ix < 0x40360000u) {                                     /* |x| < 22 */
        if ((ix | or_lo(x)) == 0) return x;
        if (ix < 0x3c800000u) return x * (1.0 + x);             /* |x| < 2^-55 */
        const double ax = or_make(ix, or_lo(x));
        if (ix >= 0x3ff00000u) { const double t = oracle_expm1(2.0 * ax); z = 1.0 - 2.0 / (t + 2.0); }
        else { const double t = oracle_expm1(-2.0 * ax); z = -t / (t + 2.0); }
    } else z = 1.0 - 1.0e-300;
    return negative ? -z : z;
}
