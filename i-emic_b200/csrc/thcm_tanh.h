// =============================================================================
// thcm_tanh.h -- one exactly specified tanh for the tracer-mixing taper (mix_imp.f:837-857, tprstb).
//
// tprstb is the only state-dependent transcendental of the hot path, and vmix_jac differences it with eps = 1e-8
// (mix_imp.f:729-815): one ulp of tanh becomes 1e-8 / 1e-16 = 1e8 ulp of a Jacobian entry.  The reference takes tanh from
// the platform libm (gfortran intrinsic -> glibc), whose last bit differs between glibc builds (the x86-64 multiarch
// variants contract to FMA, the baseline build does not) and from CUDA's tanh.  To make the device path bit-reproducible
// against the CPU oracle, both use THIS algorithm: the classic fdlibm formulation (Sun Microsystems' freely distributable
// libm, the ancestor of glibc's __tanh / __expm1): tanh through expm1 with the same branches, thresholds and coefficients,
// every operation a plain IEEE double operation (compile without FMA contraction: nvcc --fmad=false, gcc -ffp-contract=off).
// tests/test_oracle_pins.py checks it against the platform libm (<= 1 ulp over a dense sample).
// =============================================================================
#pragma once
#include <cstdint>
#include <cstring>

#ifndef THCM_HD
#ifdef __CUDACC__
#define THCM_HD __host__ __device__ __forceinline__
#else
#define THCM_HD inline
#endif
#endif

// The tanh itself is NOT inlined on the device: the mixing Jacobian calls it ten times per tracer row, and ten inlined copies of these
// branches (two FP64 divisions each) in every unrolled row evaluation pushed the assembly kernels out of the instruction cache.
#ifdef __CUDACC__
#define THCM_TANH_FN static __host__ __device__ __noinline__
#else
#define THCM_TANH_FN inline
#endif

namespace thcm {

THCM_HD uint32_t f64_hi(double x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2hiint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32);
#endif
}
THCM_HD uint32_t f64_lo(double x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}
THCM_HD double f64_from(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t u = ((uint64_t)hi << 32) | lo; double x; memcpy(&x, &u, 8); return x;
#endif
}

// expm1(x) = exp(x) - 1: argument reduction x = k ln2 + r, |r| <= 0.5 ln2, rational approximation of expm1(r) in r^2 / 2
// (five coefficients Q1..Q5), reconstruction by cases on k.
THCM_HD double fd_expm1(double x) {
    const double one = 1.0, huge = 1.0e+300, tiny = 1.0e-300;
    const double o_threshold = 7.09782712893383973096e+02;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, invln2 = 1.44269504088896338700e+00;
    const double Q1 = -3.33333333333331316428e-02, Q2 = 1.58730158725481460165e-03, Q3 = -7.93650757867487942473e-05,
                 Q4 = 4.00821782732936239552e-06, Q5 = -2.01099218183624371326e-07;
    double y, hi, lo, c = 0.0, t, e, hxs, hfx, r1;
    int k;
    uint32_t hx = f64_hi(x);
    const uint32_t xsb = hx & 0x80000000u;     // sign bit
    hx &= 0x7fffffffu;
    if (hx >= 0x4043687Au) {                   // |x| >= 56 ln2
        if (hx >= 0x40862E42u) {               // |x| >= 709.78
            if (hx >= 0x7ff00000u) {
                if (((hx & 0xfffffu) | f64_lo(x)) != 0) return x + x;   // NaN
                return xsb == 0 ? x : -1.0;                              // exp(+-inf) - 1
            }
            if (x > o_threshold) return huge * huge;                     // overflow
        }
        if (xsb != 0) { if (x + tiny < 0.0) return tiny - one; }         // -1 with inexact
    }
    if (hx > 0x3fd62e42u) {                    // |x| > 0.5 ln2
        if (hx < 0x3FF0A2B2u) {                // |x| < 1.5 ln2
            if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
            else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
        } else {
            k = (int)(invln2 * x + (xsb == 0 ? 0.5 : -0.5));
            t = (double)k;
            hi = x - t * ln2_hi;               // t * ln2_hi is exact here
            lo = t * ln2_lo;
        }
        x = hi - lo;
        c = (hi - x) - lo;
    } else if (hx < 0x3c900000u) {             // |x| < 2^-54: x itself
        t = huge + x;
        return x - (t - (huge + x));
    } else k = 0;
    hfx = 0.5 * x;
    hxs = x * hfx;
    r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    t = 3.0 - r1 * hfx;
    e = hxs * ((r1 - t) / (6.0 - x * t));
    if (k == 0) return x - (x * e - hxs);
    e = (x * (e - c) - c);
    e -= hxs;
    if (k == -1) return 0.5 * (x - e) - 0.5;
    if (k == 1) {
        if (x < -0.25) return -2.0 * (e - (x + 0.5));
        return one + 2.0 * (x - e);
    }
    if (k <= -2 || k > 56) {                   // exp(x) - 1 = 2^k (1 - (e - x)) - 1
        y = one - (e - x);
        y = f64_from(f64_hi(y) + ((uint32_t)k << 20), f64_lo(y));      // add k to y's exponent
        return y - one;
    }
    if (k < 20) {
        t = f64_from(0x3ff00000u - (0x200000u >> k), 0u);               // 1 - 2^-k
        y = t - (e - x);
    } else {
        t = f64_from((uint32_t)(0x3ff - k) << 20, 0u);                  // 2^-k
        y = x - (e + t);
        y += one;
    }
    return f64_from(f64_hi(y) + ((uint32_t)k << 20), f64_lo(y));
}

// tanh(x): |x| < 2^-55 -> x (1 + x); |x| < 1 -> -t / (t + 2), t = expm1(-2|x|); |x| < 22 -> 1 - 2 / (t + 2), t = expm1(2|x|); else 1 - tiny
THCM_TANH_FN double fd_tanh(double x) {
    const double one = 1.0, two = 2.0, tiny = 1.0e-300;
    double t, z;
    const uint32_t jx = f64_hi(x), ix = jx & 0x7fffffffu;
    if (ix >= 0x7ff00000u) return (jx >> 31) ? one / x - one : one / x + one;      // inf or NaN
    if (ix < 0x40360000u) {                                                         // |x| < 22
        if ((ix | f64_lo(x)) == 0) return x;                                        // +-0
        if (ix < 0x3c800000u) return x * (one + x);                                 // |x| < 2^-55
        const double ax = f64_from(ix, f64_lo(x));
        if (ix >= 0x3ff00000u) { t = fd_expm1(two * ax); z = one - two / (t + two); }
        else { t = fd_expm1(-two * ax); z = -t / (t + two); }
    } else z = one - tiny;
    return (jx >> 31) ? -z : z;
}

}  // namespace thcm
