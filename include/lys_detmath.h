/* lys_detmath.h -- the arithmetic contract of the B200 tracer.
 *
 * The reference (bryal/msc-futhark-ray-tracer) evaluates sin/cos/exp/log/pow/acos
 * through whatever libm its Futhark backend links (glibc for `c`, the OpenCL /
 * CUDA built-ins for the GPU backends), so the last ulp of those calls is
 * backend-defined, not reference-defined.  To make "same inputs -> same bits"
 * a testable statement between a CPU restatement and sm_100a kernels, every
 * transcendental that is evaluated PER PATH is defined here once, using only
 * IEEE-754 correctly rounded primitives (+ - * / sqrt fma, int<->float
 * conversions), in a fixed evaluation order.  The same text compiles under g++
 * (-ffp-contract=off) and nvcc (-fmad=false) to the same results.
 *
 * Call sites in the reference (paths relative to /root/reference):
 *   det_sinf / det_cosf : src/rand.fut:25 (unit disk), src/material.fut:270-271
 *   det_expf            : src/material.fut:222 (Beckmann D)
 *   det_logf            : src/material.fut:286 (sample_wh), probit tails
 *   det_pow5f           : src/material.fut:211 (`** 5`, Schlick)
 *   det_acosf           : src/light.fut:42 (frustum light cone test)
 *   det_probitf         : src/camera.fut:78 (statistics pkg `sample (mk_normal ..) p`;
 *                         package source is not vendored -> "parity unpinned")
 *   lys_fminf/lys_fmaxf : every `f32.min` / `f32.max` (fminf/fmaxf NaN semantics,
 *                         first operand returned on ties so +-0 is deterministic)
 *
 * Polynomial coefficients are the classic single-precision Cephes sets
 * (public domain, S. Moshier); accuracy against float64 libm is checked in
 * tests/test_detmath.py (max error bound stated there).
 */
#ifndef LYS_DETMATH_H
#define LYS_DETMATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define LYS_HD __host__ __device__ __forceinline__
#else
#define LYS_HD static inline
#endif

LYS_HD uint32_t lys_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LYS_HD float lys_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

/* fminf/fmaxf semantics (a NaN operand is ignored), ties return `a`. */
LYS_HD float lys_fminf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    return (b < a) ? b : a;
}
LYS_HD float lys_fmaxf(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    return (b > a) ? b : a;
}
LYS_HD float lys_fabsf(float a) { return lys_u2f(lys_f2u(a) & 0x7fffffffu); }
LYS_HD int lys_isinff(float a) { return (lys_f2u(a) & 0x7fffffffu) == 0x7f800000u; }

/* Futhark f32.sgn: -1, 0, 1 (NaN -> 1). */
LYS_HD float lys_sgnf(float x) { return (x < 0.0f) ? -1.0f : ((x == 0.0f) ? 0.0f : 1.0f); }

/* round-to-nearest-even to an integer-valued float for |x| < 2^22 */
LYS_HD float det_rintf_small(float x) {
    const float magic = 12582912.0f; /* 1.5 * 2^23 */
    float t = x + magic;
    return t - magic;
}

/* Shared quadrant reduction for sin/cos, valid for |x| <= ~1e4 (call sites
 * stay within [0, 2*pi]).  r in [-pi/4, pi/4], q = quadrant. */
LYS_HD float det_reduce_pio2(float x, int *q) {
    const float two_over_pi = 0.636619772367581343f;
    const float p1 = 1.5703125f;                 /* pi/2 split in 3 (Cody-Waite) */
    const float p2 = 4.837512969970703125e-4f;
    const float p3 = 7.54978995489188216e-8f;
    float k = det_rintf_small(x * two_over_pi);
    float r = fmaf(-k, p1, x);
    r = fmaf(-k, p2, r);
    r = fmaf(-k, p3, r);
    *q = (int)k;
    return r;
}
LYS_HD float det_sin_poly(float r) {
    float z = r * r;
    float p = -1.9515295891e-4f;
    p = fmaf(p, z, 8.3321608736e-3f);
    p = fmaf(p, z, -1.6666654611e-1f);
    return fmaf(p * z, r, r);
}
LYS_HD float det_cos_poly(float r) {
    float z = r * r;
    float p = 2.443315711809948e-5f;
    p = fmaf(p, z, -1.388731625493765e-3f);
    p = fmaf(p, z, 4.166664568298827e-2f);
    float t = fmaf(-0.5f, z, 1.0f);
    return fmaf(p * z, z, t);
}
LYS_HD float det_sinf(float x) {
    int q; float r = det_reduce_pio2(x, &q);
    float s = (q & 1) ? det_cos_poly(r) : det_sin_poly(r);
    return (q & 2) ? -s : s;
}
LYS_HD float det_cosf(float x) {
    int q; float r = det_reduce_pio2(x, &q);
    float c = (q & 1) ? det_sin_poly(r) : det_cos_poly(r);
    return ((q + 1) & 2) ? -c : c;
}

/* exp(x).  Overflow -> +inf, gradual underflow kept (two-step scaling). */
LYS_HD float det_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283935546875f) return lys_u2f(0x7f800000u);
    if (x < -103.972084045410f) return 0.0f;
    const float log2e = 1.44269504088896341f;
    const float c1 = 0.693359375f;
    const float c2 = -2.12194440e-4f;
    float n = det_rintf_small(x * log2e);
    float r = fmaf(-n, c1, x);
    r = fmaf(-n, c2, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float y = fmaf(p, r * r, r) + 1.0f;
    int e = (int)n;
    /* y in ~[0.7, 1.42]; scale by 2^e in one or two exact-power steps */
    if (e < -125) {
        y = y * lys_u2f((uint32_t)(e + 100 + 127) << 23);
        return y * lys_u2f((uint32_t)(-100 + 127) << 23);
    }
    if (e > 127) {
        y = y * 2.0f;
        e -= 1;
    }
    return y * lys_u2f((uint32_t)(e + 127) << 23);
}

/* natural log.  log(0) = -inf, log(<0) = NaN, denormals handled. */
LYS_HD float det_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return lys_u2f(0x7fc00000u);
    if (x == 0.0f) return lys_u2f(0xff800000u);
    if (lys_isinff(x)) return x;
    int e = 0;
    uint32_t u = lys_f2u(x);
    if (u < 0x00800000u) {            /* denormal: renormalise exactly */
        x = x * 8388608.0f;           /* 2^23 */
        u = lys_f2u(x);
        e = -23;
    }
    e += (int)(u >> 23) - 126;        /* x = m * 2^e, m in [0.5, 1) */
    float m = lys_u2f((u & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) {
        e -= 1;
        m = (m + m) - 1.0f;
    } else {
        m = m - 1.0f;
    }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float y = (p * m) * z;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(fe, 0.693359375f, r);
}

/* x**5 as the reference's powf(x, 5): evaluated in binary64 and rounded once. */
LYS_HD float det_pow5f(float x) {
    double d = (double)x;
    double d2 = d * d;
    double d4 = d2 * d2;
    return (float)(d4 * d);
}

/* acos on [-1, 1]; outside -> NaN. */
LYS_HD float det_asin_core(float a) { /* 0 <= a <= 0.5 */
    float z = a * a;
    float p = 4.2163199048e-2f;
    p = fmaf(p, z, 2.4181311049e-2f);
    p = fmaf(p, z, 4.5470025998e-2f);
    p = fmaf(p, z, 7.4953002686e-2f);
    p = fmaf(p, z, 1.6666752422e-1f);
    return fmaf(p * z, a, a);
}
LYS_HD float det_acosf(float x) {
    const float pio2_hi = 1.57079637050628662109375f;
    const float pio2_lo = -4.37113900018624283e-8f;
    const float pi_hi = 3.1415927410125732421875f;
    const float pi_lo = -8.74227800037248566e-8f;
    if (x != x) return x;
    float a = lys_fabsf(x);
    if (a > 1.0f) return lys_u2f(0x7fc00000u);
    if (a <= 0.5f) {
        float s = det_asin_core(a);
        s = (x < 0.0f) ? -s : s;
        return (pio2_hi - s) + pio2_lo;
    }
    float h = 0.5f * (1.0f - a);
    float s = det_asin_core(sqrtf(h));
    float t = s + s;                       /* acos(|x|) */
    if (x > 0.0f) return t;
    return (pi_hi - t) + pi_lo;
}

/* Standard-normal quantile (probit): Acklam's rational approximation (|rel err| < 1.2e-9)
 * evaluated in binary64 with plain (uncontracted) Horner steps and rounded once to f32;
 * the tail argument uses det_logf.  p = 0 -> -inf, p = 1 -> +inf. */
LYS_HD double det_horner6(double x, double k0, double k1, double k2, double k3, double k4, double k5) {
    double r = k0;
    r = r * x + k1; r = r * x + k2; r = r * x + k3; r = r * x + k4; r = r * x + k5;
    return r;
}
LYS_HD float det_probitf(float p) {
    const double plow = 0.02425;
    if (p != p) return p;
    if (p <= 0.0f) return (p == 0.0f) ? lys_u2f(0xff800000u) : lys_u2f(0x7fc00000u);
    if (p >= 1.0f) return (p == 1.0f) ? lys_u2f(0x7f800000u) : lys_u2f(0x7fc00000u);
    double dp = (double)p;
    if (dp < plow || dp > 1.0 - plow) {
        int upper = dp > 0.5;
        float pt = upper ? (1.0f - p) : p;            /* exact in f32 for p > 0.5 */
        double q = sqrt(-2.0 * (double)det_logf(pt));
        double num = det_horner6(q, -7.784894002430293e-03, -3.223964580411365e-01,
                                 -2.400758277161838e+00, -2.549732539343734e+00,
                                 4.374664141464968e+00, 2.938163982698783e+00);
        double den = det_horner6(q, 0.0, 7.784695709041462e-03, 3.224671290700398e-01,
                                 2.445134137142996e+00, 3.754408661907416e+00, 1.0);
        double v = num / den;
        return (float)(upper ? -v : v);
    }
    double q = dp - 0.5;
    double r = q * q;
    double num = det_horner6(r, -3.969683028665376e+01, 2.209460984245205e+02,
                             -2.759285104469687e+02, 1.383577518672690e+02,
                             -3.066479806614716e+01, 2.506628277459239e+00) * q;
    double den = det_horner6(r, -5.447609879822406e+01, 1.615858368580409e+02,
                             -1.556989798598866e+02, 6.680131188771972e+01,
                             -1.328068155288572e+01, 1.0);
    return (float)(num / den);
}

#endif /* LYS_DETMATH_H */
