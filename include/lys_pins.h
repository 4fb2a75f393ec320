/* lys_pins.h -- every third-party semantic the tracer had to restate from memory, ONE definition each.
 *
 * The reference imports five Futhark packages (futhark.pkg:2-6) whose sources are not vendored under
 * /root/reference/lib, and no Futhark compiler exists on the build or GPU boxes, so what those packages compute is
 * restated here from their published sources as remembered.  Both the CPU oracle
 * (oracle/lys_oracle.cpp) and the product (csrc/lys_device.cuh, csrc/wavefront.cu, csrc/abi.cu) call THESE functions, so
 * the two cannot drift apart, and each guess sits behind one named switch: a maintainer who has `futhark` runs
 * `futhark test oracle/pin/*.fut` (oracle/pin/README.md); a failing pin names the switch below to flip, and
 * `python tools/make_golden.py && python tools/make_pin.py` regenerates the committed vectors.
 *
 *   switch                              package / call site in the reference                         pin program
 *   LYS_PIN_HASH_SHIFT_ARITHMETIC       cpprandom 1.1.9 `hash` used by `split_rng`  integrator.fut:109   pin_rand.fut: split_first
 *   LYS_PIN_SEED_THEN_RAND              cpprandom `rng_from_seed`                    lib.fut:95           pin_rand.fut: seed_stream
 *   LYS_PIN_UNIFORM_MIN / _MAX          cpprandom `uniform_real_distribution.rand`   rand.fut:8,12,16,22  pin_rand.fut: unit_bits
 *   LYS_PIN_PROBIT_ACKLAM_F64           statistics 0.1.6 `sample (mk_normal d) p`    camera.fut:78        pin_stat.fut: normal_bits / normal_values
 *   LYS_PIN_NORMALISE_BY_RECIPROCAL     vector 0.4.5 `normalise`                     linalg.fut:4         pin_vec.fut: normalise_bits
 *   LYS_PIN_ARGB_SCALE                  matte 0.1.1 `argb.from_rgba`                 lib.fut:188-189      pin_argb.fut: pack
 *   (sorts 0.3.10 `radix_sort_by_key` is only assumed to be a stable sort, bvh.fut:95-97: pin_sort.fut)
 */
#ifndef LYS_PINS_H
#define LYS_PINS_H

#include "lys_detmath.h"

/* cpprandom `hash` (stackoverflow 12996028) is declared on i32 (`hash (x: i32): i32`) and written with `>>`, which Futhark
 * defines as the ARITHMETIC shift for signed types (`>>>` is the logical one): after the first multiply the value is negative
 * for about half of the pixel indices and the two readings differ.  1 = i32 with arithmetic shifts (the source as remembered),
 * 0 = the stackoverflow original on unsigned int (logical shifts; what round 1 of this repository assumed). */
#ifndef LYS_PIN_HASH_SHIFT_ARITHMETIC
#define LYS_PIN_HASH_SHIFT_ARITHMETIC 1
#endif
/* rng_from_seed [s]: fold  s' = ((s' >> 16) ^ s') ^ (s ^ 0b1010101010101)  over the seeds from s' = 1 (u32, logical shift),
 * then 1 = the state after one `rand` of that value, 0 = the folded value itself. */
#ifndef LYS_PIN_SEED_THEN_RAND
#define LYS_PIN_SEED_THEN_RAND 1
#endif
/* uniform_real_distribution: lo + ((f32 x - f32 min) / (f32 max - f32 min)) * (hi - lo) with the engine's min / max.
 * minstd_rand in cpprandom: min = 0, max = m = 2^31 - 1 (C++'s std::minstd_rand has min 1, max m - 1). */
#ifndef LYS_PIN_UNIFORM_MIN
#define LYS_PIN_UNIFORM_MIN 0u
#endif
#ifndef LYS_PIN_UNIFORM_MAX
#define LYS_PIN_UNIFORM_MAX 2147483647u
#endif
/* statistics `sample (mk_normal {mu, sigma}) p` = mu + sigma * probit p; 1 = probit by Acklam's rational approximation
 * evaluated in f64 and rounded once (lys_detmath.h det_probitf).  There is no second implementation here: a failing
 * pin_stat.fut `normal_bits` with a passing `normal_values` means "same function, other rounding" (bits differ, images agree
 * to 1e-6 relative); both failing means another quantile algorithm. */
#ifndef LYS_PIN_PROBIT_ACKLAM_F64
#define LYS_PIN_PROBIT_ACKLAM_F64 1
#endif
/* vector `normalise v`: 1 = scale (1 / norm v) v (one division, three multiplications), 0 = v / norm v per component. */
#ifndef LYS_PIN_NORMALISE_BY_RECIPROCAL
#define LYS_PIN_NORMALISE_BY_RECIPROCAL 1
#endif
/* matte `argb.from_rgba r g b a`: each channel clamped to [0, 1], multiplied by this scale and truncated to u8. */
#ifndef LYS_PIN_ARGB_SCALE
#define LYS_PIN_ARGB_SCALE 255.0f
#endif

/* ---- cpprandom: minstd_rand = linear_congruential_engine u32 {a = 48271, c = 0, m = 2^31 - 1}, wrapping u32 arithmetic ---- */
LYS_HD uint32_t lys_pin_lcg(uint32_t s) { return (48271u * s) % 2147483647u; }
LYS_HD uint32_t lys_pin_split_hash(uint32_t x) {                     /* split_rng n rng = map (\i -> rng ^ hash i) (iota n) */
#if LYS_PIN_HASH_SHIFT_ARITHMETIC
    /* (x >> 16) on i32: the top 16 bits are copies of the sign bit */
#define LYS_PIN_SHR16(v) (((v) >> 16) | (((v) & 0x80000000u) ? 0xffff0000u : 0u))
#else
#define LYS_PIN_SHR16(v) ((v) >> 16)
#endif
    x = (LYS_PIN_SHR16(x) ^ x) * 0x45d9f3bu;
    x = (LYS_PIN_SHR16(x) ^ x) * 0x45d9f3bu;
    x = LYS_PIN_SHR16(x) ^ x;
#undef LYS_PIN_SHR16
    return x;
}
LYS_HD uint32_t lys_pin_rng_from_seed(int32_t seed) {
    uint32_t sp = 1u;
    sp = ((sp >> 16) ^ sp) ^ ((uint32_t)seed ^ 0x1555u);
#if LYS_PIN_SEED_THEN_RAND
    return lys_pin_lcg(sp);
#else
    return sp;
#endif
}
/* the value drawn from state x (x is the NEW state: for an LCG the output is the state) */
LYS_HD float lys_pin_uniform(uint32_t x, float lo, float hi) {
    const float xf = (float)x, mn = (float)LYS_PIN_UNIFORM_MIN, mx = (float)LYS_PIN_UNIFORM_MAX;      /* f32 (2^31 - 1) == 2^31 */
    const float xp = (xf - mn) / (mx - mn);
    return lo + xp * (hi - lo);
}
/* ---- vector: normalise ---- */
LYS_HD void lys_pin_normalise(float x, float y, float z, float len, float *ox, float *oy, float *oz) {
#if LYS_PIN_NORMALISE_BY_RECIPROCAL
    const float s = 1.0f / len;
    *ox = s * x; *oy = s * y; *oz = s * z;
#else
    *ox = x / len; *oy = y / len; *oz = z / len;
#endif
}
/* ---- matte: one channel of argb.from_rgba ---- */
LYS_HD uint32_t lys_pin_argb_channel(float v) {
    const float c = (v < 0.0f) ? 0.0f : ((v > 1.0f) ? 1.0f : v);
    const float q = c * LYS_PIN_ARGB_SCALE;
    return (q > 0.0f) ? (uint32_t)q : 0u;
}

#endif /* LYS_PINS_H */
