/* lys_device.cuh -- device-side primitives of the B200 tracer.
 *
 * Arithmetic follows the reference's Futhark sources expression by expression (same operand
 * order, f32, no FMA contraction: this translation unit is compiled with -fmad=false) so that
 * results can be compared bit-for-bit with the CPU restatement.  Transcendentals come from
 * include/lys_detmath.h.  Citations are reference paths relative to /root/reference/src.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/lys_detmath.h"
#include "../../include/lys_pins.h"      /* third-party package semantics: one shared definition each (oracle + device) */

#define LYS_D __device__ __forceinline__
/* The big shading routines.  They were out of line while k_shade was instruction-fetch bound (6000 SASS instructions with
 * a one-lane reflection path in every warp); with that path compacted (wavefront.cu) inlining wins again: no caller-saved
 * spills around the calls, 2643 -> 2794 Mpaths/s (profiles/README.md 4.5).  -DLYS_DN=... restores the calls for experiments. */
#ifndef LYS_DN
#define LYS_DN static __device__ __forceinline__
#endif
#define LYS_HDI __host__ __device__ __forceinline__
/* per-routine choice (experiments: -DLYS_DN_REFR=LYS_D ...) */
#ifndef LYS_DN_REFR
#define LYS_DN_REFR LYS_DN
#endif
#ifndef LYS_DN_UBER
#define LYS_DN_UBER LYS_DN
#endif
#ifndef LYS_DN_SREFL
#define LYS_DN_SREFL LYS_DN
#endif
#ifndef LYS_DN_RTERMS
#define LYS_DN_RTERMS LYS_DN
#endif

/* evict-first cache operators (ld.global.cs / st.global.cs) for path-state records, used by the traversal kernels of LARGE
 * scenes only: there the records of a bounce (read once, written once) would otherwise displace the traversal records from L2
 * (+2 % on the 1 M-triangle scene; -0.5 % on CornellBox when used everywhere, hence the template switch; a persisting
 * access-policy window on the records changed nothing: profiles/README.md 8.4) */
template <bool CS, class T> LYS_D T ld_state(const T *p) {
#ifdef __CUDACC__
    if (CS) return __ldcs(p);
#endif
    return *p;
}
template <bool CS, class T> LYS_D void st_state(T *p, T v) {
#ifdef __CUDACC__
    if (CS) { __stcs(p, v); return; }
#endif
    *p = v;
}
#define LYS_PI 3.14159265358979323846f
#define LYS_INV_PI (1.0f / LYS_PI)          /* linalg.fut:55 */
#define LYS_INF (lys_u2f(0x7f800000u))
#define LYS_MAX_PATH_LEN 16

/* ---- vec3: athas/vector mk_vspace_3d f32 (linalg.fut:4) ------------------------------- */
struct V3 { float x, y, z; };
LYS_HDI V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
LYS_HDI V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
LYS_HDI V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
LYS_HDI V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
LYS_HDI V3 operator*(float s, V3 v) { return v3(s * v.x, s * v.y, s * v.z); }       /* vec3.scale */
LYS_HDI float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LYS_HDI V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
LYS_HDI float quadrance(V3 v) { return dot(v, v); }
LYS_HDI float norm(V3 v) { return sqrtf(quadrance(v)); }
LYS_HDI V3 normalise(V3 v) { V3 r; lys_pin_normalise(v.x, v.y, v.z, norm(v), &r.x, &r.y, &r.z); return r; }    /* vector `normalise` (lys_pins.h) */
LYS_HDI V3 same_side(V3 dominant, V3 w) { return lys_sgnf(dot(dominant, w)) * w; }     /* linalg.fut:30-31 */
LYS_HDI V3 vmin3(V3 a, V3 b) { return v3(lys_fminf(a.x, b.x), lys_fminf(a.y, b.y), lys_fminf(a.z, b.z)); }
LYS_HDI V3 vmax3(V3 a, V3 b) { return v3(lys_fmaxf(a.x, b.x), lys_fmaxf(a.y, b.y), lys_fmaxf(a.z, b.z)); }
LYS_HDI float lerpf(float a, float b, float t) { return a + (b - a) * t; }             /* f32.lerp */

/* ---- boxes: shapes.fut:88-110 ---------------------------------------------------------- */
struct Box { V3 c, h; };                                                              /* {center, half_dims} */
LYS_HDI Box contain(Box b1, Box b2) {                                                 /* containing_aabb :96-101 */
    V3 mn = vmin3(b1.c - b1.h, b2.c - b2.h);
    V3 mx = vmax3(b1.c + b1.h, b2.c + b2.h);
    Box r; r.c = 0.5f * (mn + mx); r.h = mx - r.c; return r;
}
LYS_HDI Box point_box(V3 p) { Box b; b.c = p; b.h = v3(0.0f, 0.0f, 0.0f); return b; } /* :103-104 */
LYS_HDI Box triangle_box(V3 a, V3 b, V3 c) { return contain(point_box(a), contain(point_box(b), point_box(c))); } /* :106-110 */
LYS_HDI bool box_bits_equal(Box a, Box b) {
    return lys_f2u(a.c.x) == lys_f2u(b.c.x) && lys_f2u(a.c.y) == lys_f2u(b.c.y) && lys_f2u(a.c.z) == lys_f2u(b.c.z) &&
           lys_f2u(a.h.x) == lys_f2u(b.h.x) && lys_f2u(a.h.y) == lys_f2u(b.h.y) && lys_f2u(a.h.z) == lys_f2u(b.h.z);
}

/* ---- Morton codes: bvh.fut:45-73 ------------------------------------------------------- */
LYS_HDI uint32_t expand_bits10(uint32_t x) {
    x = (x * 0x00010001u) & 0xFF0000FFu;
    x = (x * 0x00000101u) & 0x0F00F00Fu;
    x = (x * 0x00000011u) & 0xC30C30C3u;
    x = (x * 0x00000005u) & 0x49249249u;
    return x;
}
LYS_HDI uint32_t trunc_u32(float x) { return (x > 0.0f) ? (uint32_t)x : 0u; }            /* u32.f32 */
LYS_HDI uint32_t morton30(V3 v) {
    const float top = 1023.0f;
    V3 s = vmin3((top + 1.0f) * v, v3(top, top, top));
    return expand_bits10(trunc_u32(s.x)) * 4u + expand_bits10(trunc_u32(s.y)) * 2u + expand_bits10(trunc_u32(s.z));
}

/* ---- RNG: rand.fut + cpprandom minstd_rand (wrapping-u32 LCG) -------------------------- */
LYS_HDI uint32_t lcg_next(uint32_t &s) { s = lys_pin_lcg(s); return s; }
LYS_HDI float lcg_uniform(uint32_t &s, float lo, float hi) {
    return lys_pin_uniform(lcg_next(s), lo, hi);                 /* uniform_real_distribution (lys_pins.h) */
}
LYS_HDI void rng_advance(uint32_t &s) { (void)lcg_next(s); }                               /* rand.fut:11-12 */
LYS_HDI float rng_unit(uint32_t &s) { return lcg_uniform(s, 0.0f, 0.9999f); }              /* rand.fut:15-16 */
LYS_HDI uint32_t rng_split_hash(uint32_t x) { return lys_pin_split_hash(x); }              /* cpprandom hash of split_rng (lys_pins.h) */
LYS_HDI V3 rng_unit_disk(uint32_t &s) {                                                    /* rand.fut:21-25 */
    float theta = lcg_uniform(s, 0.0f, 2.0f * LYS_PI);
    float u = rng_unit(s);
    float r = sqrtf(u);
    return r * v3(det_cosf(theta), det_sinf(theta), 0.0f);
}

/* ---- spectrum.fut:30-49 ---------------------------------------------------------------- */
LYS_HDI float spectrum_lookup12(float v, const float *k /* 6 x (wavelength, value) */) {
    float wb = -1.0f, xb = 0.0f, wa = LYS_INF, xa = 0.0f;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float w = k[2 * i], x = k[2 * i + 1];
        if (w > wb && w <= v) { wb = w; xb = x; }
        else if (w < wa && w > v) { wa = w; xa = x; }
    }
    bool nb = wb < 0.0f, na = lys_isinff(wa);
    if (nb && na) return 0.0f;
    if (nb) return xa;
    if (na) return xb;
    return lerpf(xb, xa, (v - wb) / (wa - wb));
}

/* ---- ray / triangle: shapes.fut -------------------------------------------------------- */
struct Hit { float t; V3 pos, n; };
/* hit_triangle (shapes.fut:66-86) on a triangle stored as vertex a and edges e1 = b - a, e2 = c - a */
LYS_D bool tri_test(V3 o, V3 d, V3 a, V3 e1, V3 e2, float tmax, float &t_out, V3 &ncross) {
    /* The reference computes (t, u, v) = (1/det) * (n.s, m.e2, -(m.e1)) and then tests
     * `u >= 0 && v >= 0 && u + v <= 1 && t < tmax && t > 0`.  The values and the boolean below are identical; only
     * the order of evaluation is chosen so that the common rejections leave early:
     *  - t = fl(inv * dn) with sign(inv) == sign(det): if dn == 0 or the signs differ, t is <= 0 or -0 -> reject
     *    before the IEEE division;
     *  - t is tested against (0, tmax) before u and v are computed. */
    V3 n = cross(e1, e2);
    float det = -(dot(n, d));
    if (det > -0.00001f && det < 0.00001f) return false;                                /* approx_zero common.fut:35 */
    V3 s = o - a;
    float dn = dot(n, s);
    if (!((dn > 0.0f && det > 0.0f) || (dn < 0.0f && det < 0.0f))) return false;        /* t > 0 impossible */
    float inv = 1.0f / det;
    float t = inv * dn;
    if (!(t < tmax && t > 0.0f)) return false;                                          /* in_bounds :64 */
    V3 m = cross(s, d);
    float u = inv * dot(m, e2);
    float v = inv * (-(dot(m, e1)));
    if (!(u >= 0.0f && v >= 0.0f && u + v <= 1.0f)) return false;
    t_out = t; ncross = n;
    return true;
}
/* The same test split for the traversal loop: the plane part on (a, n = e1 x e2 precomputed by the build with the
 * same three products and differences), then the barycentric part on the edges.  Same values, same boolean. */
LYS_D bool tri_plane_test(V3 o, V3 d, V3 a, V3 n, float tmax, float &t_out, float &inv_out, V3 &s_out) {
    float det = -(dot(n, d));
    if (det > -0.00001f && det < 0.00001f) return false;                                /* approx_zero common.fut:35 */
    V3 s = o - a;
    float dn = dot(n, s);
    if (!((dn > 0.0f && det > 0.0f) || (dn < 0.0f && det < 0.0f))) return false;        /* t > 0 impossible */
    float inv = 1.0f / det;
    float t = inv * dn;
    if (!(t < tmax && t > 0.0f)) return false;                                          /* in_bounds :64 */
    t_out = t; inv_out = inv; s_out = s;
    return true;
}
LYS_D bool tri_uv_test(V3 d, V3 s, float inv, V3 e1, V3 e2) {
    V3 m = cross(s, d);
    float u = inv * dot(m, e2);
    float v = inv * (-(dot(m, e1)));
    return u >= 0.0f && v >= 0.0f && u + v <= 1.0f;
}
/* mkray_adjust_acne (shapes.fut:41-46) */
LYS_D void ray_from_hit(V3 pos, V3 n, V3 wi, V3 &o, V3 &d) {
    o = pos + 0.001f * same_side(wi, n);
    d = normalise(wi);
}

/* ---- material.fut ----------------------------------------------------------------------- */
struct Mat1 { float color, roughness, metalness, ref_ix, opacity; };                   /* material' :25-30 */
enum { PDF_DELTA = 0, PDF_IMPOSSIBLE = 1, PDF_NONZERO = 2 };                            /* :45-54 */
struct DirSample { V3 wi; float bsdf; int kind; float pdf; };
struct Onb { V3 t, b, n; };

LYS_D Mat1 material_at(const float *row /* 28 floats */, float wavelen) {               /* :32-42 */
    Mat1 m;
    m.color = spectrum_lookup12(wavelen, row);
    m.roughness = row[12]; m.metalness = row[13];
    m.ref_ix = row[14] - (wavelen - 589.0f) / 10000.0f;
    m.opacity = row[15];
    return m;
}
LYS_D Onb make_onb(V3 n) {                                                              /* :374-379 */
    Onb o;
    o.b = (lys_fabsf(n.x) > lys_fabsf(n.z)) ? normalise(v3(-n.y, n.x, 0.0f)) : normalise(v3(0.0f, -n.z, n.y));
    o.t = cross(o.b, n); o.n = n;
    return o;
}
LYS_D V3 to_local(const Onb &o, V3 w) { return v3(dot(w, o.t), dot(w, o.b), dot(w, o.n)); }   /* :381-384 */
LYS_D V3 to_world(const Onb &o, V3 w) { return (w.x * o.t + w.y * o.b) + w.z * o.n; }         /* :388-391 */

LYS_D float sin2_theta(V3 w) { return lys_fmaxf(0.0f, 1.0f - w.z * w.z); }                     /* :70-71 */
LYS_D bool same_hemi(V3 a, V3 b) { return a.z * b.z > 0.0f; }                                  /* :85-86 */
LYS_D V3 reflect_about(V3 w, V3 n) { return (-1.0f) * w + (2.0f * dot(w, n)) * n; }            /* :90-91 */
LYS_D float beckmann_alpha(float roughness) { return 1.62142f * lys_fmaxf(0.004f, roughness); } /* :241-248 */
LYS_D float beckmann_d(float alpha, V3 wh) {                                                   /* :218-223 */
    float c2 = wh.z * wh.z;
    float t2 = sin2_theta(wh) / c2;
    if (lys_isinff(t2)) return 0.0f;
    return det_expf(-t2 / (alpha * alpha)) / (LYS_PI * alpha * alpha * c2 * c2);
}
LYS_D float beckmann_lambda(float alpha, V3 w) {                                               /* :231-238 */
    float at = lys_fabsf(sqrtf(sin2_theta(w)) / w.z);
    if (lys_isinff(at)) return 0.0f;
    float a = 1.0f / (alpha * at);
    if (a >= 1.6f) return 0.0f;
    return (1.0f - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
LYS_D float schlick(V3 wo, const Mat1 &m) {                                                    /* :207-211 */
    float x = (1.0f - m.ref_ix) / (1.0f + m.ref_ix);
    float r0 = x * x;
    return r0 + (1.0f - r0) * det_pow5f(1.0f - wo.z);
}
/* Torrance-Sparrow reflection value (:264-266) and its pdf (:298-302), sharing D(wh). */
LYS_DN_RTERMS void reflection_terms(V3 wo, V3 wi, const Mat1 &m, float &bsdf, float &pdf) {
    float alpha = beckmann_alpha(m.roughness);
    V3 wh = normalise(wi + wo);
    float D = beckmann_d(alpha, wh);
    float G = 1.0f / (1.0f + beckmann_lambda(alpha, wo) + beckmann_lambda(alpha, wi));       /* :229-239 */
    bsdf = (D * G) / (4.0f * wo.z * wi.z);
    pdf = same_hemi(wo, wi) ? (D * lys_fabsf(wh.z)) / (4.0f * dot(wo, wh)) : 0.0f;
}
/* uber_bsdf (:357-358) and uber_pdf (:360-361, operands as written in the reference) in local space */
/* have_F: F_pre = schlick(wo, m) computed by the caller (it depends on the vertex only and is needed up to three times) */
LYS_DN_UBER void uber_eval(V3 wo, V3 wi, const Mat1 &m, float &f, float &pdf, bool have_F = false, float F_pre = 0.0f) {
    float refl_f, refl_pdf;
    reflection_terms(wo, wi, m, refl_f, refl_pdf);
    float refr_f = lerpf(0.0f, m.color * LYS_INV_PI, m.opacity);                              /* :187-188 */
    float diff_pdf = same_hemi(wo, wi) ? wi.z * LYS_INV_PI : 0.0f;                             /* :117-120 */
    float refr_pdf = lerpf(0.0f, diff_pdf, m.opacity);                                         /* :190-193 */
    bool inside = wo.z <= 0.0f;
    float F = inside ? 0.0f : (have_F ? F_pre : schlick(wo, m));
    float diel_f = lerpf(refr_f, refl_f, F);                                                   /* :317-323 */
    float diel_pdf = inside ? refr_pdf : lerpf(refr_pdf, refl_pdf, F);                         /* :325-330 */
    f = lerpf(diel_f, m.color * refl_f, m.metalness);
    pdf = lerpf(refl_pdf, diel_pdf, m.metalness);
}
/* dielectric_reflection_sample_dir (:305-315), with sample_wh (:283-296) */
LYS_DN_SREFL DirSample sample_reflection(V3 wo, const Mat1 &m, uint32_t &rng) {
    float u0 = rng_unit(rng), u1 = rng_unit(rng);
    float ls = det_logf(1.0f - u0);
    V3 wh; float pdf_wh;
    float alpha = beckmann_alpha(m.roughness);
    if (lys_isinff(ls)) { wh = v3(0.0f, 0.0f, 0.0f); pdf_wh = 0.0f; }
    else {
        float tan2 = -alpha * alpha * ls;
        float phi = u1 * 2.0f * LYS_PI;
        float ct = 1.0f / sqrtf(1.0f + tan2);
        float st = sqrtf(lys_fmaxf(0.0f, 1.0f - ct * ct));
        wh = v3(st * det_cosf(phi), st * det_sinf(phi), ct);                                   /* :269-272 */
        if (!same_hemi(wo, wh)) wh = -wh;
        pdf_wh = beckmann_d(alpha, wh) * lys_fabsf(ct);
    }
    V3 wi = reflect_about(wo, wh);
    DirSample s;
    if (!same_hemi(wo, wi)) { s.wi = v3(0.0f, 0.0f, 0.0f); s.bsdf = 0.0f; s.kind = PDF_IMPOSSIBLE; s.pdf = 0.0f; return s; }
    if (pdf_wh > 0.0f) { s.kind = PDF_NONZERO; s.pdf = pdf_wh / (4.0f * dot(wo, wh)); }
    else { s.kind = PDF_IMPOSSIBLE; s.pdf = 0.0f; }
    float rf, rp; reflection_terms(wo, wi, m, rf, rp);
    s.wi = wi; s.bsdf = rf;
    return s;
}
/* dielectric_refraction_sample_dir (:195-200): Lambert (:106-129) or delta transmission (:132-183) */
LYS_DN_REFR DirSample sample_refraction(V3 wo, const Mat1 &m, uint32_t &rng) {
    DirSample s;
    float p = rng_unit(rng);
    if (p < m.opacity) {
        V3 d = rng_unit_disk(rng);
        float z = sqrtf(lys_fmaxf(0.0f, 1.0f - (d.x * d.x + d.y * d.y)));
        s.wi = v3(d.x, d.y, z); s.bsdf = m.color * LYS_INV_PI; s.kind = PDF_NONZERO; s.pdf = z * LYS_INV_PI;
        return s;
    }
    bool entering = wo.z > 0.0f;
    V3 n = entering ? v3(0.0f, 0.0f, 1.0f) : v3(-0.0f, -0.0f, -1.0f);
    float eta = entering ? (1.0f / m.ref_ix) : (m.ref_ix / 1.0f);
    float ci = dot(n, wo);
    float s2i = lys_fmaxf(0.0f, 1.0f - ci * ci);
    float s2t = eta * eta * s2i;
    V3 wi;
    if (s2t >= 1.0f) wi = reflect_about(wo, n);
    else { float ctt = sqrtf(1.0f - s2t); wi = (-eta) * wo + (eta * ci - ctt) * n; }
    s.wi = wi; s.bsdf = 1.0f / lys_fabsf(wi.z); s.kind = PDF_DELTA; s.pdf = 0.0f;
    return s;
}
/* The draws of uber_sample_dir (:365-370) that come before its branch: metal (:352-355), else Fresnel coin (:338-344).
 * Returns true for the reflection branch; rng is left where the chosen sampler starts. */
LYS_D bool bsdf_choose(V3 wo, const Mat1 &m, uint32_t &rng, bool &metal, bool have_F = false, float F_pre = 0.0f) {
    float p = rng_unit(rng);
    metal = p < m.metalness;
    if (metal) return true;
    if (wo.z <= 0.0f) return false;
    float r = have_F ? F_pre : schlick(wo, m); float q = rng_unit(rng);
    return q < r;
}
/* sample_dir (:406-410) = uber_sample_dir (:365-370) in the local frame */
LYS_DN DirSample sample_bsdf(V3 wo_world, const Onb &onb, const Mat1 &m, uint32_t &rng) {
    V3 wo = to_local(onb, wo_world);
    DirSample s;
    float p = rng_unit(rng);
    bool metal = p < m.metalness;                                                              /* metal :352-355 */
    bool reflect;
    if (metal) reflect = true;
    else if (wo.z <= 0.0f) reflect = false;                                                    /* :338-339 */
    else { float r = schlick(wo, m); float q = rng_unit(rng); reflect = q < r; }               /* :340-344 */
    if (reflect) { s = sample_reflection(wo, m, rng); if (metal) s.bsdf = m.color * s.bsdf; }
    else s = sample_refraction(wo, m, rng);
    s.wi = to_world(onb, s.wi);
    return s;
}
