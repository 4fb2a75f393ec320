/* lys_oracle.cpp -- CPU restatement of the reference's Futhark program.
 *
 * TEST INFRASTRUCTURE ONLY (see lys_oracle.h).  PARITY UNPINNED (see lys_oracle.h).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  Semantics assumed for the Futhark `c` backend: `reduce`
 * is a sequential left fold, no FMA contraction, f32 arithmetic throughout.
 * Third-party package semantics (cpprandom 1.1.9, sorts 0.3.10, statistics 0.1.6,
 * vector 0.4.5, matte 0.1.1 -- pinned in futhark.pkg:2-6, sources NOT vendored) are
 * restated from their published algorithms; see DESIGN.md "Third-party semantics".
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 */
#include "lys_oracle.h"
#include "../include/lys_detmath.h"
#include "../include/lys_pins.h"       /* third-party package semantics: the ONE definition shared with the device code */

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* ------------------------------------------------------------------ knobs */
int g_path_len = 16;    /* integrator.fut:23 */
int g_refit_mode = 0;
int g_math_mode = 0;
int g_threads = 0;

constexpr int MAX_PATH_LEN = 16;
constexpr float F32_PI = 3.14159265358979323846f;   /* f32.pi */
constexpr float F32_HIGHEST = FLT_MAX;              /* f32.highest */
const float F32_INF = INFINITY;

/* transcendentals: deterministic contract or glibc (math_mode 1, for the
 * "does the contract change the picture" comparison only) */
inline float m_sin(float x) { return g_math_mode ? sinf(x) : det_sinf(x); }
inline float m_cos(float x) { return g_math_mode ? cosf(x) : det_cosf(x); }
inline float m_exp(float x) { return g_math_mode ? expf(x) : det_expf(x); }
inline float m_log(float x) { return g_math_mode ? logf(x) : det_logf(x); }
inline float m_pow5(float x) { return g_math_mode ? powf(x, 5.0f) : det_pow5f(x); }
inline float m_acos(float x) { return g_math_mode ? acosf(x) : det_acosf(x); }
inline float f_max(float a, float b) { return lys_fmaxf(a, b); }
inline float f_min(float a, float b) { return lys_fminf(a, b); }
/* f32.lerp v0 v1 t = v0 + (v1 - v0) * t */
inline float f_lerp(float a, float b, float t) { return a + (b - a) * t; }

/* ------------------------------------------------------------------ counters */
struct Counters { uint64_t paths = 0, vertices = 0, closest_rays = 0, shadow_rays = 0, node_visits = 0,
                  box_tests = 0, tri_tests = 0, loop_iters = 0, closest_box = 0, closest_tri = 0, shadow_box = 0, shadow_tri = 0; };
Counters g_counters;
thread_local Counters t_counters;
void flush_counters() {
#pragma omp critical(orc_counters)
    {
        g_counters.paths += t_counters.paths; g_counters.vertices += t_counters.vertices;
        g_counters.closest_rays += t_counters.closest_rays; g_counters.shadow_rays += t_counters.shadow_rays;
        g_counters.node_visits += t_counters.node_visits; g_counters.box_tests += t_counters.box_tests;
        g_counters.tri_tests += t_counters.tri_tests; g_counters.loop_iters += t_counters.loop_iters;
        g_counters.closest_box += t_counters.closest_box; g_counters.closest_tri += t_counters.closest_tri;
        g_counters.shadow_box += t_counters.shadow_box; g_counters.shadow_tri += t_counters.shadow_tri;
    }
    t_counters = Counters();
}

/* ------------------------------------------------------------------ linalg.fut / athas/vector vspace
 * mk_vspace_3d f32 (linalg.fut:4): elementwise + - * /, dot = x*x' + y*y' + z*z', quadrance = dot v v,
 * norm = sqrt quadrance, scale s v, normalise v = scale (1/norm v) v, cross, rot_z. */
struct vec3 { float x, y, z; };
inline vec3 mkvec3(float x, float y, float z) { return {x, y, z}; }
inline vec3 vadd(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 vsub(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 vdiv(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float vdot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 vcross(vec3 a, vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline vec3 vscale(float s, vec3 v) { return {s * v.x, s * v.y, s * v.z}; }
inline float vquadrance(vec3 v) { return vdot(v, v); }
inline float vnorm(vec3 v) { return sqrtf(vquadrance(v)); }
inline vec3 vnormalise(vec3 v) { vec3 r; lys_pin_normalise(v.x, v.y, v.z, vnorm(v), &r.x, &r.y, &r.z); return r; }
inline vec3 vneg(vec3 v) { return {-v.x, -v.y, -v.z}; }                       /* linalg.fut:19 */
inline vec3 same_side(vec3 dominant, vec3 w) { return vscale(lys_sgnf(vdot(dominant, w)), w); } /* linalg.fut:30-31 */
inline vec3 vmax(vec3 u, vec3 v) { return {f_max(u.x, v.x), f_max(u.y, v.y), f_max(u.z, v.z)}; } /* linalg.fut:35 */
inline vec3 vmin(vec3 u, vec3 v) { return {f_min(u.x, v.x), f_min(u.y, v.y), f_min(u.z, v.z)}; } /* linalg.fut:41 */
const vec3 world_up = {0, 1, 0};                                              /* linalg.fut:47 */
const float inv_pi = 1.0f / F32_PI;                                           /* linalg.fut:55 */
inline float clampf(float lo, float hi, float x) { return f_max(lo, f_min(hi, x)); } /* common.fut:37-38 */
inline bool approx_zero(float a, float eps) { return a > -eps && a < eps; }   /* common.fut:35 */

/* ------------------------------------------------------------------ rand.fut + cpprandom minstd_rand
 * linear_congruential_engine u32 {a=48271, c=0, m=2147483647}: rand x = (a*x + c) %% m in wrapping u32. */
typedef uint32_t rnge;
inline uint32_t rng_rand(rnge &s) { s = lys_pin_lcg(s); return s; }
/* uniform_real_distribution f32: x' = (f32 x - f32 min) / (f32 max - f32 min); lo + x' * (hi - lo) */
inline float dist_rand(rnge &s, float lo, float hi) {
    return lys_pin_uniform(rng_rand(s), lo, hi);
}
inline void advance_rng(rnge &s) { (void)dist_rand(s, 0.0f, 1.0f); }            /* rand.fut:11-12 */
inline float random_unit_exclusive(rnge &s) { return dist_rand(s, 0.0f, 0.9999f); } /* rand.fut:15-16 */
inline vec3 random_in_unit_disk(rnge &s) {                                      /* rand.fut:21-25 */
    float theta = dist_rand(s, 0.0f, 2.0f * F32_PI);
    float u = random_unit_exclusive(s);
    float r = sqrtf(u);
    return vscale(r, mkvec3(m_cos(theta), m_sin(theta), 0.0f));
}
inline void random_in_unit_square(rnge &s, float &x, float &y) {               /* rand.fut:28-31 */
    x = random_unit_exclusive(s); y = random_unit_exclusive(s);
}
inline void random_in_triangle(rnge &s, float &u, float &v) {                  /* rand.fut:34-37 */
    float a, b; random_in_unit_square(s, a, b);
    float su = sqrtf(a);
    u = 1.0f - su; v = b * su;
}
/* cpprandom hash (stackoverflow 12996028) as split_rng uses it: lys_pins.h (LYS_PIN_HASH_SHIFT_ARITHMETIC) */
inline uint32_t rng_hash(int32_t xi) { return lys_pin_split_hash((uint32_t)xi); }
/* rng_from_seed [seed]: seed' = fold ((s'>>16)^s') ^ (seed ^ 0b1010101010101) from 1; then one rand */
inline rnge rng_from_seed1(int32_t seed) {
    return lys_pin_rng_from_seed(seed);
}

/* ------------------------------------------------------------------ spectrum.fut */
struct spectrum { float w[6], x[6]; };
inline spectrum spectrum_from12(const float *p) { spectrum s; for (int i = 0; i < 6; i++) { s.w[i] = p[2 * i]; s.x[i] = p[2 * i + 1]; } return s; }
float spectrum_lookup(float v, const spectrum &s) {                            /* spectrum.fut:30-49 */
    float w_below = -1, x_below = 0, w_above = F32_INF, x_above = 0;
    for (int i = 0; i < 6; i++) {
        float w = s.w[i], x = s.x[i];
        if (w > w_below && w <= v) { w_below = w; x_below = x; }
        else if (w < w_above && w > v) { w_above = w; x_above = x; }
    }
    bool nb = w_below < 0, na = lys_isinff(w_above);
    if (nb && na) return 0;
    if (nb) return x_above;
    if (na) return x_below;
    return f_lerp(x_below, x_above, (v - w_below) / (w_above - w_below));
}
spectrum uniform_spectrum(float k) {                                           /* spectrum.fut:81-87 */
    spectrum s; s.w[0] = 0; s.x[0] = k; for (int i = 1; i < 6; i++) { s.w[i] = -1; s.x[i] = 0; } return s;
}
spectrum blackbody(float T) {                                                  /* spectrum.fut:64-72 (host libm) */
    const float c = 299792458.0f, h = 6.62606957e-34f, kb = 1.3806488e-23f;
    const float nm[6] = {150, 460, 550, 610, 1000, 2000};
    spectrum s;
    for (int i = 0; i < 6; i++) {
        float l = nm[i] * 1e-9f;
        float planck = (2 * h * c * c) / (powf(l, 5.0f) * (expf((h * c) / (l * kb * T)) - 1));
        s.w[i] = l * 1e9f; s.x[i] = planck;
    }
    return s;
}
spectrum blackbody_normalized(float T) {                                       /* spectrum.fut:74-79 */
    spectrum r = blackbody(T);
    float lambda_max = (2.8977721e-3f / T) * 1e9f;
    float mx = spectrum_lookup(lambda_max, r);
    for (int i = 0; i < 6; i++) r.x[i] = r.x[i] / mx;
    return r;
}
spectrum map_intensities_mul(spectrum s, float k) { for (int i = 0; i < 6; i++) s.x[i] = s.x[i] * k; return s; }
spectrum bright_blue_sky() { return map_intensities_mul(blackbody_normalized(17000.0f), 5.0f); } /* spectrum.fut:89 */

/* ------------------------------------------------------------------ shapes.fut */
struct ray { vec3 origin, dir; };
struct hit { float t; vec3 pos, normal; };
struct triangle { vec3 a, b, c; };
struct aabb { vec3 center, half_dims; };

inline ray mkray(vec3 o, vec3 d) { return {o, vnormalise(d)}; }                /* shapes.fut:37-38 */
inline ray mkray_adjust_acne(const hit &h, vec3 wi) {                          /* shapes.fut:41-46 */
    const float eps = 0.001f;
    vec3 acne_offset = vscale(eps, same_side(wi, h.normal));
    return mkray(vadd(h.pos, acne_offset), wi);
}
inline vec3 triangle_normal(const triangle &t) {                              /* shapes.fut:59-62 */
    return vnormalise(vcross(vsub(t.b, t.a), vsub(t.c, t.a)));
}
inline bool hit_triangle(float tmax, const ray &ra, const triangle &tr, hit &out) { /* shapes.fut:66-86 */
    t_counters.tri_tests++;
    const float eps = 0.00001f;
    vec3 e1 = vsub(tr.b, tr.a), e2 = vsub(tr.c, tr.a);
    vec3 n = vcross(e1, e2);
    float a = -(vdot(n, ra.dir));
    if (approx_zero(a, eps)) return false;
    vec3 s = vsub(ra.origin, tr.a);
    vec3 m = vcross(s, ra.dir);
    vec3 tuv = vscale(1.0f / a, mkvec3(vdot(n, s), vdot(m, e2), -(vdot(m, e1))));
    float t = tuv.x, u = tuv.y, v = tuv.z;
    bool in_triangle = u >= 0 && v >= 0 && u + v <= 1;
    bool in_bounds = t < tmax && t > 0;                                        /* shapes.fut:64 */
    if (!(in_triangle && in_bounds)) return false;
    out.t = t;
    out.pos = vadd(ra.origin, vscale(t, ra.dir));                              /* point_at_param shapes.fut:48-49 */
    out.normal = vnormalise(n);
    return true;
}
inline vec3 aabb_min_corner(const aabb &b) { return vsub(b.center, b.half_dims); }  /* shapes.fut:88-89 */
inline vec3 aabb_max_corner(const aabb &b) { return vadd(b.center, b.half_dims); }  /* shapes.fut:91-92 */
inline aabb containing_aabb(const aabb &b1, const aabb &b2) {                 /* shapes.fut:96-101 */
    vec3 mn = vmin(aabb_min_corner(b1), aabb_min_corner(b2));
    vec3 mx = vmax(aabb_max_corner(b1), aabb_max_corner(b2));
    vec3 center = vscale(0.5f, vadd(mn, mx));
    return {center, vsub(mx, center)};
}
inline aabb bounding_box_point(vec3 p) { return {p, mkvec3(0, 0, 0)}; }        /* shapes.fut:103-104 */
inline aabb bounding_box_triangle(const triangle &t) {                        /* shapes.fut:106-110 */
    return containing_aabb(bounding_box_point(t.a),
                           containing_aabb(bounding_box_point(t.b), bounding_box_point(t.c)));
}
inline bool hit_aabb(float tmax, const ray &r, const aabb &b) {               /* shapes.fut:114-135 */
    t_counters.box_tests++;
    const float eps = 0.001f;
    vec3 mn = aabb_min_corner(b), mx = aabb_max_corner(b);
    float tmin = 0;
    const float mns[3] = {mn.x, mn.y, mn.z}, mxs[3] = {mx.x, mx.y, mx.z};
    const float os[3] = {r.origin.x, r.origin.y, r.origin.z}, ds[3] = {r.dir.x, r.dir.y, r.dir.z};
    for (int k = 0; k < 3; k++) {
        float invD = 1.0f / ds[k];
        float t0 = (mns[k] - os[k]) * invD;
        float t1 = (mxs[k] - os[k]) * invD;
        if (invD < 0) { float tmp = t0; t0 = t1; t1 = tmp; }
        t1 = t1 * (1 + eps);
        tmin = f_max(t0, tmin);
        tmax = f_min(t1, tmax);
        if (tmax <= tmin) return false;
    }
    return true;
}
/* disk (shapes.fut:17-35): n_sectors sector triangles around p facing `normal`.
 * vec3.rot_z b (1,0,0) = (1*cos b - 0*sin b, 1*sin b + 0*cos b, 0); sector angles use host libm. */
void disk(vec3 p, vec3 normal, float radius, int n_sectors, triangle *out) {
    float a = 2 * F32_PI / (float)n_sectors;
    vec3 c = vcross(normal, world_up);
    vec3 right = (vnorm(c) == 0) ? mkvec3(1, 0, 0) : vnormalise(c);
    vec3 up = vnormalise(vcross(right, normal));
    for (int i = 0; i < n_sectors; i++) {
        float fi = (float)i;
        float b0 = a * fi, b1 = a * (fi + 1);
        auto angle_to_vec = [&](float b) {
            float x = 1.0f * cosf(b) - 0.0f * sinf(b);
            float y = 1.0f * sinf(b) + 0.0f * cosf(b);
            return vadd(vscale(x, right), vscale(y, up));
        };
        vec3 v0 = angle_to_vec(b0), v1 = angle_to_vec(b1);
        out[i] = {p, vadd(p, vscale(radius, v1)), vadd(p, vscale(radius, v0))};
    }
}

/* ------------------------------------------------------------------ radix_tree.fut */
/* ptr = #internal i | #leaf i  ->  i | ~i */
inline int32_t ptr_internal(int32_t i) { return i; }
inline int32_t ptr_leaf(int32_t i) { return ~i; }
inline bool ptr_is_leaf(int32_t p) { return p < 0; }
inline int32_t ptr_leaf_ix(int32_t p) { return ~p; }
/* `#internal (-1)`, the initial `prev` of the walks (bvh.fut:126,151): equal to no real pointer */
constexpr int32_t PREV_NONE = INT32_MIN;
inline int clz32(uint32_t x) { return x == 0 ? 32 : __builtin_clz(x); }

void radix_tree_mk(const uint32_t *L, int64_t n, int32_t *left, int32_t *right, int32_t *parent) { /* radix_tree.fut:21-89 */
    auto delta = [&](int32_t i, int32_t j) -> int32_t {                         /* :22-29 */
        if (j >= 0 && j < n) {
            uint32_t Li = L[i], Lj = L[j];
            if (Li == Lj) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
            return clz32(Li ^ Lj);
        }
        return -1;
    };
    int64_t n_nodes = n - 1;
    for (int64_t k = 0; k < n_nodes; k++) parent[k] = -1;                       /* :83 */
    std::vector<int32_t> lchild(n_nodes), rchild(n_nodes);
    for (int32_t i = 0; i < n_nodes; i++) {                                    /* mk_node :31-71 */
        int32_t dd = delta(i, i + 1) - delta(i, i - 1);
        int32_t d = (dd > 0) - (dd < 0);                                       /* i32.sgn */
        int32_t delta_min = delta(i, i - d);
        int32_t l_max = 2;
        while (delta(i, i + l_max * d) > delta_min) l_max *= 2;
        int32_t l = 0;
        for (int32_t t = l_max / 2; t >= 1; t /= 2)
            if (delta(i, i + (l + t) * d) > delta_min) l += t;
        int32_t j = i + l * d;
        int32_t delta_node = delta(i, j);
        int32_t s = 0;
        for (int32_t q = 1; q <= l; q *= 2) {
            int32_t t = (l + q * 2 - 1) / (q * 2);                             /* div_rounding_up :10 */
            if (delta(i, i + (s + t) * d) > delta_node) s += t;
        }
        int32_t gamma = i + s * d + std::min(d, 0);
        if (std::min(i, j) == gamma) { left[i] = ptr_leaf(gamma); lchild[i] = -1; }
        else { left[i] = ptr_internal(gamma); lchild[i] = gamma; }
        if (std::max(i, j) == gamma + 1) { right[i] = ptr_leaf(gamma + 1); rchild[i] = -1; }
        else { right[i] = ptr_internal(gamma + 1); rchild[i] = gamma + 1; }
    }
    /* scatter (:83-85): children = left ++ right, negative indices ignored */
    for (int32_t i = 0; i < n_nodes; i++) if (lchild[i] >= 0) parent[lchild[i]] = i;
    for (int32_t i = 0; i < n_nodes; i++) if (rchild[i] >= 0) parent[rchild[i]] = i;
}

/* ------------------------------------------------------------------ bvh.fut */
inline uint32_t expand_bits(uint32_t x) {                                      /* bvh.fut:52-57 */
    x = (x * 0x00010001u) & 0xFF0000FFu;
    x = (x * 0x00000101u) & 0x0F00F00Fu;
    x = (x * 0x00000011u) & 0xC30C30C3u;
    x = (x * 0x00000005u) & 0x49249249u;
    return x;
}
/* u32.f32: C cast; defined here as 0 for x <= 0 / NaN (the reference value is UB there) */
inline uint32_t u32_of_f32(float x) { return (x > 0.0f) ? (uint32_t)x : 0u; }
inline uint32_t morton3D(vec3 v) {                                             /* bvh.fut:67-73 */
    const float maxv = 1023.0f;                                                /* :45-48 */
    vec3 s = vmin(vscale(maxv + 1, v), mkvec3(maxv, maxv, maxv));
    uint32_t xx = expand_bits(u32_of_f32(s.x)), yy = expand_bits(u32_of_f32(s.y)), zz = expand_bits(u32_of_f32(s.z));
    return xx * 4 + yy * 2 + zz;
}

struct obj { triangle geom; uint32_t mat_ix; int32_t src_index; };            /* scene.fut:6 (+ provenance) */
struct node { aabb box; int32_t left, right, parent; };                       /* bvh.fut:76 */
struct bvh_t {                                                                 /* bvh.fut:79-84 */
    aabb bounds;
    std::vector<obj> leaves;
    std::vector<node> nodes;
    std::vector<aabb> leaf_aabbs;      /* sorted aabbs (kept for introspection) */
    std::vector<uint32_t> mortons;     /* sorted keys  (kept for introspection) */
};

void bvh_build(const std::vector<obj> &xs_in, bvh_t &out) {                    /* bvh.fut:86-121 */
    int64_t n = (int64_t)xs_in.size();
    std::vector<aabb> aabbs(n);
    for (int64_t i = 0; i < n; i++) aabbs[i] = bounding_box_triangle(xs_in[i].geom);   /* :87 */
    aabb bounds = {mkvec3(0, 0, 0), mkvec3(-F32_INF, -F32_INF, -F32_INF)};            /* :88-89 */
    for (int64_t i = 0; i < n; i++) bounds = containing_aabb(bounds, aabbs[i]);        /* :90 left fold */
    vec3 bmin = aabb_min_corner(bounds);
    vec3 bdim = vscale(2, bounds.half_dims);                                           /* shapes.fut:94 */
    std::vector<uint32_t> mortons(n);
    for (int64_t i = 0; i < n; i++)
        mortons[i] = morton3D(vdiv(vsub(aabbs[i].center, bmin), bdim));                /* :91-94 */
    /* radix_sort_by_key over all 32 bits, stable (:95-97) == stable sort by key */
    std::vector<int32_t> perm(n);
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return mortons[a] < mortons[b]; });
    out.leaves.resize(n); out.leaf_aabbs.resize(n); out.mortons.resize(n);
    for (int64_t i = 0; i < n; i++) {
        out.leaves[i] = xs_in[perm[i]]; out.leaf_aabbs[i] = aabbs[perm[i]]; out.mortons[i] = mortons[perm[i]];
    }
    int64_t n_nodes = n - 1;
    std::vector<int32_t> L(n_nodes), R(n_nodes), P(n_nodes);
    radix_tree_mk(out.mortons.data(), n, L.data(), R.data(), P.data());                /* :98 */
    std::vector<node> I(n_nodes);
    for (int64_t i = 0; i < n_nodes; i++) I[i] = {{mkvec3(0, 0, 0), mkvec3(0, 0, 0)}, L[i], R[i], P[i]}; /* :105-108 */
    int depth = (int32_t)(log2f((float)(int32_t)n)) + 2;                               /* :109 */
    auto get_aabb = [&](const std::vector<node> &inners, int32_t p) -> aabb {          /* :110-113 */
        return ptr_is_leaf(p) ? out.leaf_aabbs[ptr_leaf_ix(p)] : inners[p].box;
    };
    if (g_refit_mode == 1) depth = 1 << 30;
    std::vector<node> J(n_nodes);
    for (int it = 0; it < depth; it++) {                                               /* :118-120 Jacobi sweeps */
        bool changed = false;
        for (int64_t i = 0; i < n_nodes; i++) {
            J[i] = I[i];
            J[i].box = containing_aabb(get_aabb(I, I[i].left), get_aabb(I, I[i].right));
            if (memcmp(&J[i].box, &I[i].box, sizeof(aabb)) != 0) changed = true;
        }
        I.swap(J);
        if (!changed) break;      /* fixed point: further sweeps are identities */
    }
    out.bounds = bounds; out.nodes = I;
}

struct trav_hit { int32_t leaf; hit h; };
/* closest_hit (bvh.fut:123-145): stackless parent-pointer walk, always left first */
bool bvh_closest_hit(float tmax_in, const ray &r, const bvh_t &bvh, trav_hit &out) {
    t_counters.closest_rays++;
    const uint64_t box0 = t_counters.box_tests, tri0 = t_counters.tri_tests;
    int32_t closest = -1; float tmax = tmax_in; int32_t current = 0; int32_t prev = PREV_NONE;
    if (bvh.nodes.empty()) current = -1;
    while (current != -1) {
        t_counters.loop_iters++;
        const node &nd = bvh.nodes[current];
        int32_t rec_child; bool have = false;
        if (prev == nd.left) { rec_child = nd.right; have = true; }
        else if (prev != nd.right) { t_counters.node_visits++; if (hit_aabb(tmax, r, nd.box)) { rec_child = nd.left; have = true; } }
        if (!have) { prev = ptr_internal(current); current = nd.parent; continue; }
        if (!ptr_is_leaf(rec_child)) { prev = ptr_internal(current); current = rec_child; continue; }
        int32_t i = ptr_leaf_ix(rec_child);
        hit h;
        if (hit_triangle(tmax, r, bvh.leaves[i].geom, h)) { closest = i; tmax = h.t; }
        prev = rec_child;
    }
    t_counters.closest_box += t_counters.box_tests - box0; t_counters.closest_tri += t_counters.tri_tests - tri0;
    if (closest < 0) return false;
    out.leaf = closest;
    { const uint64_t keep = t_counters.tri_tests; bool ok = hit_triangle(tmax_in, r, bvh.leaves[closest].geom, out.h); t_counters.tri_tests = keep; return ok; } /* :143-145 (outer tmax); not counted as traversal work */
}
/* any_hit (bvh.fut:149-167) */
bool bvh_any_hit(float tmax, const ray &r, const bvh_t &bvh) {
    t_counters.shadow_rays++;
    const uint64_t box0 = t_counters.box_tests, tri0 = t_counters.tri_tests;
    bool found = false; int32_t current = 0; int32_t prev = PREV_NONE;
    if (bvh.nodes.empty()) current = -1;
    while (!found && current != -1) {
        t_counters.loop_iters++;
        const node &nd = bvh.nodes[current];
        int32_t rec_child; bool have = false;
        if (prev == nd.left) { rec_child = nd.right; have = true; }
        else if (prev != nd.right) { t_counters.node_visits++; if (hit_aabb(tmax, r, nd.box)) { rec_child = nd.left; have = true; } }
        if (!have) { prev = ptr_internal(current); current = nd.parent; continue; }
        if (!ptr_is_leaf(rec_child)) { prev = ptr_internal(current); current = rec_child; continue; }
        hit h;
        if (hit_triangle(tmax, r, bvh.leaves[ptr_leaf_ix(rec_child)].geom, h)) found = true;
        prev = rec_child;
    }
    t_counters.shadow_box += t_counters.box_tests - box0; t_counters.shadow_tri += t_counters.tri_tests - tri0;
    return found;
}

/* ------------------------------------------------------------------ material.fut */
struct material { spectrum color; float roughness, metalness, ref_ix, opacity; spectrum emission; }; /* :12-18 */
struct material1 { float color, roughness, metalness, ref_ix, opacity; };                            /* material' :25-30 */
struct interaction { hit h; material mat; float wavelen; };                                           /* :22 */
enum pdf_kind { PDF_DELTA = 0, PDF_IMPOSSIBLE = 1, PDF_NONZERO = 2 };                                 /* :45-54 */
struct dir_sample { vec3 wi; float bsdf; pdf_kind kind; float pdf; };                                 /* :56 */
const dir_sample null_sample = {{0, 0, 0}, 0, PDF_IMPOSSIBLE, 0};                                     /* :58-59 */

material parse_mat(const float *m) {                                          /* scene.fut:37-53 */
    material r; r.color = spectrum_from12(m); r.roughness = m[12]; r.metalness = m[13];
    r.ref_ix = m[14]; r.opacity = m[15]; r.emission = spectrum_from12(m + 16); return r;
}
inline material1 material_at_wavelen(const material &m, float wavelen) {      /* material.fut:32-42 */
    material1 r; r.color = spectrum_lookup(wavelen, m.color); r.roughness = m.roughness; r.metalness = m.metalness;
    float delta = wavelen - 589.0f;
    r.ref_ix = m.ref_ix - delta / 10000.0f; r.opacity = m.opacity; return r;
}
inline float cos_theta(vec3 w) { return w.z; }                                 /* :69-74 */
inline float cos2_theta(vec3 w) { return w.z * w.z; }
inline float sin2_theta(vec3 w) { return f_max(0, 1 - cos2_theta(w)); }
inline float sin_theta(vec3 w) { return sqrtf(sin2_theta(w)); }
inline float tan_theta(vec3 w) { return sin_theta(w) / cos_theta(w); }
inline float tan2_theta(vec3 w) { return sin2_theta(w) / cos2_theta(w); }
inline bool same_hemisphere(vec3 w, vec3 u) { return w.z * u.z > 0; }          /* :85-86 */
inline vec3 reflect(vec3 w, vec3 n) { return vadd(vscale(-1, w), vscale(2 * vdot(w, n), n)); } /* :90-91 */
inline vec3 cosine_sample_hemisphere(rnge &rng) {                             /* :106-112 */
    vec3 d = random_in_unit_disk(rng);
    float sin2theta = d.x * d.x + d.y * d.y;
    float cos2theta = f_max(0, 1 - sin2theta);
    return mkvec3(d.x, d.y, sqrtf(cos2theta));
}
inline float diffuse_bsdf(float color) { return color * inv_pi; }              /* :114-115 */
inline float diffuse_pdf(vec3 wo, vec3 wi) { return same_hemisphere(wo, wi) ? cos_theta(wi) * inv_pi : 0; } /* :117-120 */
inline dir_sample diffuse_sample_dir(const material1 &m, rnge &rng) {          /* :125-129 */
    vec3 wi = cosine_sample_hemisphere(rng);
    return {wi, diffuse_bsdf(m.color), PDF_NONZERO, cos_theta(wi) * inv_pi};
}
inline vec3 refract(vec3 wi, vec3 n, float eta) {                              /* :132-142 (both cases return a vec3) */
    float cos_theta_i = vdot(n, wi);
    float sin2_theta_i = f_max(0, 1 - cos_theta_i * cos_theta_i);
    float sin2_theta_t = eta * eta * sin2_theta_i;
    if (sin2_theta_t >= 1) return reflect(wi, n);
    float cos_theta_t = sqrtf(1 - sin2_theta_t);
    return vadd(vscale(-eta, wi), vscale(eta * cos_theta_i - cos_theta_t, n));
}
inline dir_sample transmission_sample_dir(vec3 wo, const material1 &m) {      /* :166-183 */
    bool entering = cos_theta(wo) > 0;
    const vec3 local_normal = {0, 0, 1};
    vec3 n = entering ? local_normal : vneg(local_normal);
    float eta = entering ? (1.0f / m.ref_ix) : (m.ref_ix / 1.0f);
    vec3 wi = refract(wo, n, eta);
    return {wi, 1 / lys_fabsf(cos_theta(wi)), PDF_DELTA, 0};
}
inline float dielectric_refraction_bsdf(const material1 &m) { return f_lerp(0, diffuse_bsdf(m.color), m.opacity); } /* :187-188 */
inline float dielectric_refraction_pdf(vec3 wo, vec3 wh, const material1 &m) { return f_lerp(0, diffuse_pdf(wo, wh), m.opacity); } /* :190-193 */
inline dir_sample dielectric_refraction_sample_dir(vec3 wo, const material1 &m, rnge &rng) { /* :195-200 */
    float p = random_unit_exclusive(rng);
    if (p < m.opacity) return diffuse_sample_dir(m, rng);
    return transmission_sample_dir(wo, m);
}
inline float fresnel_reflectance(vec3 wo, const material1 &m) {               /* :207-211 */
    float ix_1 = 1, ix_2 = m.ref_ix;
    float x = (ix_1 - ix_2) / (ix_1 + ix_2);
    float r0 = x * x;
    return r0 + (1 - r0) * m_pow5(1 - cos_theta(wo));
}
inline float microfacet_distribution(float alpha, vec3 wh) {                  /* :218-223 */
    float t2 = tan2_theta(wh);
    if (lys_isinff(t2)) return 0;
    return m_exp(-t2 / (alpha * alpha)) / (F32_PI * alpha * alpha * cos2_theta(wh) * cos2_theta(wh));
}
inline float ssf_lambda(float alpha, vec3 w) {                                 /* :231-238 */
    float abs_tan_theta = lys_fabsf(tan_theta(w));
    if (lys_isinff(abs_tan_theta)) return 0;
    float a = 1 / (alpha * abs_tan_theta);
    if (a >= 1.6f) return 0;
    return (1 - 1.259f * a + 0.396f * a * a) / (3.535f * a + 2.181f * a * a);
}
inline float self_shadowing_factor(float alpha, vec3 wo, vec3 wi) { return 1 / (1 + ssf_lambda(alpha, wo) + ssf_lambda(alpha, wi)); } /* :229-239 */
inline float beckmann_alpha(float roughness) { return 1.62142f * f_max(0.004f, roughness); } /* :241-248 */
inline float microfacet_factor(vec3 wo, vec3 wi, const material1 &m) {        /* :252-256 */
    vec3 wh = vnormalise(vadd(wi, wo));
    float alpha = beckmann_alpha(m.roughness);
    return microfacet_distribution(alpha, wh) * self_shadowing_factor(alpha, wo, wi);
}
inline float dielectric_reflection_bsdf(vec3 wo, vec3 wi, const material1 &m) { /* :264-266 */
    return microfacet_factor(wo, wi, m) / (4 * cos_theta(wo) * cos_theta(wi));
}
inline float dielectric_reflection_wh_pdf(vec3 wh, const material1 &m) {      /* :274-276 */
    float alpha = beckmann_alpha(m.roughness);
    return microfacet_distribution(alpha, wh) * lys_fabsf(cos_theta(wh));
}
inline void dielectric_reflection_sample_wh(vec3 wo, const material1 &m, rnge &rng, vec3 &wh, float &pdf_wh) { /* :283-296 */
    float u0, u1; random_in_unit_square(rng, u0, u1);
    float log_sample = m_log(1 - u0);
    if (lys_isinff(log_sample)) { wh = mkvec3(0, 0, 0); pdf_wh = 0; return; }
    float alpha = beckmann_alpha(m.roughness);
    float tan2 = -alpha * alpha * log_sample;
    float phi = u1 * 2 * F32_PI;
    float ct = 1 / sqrtf(1 + tan2);
    float st = sqrtf(f_max(0, 1 - ct * ct));
    wh = mkvec3(st * m_cos(phi), st * m_sin(phi), ct);                          /* spherical_direction :269-272 */
    if (!same_hemisphere(wo, wh)) wh = vneg(wh);
    pdf_wh = microfacet_distribution(alpha, wh) * lys_fabsf(ct);
}
inline float dielectric_reflection_pdf(vec3 wo, vec3 wi, const material1 &m) { /* :298-302 */
    if (!same_hemisphere(wo, wi)) return 0;
    vec3 wh = vnormalise(vadd(wo, wi));
    return dielectric_reflection_wh_pdf(wh, m) / (4 * vdot(wo, wh));
}
inline dir_sample dielectric_reflection_sample_dir(vec3 wo, const material1 &m, rnge &rng) { /* :305-315 */
    vec3 wh; float pdf_wh; dielectric_reflection_sample_wh(wo, m, rng, wh, pdf_wh);
    vec3 wi = reflect(wo, wh);
    dir_sample s;
    if (pdf_wh > 0) { s.kind = PDF_NONZERO; s.pdf = pdf_wh / (4 * vdot(wo, wh)); } else { s.kind = PDF_IMPOSSIBLE; s.pdf = 0; }
    if (!same_hemisphere(wo, wi)) return null_sample;
    s.wi = wi; s.bsdf = dielectric_reflection_bsdf(wo, wi, m); return s;
}
inline float dielectric_bsdf(vec3 wo, vec3 wi, const material1 &m) {          /* :317-323 */
    float reflectance = (cos_theta(wo) <= 0) ? 0 : fresnel_reflectance(wo, m);
    return f_lerp(dielectric_refraction_bsdf(m), dielectric_reflection_bsdf(wo, wi, m), reflectance);
}
inline float dielectric_pdf(vec3 wo, vec3 wi, const material1 &m) {           /* :325-330 */
    if (cos_theta(wo) <= 0) return dielectric_refraction_pdf(wo, wi, m);
    return f_lerp(dielectric_refraction_pdf(wo, wi, m), dielectric_reflection_pdf(wo, wi, m), fresnel_reflectance(wo, m));
}
inline dir_sample dielectric_sample_dir(vec3 wo, const material1 &m, rnge &rng) { /* :336-344 */
    if (cos_theta(wo) <= 0) return dielectric_refraction_sample_dir(wo, m, rng);
    float r = fresnel_reflectance(wo, m);
    float p = random_unit_exclusive(rng);
    if (p < r) return dielectric_reflection_sample_dir(wo, m, rng);
    return dielectric_refraction_sample_dir(wo, m, rng);
}
inline float metal_bsdf(vec3 wo, vec3 wi, const material1 &m) { return m.color * dielectric_reflection_bsdf(wo, wi, m); } /* :346-347 */
inline float metal_pdf(vec3 wo, vec3 wi, const material1 &m) { return dielectric_reflection_pdf(wo, wi, m); }            /* :349-350 */
inline dir_sample metal_sample_dir(vec3 wo, const material1 &m, rnge &rng) {  /* :352-355 */
    dir_sample s = dielectric_reflection_sample_dir(wo, m, rng); s.bsdf = m.color * s.bsdf; return s;
}
inline float uber_bsdf(vec3 wo, vec3 wi, const material1 &m) { return f_lerp(dielectric_bsdf(wo, wi, m), metal_bsdf(wo, wi, m), m.metalness); } /* :357-358 */
inline float uber_pdf(vec3 wo, vec3 wi, const material1 &m) { return f_lerp(metal_pdf(wo, wi, m), dielectric_pdf(wo, wi, m), m.metalness); }   /* :360-361 (operands as written) */
inline dir_sample uber_sample_dir(vec3 wo, const material1 &m, rnge &rng) {    /* :365-370 */
    float p = random_unit_exclusive(rng);
    if (p < m.metalness) return metal_sample_dir(wo, m, rng);
    return dielectric_sample_dir(wo, m, rng);
}
struct onb_t { vec3 tangent, binormal, normal; };                               /* :372 */
inline onb_t mk_orthonormal_basis(vec3 normal) {                              /* :374-379 */
    vec3 binormal = (lys_fabsf(normal.x) > lys_fabsf(normal.z))
        ? vnormalise(mkvec3(-normal.y, normal.x, 0)) : vnormalise(mkvec3(0, -normal.z, normal.y));
    return {vcross(binormal, normal), binormal, normal};
}
inline vec3 world_to_local(const onb_t &o, vec3 w) { return mkvec3(vdot(w, o.tangent), vdot(w, o.binormal), vdot(w, o.normal)); } /* :381-384 */
inline vec3 local_to_world(const onb_t &o, vec3 w) {                          /* :388-391 */
    return vadd(vadd(vscale(w.x, o.tangent), vscale(w.y, o.binormal)), vscale(w.z, o.normal));
}
inline float bsdf_f(vec3 wo, vec3 wi, const interaction &i) {                 /* :393-396 */
    onb_t o = mk_orthonormal_basis(i.h.normal);
    return uber_bsdf(world_to_local(o, wo), world_to_local(o, wi), material_at_wavelen(i.mat, i.wavelen));
}
inline float bsdf_pdf(vec3 wo, vec3 wi, const interaction &i) {               /* :398-401 */
    onb_t o = mk_orthonormal_basis(i.h.normal);
    return uber_pdf(world_to_local(o, wo), world_to_local(o, wi), material_at_wavelen(i.mat, i.wavelen));
}
inline dir_sample sample_dir(vec3 wo, const interaction &i, rnge &rng) {      /* :406-410 */
    onb_t o = mk_orthonormal_basis(i.h.normal);
    vec3 wol = world_to_local(o, wo);
    dir_sample s = uber_sample_dir(wol, material_at_wavelen(i.mat, i.wavelen), rng);
    s.wi = local_to_world(o, s.wi);
    return s;
}

/* ------------------------------------------------------------------ light.fut */
enum light_kind { LIGHT_DIFFUSE = 0, LIGHT_FRUSTUM = 1 };
struct light { light_kind kind; triangle geom; float theta; spectrum emission; int32_t src_index; }; /* light.fut:4-11 */
inline float arealight_incident_radiance(const light &l, vec3 hitp, vec3 lightp, float wavelen) { /* light.fut:19-55 */
    vec3 v = vsub(lightp, hitp);
    vec3 wi = vnormalise(v); float distance_sq = vquadrance(v);
    vec3 lnormal = triangle_normal(l.geom);
    float cos_theta_l = vdot(vneg(wi), lnormal);
    if (l.kind == LIGHT_DIFFUSE)
        return f_max(0, spectrum_lookup(wavelen, l.emission) * cos_theta_l / distance_sq);
    if (m_acos(cos_theta_l) <= l.theta) return spectrum_lookup(wavelen, l.emission) / distance_sq;
    return 0;
}

/* ------------------------------------------------------------------ scene.fut */
struct scene_t {                                                               /* accel_scene scene.fut:20-23 */
    bvh_t objs; std::vector<material> mats; std::vector<light> lights;
};
std::shared_ptr<scene_t> make_scene(const float *tris, const uint32_t *tri_mats, int64_t n, const float *mat_data, int64_t m) {
    auto sc = std::make_shared<scene_t>();
    std::vector<obj> objs(n);
    for (int64_t i = 0; i < n; i++) {                                          /* parse_triangles :26-35 */
        const float *t = tris + 9 * i;
        objs[i].geom = {{t[0], t[1], t[2]}, {t[3], t[4], t[5]}, {t[6], t[7], t[8]}};
        objs[i].mat_ix = tri_mats[i]; objs[i].src_index = (int32_t)i;
    }
    sc->mats.resize(m);
    for (int64_t i = 0; i < m; i++) sc->mats[i] = parse_mat(mat_data + 28 * i); /* :55-56 */
    for (int64_t i = 0; i < n; i++) {                                          /* get_lights :58-66 */
        const spectrum &e = sc->mats[(int32_t)objs[i].mat_ix].emission;
        bool nonzero = false;
        for (int k = 0; k < 6; k++) if (e.w[k] >= 0 && e.x[k] > 0) nonzero = true;
        if (nonzero) sc->lights.push_back({LIGHT_DIFFUSE, objs[i].geom, 0, e, (int32_t)i});
    }
    bvh_build(objs, sc->objs);                                                 /* accelerate_scene :75-76 */
    return sc;
}
inline bool closest_interaction(float tmax, const ray &r, float wavelen, const scene_t &sc, interaction &out) { /* :68-73 */
    trav_hit th;
    if (!bvh_closest_hit(tmax, r, sc.objs, th)) return false;
    out.h = th.h; out.mat = sc.mats[(int32_t)sc.objs.leaves[th.leaf].mat_ix]; out.wavelen = wavelen;
    return true;
}

/* ------------------------------------------------------------------ camera.fut */
struct normal_dist { float mu, sigma; };
enum transmitter_kind { TX_NONE = 0, TX_FLASH = 1, TX_SCANNING = 2 };
struct camera_config {                                                         /* camera.fut:34-40 */
    float aperture, focal_dist, offset_radius, field_of_view;
    int n_sensor; normal_dist sensor_dist[3]; vec3 sensor_vis[3];
    transmitter_kind tx; float tx_radius, tx_theta; spectrum tx_emission;
};
struct camera { float pitch, yaw; vec3 origin; camera_config conf; };          /* :42-45 */
inline float from_deg(float d) { return d * F32_PI / 180.0f; }                 /* linalg.fut:53 */
camera_config lidar_conf() {                                                   /* lib.fut:10-18 */
    camera_config c{}; c.aperture = 0; c.focal_dist = 1; c.offset_radius = 0.01f; c.field_of_view = from_deg(90);
    c.n_sensor = 1; c.sensor_dist[0] = {1550, 10}; c.sensor_vis[0] = {1, 0, 0};
    c.tx = TX_SCANNING; c.tx_radius = 0.01f; c.tx_theta = from_deg(3); c.tx_emission = uniform_spectrum(1500);
    return c;
}
camera_config visual_conf() {                                                  /* lib.fut:20-28 */
    camera_config c{}; c.aperture = 0; c.focal_dist = 1; c.offset_radius = 1; c.field_of_view = from_deg(80);
    c.n_sensor = 3;
    c.sensor_dist[0] = {455, 22}; c.sensor_vis[0] = {0, 0, 1};
    c.sensor_dist[1] = {535, 32}; c.sensor_vis[1] = {0, 1, 0};
    c.sensor_dist[2] = {610, 26}; c.sensor_vis[2] = {1, 0, 0};
    c.tx = TX_NONE; c.tx_radius = 0; c.tx_theta = 0; c.tx_emission = uniform_spectrum(0);
    return c;
}
camera_config visual_flash_conf() {                                            /* lib.fut:30-33 */
    camera_config c = visual_conf(); c.tx = TX_FLASH; c.tx_radius = 0.05f;
    c.tx_emission = map_intensities_mul(blackbody_normalized(5500), 1000); return c;
}
inline vec3 cam_dir(const camera &c) { return vnormalise(mkvec3(sinf(c.yaw), sinf(c.pitch), -(cosf(c.yaw)))); } /* camera.fut:47-49 (host libm) */
inline vec3 cam_right(const camera &c) { return vnormalise(vcross(cam_dir(c), world_up)); }                       /* :51-52 */
inline vec3 cam_up(const camera &c) { return vnormalise(vcross(cam_right(c), cam_dir(c))); }                      /* :54-55 */
camera move_camera(camera cam, vec3 m) {                                       /* :57-62 */
    vec3 d = cam_dir(cam); d.y = 0;
    vec3 fwd = vnormalise(d);
    cam.origin = vadd(vadd(vadd(cam.origin, vscale(0.1f * m.z, fwd)), vscale(0.1f * m.x, cam_right(cam))), vscale(0.1f * m.y, world_up));
    return cam;
}
camera turn_camera(camera cam, float pitch, float yaw) {                       /* :64-66 */
    cam.pitch = clampf(-0.5f * F32_PI, 0.5f * F32_PI, cam.pitch + pitch);
    cam.yaw = fmodf(cam.yaw + yaw, 2 * F32_PI);
    return cam;
}
inline void sample_camera_wavelength(const camera &cam, rnge &rng, float &wavelen, int32_t &channel) { /* :68-79 */
    uint32_t n = rng_rand(rng);                                                /* random_select' rand.fut:39-42 */
    channel = (int32_t)(n % (uint32_t)cam.conf.n_sensor);
    normal_dist d = cam.conf.sensor_dist[channel];
    float p = random_unit_exclusive(rng);
    wavelen = d.mu + d.sigma * det_probitf(p);                                 /* stat.sample (mk_normal) p */
}
inline ray sample_camera_ray(const camera &cam, float w, float h, float j, float i, rnge rng /* by value: :86,102 */) { /* :81-110 */
    float ratio = w / h;
    rnge r1 = rng; float o0, o1; random_in_unit_square(r1, o0, o1);
    float offx = cam.conf.offset_radius * o0, offy = cam.conf.offset_radius * o1;
    float x = (j + offx) / w, y = (i + offy) / h;
    float lens_radius = cam.conf.aperture / 2;
    float half_height = tanf(cam.conf.field_of_view / 2.0f);
    float half_width = ratio * half_height;
    vec3 ww = vscale(-1, cam_dir(cam)), u = cam_right(cam), v = cam_up(cam);
    float focus_dist = cam.conf.focal_dist;
    vec3 llc = vsub(vsub(vsub(cam.origin, vscale(half_width * focus_dist, u)), vscale(half_height * focus_dist, v)), vscale(focus_dist, ww));
    vec3 horizontal = vscale(2 * half_width * focus_dist, u);
    vec3 vertical = vscale(2 * half_height * focus_dist, v);
    rnge r2 = rng; vec3 d = random_in_unit_disk(r2);
    vec3 lens = vscale(lens_radius, d);
    vec3 lens_offset = vadd(vscale(lens.x, u), vscale(lens.y, v));
    vec3 origin = vadd(cam.origin, lens_offset);
    return mkray(origin, vsub(vadd(vadd(llc, vscale(x, horizontal)), vscale(y, vertical)), origin));
}
/* gen_transmitter (camera.fut:112-122): 0 or 8 lights appended per ray */
inline int gen_transmitter(const camera &c, const ray &r, light *out) {
    const int n_sectors = 8; triangle tris[8];
    if (c.conf.tx == TX_NONE) return 0;
    if (c.conf.tx == TX_FLASH) {
        disk(c.origin, cam_dir(c), c.conf.tx_radius, n_sectors, tris);
        for (int k = 0; k < 8; k++) out[k] = {LIGHT_DIFFUSE, tris[k], 0, c.conf.tx_emission, -1};
    } else {
        disk(c.origin, r.dir, c.conf.tx_radius, n_sectors, tris);
        for (int k = 0; k < 8; k++) out[k] = {LIGHT_FRUSTUM, tris[k], c.conf.tx_theta, c.conf.tx_emission, -1};
    }
    return 8;
}

/* ------------------------------------------------------------------ direct.fut */
struct light_list { const light *scene_lights; int n_scene; light extra[8]; int n_extra;
    int size() const { return n_scene + n_extra; }
    const light &at(int i) const { return i < n_scene ? scene_lights[i] : extra[i - n_scene]; } };

inline bool occluded(const hit &h, vec3 lightp, const bvh_t &objs) {          /* direct.fut:7-15 */
    vec3 v = vsub(lightp, h.pos);
    vec3 w = vnormalise(v);
    const float eps = 0.01f;
    if (vdot(w, h.normal) <= 0) return true;
    float distance = vnorm(v);
    ray r = mkray_adjust_acne(h, w);
    return bvh_any_hit(distance - eps, r, objs);
}
inline float triangle_area(const triangle &t) { return vnorm(vcross(vsub(t.b, t.a), vsub(t.c, t.a))) / 2; } /* :17-20 */
inline float balance(float pdf_f, float pdf_g) {                              /* :56-58 with nf = ng = 1 */
    float nf = 1.0f, ng = 1.0f;
    return nf * pdf_f / (nf * pdf_f + ng * pdf_g);
}
float estimate_direct(rnge &rng, vec3 wo, const interaction &i, const light &l, const bvh_t &objs) { /* :63-103 */
    /* sample_light -> sample_arealight (:32-42): peeks two draws, returns the un-advanced rng */
    float light_radiance;
    {
        const triangle &t = l.geom;
        vec3 e1 = vsub(t.b, t.a), e2 = vsub(t.c, t.a);
        float area = vnorm(vcross(e1, e2)) / 2;
        rnge peek = rng; float u, v; random_in_triangle(peek, u, v);
        vec3 p = vadd(vadd(t.a, vscale(u, e1)), vscale(v, e2));
        vec3 wi = vnormalise(vsub(p, i.h.pos));
        float in_radiance = arealight_incident_radiance(l, i.h.pos, p, i.wavelen);
        float pdf = 1 / area;
        if (occluded(i.h, p, objs)) in_radiance = 0;                           /* :51-53 */
        if (pdf == 0 || in_radiance == 0) light_radiance = 0;                  /* :73-74 */
        else {
            float f = bsdf_f(wo, wi, i) * lys_fabsf(vdot(wi, i.h.normal));
            float scattering_pdf = bsdf_pdf(wo, wi, i);
            float weight = balance(pdf, scattering_pdf);
            light_radiance = f * weight * in_radiance / pdf;
        }
    }
    float bsdf_radiance;
    {                                                                          /* :83-102 */
        dir_sample s = sample_dir(wo, i, rng);
        ray r = mkray_adjust_acne(i.h, s.wi);
        hit lh;
        if (!hit_triangle(F32_HIGHEST, r, l.geom, lh)) bsdf_radiance = 0;
        else if (occluded(i.h, lh.pos, objs)) bsdf_radiance = 0;
        else {
            float in_radiance = arealight_incident_radiance(l, i.h.pos, lh.pos, i.wavelen);
            float f = s.bsdf * lys_fabsf(vdot(s.wi, i.h.normal));
            if (s.kind == PDF_IMPOSSIBLE) bsdf_radiance = 0;
            else if (s.kind == PDF_DELTA) bsdf_radiance = f * in_radiance;
            else {
                float light_pdf = 1 / triangle_area(l.geom);                  /* arealight_pdf :22 */
                float weight = balance(s.pdf, light_pdf);
                bsdf_radiance = f * in_radiance * weight / s.pdf;
            }
        }
    }
    return light_radiance + bsdf_radiance;
}
float direct_radiance(rnge &rng, vec3 wo, const interaction &i, const light_list &lights, const bvh_t &objs) { /* :111-122 */
    if (lights.size() == 0) return 0;
    uint32_t n = rng_rand(rng);                                                /* random_select */
    const light &l = lights.at((int32_t)(n % (uint32_t)lights.size()));
    float radiance = estimate_direct(rng, wo, i, l, objs);
    float light_pdf = 1 / (float)(int32_t)lights.size();
    return radiance / light_pdf;
}

/* ------------------------------------------------------------------ integrator.fut */
struct path_vertex { float distance, radiance; };
static thread_local float *t_ray_dump = nullptr;      /* divergence studies: [MAX_PATH_LEN][6] rays of the current path */
static thread_local int32_t *t_ray_dump_n = nullptr;
void path_trace(ray r, float wavelen, const scene_t &scene, const light_list &lights, const spectrum &ambience_s,
                rnge rng, path_vertex *path) {                                 /* :27-76 */
    const float tmax = F32_HIGHEST;
    float ambience = spectrum_lookup(wavelen, ambience_s);
    for (int k = 0; k < MAX_PATH_LEN; k++) path[k] = {F32_INF, 0};              /* dark_path :41 */
    int i = 0; float distance = 0; bool should_continue = true;
    t_counters.paths++;
    while (should_continue && i < g_path_len) {
        interaction inter;
        if (t_ray_dump) { float *p = t_ray_dump + 6 * i; p[0] = r.origin.x; p[1] = r.origin.y; p[2] = r.origin.z; p[3] = r.dir.x; p[4] = r.dir.y; p[5] = r.dir.z; *t_ray_dump_n = i + 1; }
        if (closest_interaction(tmax, r, wavelen, scene, inter)) {
            t_counters.vertices++;
            advance_rng(rng);                                                  /* :48 */
            vec3 wo = vneg(r.dir);
            float direct = direct_radiance(rng, wo, inter, lights, scene.objs);
            float radiance = direct + ((i == 0) ? spectrum_lookup(wavelen, inter.mat.emission) : 0); /* :51-53 */
            distance = distance + inter.h.t;
            path[i] = {distance, radiance};
            dir_sample s = sample_dir(wo, inter, rng);                        /* :56 */
            float pdf = (s.kind == PDF_IMPOSSIBLE) ? 0 : ((s.kind == PDF_DELTA) ? 1 : s.pdf);
            float cosFalloff = lys_fabsf(vdot(inter.h.normal, s.wi));
            float p_terminate = 1 - s.bsdf * cosFalloff / pdf;                 /* :67 */
            bool terminate = random_unit_exclusive(rng) < p_terminate;          /* :68 */
            if (pdf == 0 || terminate) should_continue = false;
            else { i = i + 1; r = mkray_adjust_acne(inter.h, s.wi); }
        } else {
            path[i] = {F32_INF, ambience};                                     /* :76 */
            should_continue = false;
        }
    }
}
struct pixel_out { ray r; int32_t channel; path_vertex path[MAX_PATH_LEN]; };
void sample_pixel(const scene_t &scene, const camera &cam, const spectrum &ambience, float w, float h,
                  uint32_t j, uint32_t i, rnge rng, pixel_out &out) {           /* :78-101 */
    float wl; int32_t channel;
    sample_camera_wavelength(cam, rng, wl, channel);
    ray r = sample_camera_ray(cam, w, h, (float)j, h - (float)i - 1.0f, rng);
    light_list lights; lights.scene_lights = scene.lights.data(); lights.n_scene = (int)scene.lights.size();
    lights.n_extra = gen_transmitter(cam, r, lights.extra);                    /* :96 */
    out.r = r; out.channel = channel;
    path_trace(r, wl, scene, lights, ambience, rng, out.path);
}

} // namespace

/* ------------------------------------------------------------------ state.fut / lib.fut */
struct orc_state {                                                             /* state.fut:8-19 */
    uint32_t dim_w, dim_h;
    uint32_t subsampling;
    rnge rng;
    uint32_t img_h, img_w; std::vector<vec3> img;
    uint32_t n_frames;
    spectrum ambience;
    bool mode;
    int render_mode;          /* 0 = #render_color, 1 = #render_distance */
    uint32_t cam_conf_id;
    camera cam;
    std::shared_ptr<scene_t> scene;
};

namespace {
inline void grid_dims(const orc_state &s, uint32_t &gw, uint32_t &gh) {        /* integrator.fut:175-176 */
    gw = (s.dim_w + s.subsampling - 1) / s.subsampling;
    gh = (s.dim_h + s.subsampling - 1) / s.subsampling;
}
/* sample_pixels (integrator.fut:103-116) */
rnge sample_pixels(rnge rng, uint32_t w, uint32_t h, const scene_t &scene, const camera &cam, const spectrum &amb,
                   std::vector<pixel_out> &img) {
    img.resize((size_t)w * h);
    int nt = g_threads > 0 ? g_threads : 0;
#ifdef _OPENMP
    if (nt == 0) nt = omp_get_max_threads();
#else
    nt = 1;
#endif
#pragma omp parallel num_threads(nt)
    {
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < (int64_t)h; i++)
            for (int64_t j = 0; j < (int64_t)w; j++) {
                int32_t ix = (int32_t)(i * (int64_t)w + j);
                rnge r = rng ^ rng_hash(ix);                                   /* split_rng */
                sample_pixel(scene, cam, amb, (float)w, (float)h, (uint32_t)j, (uint32_t)i, r, img[(size_t)ix]);
            }
        flush_counters();
    }
    advance_rng(rng);
    return rng;
}
vec3 hue_to_rgb(float h) {                                                     /* integrator.fut:139-148 */
    float hp = h * 6;
    float x = 1 - lys_fabsf(fmodf(hp, 2.0f) - 1);
    switch (u32_of_f32(hp)) {
        case 0: return {1, x, 0}; case 1: return {x, 1, 0}; case 2: return {0, 1, x};
        case 3: return {0, x, 1}; case 4: return {x, 0, 1}; default: return {1, 0, x};
    }
}
vec3 visualize(int render_mode, const camera_config &conf, const pixel_out &p) { /* integrator.fut:150-168 */
    if (render_mode == 1) {
        const float min_d = 0.5f, max_d = 10;
        bool any = false; float best = 0;
        for (int k = 0; k < MAX_PATH_LEN; k++) {
            const path_vertex &s = p.path[k];
            if (s.radiance > 0 && s.distance > min_d && s.distance < max_d) {
                if (!any || s.distance < best) { best = s.distance; any = true; }
            }
        }
        if (!any) return {0, 0, 0};
        return hue_to_rgb(0.85f * (best - min_d) / (max_d - min_d));
    }
    vec3 acc = {0, 0, 0};
    vec3 ch = conf.sensor_vis[p.channel];
    for (int k = 0; k < MAX_PATH_LEN; k++) acc = vadd(acc, vscale(p.path[k].radiance, ch));
    return vscale((float)(int32_t)conf.n_sensor, acc);
}
/* sample_frame (integrator.fut:172-178) */
rnge sample_frame(const orc_state &s, std::vector<vec3> &img, uint32_t &gw, uint32_t &gh) {
    grid_dims(s, gw, gh);
    std::vector<pixel_out> px;
    rnge rng = sample_pixels(s.rng, gw, gh, *s.scene, s.cam, s.ambience, px);
    img.resize(px.size());
    for (size_t k = 0; k < px.size(); k++) img[k] = visualize(s.render_mode, s.cam.conf, px[k]);
    return rng;
}
/* sample_frame_accum (integrator.fut:180-192) */
rnge sample_frame_accum(const orc_state &s, std::vector<vec3> &out, uint32_t &gw, uint32_t &gh) {
    std::vector<vec3> img_new;
    rnge rng = sample_frame(s, img_new, gw, gh);
    float n_frames = (float)s.n_frames;
    out.resize(img_new.size());
    /* `img_new :> [m][n]vec3`: shapes must agree (a Futhark run-time size error otherwise) */
    if (s.img.size() != img_new.size()) { fprintf(stderr, "orc: sample_frame_accum shape mismatch\n"); abort(); }
    for (size_t k = 0; k < out.size(); k++) {
        vec3 acc = s.img[k], c = img_new[k];
        if (s.render_mode == 1) out[k] = (vnorm(acc) > 0) ? acc : c;
        else out[k] = vadd(vscale((n_frames - 1) / n_frames, acc), vscale(1 / n_frames, c));
    }
    return rng;
}
int32_t argb_from_rgba(float r, float g, float b, float a) {                   /* matte colour.fut argb.from_rgba */
    auto ch = [](float x) -> uint32_t {
        float c = (x < 0.0f) ? 0.0f : ((x > 1.0f) ? 1.0f : x);
        (void)c;
        return lys_pin_argb_channel(x);
    };
    return (int32_t)((ch(a) << 24) | (ch(r) << 16) | (ch(g) << 8) | ch(b));
}
} // namespace

extern "C" {

void orc_set_path_len(int n) { g_path_len = (n < 1) ? 1 : (n > MAX_PATH_LEN ? MAX_PATH_LEN : n); }
void orc_set_refit_mode(int mode) { g_refit_mode = mode; }
void orc_set_math_mode(int mode) { g_math_mode = mode; }
void orc_set_threads(int n) { g_threads = n; }
int orc_get_threads(void) {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

orc_state *orc_init(int32_t seed, uint32_t h, uint32_t w, uint32_t cam_conf_id, const float *tri_geoms,
                    const uint32_t *tri_mats, int64_t n_tris, const float *mat_data, int64_t n_mats,
                    float cam_pitch, float cam_yaw, const float *cam_origin) {  /* lib.fut:76-106 */
    if (n_tris < 2) return nullptr;       /* radix_tree.mk needs n >= 2 */
    orc_state *s = new orc_state();
    s->dim_w = w; s->dim_h = h; s->subsampling = 1;
    s->rng = rng_from_seed1(seed);
    s->img_h = h; s->img_w = w; s->img.assign((size_t)w * h, vec3{0, 0, 0});
    s->n_frames = 0; s->ambience = uniform_spectrum(0); s->mode = false;
    if (cam_conf_id == 0) { s->render_mode = 0; s->cam.conf = visual_conf(); }
    else if (cam_conf_id == 1) { s->render_mode = 0; s->cam.conf = visual_flash_conf(); }
    else { s->render_mode = 1; s->cam.conf = lidar_conf(); }
    s->cam_conf_id = cam_conf_id;
    s->cam.pitch = cam_pitch; s->cam.yaw = cam_yaw; s->cam.origin = {cam_origin[0], cam_origin[1], cam_origin[2]};
    s->scene = make_scene(tri_geoms, tri_mats, n_tris, mat_data, n_mats);
    return s;
}
orc_state *orc_resize(uint32_t h, uint32_t w, const orc_state *s) {            /* lib.fut:108-109 */
    orc_state *r = new orc_state(*s); r->dim_w = w; r->dim_h = h; r->mode = false; return r;
}
orc_state *orc_step(const orc_state *s) {                                      /* lib.fut:111-118 */
    orc_state *r = new orc_state(*s);
    uint32_t gw, gh;
    if (s->mode && s->n_frames > 0) { r->rng = sample_frame_accum(*s, r->img, gw, gh); r->n_frames = s->n_frames + 1; }
    else { r->rng = sample_frame(*s, r->img, gw, gh); r->n_frames = 1; }
    r->img_w = gw; r->img_h = gh;
    return r;
}
orc_state *orc_key(int32_t e, int32_t key, const orc_state *s) {               /* lib.fut:120-185 */
    orc_state *r = new orc_state(*s);
    if (e != 0) return r;
    auto mv = [&](float x, float y, float z) { r->cam = move_camera(s->cam, {x, y, z}); r->n_frames = 0; };
    auto tn = [&](float p, float y) { r->cam = turn_camera(s->cam, p, y); r->n_frames = 0; };
    switch (key) {
        case 0x32: r->subsampling = s->subsampling + 1; r->n_frames = 0; break;
        case 0x31: r->subsampling = std::max<uint32_t>(1, s->subsampling - 1); r->n_frames = 0; break;
        case 0x77: mv(0, 0, 1); break;
        case 0x61: mv(-1, 0, 0); break;
        case 0x73: mv(0, 0, -1); break;
        case 0x64: mv(1, 0, 0); break;
        case 0x40000052: tn(-0.1f, 0.0f); break;
        case 0x40000051: tn(0.1f, 0.0f); break;
        case 0x4000004F: tn(0.0f, 0.1f); break;
        case 0x40000050: tn(0.0f, -0.1f); break;
        case 0x78: mv(0, 1, 0); break;
        case 0x7A: mv(0, -1, 0); break;
        case 0x20: r->mode = !s->mode; r->n_frames = 0; break;
        case 0x6E: r->mode = false; r->n_frames = 0; break;
        case 0x6D: r->mode = true; break;
        case 0x69: r->cam.conf.aperture = f_min(2, s->cam.conf.aperture + 0.08f); break;
        case 0x6B: r->cam.conf.aperture = f_max(0, s->cam.conf.aperture - 0.08f); break;
        case 0x6F: r->cam.conf.focal_dist = s->cam.conf.focal_dist * 1.14f; break;
        case 0x6C: r->cam.conf.focal_dist = f_max(0.1f, s->cam.conf.focal_dist / 1.14f); break;
        case 0x74:
            if (s->cam_conf_id == 0) { r->cam.conf = visual_flash_conf(); r->cam_conf_id = 1; r->render_mode = 0; }
            else if (s->cam_conf_id == 1) { r->cam.conf = lidar_conf(); r->cam_conf_id = 2; r->render_mode = 1; }
            else { r->cam.conf = visual_conf(); r->cam_conf_id = 0; r->render_mode = 0; }
            r->n_frames = 0; break;
        case 0x70: r->ambience = (s->ambience.x[0] == 0) ? bright_blue_sky() : uniform_spectrum(0); break;
        default: break;
    }
    return r;
}
void orc_render(const orc_state *s, int32_t *out) {                            /* lib.fut:187-196 */
    int32_t sub = (int32_t)s->subsampling;
    for (int32_t i = 0; i < (int32_t)s->dim_h; i++)
        for (int32_t j = 0; j < (int32_t)s->dim_w; j++) {
            size_t ii = (size_t)(i / sub), jj = (size_t)(j / sub);
            vec3 c = {0, 0, 0};
            if (ii < s->img_h && jj < s->img_w) c = s->img[ii * s->img_w + jj];  /* `unsafe` index in the reference */
            out[(size_t)i * s->dim_w + j] = argb_from_rgba(c.x, c.y, c.z, 1.0f);
        }
}
void orc_sample_n_frames(const orc_state *s0, uint32_t n, float *out) {        /* lib.fut:67-74 */
    orc_state s = *s0;
    uint32_t gw, gh;
    { std::vector<vec3> img; rnge rng = sample_frame(s, img, gw, gh); s.n_frames = 1; s.rng = rng; s.img = img; s.img_w = gw; s.img_h = gh; }
    while (s.n_frames < n) {
        std::vector<vec3> img; rnge rng = sample_frame_accum(s, img, gw, gh);
        s.img = img; s.rng = rng; s.n_frames = s.n_frames + 1;
    }
    for (size_t k = 0; k < s.img.size(); k++) { out[3 * k] = s.img[k].x; out[3 * k + 1] = s.img[k].y; out[3 * k + 2] = s.img[k].z; }
}
orc_state *orc_sample_points_n(const orc_state *s0, uint32_t spp, float *out) { /* lib.fut:35-63 */
    struct cloud_point { vec3 pos; float distance, intensity; };
    orc_state *res = new orc_state(*s0);
    uint32_t gw, gh; grid_dims(*s0, gw, gh);
    float factor = 1 / (float)spp;
    const float min_d = 0.5f, max_d = 10;
    auto closest = [&](const pixel_out &p) -> cloud_point {                    /* :41-47 */
        cloud_point best = {{-1, -1, -1}, F32_INF, 0}; bool any = false;
        for (int k = 0; k < MAX_PATH_LEN; k++) {
            float dist = p.path[k].distance, inten = p.path[k].radiance * factor;
            if (inten > 0 && dist > min_d && dist < max_d) {
                if (!any || dist < best.distance) {                            /* minimum_by common.fut:40-50 (strict <) */
                    best.pos = vadd(p.r.origin, vscale(dist, p.r.dir));        /* to_cloud_points integrator.fut:122-126 */
                    best.distance = dist; best.intensity = inten; any = true;
                }
            }
        }
        return best;
    };
    std::vector<pixel_out> px;
    rnge rng = sample_pixels(s0->rng, gw, gh, *s0->scene, s0->cam, s0->ambience, px);   /* sample_points :118-128 */
    std::vector<cloud_point> points(px.size());
    for (size_t k = 0; k < px.size(); k++) points[k] = closest(px[k]);
    for (int32_t it = 0; it < (int32_t)spp - 1; it++) {                        /* :55-59 */
        rng = sample_pixels(rng, gw, gh, *s0->scene, s0->cam, s0->ambience, px);
        for (size_t k = 0; k < px.size(); k++) {
            cloud_point p2 = closest(px[k]);
            if (!(points[k].distance < p2.distance)) points[k] = p2;           /* merge :48-51 */
        }
    }
    res->rng = rng;
    for (size_t k = 0; k < points.size(); k++) {
        out[4 * k] = points[k].pos.x; out[4 * k + 1] = points[k].pos.y; out[4 * k + 2] = points[k].pos.z; out[4 * k + 3] = points[k].intensity;
    }
    return res;
}
orc_state *orc_advance_rng(const orc_state *s, uint32_t k) { orc_state *r = new orc_state(*s); for (uint32_t i = 0; i < k; i++) advance_rng(r->rng); return r; }
void orc_free_state(orc_state *s) { delete s; }

void orc_state_dims(const orc_state *s, uint32_t *w, uint32_t *h, uint32_t *gw, uint32_t *gh) {
    *w = s->dim_w; *h = s->dim_h; grid_dims(*s, *gw, *gh);
}
void orc_state_image(const orc_state *s, float *out, uint32_t *img_h, uint32_t *img_w) {
    if (img_h) *img_h = s->img_h;
    if (img_w) *img_w = s->img_w;
    if (out) for (size_t k = 0; k < s->img.size(); k++) { out[3 * k] = s->img[k].x; out[3 * k + 1] = s->img[k].y; out[3 * k + 2] = s->img[k].z; }
}
void orc_state_scalars(const orc_state *s, uint32_t *rng, uint32_t *n_frames, uint32_t *subsampling, int32_t *mode,
                       int32_t *render_mode, uint32_t *cam_conf_id, float *cam, float *ambience12) {
    *rng = s->rng; *n_frames = s->n_frames; *subsampling = s->subsampling; *mode = s->mode ? 1 : 0;
    *render_mode = s->render_mode; *cam_conf_id = s->cam_conf_id;
    cam[0] = s->cam.pitch; cam[1] = s->cam.yaw; cam[2] = s->cam.origin.x; cam[3] = s->cam.origin.y; cam[4] = s->cam.origin.z;
    cam[5] = s->cam.conf.aperture; cam[6] = s->cam.conf.focal_dist;
    for (int i = 0; i < 6; i++) { ambience12[2 * i] = s->ambience.w[i]; ambience12[2 * i + 1] = s->ambience.x[i]; }
}

int64_t orc_bvh_size(const orc_state *s) { return (int64_t)s->scene->objs.leaves.size(); }
int64_t orc_n_lights(const orc_state *s) { return (int64_t)s->scene->lights.size(); }
void orc_bvh_get(const orc_state *s, float *bounds6, uint32_t *sorted_morton, int32_t *sorted_src_index, int32_t *left,
                 int32_t *right, int32_t *parent, float *node_aabb, float *leaf_aabb) {
    const bvh_t &b = s->scene->objs;
    auto put = [](float *p, const aabb &a) { p[0] = a.center.x; p[1] = a.center.y; p[2] = a.center.z; p[3] = a.half_dims.x; p[4] = a.half_dims.y; p[5] = a.half_dims.z; };
    if (bounds6) put(bounds6, b.bounds);
    for (size_t i = 0; i < b.leaves.size(); i++) {
        if (sorted_morton) sorted_morton[i] = b.mortons[i];
        if (sorted_src_index) sorted_src_index[i] = b.leaves[i].src_index;
        if (leaf_aabb) put(leaf_aabb + 6 * i, b.leaf_aabbs[i]);
    }
    for (size_t i = 0; i < b.nodes.size(); i++) {
        if (left) left[i] = b.nodes[i].left;
        if (right) right[i] = b.nodes[i].right;
        if (parent) parent[i] = b.nodes[i].parent;
        if (node_aabb) put(node_aabb + 6 * i, b.nodes[i].box);
    }
}
void orc_light_indices(const orc_state *s, int32_t *src_index) {
    for (size_t i = 0; i < s->scene->lights.size(); i++) src_index[i] = s->scene->lights[i].src_index;
}

uint32_t orc_expand_bits(uint32_t x) { return expand_bits(x); }
uint32_t orc_morton3d(float x, float y, float z) { return morton3D({x, y, z}); }
uint32_t orc_hash(int32_t x) { return rng_hash(x); }
uint32_t orc_rng_from_seed(int32_t seed) { return rng_from_seed1(seed); }
uint32_t orc_rng_next(uint32_t s) { rng_rand(s); return s; }
float orc_rng_uniform(uint32_t s, float lo, float hi, uint32_t *s_out) { float v = dist_rand(s, lo, hi); if (s_out) *s_out = s; return v; }
void orc_radix_tree(const uint32_t *keys, int64_t n, int32_t *left, int32_t *right, int32_t *parent) { radix_tree_mk(keys, n, left, right, parent); }
float orc_spectrum_lookup(float v, const float *spectrum12) { return spectrum_lookup(v, spectrum_from12(spectrum12)); }
int32_t orc_hit_aabb(float tmax, const float *ray6, const float *box6) {          /* box6 = center xyz, half_dims xyz */
    ray r = {{ray6[0], ray6[1], ray6[2]}, {ray6[3], ray6[4], ray6[5]}};
    aabb b = {{box6[0], box6[1], box6[2]}, {box6[3], box6[4], box6[5]}};
    return hit_aabb(tmax, r, b) ? 1 : 0;
}
int32_t orc_hit_triangle(float tmax, const float *ray6, const float *tri9, float *t_pos_normal7) {
    ray r = {{ray6[0], ray6[1], ray6[2]}, {ray6[3], ray6[4], ray6[5]}};
    triangle tr = {{tri9[0], tri9[1], tri9[2]}, {tri9[3], tri9[4], tri9[5]}, {tri9[6], tri9[7], tri9[8]}};
    hit h;
    if (!hit_triangle(tmax, r, tr, h)) return 0;
    if (t_pos_normal7) { float o[7] = {h.t, h.pos.x, h.pos.y, h.pos.z, h.normal.x, h.normal.y, h.normal.z}; memcpy(t_pos_normal7, o, sizeof o); }
    return 1;
}
void orc_normalise(const float *v3, float *out3) { vec3 r = vnormalise({v3[0], v3[1], v3[2]}); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; }
int32_t orc_argb_from_rgba(float r, float g, float b, float a) { return argb_from_rgba(r, g, b, a); }
void orc_eval_math(int fn, const float *in, float *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        float x = in[i], y;
        switch (fn) {
            case 0: y = det_sinf(x); break; case 1: y = det_cosf(x); break; case 2: y = det_expf(x); break;
            case 3: y = det_logf(x); break; case 4: y = det_pow5f(x); break; case 5: y = det_acosf(x); break;
            case 6: y = det_probitf(x); break; default: y = 0;
        }
        out[i] = y;
    }
}
void orc_material_probe(const float *mat28, float wavelen, const float *wo, const float *wi, const float *normal,
                        uint32_t rng, float *out) {
    interaction it; it.mat = parse_mat(mat28); it.wavelen = wavelen;
    it.h.t = 1; it.h.pos = {0, 0, 0}; it.h.normal = {normal[0], normal[1], normal[2]};
    vec3 o = {wo[0], wo[1], wo[2]}, i = {wi[0], wi[1], wi[2]};
    out[0] = bsdf_f(o, i, it); out[1] = bsdf_pdf(o, i, it);
    rnge r = rng; dir_sample s = sample_dir(o, it, r);
    out[2] = s.wi.x; out[3] = s.wi.y; out[4] = s.wi.z; out[5] = s.bsdf; out[6] = (float)s.kind; out[7] = s.pdf;
    memcpy(out + 8, &r, 4);
}

void orc_probe_primary(const orc_state *s, int32_t *leaf, int32_t *src_tri, float *t, float *ray_out, float *wavelen_out) {
    uint32_t gw, gh; grid_dims(*s, gw, gh);
    const scene_t &sc = *s->scene;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)gh; i++)
        for (int64_t j = 0; j < (int64_t)gw; j++) {
            int32_t ix = (int32_t)(i * (int64_t)gw + j);
            rnge r = s->rng ^ rng_hash(ix);
            float wl; int32_t channel; sample_camera_wavelength(s->cam, r, wl, channel);
            ray ry = sample_camera_ray(s->cam, (float)gw, (float)gh, (float)(uint32_t)j, (float)gh - (float)(uint32_t)i - 1.0f, r);
            trav_hit th; bool ok = bvh_closest_hit(F32_HIGHEST, ry, sc.objs, th);
            leaf[ix] = ok ? th.leaf : -1;
            if (src_tri) src_tri[ix] = ok ? sc.objs.leaves[th.leaf].src_index : -1;
            if (t) t[ix] = ok ? th.h.t : F32_INF;
            if (ray_out) { float *p = ray_out + 6 * (size_t)ix; p[0] = ry.origin.x; p[1] = ry.origin.y; p[2] = ry.origin.z; p[3] = ry.dir.x; p[4] = ry.dir.y; p[5] = ry.dir.z; }
            if (wavelen_out) wavelen_out[ix] = wl;
        }
}
void orc_probe_pass(const orc_state *s, float *radiance, float *distance, int32_t *channel) {
    uint32_t gw, gh; grid_dims(*s, gw, gh);
    std::vector<pixel_out> px;
    (void)sample_pixels(s->rng, gw, gh, *s->scene, s->cam, s->ambience, px);
    for (size_t k = 0; k < px.size(); k++) {
        for (int v = 0; v < MAX_PATH_LEN; v++) {
            if (radiance) radiance[k * MAX_PATH_LEN + v] = px[k].path[v].radiance;
            if (distance) distance[k * MAX_PATH_LEN + v] = px[k].path[v].distance;
        }
        if (channel) channel[k] = px[k].channel;
    }
}
void orc_brute_force_hits(const orc_state *s, const float *rays, int64_t n, int32_t *src_tri, float *t) {
    const bvh_t &b = s->scene->objs;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n; k++) {
        ray r = {{rays[6 * k], rays[6 * k + 1], rays[6 * k + 2]}, {rays[6 * k + 3], rays[6 * k + 4], rays[6 * k + 5]}};
        int32_t best = -1; float bt = F32_INF;
        for (size_t i = 0; i < b.leaves.size(); i++) {            /* select_min_hit: strictly smaller t wins, earlier kept on ties */
            hit h;
            if (hit_triangle(F32_HIGHEST, r, b.leaves[i].geom, h) && (best < 0 || h.t < bt)) { best = (int32_t)i; bt = h.t; }
        }
        src_tri[k] = best < 0 ? -1 : b.leaves[best].src_index; if (t) t[k] = bt;
    }
}
void orc_closest_hits(const orc_state *s, const float *rays, int64_t n, int32_t *leaf, float *t) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n; k++) {
        ray r = {{rays[6 * k], rays[6 * k + 1], rays[6 * k + 2]}, {rays[6 * k + 3], rays[6 * k + 4], rays[6 * k + 5]}};
        trav_hit th; bool ok = bvh_closest_hit(F32_HIGHEST, r, s->scene->objs, th);
        leaf[k] = ok ? th.leaf : -1; if (t) t[k] = ok ? th.h.t : F32_INF;
    }
}
void orc_closest_hits_steps(const orc_state *s, const float *rays, int64_t n, int32_t *box_tests, int32_t *tri_tests) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n; k++) {
        ray r = {{rays[6 * k], rays[6 * k + 1], rays[6 * k + 2]}, {rays[6 * k + 3], rays[6 * k + 4], rays[6 * k + 5]}};
        uint64_t b0 = t_counters.box_tests, t0 = t_counters.tri_tests;
        trav_hit th; (void)bvh_closest_hit(F32_HIGHEST, r, s->scene->objs, th);
        box_tests[k] = (int32_t)(t_counters.box_tests - b0); tri_tests[k] = (int32_t)(t_counters.tri_tests - t0);
    }
}
/* the closest-hit rays of every path of the state's NEXT pass: rays [gh][gw][MAX_PATH_LEN][6], n_rays [gh][gw] */
void orc_probe_path_rays(const orc_state *s, float *rays, int32_t *n_rays) {
    uint32_t gw, gh; grid_dims(*s, gw, gh);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t i = 0; i < (int64_t)gh; i++)
        for (int64_t j = 0; j < (int64_t)gw; j++) {
            int32_t ix = (int32_t)(i * (int64_t)gw + j);
            rnge r = s->rng ^ rng_hash(ix);
            pixel_out px;
            n_rays[ix] = 0;
            t_ray_dump = rays + (size_t)ix * MAX_PATH_LEN * 6; t_ray_dump_n = n_rays + ix;
            sample_pixel(*s->scene, s->cam, s->ambience, (float)gw, (float)gh, (uint32_t)j, (uint32_t)i, r, px);
            t_ray_dump = nullptr; t_ray_dump_n = nullptr;
        }
}
/* per-ray visit pattern of the closest-hit walk in left-first order: 0 = box test failed, 1 = box test passed,
 * 2 = triangle test missed, 3 = triangle test hit (closest updated); at most max_steps entries per ray, len = true length */
void orc_closest_hits_pattern(const orc_state *s, const float *rays, int64_t n, int32_t max_steps, uint8_t *pattern, int32_t *len) {
    const bvh_t &bvh = s->scene->objs;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n; k++) {
        ray r = {{rays[6 * k], rays[6 * k + 1], rays[6 * k + 2]}, {rays[6 * k + 3], rays[6 * k + 4], rays[6 * k + 5]}};
        uint8_t *pat = pattern + (size_t)k * max_steps; int32_t m = 0;
        float tmax = F32_HIGHEST;
        std::vector<int32_t> stack; int32_t cur = bvh.nodes.empty() ? -1 : 0; bool go = !bvh.nodes.empty();
        while (go) {                                   /* same decisions as bvh_closest_hit, written as a DFS */
            uint8_t code;
            if (!ptr_is_leaf(cur)) {
                const node &nd = bvh.nodes[cur];
                if (hit_aabb(tmax, r, nd.box)) { code = 1; stack.push_back(nd.right); cur = nd.left; if (m < max_steps) pat[m] = code; m++; continue; }
                code = 0;
            } else {
                hit h;
                if (hit_triangle(tmax, r, bvh.leaves[ptr_leaf_ix(cur)].geom, h)) { tmax = h.t; code = 3; } else code = 2;
            }
            if (m < max_steps) pat[m] = code; m++;
            if (stack.empty()) go = false; else { cur = stack.back(); stack.pop_back(); }
        }
        len[k] = m;
    }
}
void orc_any_hits(const orc_state *s, const float *rays, const float *tmax, int64_t n, int32_t *hit_out) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < n; k++) {
        ray r = {{rays[6 * k], rays[6 * k + 1], rays[6 * k + 2]}, {rays[6 * k + 3], rays[6 * k + 4], rays[6 * k + 5]}};
        hit_out[k] = bvh_any_hit(tmax[k], r, s->scene->objs) ? 1 : 0;
    }
}

void orc_counters_reset(void) { g_counters = Counters(); t_counters = Counters(); }
void orc_counters_get(orc_counters *out) {
    flush_counters();
    out->paths = g_counters.paths; out->vertices = g_counters.vertices; out->closest_rays = g_counters.closest_rays;
    out->shadow_rays = g_counters.shadow_rays; out->node_visits = g_counters.node_visits; out->box_tests = g_counters.box_tests;
    out->tri_tests = g_counters.tri_tests; out->loop_iters = g_counters.loop_iters;
    out->closest_box = g_counters.closest_box; out->closest_tri = g_counters.closest_tri; out->shadow_box = g_counters.shadow_box; out->shadow_tri = g_counters.shadow_tri;
}

} // extern "C"
