/* wavefront.cu -- the sample pass as a wavefront path tracer on sm_100a.
 *
 * Replaces the per-pixel megakernel the Futhark compiler generates for `sample_pixels`
 * (reference src/integrator.fut:103-116, one thread running `path_trace` :27-76 to completion) by
 * queue-driven stages:
 *   k_generate_trace  camera.fut:68-110, integrator.fut:78-93,109-115   wavelength + camera ray per pixel, and its closest hit
 *   k_shade(b)        integrator.fut:46-76, direct.fut:32-122, material.fut   vertex shading; emits <= 2 shadow rays and the
 *                                                                   continuation ray, compacts live paths (warp ballot)
 *   k_trace(b)        bvh.fut:123-145 (closest_hit) of bounce b+1 and direct.fut:7-15 + bvh.fut:149-167 (any_hit) of the
 *                     shadow rays of bounce b, with the radiance accumulation, in ONE persistent launch
 *                     (grid = SMs x resident CTAs, warp-uniform grid-stride over device-side counts)
 *   k_tail(b0)        all bounces from the first sparse one on in one launch (CTA-local wavefront)
 *   k_accumulate      integrator.fut:133-192                             channel resolve + running average
 *   k_render          lib.fut:187-196                                    upscale + ARGB pack
 * Per-path arithmetic is the reference's, operation for operation; only the scheduling differs.
 */
#include "lys_wavefront.h"
#include <cstdio>
#include <cstdlib>

namespace lys {

/* ------------------------------------------------------------------ BVH traversal
 * The reference walks the tree without a stack: parent pointers, always left child first, a node's box
 * tested once on entry against the CURRENT tmax, leaves never box-tested (bvh.fut:126-142).  A depth-first
 * walk that pushes the right child while descending left takes exactly the same decisions in the same
 * order, so hits, ties (strict t < tmax, shapes.fut:64) and culling are identical. */
struct RayInv { V3 o, d, inv; };
/* one 32-byte sector of a traversal record in ONE request: sm_100 has 256-bit global loads (LDG.E.256), half the LSU
 * instructions and L1 tag look-ups of two 128-bit loads (the large scene runs the L1 data stage at 72 % of its peak) */
LYS_D void ld_sector(const float4 *__restrict__ p, float4 &a, float4 &b) {
#ifdef __CUDACC__
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    a = p[0]; b = p[1];
#endif
}
/* hit_aabb (shapes.fut:114-135).  The reference leaves after the first axis with tmax <= tmin; tmin only grows and
 * tmax only shrinks from axis to axis (fmaxf/fminf drop NaN operands), so an axis that fails keeps failing and the
 * single test after the third axis gives the same boolean.  No per-axis branch: the lanes of a warp stay together.
 * `tn` = tmin after the third axis = max(0, the three entry distances): it does not depend on the ray's tmax. */
LYS_D bool slab_test(const RayInv &r, float4 lo, float4 hi, float tmax, float &tn) {
    float tmin = 0.0f;
    {
        float t0 = (lo.x - r.o.x) * r.inv.x, t1 = (hi.x - r.o.x) * r.inv.x;
        if (r.inv.x < 0.0f) { float s = t0; t0 = t1; t1 = s; }
        t1 = t1 * (1.0f + 0.001f);
        tmin = fmaxf(t0, tmin); tmax = fminf(t1, tmax);
    }
    {
        float t0 = (lo.y - r.o.y) * r.inv.y, t1 = (hi.y - r.o.y) * r.inv.y;
        if (r.inv.y < 0.0f) { float s = t0; t0 = t1; t1 = s; }
        t1 = t1 * (1.0f + 0.001f);
        tmin = fmaxf(t0, tmin); tmax = fminf(t1, tmax);
    }
    {
        float t0 = (lo.z - r.o.z) * r.inv.z, t1 = (hi.z - r.o.z) * r.inv.z;
        if (r.inv.z < 0.0f) { float s = t0; t0 = t1; t1 = s; }
        t1 = t1 * (1.0f + 0.001f);
        tmin = fmaxf(t0, tmin); tmax = fminf(t1, tmax);
    }
    tn = tmin;
    return !(tmax <= tmin);
}
/* the same test on an octant box (lys_scene.h: near/far picked at build time as the swap above would) */
LYS_D bool slab_test_oct(const RayInv &r, float4 nr, float4 fr, float tmax, float &tn) {
    float tmin = 0.0f;
    tmin = fmaxf((nr.x - r.o.x) * r.inv.x, tmin); tmax = fminf(((fr.x - r.o.x) * r.inv.x) * (1.0f + 0.001f), tmax);
    tmin = fmaxf((nr.y - r.o.y) * r.inv.y, tmin); tmax = fminf(((fr.y - r.o.y) * r.inv.y) * (1.0f + 0.001f), tmax);
    tmin = fmaxf((nr.z - r.o.z) * r.inv.z, tmin); tmax = fminf(((fr.z - r.o.z) * r.inv.z) * (1.0f + 0.001f), tmax);
    tn = tmin;
    return !(tmax <= tmin);
}
/* hit_triangle (shapes.fut:66-86) against sorted leaf `leaf`: the plane part needs only (a, n = e1 x e2), one 32-byte
 * sector; the edges are fetched only by the lanes whose t lies in (0, tmax). */
LYS_D bool leaf_test_at(const RayInv &r, const float4 *__restrict__ q, float tmax, float &t, int &escape) {
    float4 q0, q1; ld_sector(q, q0, q1);
    escape = __float_as_int(q1.w);                       /* where the walk goes after this leaf */
    float inv; V3 s;
    if (!tri_plane_test(r.o, r.d, v3(q0.x, q0.y, q0.z), v3(q1.x, q1.y, q1.z), tmax, t, inv, s)) return false;
    float4 q2, q3; ld_sector(q + 2, q2, q3);
    return tri_uv_test(r.d, s, inv, v3(q2.x, q2.y, q2.z), v3(q3.x, q3.y, q3.z));
}
LYS_D bool leaf_test(const RayInv &r, const float4 *__restrict__ leaf_tri, int leaf, float tmax, float &t, int &escape) {
    return leaf_test_at(r, leaf_tri + 4ll * leaf, tmax, t, escape);
}
/* ---- the walk.  Reference order (bvh.fut:126-142): enter a node = test ITS box against the CURRENT tmax; if it passes go
 * left, and come back for the right child when the left subtree is done; leaves are never box-tested.  The walk is left-first
 * WHATEVER THE RAY, so the node that follows a failed box test or a finished leaf is a property of the tree: the right child of
 * the nearest ancestor-or-self that is a left child, or the end marker on the right spine.  The build stores that escape link in
 * every record (lbvh.cu: k_pack_records), and the walk needs neither the reference's parent pointers nor a stack:
 *     internal node:  cur = box passes ? left child : escape link          leaf:  triangle test, then cur = escape link
 * Same tests against the same tmax in the same order, so hits, ties (strict t < tmax, shapes.fut:64) and the culling by the
 * non-conservative truncated boxes (bvh.fut:105-120) are identical (tests/test_traversal_algebra.py: escape links vs the
 * parent-pointer walk; parity tests).  No local memory, 23 instructions per box visit.
 *
 * One loop iteration = TRAV_NB node stages, then one triangle stage, then a warp vote.  A lane takes part in a stage if its
 * next visit has that type, so a node whose left child is a leaf is entered and the leaf triangle-tested in the same
 * iteration, and the lanes of a warp meet in few, well filled stages (model: tools/simt_model.py).  The vote keeps the warp
 * converged at the loop head -- without it the compiler threads "still at an internal node" back into the node stage, i.e.
 * builds a while-while loop.  ALL 32 LANES OF A WARP MUST CALL THIS TOGETHER; lanes without a ray pass active = false.
 *
 * Record layouts (lys_scene.h), picked by scene size in futhark_entry_init:
 *   LAY_OCT  one copy of the records per ray-direction octant, near / far picked at build time: box test without selects
 *            (scenes up to LYS_OCT_MAX_NODES nodes: the copies stay cache resident)
 *   LAY_SEL  one copy, box test with selects (larger scenes)
 * Pair records (both children's boxes in one 64-byte record, right child pushed with its entry distance and re-checked at pop
 * time) were the round-2 layout for scenes above 1024 triangles until the escape links made the stack unnecessary: single-box
 * records now win on every scene size (profiles/README.md 8.10) and the pair code is gone (history: commit 3c87084). */
#define TRAV_DONE ((int)0x80000000)
enum { LAY_OCT = 0, LAY_SEL = 1, LAY_OCT_RF = 2 };      /* LAY_OCT_RF: LAY_OCT records, k_trace with lane refill (scene size, launch_trace) */
#define LYS_LAY_IS_OCT(LAY) ((LAY) != LAY_SEL)
#ifndef TRAV_NB
#define TRAV_NB 2          /* node stages per loop iteration: 2 measured best on every layout and scene size (profiles/README.md 8.2, 8.10) */
#endif
template <bool ANY, int LAY>
LYS_D int traverse(const float4 *__restrict__ nodes, const float4 *__restrict__ leaf_tri, int n_nodes, bool active,
                   V3 o, V3 d, float tmax, float &t_hit) {
    RayInv r; r.o = o; r.d = d; r.inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int closest = -1;
    unsigned long long nbase = reinterpret_cast<unsigned long long>(nodes);        /* per-lane base: node i at nbase + 32 i */
    if (LYS_LAY_IS_OCT(LAY)) nbase += 32ull * (unsigned)n_nodes * (unsigned)(((r.inv.x < 0.0f) ? 4 : 0) | ((r.inv.y < 0.0f) ? 2 : 0) | ((r.inv.z < 0.0f) ? 1 : 0));
    asm volatile("" : "+l"(nbase));      /* keep the sum in a register pair: one IMAD.WIDE per node address */
    int cur = (active && n_nodes > 0) ? 0 : TRAV_DONE;       /* internal node to enter (>= 0), leaf pointer (~leaf) or TRAV_DONE */
    do {
#pragma unroll
        for (int k = 0; k < TRAV_NB; k++) {
            if (cur >= 0) {
                const float4 *q = reinterpret_cast<const float4 *>(nbase + 32ull * (unsigned)cur);
                float4 lo, hi; ld_sector(q, lo, hi);
                float tn;
                cur = __float_as_int((LYS_LAY_IS_OCT(LAY) ? slab_test_oct(r, lo, hi, tmax, tn) : slab_test(r, lo, hi, tmax, tn)) ? lo.w : hi.w);      /* left child, or the escape link */
            }
        }
        if ((unsigned)cur > (unsigned)TRAV_DONE) {            /* a leaf pointer */
            float t; int next;
            if (leaf_test(r, leaf_tri, ~cur, tmax, t, next)) { closest = ~cur; tmax = t; }
            cur = (ANY && closest >= 0) ? TRAV_DONE : next;             /* any_hit stops at the first hit (bvh.fut:152) */
        }
    } while (__any_sync(0xffffffffu, cur != TRAV_DONE));
    t_hit = tmax;
    return closest;
}
/* ------------------------------------------------------------------ camera: camera.fut:68-110, integrator.fut:85-90 */
LYS_DN void camera_sample(const FrameParams &fp, int col, int row, uint32_t &rng, V3 &o, V3 &d, float &wavelen, int &chan) {
    uint32_t x = lcg_next(rng);                                   /* random_select' rand.fut:39-42 */
    chan = (int)(x % (uint32_t)fp.n_sensor);
    float p = rng_unit(rng);
    wavelen = fp.sensor_mu[chan] + fp.sensor_sigma[chan] * det_probitf(p);
    float j = (float)(uint32_t)col;
    float i = fp.fh - (float)(uint32_t)row - 1.0f;
    uint32_t r1 = rng;                                            /* camera.fut:86: rng is only peeked */
    float o0 = rng_unit(r1), o1 = rng_unit(r1);
    float px = (j + fp.offset_radius * o0) / fp.fw;
    float py = (i + fp.offset_radius * o1) / fp.fh;
    uint32_t r2 = rng;                                            /* camera.fut:102: same draws again */
    V3 dk = rng_unit_disk(r2);
    V3 lens = fp.lens_radius * dk;
    V3 lens_offset = lens.x * fp.cam_u + lens.y * fp.cam_v;
    o = fp.cam_origin + lens_offset;
    d = normalise(((fp.llc + px * fp.horizontal) + py * fp.vertical) - o);
}
LYS_D int local_to_pixel(const FrameParams &fp, int pid, int &col, int &row) {
    int rl; path_tile(fp.gw, fp.n_local / fp.gw, pid, col, rl);
    row = rl * fp.world + fp.rank;
    return row * fp.gw + col;
}
LYS_D size_t probe_index(const FrameParams &fp, int pid) { int c, r; return (size_t)local_to_pixel(fp, pid, c, r); }   /* probes are exported in pixel order */
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ FrameParams fp, PassBuffers b) {
    int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid == 0) {
        b.counts[0] = fp.n_local;
        for (int k = 1; k <= LYS_MAX_PATH_LEN; k++) b.counts[k] = 0;
        for (int k = 0; k < 2 * (LYS_MAX_PATH_LEN + 1); k++) b.split[k] = 0;
    }
    if (pid >= fp.n_local) return;
    int col, row; int ix = local_to_pixel(fp, pid, col, row);
    uint32_t rng = fp.frame_rng ^ rng_split_hash((uint32_t)ix);   /* split_rng integrator.fut:109-114 */
    V3 o, d; float wl; int ch;
    camera_sample(fp, col, row, rng, o, d, wl, ch);
    b.ray_o[0][pid] = make_float4(o.x, o.y, o.z, wl);          /* bounce 0: slot == path id */
    b.ray_d[0][pid] = make_float4(d.x, d.y, d.z, __uint_as_float(rng));
    b.dist[0][pid] = 0.0f;
    b.acc[pid] = make_float4(0.0f, 0.0f, LYS_INF, 0.0f);
    b.chan[pid] = (uint8_t)ch;
    b.queue[0][pid] = pid;
    if (b.probe_rad) for (int k = 0; k < LYS_MAX_PATH_LEN; k++) { b.probe_rad[(size_t)ix * 16 + k] = 0.0f; b.probe_dist[(size_t)ix * 16 + k] = LYS_INF; }
}

/* ------------------------------------------------------------------ lights */
/* one light as a vertex sees it: geometry + E = emission at the path's wavelength (light.fut:25,38; the lookup depends on
 * the light and the wavelength only, so it is done once per vertex instead of once per incident_radiance call) */
struct LightD { V3 a, e1, e2, n; float inv_area, theta; int kind; float E; };
LYS_D void load_light(const LightRec *__restrict__ L, float wavelen, LightD &l) {
    const float4 *q = reinterpret_cast<const float4 *>(L);
    float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
    l.a = v3(q0.x, q0.y, q0.z); l.e1 = v3(q1.x, q1.y, q1.z); l.inv_area = q1.w;
    l.e2 = v3(q2.x, q2.y, q2.z); l.theta = q2.w; l.n = v3(q3.x, q3.y, q3.z); l.kind = __float_as_int(q3.w);
    float4 e0 = __ldg(q + 4), e1 = __ldg(q + 5), e2 = __ldg(q + 6);
    float em[12] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z, e2.w};
    l.E = spectrum_lookup12(wavelen, em);
}
/* light k of the scanning transmitter for a primary ray direction (camera.fut:119-121, shapes.fut:17-35) */
LYS_DN void scanning_light(const FrameParams &fp, V3 prim_dir, int k, float wavelen, LightD &l) {
    V3 c = cross(prim_dir, v3(0.0f, 1.0f, 0.0f));
    V3 right = (norm(c) == 0.0f) ? v3(1.0f, 0.0f, 0.0f) : normalise(c);
    V3 up = normalise(cross(right, prim_dir));
    V3 v0 = fp.sector_x[k] * right + fp.sector_y[k] * up;
    V3 v1 = fp.sector_x[k + 1] * right + fp.sector_y[k + 1] * up;
    V3 A = fp.cam_origin, B = fp.cam_origin + fp.tx_radius * v1, C = fp.cam_origin + fp.tx_radius * v0;
    l.a = A; l.e1 = B - A; l.e2 = C - A;
    V3 nc = cross(l.e1, l.e2);
    float area = norm(nc) / 2.0f;
    l.inv_area = 1.0f / area; l.n = normalise(nc); l.theta = fp.tx_theta; l.kind = 1;
    l.E = spectrum_lookup12(wavelen, fp.tx_emission);
}
/* arealight_incident_radiance (light.fut:19-55) */
LYS_D float incident_radiance(const LightD &l, V3 hitp, V3 lightp) {
    V3 v = lightp - hitp;
    V3 wi = normalise(v);
    float d2 = quadrance(v);
    float cl = dot(-wi, l.n);
    const float E = l.E;
    if (l.kind == 0) return lys_fmaxf(0.0f, E * cl / d2);
    return (det_acosf(cl) <= l.theta) ? E / d2 : 0.0f;
}
LYS_D float balance1(float pf, float pg) { return 1.0f * pf / (1.0f * pf + 1.0f * pg); }   /* direct.fut:56-58, nf = ng = 1 */

/* ------------------------------------------------------------------ shade
 * One vertex = prologue (t of the winning leaf, material, frame) + light sample + BSDF-MIS sample + continuation.
 * k_shade runs all of it (see its header for the CTA-level arrangements). */
struct VertexCtx {
    int pid, leaf;
    V3 o, d, pos, n, wo, wo_l;
    float wavelen, t;
    uint32_t rng;                 /* state after the per-vertex advance_rng (integrator.fut:48) */
    Onb onb; Mat1 m; const float *mrow;
    bool dark;                    /* the material's emission spectrum is +0.0 everywhere (SceneDev::mat_flag bit 1) */
    float F;                      /* schlick(wo_l, m) (material.fut:207-211) when wo_l.z > 0: used by the light sample's bsdf_f and by both sample_dir calls */
};
LYS_D bool shade_prologue(const SceneDev &sc, const PassBuffers &b, int bounce, int i, VertexCtx &v) {
    v.pid = b.queue[bounce & 1][i];
    float4 ro4 = b.ray_o[bounce & 1][i], rd4 = b.ray_d[bounce & 1][i];
    v.o = v3(ro4.x, ro4.y, ro4.z); v.d = v3(rd4.x, rd4.y, rd4.z);
    v.wavelen = ro4.w;
    v.rng = __float_as_uint(rd4.w);
    v.leaf = b.hit[i];
    if (v.leaf < 0) return false;
    const float4 *lq = sc.leaf_tri + 4ll * v.leaf;
    float4 q0 = __ldg(lq), q1 = __ldg(lq + 1);
    v.mrow = sc.mats + 28ll * (int)__float_as_uint(q0.w);
    v.dark = (__ldg(sc.mat_flag + (int)__float_as_uint(q0.w)) & 2) != 0;
    V3 nc = v3(q1.x, q1.y, q1.z);                                          /* e1 x e2, stored by the build */
    { float inv; V3 s; (void)tri_plane_test(v.o, v.d, v3(q0.x, q0.y, q0.z), nc, FLT_MAX, v.t, inv, s); }   /* bvh.fut:143-145: t of the winner */
    v.pos = v.o + v.t * v.d;
    const float4 *fq = sc.leaf_frame + 3ll * v.leaf;                       /* normalise(nc) and mk_orthonormal_basis of it, stored by the build */
    float4 f0 = __ldg(fq), f1 = __ldg(fq + 1), f2 = __ldg(fq + 2);
    v.n = v3(f0.x, f0.y, f0.z);
    v.onb.n = v.n; v.onb.b = v3(f1.x, f1.y, f1.z); v.onb.t = v3(f2.x, f2.y, f2.z);
    rng_advance(v.rng);                                                   /* integrator.fut:48 */
    v.wo = -v.d;
    v.m = material_at(v.mrow, v.wavelen);
    v.wo_l = to_local(v.onb, v.wo);
    v.F = (v.wo_l.z <= 0.0f) ? 0.0f : schlick(v.wo_l, v.m);
    return true;
}
/* direct.fut:116-119: one raw draw picks the light (scene lights, then the 8 transmitter lights of this ray) */
LYS_D void shade_pick_light(const SceneDev &sc, const FrameParams &fp, const PassBuffers &b, VertexCtx &v, int nl, LightD &l) {
    uint32_t pick = lcg_next(v.rng) % (uint32_t)nl;
    if ((int)pick < fp.n_scene_lights) load_light(sc.lights + pick, v.wavelen, l);
    else if (fp.tx_kind == 1) load_light(b.tx_lights + (pick - fp.n_scene_lights), v.wavelen, l);
    else {
        int col, row; int ix = local_to_pixel(fp, v.pid, col, row);
        uint32_t r0 = fp.frame_rng ^ rng_split_hash((uint32_t)ix);
        V3 po, pd; float pw; int pc;
        camera_sample(fp, col, row, r0, po, pd, pw, pc);
        scanning_light(fp, pd, (int)pick - fp.n_scene_lights, v.wavelen, l);
    }
}
/* light sample: sample_arealight peeks two draws (direct.fut:32-42), MIS weight (direct.fut:70-78) */
LYS_D void shade_light_sample(const VertexCtx &v, const LightD &l, float &cL, float4 &rec_d1, int &flags) {
    uint32_t pk = v.rng;
    float u0 = rng_unit(pk), u1 = rng_unit(pk);
    float su = sqrtf(u0);
    float lu = 1.0f - su, lv = u1 * su;                                    /* rand.fut:34-37 */
    V3 p = (l.a + lu * l.e1) + lv * l.e2;
    V3 vv = p - v.pos;
    V3 wi = normalise(vv);
    float in_rad = incident_radiance(l, v.pos, p);
    float pdf = l.inv_area;
    bool facing = !(dot(wi, v.n) <= 0.0f);                                 /* direct.fut:12 */
    if (facing && !(pdf == 0.0f || in_rad == 0.0f)) {                      /* direct.fut:51-53,73-74 */
        float f, spdf;
        uber_eval(v.wo_l, to_local(v.onb, wi), v.m, f, spdf, true, v.F);
        f = f * lys_fabsf(dot(wi, v.n));
        float weight = balance1(pdf, spdf);
        cL = f * weight * in_rad / pdf;
        V3 sd = normalise(wi);
        rec_d1 = make_float4(sd.x, sd.y, sd.z, norm(vv) - 0.01f);
        flags |= 1;
    }
}
/* BSDF sample towards the same light (direct.fut:83-102) for an already drawn sample s (world space) */
LYS_D void shade_bsdf_light_use(const VertexCtx &v, const LightD &l, const DirSample &s, float &cB, float4 &rec_d2, int &flags) {
    V3 bo, bd; ray_from_hit(v.pos, v.n, s.wi, bo, bd);
    float tl; V3 ncl;
    if (tri_test(bo, bd, l.a, l.e1, l.e2, FLT_MAX, tl, ncl)) {
        V3 lp = bo + tl * bd;
        V3 vv = lp - v.pos;
        V3 w = normalise(vv);
        if (!(dot(w, v.n) <= 0.0f) && s.kind != PDF_IMPOSSIBLE) {
            float in_rad = incident_radiance(l, v.pos, lp);
            float f = s.bsdf * lys_fabsf(dot(s.wi, v.n));
            if (s.kind == PDF_DELTA) cB = f * in_rad;
            else { float weight = balance1(s.pdf, l.inv_area); cB = f * in_rad * weight / s.pdf; }
            V3 sd = normalise(w);
            rec_d2 = make_float4(sd.x, sd.y, sd.z, norm(vv) - 0.01f);
            flags |= 2;
        }
    }
}
/* the same with the sample drawn here; advances v.rng */
LYS_D void shade_bsdf_light_sample(VertexCtx &v, const LightD &l, float &cB, float4 &rec_d2, int &flags) {
    DirSample s = sample_bsdf(v.wo, v.onb, v.m, v.rng);
    shade_bsdf_light_use(v, l, s, cB, rec_d2, flags);
}
/* emission, distance, shadow record, continuation + roulette (integrator.fut:51-75); returns true if the path lives on */
LYS_D bool shade_finish_use(const FrameParams &fp, const PassBuffers &b, int bounce, int i, VertexCtx &v, const DirSample &s, float cL, float cB, int flags,
                            float4 &next_o, float4 &next_d, float &next_dist) {
    float em = (bounce == 0 && !v.dark) ? spectrum_lookup12(v.wavelen, v.mrow + 16) : 0.0f;      /* integrator.fut:51-53 */
    const float dist = b.dist[bounce & 1][i] + v.t;                                     /* :54 */
    next_dist = dist;
    V3 so = v.pos + 0.001f * v.n;       /* mkray_adjust_acne with dot(w, n) > 0: same_side = 1 * n */
    b.sh_o[i] = make_float4(so.x, so.y, so.z, __int_as_float(flags));
    b.sh_c[i] = make_float4(cL, cB, em, dist);
    float pdf = (s.kind == PDF_IMPOSSIBLE) ? 0.0f : ((s.kind == PDF_DELTA) ? 1.0f : s.pdf);
    float cosf = lys_fabsf(dot(v.n, s.wi));
    float p_term = 1.0f - s.bsdf * cosf / pdf;
    bool terminate = rng_unit(v.rng) < p_term;
    if (!(pdf == 0.0f || terminate) && bounce + 1 < fp.path_len) {
        V3 no, nd; ray_from_hit(v.pos, v.n, s.wi, no, nd);
        next_o = make_float4(no.x, no.y, no.z, v.wavelen);
        next_d = make_float4(nd.x, nd.y, nd.z, __uint_as_float(v.rng));
        return true;
    }
    return false;
}
LYS_D bool shade_finish(const FrameParams &fp, const PassBuffers &b, int bounce, int i, VertexCtx &v, float cL, float cB, int flags,
                        float4 &next_o, float4 &next_d, float &next_dist) {
    DirSample s = sample_bsdf(v.wo, v.onb, v.m, v.rng);       /* integrator.fut:56 */
    return shade_finish_use(fp, b, bounce, i, v, s, cL, cB, flags, next_o, next_d, next_dist);
}
LYS_D void shade_miss(const FrameParams &fp, const PassBuffers &b, int i, const VertexCtx &v) {
    float amb = spectrum_lookup12(v.wavelen, fp.ambience);                 /* integrator.fut:76 */
    b.sh_o[i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(4));          /* bit 2: miss vertex, radiance = sh_c.z */
    b.sh_c[i] = make_float4(0.0f, 0.0f, amb, LYS_INF);
}
/* compaction of live paths (warp ballot + prefix popcount, one atomic per warp) */
LYS_D void shade_compact(const PassBuffers &b, int bounce, bool alive, int pid, float4 next_o, float4 next_d, float next_dist) {
    const int lane = threadIdx.x & 31;
    unsigned mask = __ballot_sync(0xffffffffu, alive);
    int base = 0;
    if (lane == 0 && mask) base = atomicAdd(&b.counts[bounce + 1], __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (alive) {
        const int slot = base + __popc(mask & ((1u << lane) - 1u));
        b.queue[(bounce + 1) & 1][slot] = pid;
        b.ray_o[(bounce + 1) & 1][slot] = next_o; b.ray_d[(bounce + 1) & 1][slot] = next_d; b.dist[(bounce + 1) & 1][slot] = next_dist;       /* coalesced: consecutive lanes, consecutive slots */
    }
}
/* statistics (vertices, shadow rays) of a thread's whole grid-stride loop: one pair of atomics per warp and launch */
LYS_D void shade_stats(const PassBuffers &b, unsigned n_vert, unsigned n_shadow) {
    unsigned vsum = __reduce_add_sync(0xffffffffu, n_vert), ssum = __reduce_add_sync(0xffffffffu, n_shadow);
    if ((threadIdx.x & 31) == 0 && (vsum | ssum)) { atomicAdd(&b.stats[0], (unsigned long long)vsum); atomicAdd(&b.stats[2], (unsigned long long)ssum); }
}

/* monolithic: everything for one vertex in one thread, with these arrangements (profiles/README.md 4.2, 4.5):
 *  - the reflection branch of uber_sample_dir (material.fut:365-370) is taken by a few percent of the vertices of a
 *    dielectric scene, i.e. by one or two lanes of almost every warp, and costs ~560 instructions (sample_wh,
 *    Torrance-Sparrow terms).  Both sample_dir calls of a vertex therefore only DECIDE their branch in place
 *    (bsdf_choose); refraction samples are drawn in place, reflection samples are queued in shared memory and drawn
 *    after a barrier by the first threads of the CTA, one queue entry per thread (dense warps).  Same inputs, same
 *    arithmetic, another thread.  The rng state after a reflection sample is its two draws further (:283-286);
 *  - for bounces >= 1 the slots are walked in the order k_trace left (hits first): warps are all-hit or all-miss.
 * Two block barriers per iteration (queue filled / queue drained).  What a third one at the end would protect is the queue and its
 * counter only (sample slots are private to their thread outside the drain phase), so those two alternate between iterations
 * (+1 KB of shared memory; double-buffering the whole 13 KB exchange area was measured 1.4 % slower: it costs L1). */
template <int T>
struct ShadeShared {
    float res[6][2 * T];      /* slot k * T + tid: sample k of the thread (in: wo_l, roughness, rng; out: DirSample local) */
    unsigned short q[2][2 * T];  /* slots waiting for a reflection sample; two queues, alternating between iterations */
    int qn[2];
};
template <class SH>
LYS_D void shade_put_sample(SH &sh, int slot, const DirSample &s) {
    sh.res[0][slot] = s.wi.x; sh.res[1][slot] = s.wi.y; sh.res[2][slot] = s.wi.z;
    sh.res[3][slot] = s.bsdf; sh.res[4][slot] = __int_as_float(s.kind); sh.res[5][slot] = s.pdf;
}
template <class SH>
LYS_D DirSample shade_get_sample(const SH &sh, int slot) {
    DirSample s;
    s.wi = v3(sh.res[0][slot], sh.res[1][slot], sh.res[2][slot]);
    s.bsdf = sh.res[3][slot]; s.kind = __float_as_int(sh.res[4][slot]); s.pdf = sh.res[5][slot];
    return s;
}
/* one sample_dir call up to its branch: refraction sample drawn here, reflection sample queued.  Warp-collective. */
template <class SH>
LYS_D void shade_draw_or_queue(SH &sh, int par, bool want, int slot, const VertexCtx &v, uint32_t &rng, bool &metal) {
    bool reflect = false; metal = false;
    if (want) {
        reflect = bsdf_choose(v.wo_l, v.m, rng, metal, true, v.F);
        if (reflect) {
            sh.res[0][slot] = v.wo_l.x; sh.res[1][slot] = v.wo_l.y; sh.res[2][slot] = v.wo_l.z;
            sh.res[3][slot] = v.m.roughness; sh.res[4][slot] = __uint_as_float(rng);
            (void)lcg_next(rng); (void)lcg_next(rng);                 /* u0, u1 of sample_wh */
        } else shade_put_sample(sh, slot, sample_refraction(v.wo_l, v.m, rng));
    }
    const unsigned mask = __ballot_sync(0xffffffffu, reflect);
    if (mask) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(mask) - 1) base = atomicAdd(&sh.qn[par], __popc(mask));
        base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
        if (reflect) sh.q[par][base + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)slot;
    }
}
#ifndef LYS_SHADE_MINB
#define LYS_SHADE_MINB(T) 3      /* 256 threads: 3 CTAs per SM / 80 registers measured best */
#endif
#ifndef LYS_SHADE_T
#define LYS_SHADE_T 256          /* CTA size of k_shade (128 / 256 / 512 measured: profiles/README.md 4.4, 4.5, 8.10) */
#endif
template <int SHADE_THREADS>
__global__ void __launch_bounds__(SHADE_THREADS, LYS_SHADE_MINB(SHADE_THREADS)) k_shade(SceneDev sc, const __grid_constant__ FrameParams fp, PassBuffers b, int bounce, int ordered) {
    __shared__ ShadeShared<SHADE_THREADS> sh;
    const int count = b.counts[bounce];
    const int stride = gridDim.x * blockDim.x;
    const int nl = fp.n_scene_lights + ((fp.tx_kind == 0) ? 0 : 8);
    if (threadIdx.x == 0) { sh.qn[0] = 0; sh.qn[1] = 0; }
    __syncthreads();
    unsigned tot_vert = 0, tot_shadow = 0;
    int par = 0;
    for (int b0 = blockIdx.x * blockDim.x; b0 < count; b0 += stride, par ^= 1) {
        const bool valid = b0 + (int)threadIdx.x < count;
        const int i = !valid ? 0 : ordered ? b.order[bounce & 1][b0 + threadIdx.x] : b0 + (int)threadIdx.x;      /* hits first: warps are all-hit or all-miss */
        bool alive = false, hit = false; int pid = -1;
        float4 next_o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), next_d = next_o; float next_dist = 0.0f;
        VertexCtx v;
        float cL = 0.0f, cB = 0.0f; int flags = 0;
        float4 rec_d1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rec_d2 = rec_d1;
        LightD l;
        if (valid) hit = shade_prologue(sc, b, bounce, i, v);
        const bool lit = hit && nl > 0;
        if (lit) shade_pick_light(sc, fp, b, v, nl, l);
        if (lit) shade_light_sample(v, l, cL, rec_d1, flags);
        if (hit) b.sh_d1[i] = rec_d1;                                     /* records leave the registers as soon as they are complete */
        /* both sample_dir calls of the vertex: the MIS sample (direct.fut:83) and the continuation (integrator.fut:56) */
        bool metal1, metal2;
        shade_draw_or_queue(sh, par, lit, threadIdx.x, v, v.rng, metal1);
        shade_draw_or_queue(sh, par, hit, SHADE_THREADS + threadIdx.x, v, v.rng, metal2);
        __syncthreads();
        for (int j = threadIdx.x; j < sh.qn[par]; j += SHADE_THREADS) {
            const int slot = sh.q[par][j];
            Mat1 m; m.color = 0.0f; m.roughness = sh.res[3][slot]; m.metalness = 0.0f; m.ref_ix = 0.0f; m.opacity = 0.0f;
            uint32_t rng = __float_as_uint(sh.res[4][slot]);
            shade_put_sample(sh, slot, sample_reflection(v3(sh.res[0][slot], sh.res[1][slot], sh.res[2][slot]), m, rng));
        }
        __syncthreads();
        if (threadIdx.x == 0) sh.qn[par] = 0;   /* queue drained: empty for the iteration after the next (the next one uses the other queue) */
        if (lit) {
            DirSample s = shade_get_sample(sh, threadIdx.x);
            if (metal1) s.bsdf = v.m.color * s.bsdf;                      /* metal :352-355 */
            s.wi = to_world(v.onb, s.wi);
            shade_bsdf_light_use(v, l, s, cB, rec_d2, flags);
        }
        if (hit) b.sh_d2[i] = rec_d2;
        if (valid) {
            if (hit) {
                DirSample s = shade_get_sample(sh, SHADE_THREADS + threadIdx.x);
                if (metal2) s.bsdf = v.m.color * s.bsdf;
                s.wi = to_world(v.onb, s.wi);
                alive = shade_finish_use(fp, b, bounce, i, v, s, cL, cB, flags, next_o, next_d, next_dist);
                tot_vert += 1; tot_shadow += (flags & 1) + ((flags >> 1) & 1);
            } else shade_miss(fp, b, i, v);
            pid = v.pid;
        }
        shade_compact(b, bounce, alive, pid, next_o, next_d, next_dist);
        /* no barrier here: after the second barrier a thread only reads its OWN two sample slots, and the next iteration
         * writes its own slots and the OTHER queue; this iteration's queue is touched again two barriers later */
    }
    shade_stats(b, tot_vert, tot_shadow);
}
/* ------------------------------------------------------------------ trace: all BVH traversal of one bounce boundary
 * One persistent launch resolves the shadow rays of bounce `bounce` (connect) and the closest hits of bounce
 * `bounce + 1` (extend); both only read the scene and touch disjoint path state.  bounce = -1: primary rays only.
 * A connect item carries up to two shadow rays; after them the vertex radiance is accumulated (direct.fut:121-122,
 * integrator.fut:51-55). */
LYS_D void connect_finish(const FrameParams &fp, const PassBuffers &b, int bounce, int slot, int flags, float L, float B, float em, float dist, float miss_r) {
    int pid = b.queue[bounce & 1][slot];
    float r;
    if (flags & 4) r = miss_r;                        /* miss vertex: {inf, ambience} (integrator.fut:76) */
    else {
        const int nl = fp.n_scene_lights + ((fp.tx_kind == 0) ? 0 : 8);
        float direct = 0.0f;
        if (nl > 0) { float light_pdf = 1.0f / (float)nl; direct = (L + B) / light_pdf; }   /* direct.fut:121-122 */
        r = direct + ((bounce == 0) ? em : 0.0f);                                          /* integrator.fut:51-53 */
    }
    float4 a = b.acc[pid];
    a.x = a.x + r * 1.0f;
    a.y = a.y + r * 0.0f;
    float inten = r * fp.intensity_factor;
    if (inten > 0.0f && dist > 0.5f && dist < 10.0f && dist < a.z) { a.z = dist; a.w = inten; }
    b.acc[pid] = a;
    if (b.probe_rad) { const size_t ix = probe_index(fp, pid); b.probe_rad[ix * 16 + bounce] = r; b.probe_dist[ix * 16 + bounce] = dist; }
}
/* Batch loop (LAY_OCT, and the shadow rays of LAY_OCT_RF): a warp owns 32 consecutive items per grid-stride step and runs the
 * vote-synchronised traverse<> loops on them.  With a per-ray stack, refilling idle lanes inside the loop did not pay (round 1 and
 * early round 2: -38 % and +3 % on the 1 M-triangle scene, profiles/README.md 7.2, 8.1); with the stackless walk it does on scenes
 * whose walks are long -- trace_ext_refill / trace_con_refill below, selected by scene size (profiles/README.md 8.11).
 * `ordered` bit 0: append the slots of bounce + 1 to the hits-first order list and walk this bounce's vertices in theirs. */
/* resident CTAs of 128 threads per SM the traversal kernels are compiled for.  The walk waits on dependent record loads, so
 * more warps in flight win even with a few registers spilled to L1: with octant copies 12 CTAs (40 registers) against 10 (48, no
 * spills) give +0.5 % on CornellBox and +3-4 % from 2 K to 64 K triangles (16: no further gain); the large scenes run 1106 /
 * 1147 / 1176 Mpaths/s at 10 / 12 / 16 CTAs (32 registers) on the 1 M-triangle scene (profiles/README.md 8.10). */
#ifndef LYS_TRACE_MINB
#define LYS_TRACE_MINB(LAY) ((LAY) == LAY_SEL ? 16 : 12)
#endif
#ifndef LYS_REFILL_MIN_TRIS
#define LYS_REFILL_MIN_TRIS 1024     /* lane refill from here on: +2.4 / +2.9 % on SpectrumSphere / High, -2 % on CornellBox (profiles/README.md 8.11) */
#endif
/* Closest hits with LANE REFILL (large scenes, LAY_SEL).  The stackless walk's state is (cur, tmax, closest) plus the ray, so a
 * lane that has finished its walk can take the warp's next item instead of idling until the longest walk of its batch of 32 ends
 * (walk lengths differ 20x on the 1 M-triangle scene: 11-15 of 32 lanes were busy).  The warp owns the items of its grid-stride
 * batches in sequence; a refill event -- at most LYS_TRACE_REFILL lanes busy -- writes the finished hits and hands the next items to
 * the idle lanes, rays straight from global memory (the next batch's lines are requested into L1 when a batch is started).
 * The hits-first order list of shade(bounce + 1) is appended 32 finished items at a time from a warp-private list in shared memory:
 * one pair of atomics per 32 items as in the batch loop.  Per-ray arithmetic and decisions are traverse<>'s; only which lane
 * walks which ray, and when, differs.  Measured (profiles/README.md 8.11): +6 % on the 1 M-triangle scene; -2.5 % on CornellBox,
 * where the walks are short and the refill's divisions and votes cost more than the idle lanes -- so LAY_OCT keeps the batch loop. */
#ifndef LYS_TRACE_REFILL
#define LYS_TRACE_REFILL 24
#endif
#ifndef LYS_REFILL_NB
#define LYS_REFILL_NB 4         /* node stages per iteration of the refill loops: 1 / 2 / 3 / 4 -> 1185 / 1286 / 1284 / 1296 Mpaths/s on the 1 M-triangle scene (an iteration also pays the refill vote) */
#endif
template <int LAY>
LYS_D void trace_ext_refill(const SceneDev &sc, const PassBuffers &b, int bounce, int n_ext, int stride, int ordered) {
    __shared__ int s_fin[4][64];                          /* per warp: finished items waiting for their order-list slot, (item << 1) | hit */
    constexpr bool CS = LAY == LAY_SEL;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int n_nodes = (int)sc.n_tris - 1;
    const int wbase = (blockIdx.x * blockDim.x + (threadIdx.x & ~31));          /* first item of the warp's batch 0 */
    if (wbase >= n_ext) return;
    const int n_batches = (n_ext - wbase - 1) / stride + 1;                       /* batches j with wbase + j * stride < n_ext */
    const int s_end = n_batches * 32;
    const float4 *__restrict__ ray_o = b.ray_o[(bounce + 1) & 1], *__restrict__ ray_d = b.ray_d[(bounce + 1) & 1];
    int *fin = s_fin[(threadIdx.x >> 5) & 3];
    int fn = 0;                                           /* warp-uniform */
    int s_next = 0;                                       /* warp-uniform: next position of the warp's item sequence */
    int it = -1, cur = TRAV_DONE, closest = -1;
    float tmax = 0.0f;
    RayInv r; r.o = v3(0.0f, 0.0f, 0.0f); r.d = v3(1.0f, 1.0f, 1.0f); r.inv = r.d;
    unsigned long long nbase = reinterpret_cast<unsigned long long>(sc.nodes);
    /* order list of shade(bounce + 1): hits from the front, misses from the back; the last cnt <= 32 entries of the warp's list */
    auto flush = [&](int cnt) {
        const int code = (lane < cnt) ? fin[fn - cnt + lane] : 0;
        __syncwarp();
        fn -= cnt;
        const bool isH = lane < cnt && (code & 1), isM = lane < cnt && !(code & 1);
        const unsigned mh = __ballot_sync(0xffffffffu, isH), mm = __ballot_sync(0xffffffffu, isM);
        int bh = 0, bm = 0;
        if (lane == 0) { if (mh) bh = atomicAdd(&b.split[2 * (bounce + 1)], __popc(mh)); if (mm) bm = atomicAdd(&b.split[2 * (bounce + 1) + 1], __popc(mm)); }
        bh = __shfl_sync(0xffffffffu, bh, 0); bm = __shfl_sync(0xffffffffu, bm, 0);
        if (isH) st_state<CS>(&b.order[(bounce + 1) & 1][bh + __popc(mh & lt)], code >> 1);
        if (isM) st_state<CS>(&b.order[(bounce + 1) & 1][n_ext - 1 - (bm + __popc(mm & lt))], code >> 1);
    };
    for (;;) {
        const unsigned busy = __ballot_sync(0xffffffffu, cur != TRAV_DONE);
        if (__popc(busy) <= LYS_TRACE_REFILL && (s_next < s_end || busy == 0u)) {
            const bool fin_now = cur == TRAV_DONE && it >= 0;
            if (fin_now) st_state<CS>(&b.hit[it], closest);
            if (ordered) {
                const unsigned mf = __ballot_sync(0xffffffffu, fin_now);
                if (mf) {
                    if (fin_now) fin[fn + __popc(mf & lt)] = (it << 1) | (closest >= 0 ? 1 : 0);
                    fn += __popc(mf);
                    __syncwarp();
                    if (fn >= 32) flush(32);
                }
            }
            if (fin_now) it = -1;
            if (s_next >= s_end) { if (busy == 0u) break; }
            else {
                const unsigned idle = ~busy;
                const int s = s_next + __popc(idle & lt);
                s_next = min(s_next + __popc(idle), s_end);
                if (cur == TRAV_DONE && s < s_end) {
                    const int g = wbase + (s >> 5) * stride + (s & 31);
                    if (g < n_ext) {
                        const float4 ro = ld_state<CS>(&ray_o[g]), rd = ld_state<CS>(&ray_d[g]);
                        r.o = v3(ro.x, ro.y, ro.z); r.d = v3(rd.x, rd.y, rd.z); r.inv = v3(1.0f / rd.x, 1.0f / rd.y, 1.0f / rd.z);
                        nbase = reinterpret_cast<unsigned long long>(sc.nodes);
                        if (LYS_LAY_IS_OCT(LAY)) nbase += 32ull * (unsigned)n_nodes * (unsigned)(((r.inv.x < 0.0f) ? 4 : 0) | ((r.inv.y < 0.0f) ? 2 : 0) | ((r.inv.z < 0.0f) ? 1 : 0));
                        it = g; cur = 0; closest = -1; tmax = FLT_MAX;
#ifdef __CUDACC__
                        if ((s & 7) == 0) {       /* one lane per 128-byte line of the batch started: request the next batch's records */
                            const int gn = g + stride;
                            if (gn < n_ext) { asm volatile("prefetch.global.L1 [%0];" :: "l"(&ray_o[gn])); asm volatile("prefetch.global.L1 [%0];" :: "l"(&ray_d[gn])); }
                        }
#endif
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < LYS_REFILL_NB; k++) {
            if (cur >= 0) {
                const float4 *q = reinterpret_cast<const float4 *>(nbase + 32ull * (unsigned)cur);
                float4 lo, hi; ld_sector(q, lo, hi);
                float tn;
                cur = __float_as_int((LYS_LAY_IS_OCT(LAY) ? slab_test_oct(r, lo, hi, tmax, tn) : slab_test(r, lo, hi, tmax, tn)) ? lo.w : hi.w);
            }
        }
        if ((unsigned)cur > (unsigned)TRAV_DONE) {
            float t; int next;
            if (leaf_test(r, sc.leaf_tri, ~cur, tmax, t, next)) { closest = ~cur; tmax = t; }
            cur = next;
        }
    }
    if (ordered && fn > 0) flush(fn);
}
/* Shadow rays with LANE REFILL (large scenes, LAY_SEL).  A lane owns one vertex of shade(bounce) at a time and takes it through
 * light-sample ray -> BSDF-sample ray -> radiance (connect_finish); at a refill event (at most LYS_CON_REFILL lanes walking) every lane
 * whose walk has ended moves its vertex on -- starts its second ray, or finishes it and takes the warp's next vertex.  Which lane walks
 * which ray differs from the batch loop, nothing else: a vertex is finished exactly once, with the same L and B.  Measured
 * (profiles/README.md 8.11): 1 M triangles +4 % (threshold 16); -1 ... -3 % with octant copies from 2 K to 9 K triangles, which keep the
 * batch loop for their shadow rays. */
#ifndef LYS_CON_REFILL
#define LYS_CON_REFILL 16
#endif
template <int LAY>
LYS_D void trace_con_refill(const SceneDev &sc, const FrameParams &fp, const PassBuffers &b, int bounce, int n_con, int stride, int ordered) {
    constexpr bool CS = LAY == LAY_SEL;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int n_nodes = (int)sc.n_tris - 1;
    const int wbase = (blockIdx.x * blockDim.x + (threadIdx.x & ~31));
    if (wbase >= n_con) return;
    const int n_batches = (n_con - wbase - 1) / stride + 1;
    const int s_end = n_batches * 32;
    const bool via_order = ordered && bounce >= 1;
    const int *__restrict__ order = b.order[bounce & 1];
    int s_next = 0;
    int slot = -1, st = 0, cur = TRAV_DONE;            /* st: bit 0 second ray still to walk, bit 1 first ray unoccluded, bit 2 walking the second ray, bit 3 this walk hit */
    float tmax = 0.0f;
    RayInv r; r.o = v3(0.0f, 0.0f, 0.0f); r.d = v3(1.0f, 1.0f, 1.0f); r.inv = r.d;
    unsigned long long nbase = reinterpret_cast<unsigned long long>(sc.nodes);
    for (;;) {
        const unsigned busy = __ballot_sync(0xffffffffu, cur != TRAV_DONE);
        if (__popc(busy) <= LYS_CON_REFILL && (s_next < s_end || busy == 0u)) {
            const float4 *dir_from = nullptr;                                   /* the ray this lane starts in this event */
            if (cur == TRAV_DONE && slot >= 0) {                                /* a walk has ended */
                bool finish = true;
                if (!(st & 4)) {                                                /* ... the light-sample ray */
                    if (!(st & 8)) st |= 2;
                    if (st & 1) { st = (st & 2) | 4; dir_from = &b.sh_d2[slot]; finish = false; }
                }
                if (finish) {
                    const float4 rc = ld_state<CS>(&b.sh_c[slot]);
                    const float L = (st & 2) ? rc.x : 0.0f, B = ((st & 4) && !(st & 8)) ? rc.y : 0.0f;
                    connect_finish(fp, b, bounce, slot, 0, L, B, rc.z, rc.w, rc.z);
                    slot = -1;
                }
            }
            /* idle lanes take the warp's next vertices */
            const unsigned idle = __ballot_sync(0xffffffffu, cur == TRAV_DONE && slot < 0);
            if (s_next < s_end && idle) {
                const int s = s_next + __popc(idle & lt);
                s_next = min(s_next + __popc(idle), s_end);
                if (cur == TRAV_DONE && slot < 0 && s < s_end) {
                    const int c = wbase + (s >> 5) * stride + (s & 31);
                    if (c < n_con) {
                        const int sl = via_order ? ld_state<CS>(&order[c]) : c;
                        const float4 ro = ld_state<CS>(&b.sh_o[sl]);
                        const int flags = __float_as_int(ro.w);
                        const bool need1 = !(flags & 4) && (flags & 1), need2 = !(flags & 4) && (flags & 2);
                        if (!need1 && !need2) {
                            const float4 rc = ld_state<CS>(&b.sh_c[sl]);
                            connect_finish(fp, b, bounce, sl, flags, 0.0f, 0.0f, rc.z, rc.w, rc.z);
                        } else {
                            slot = sl; r.o = v3(ro.x, ro.y, ro.z);
                            if (need1) { st = need2 ? 1 : 0; dir_from = &b.sh_d1[sl]; }
                            else { st = 4; dir_from = &b.sh_d2[sl]; }
                        }
#ifdef __CUDACC__
                        if ((s & 7) == 0) {                                     /* next batch: its order entries, or its records */
                            const int cn = c + stride;
                            if (cn < n_con) {
                                if (via_order) asm volatile("prefetch.global.L1 [%0];" :: "l"(&order[cn]));
                                else { asm volatile("prefetch.global.L1 [%0];" :: "l"(&b.sh_o[cn])); asm volatile("prefetch.global.L1 [%0];" :: "l"(&b.sh_d1[cn])); }
                            }
                        }
#endif
                    }
                }
            }
            if (dir_from) {
                const float4 d = ld_state<CS>(dir_from);
                r.d = v3(d.x, d.y, d.z); r.inv = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                nbase = reinterpret_cast<unsigned long long>(sc.nodes);
                if (LYS_LAY_IS_OCT(LAY)) nbase += 32ull * (unsigned)n_nodes * (unsigned)(((r.inv.x < 0.0f) ? 4 : 0) | ((r.inv.y < 0.0f) ? 2 : 0) | ((r.inv.z < 0.0f) ? 1 : 0));
                tmax = d.w; cur = 0; st &= ~8;
            }
            if (s_next >= s_end && __ballot_sync(0xffffffffu, cur != TRAV_DONE || slot >= 0) == 0u) break;
        }
#pragma unroll
        for (int k = 0; k < LYS_REFILL_NB; k++) {
            if (cur >= 0) {
                const float4 *q = reinterpret_cast<const float4 *>(nbase + 32ull * (unsigned)cur);
                float4 lo, hi; ld_sector(q, lo, hi);
                float tn;
                cur = __float_as_int((LYS_LAY_IS_OCT(LAY) ? slab_test_oct(r, lo, hi, tmax, tn) : slab_test(r, lo, hi, tmax, tn)) ? lo.w : hi.w);
            }
        }
        if ((unsigned)cur > (unsigned)TRAV_DONE) {
            float t; int next;
            if (leaf_test(r, sc.leaf_tri, ~cur, tmax, t, next)) { st |= 8; cur = TRAV_DONE; }       /* any_hit stops at the first hit (bvh.fut:152) */
            else cur = next;
        }
    }
}
template <int LAY>
__global__ void __launch_bounds__(128, LYS_TRACE_MINB(LAY)) k_trace(SceneDev sc, const __grid_constant__ FrameParams fp, PassBuffers b, int bounce, int ordered) {
    const int n_ext = (bounce + 1 < fp.path_len) ? b.counts[bounce + 1] : 0;
    const int n_con = (bounce >= 0) ? b.counts[bounce] : 0;
    const int total = n_ext + n_con;
    const int stride = gridDim.x * blockDim.x;
    const int n_nodes = (int)sc.n_tris - 1;
    const int lane = threadIdx.x & 31;
    const float4 *__restrict__ nodes = sc.nodes;
    constexpr bool CS = LAY == LAY_SEL;      /* large scenes: path-state records bypass L2 residency (lys_device.cuh: ld_state) */
    /* large scenes: the closest hits first, with lane refill; the loop below is then left with the shadow rays */
    constexpr bool refill_ext = LAY != LAY_OCT;
    if (refill_ext) trace_ext_refill<LAY>(sc, b, bounce, n_ext, stride, ordered);
    if (LAY == LAY_SEL) { trace_con_refill<LAY>(sc, fp, b, bounce, n_con, stride, ordered); return; }      /* ... and the shadow rays */
    /* warp-uniform loop (traverse<> votes): a warp owns 32 consecutive items; only the warp that straddles n_ext mixes kinds */
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31) + (refill_ext ? n_ext : 0); i0 < total; i0 += stride) {
        const int i = i0 + lane;
        const bool is_ext = i < n_ext, is_con = !is_ext && i < total;
        if (i0 < n_ext) {
            float4 ro = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rd = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            if (is_ext) { ro = ld_state<CS>(&b.ray_o[(bounce + 1) & 1][i]); rd = ld_state<CS>(&b.ray_d[(bounce + 1) & 1][i]); }
            float t;
            int h = traverse<false, LAY>(nodes, sc.leaf_tri, n_nodes, is_ext, v3(ro.x, ro.y, ro.z), v3(rd.x, rd.y, rd.z), FLT_MAX, t);
            if (is_ext) st_state<CS>(&b.hit[i], h);
            /* processing order of shade(bounce + 1): hits from the front, misses from the back (one atomic per warp and kind) */
            if (ordered) {                     /* not for camera rays (their misses are whole warps already: shade(0) walks the slots in order) */
                const bool isH = is_ext && h >= 0, isM = is_ext && h < 0;
                const unsigned mh = __ballot_sync(0xffffffffu, isH), mm = __ballot_sync(0xffffffffu, isM);
                int bh = 0, bm = 0;
                if (lane == 0) { if (mh) bh = atomicAdd(&b.split[2 * (bounce + 1)], __popc(mh)); if (mm) bm = atomicAdd(&b.split[2 * (bounce + 1) + 1], __popc(mm)); }
                bh = __shfl_sync(0xffffffffu, bh, 0); bm = __shfl_sync(0xffffffffu, bm, 0);
                const unsigned lt = (1u << lane) - 1u;
                if (isH) st_state<CS>(&b.order[(bounce + 1) & 1][bh + __popc(mh & lt)], i);
                if (isM) st_state<CS>(&b.order[(bounce + 1) & 1][n_ext - 1 - (bm + __popc(mm & lt))], i);
            }
        }
        if (i0 + 31 >= n_ext) {
            /* vertices of this bounce in the order shade(bounce) took them: hit vertices (shadow rays) first, miss vertices last */
            const int slot = !is_con ? 0 : (ordered && bounce >= 1) ? ld_state<CS>(&b.order[bounce & 1][i - n_ext]) : i - n_ext;
            float4 ro = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(4)), rc = ro;
            if (is_con) { ro = ld_state<CS>(&b.sh_o[slot]); rc = ld_state<CS>(&b.sh_c[slot]); }
            const int flags = __float_as_int(ro.w);
            float L = 0.0f, B = 0.0f, t;
            V3 o = v3(ro.x, ro.y, ro.z);
            const bool need1 = is_con && !(flags & 4) && (flags & 1), need2 = is_con && !(flags & 4) && (flags & 2);
            if (__any_sync(0xffffffffu, need1)) {
                float4 d1 = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                if (need1) d1 = ld_state<CS>(&b.sh_d1[slot]);
                int h = traverse<true, LAY>(nodes, sc.leaf_tri, n_nodes, need1, o, v3(d1.x, d1.y, d1.z), d1.w, t);
                if (need1 && h < 0) L = rc.x;
            }
            if (__any_sync(0xffffffffu, need2)) {
                float4 d2 = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                if (need2) d2 = ld_state<CS>(&b.sh_d2[slot]);
                int h = traverse<true, LAY>(nodes, sc.leaf_tri, n_nodes, need2, o, v3(d2.x, d2.y, d2.z), d2.w, t);
                if (need2 && h < 0) B = rc.y;
            }
            if (is_con) connect_finish(fp, b, bounce, slot, flags, L, B, rc.z, rc.w, rc.z);
        }
    }
}

/* ------------------------------------------------------------------ generate + trace(-1) in one launch
 * The camera ray of a pixel goes straight from the registers into the traversal loop: one launch less per pass and no
 * read-back of the 32-byte ray records just written (k_shade(0) still needs them, so they are written once). */
template <int LAY>
__global__ void __launch_bounds__(128, LYS_TRACE_MINB(LAY)) k_generate_trace(SceneDev sc, const __grid_constant__ FrameParams fp, PassBuffers b) {
    const int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid == 0) {
        b.counts[0] = fp.n_local;
        for (int k = 1; k <= LYS_MAX_PATH_LEN; k++) b.counts[k] = 0;
        for (int k = 0; k < 2 * (LYS_MAX_PATH_LEN + 1); k++) b.split[k] = 0;
    }
    const bool act = pid < fp.n_local;                     /* no early return: traverse<> votes per warp */
    V3 o = v3(0.0f, 0.0f, 0.0f), d = v3(1.0f, 1.0f, 1.0f);
    if (act) {
        int col, row; int ix = local_to_pixel(fp, pid, col, row);
        uint32_t rng = fp.frame_rng ^ rng_split_hash((uint32_t)ix);   /* split_rng integrator.fut:109-114 */
        float wl; int ch;
        camera_sample(fp, col, row, rng, o, d, wl, ch);
        constexpr bool CS = LAY == LAY_SEL;
        st_state<CS>(&b.ray_o[0][pid], make_float4(o.x, o.y, o.z, wl));          /* bounce 0: slot == path id */
        st_state<CS>(&b.ray_d[0][pid], make_float4(d.x, d.y, d.z, __uint_as_float(rng)));
        b.dist[0][pid] = 0.0f;
        b.acc[pid] = make_float4(0.0f, 0.0f, LYS_INF, 0.0f);
        b.chan[pid] = (uint8_t)ch;
        b.queue[0][pid] = pid;
        if (b.probe_rad) for (int k = 0; k < LYS_MAX_PATH_LEN; k++) { b.probe_rad[(size_t)ix * 16 + k] = 0.0f; b.probe_dist[(size_t)ix * 16 + k] = LYS_INF; }
    }
    float t;
    const int h = traverse<false, LAY>(sc.nodes, sc.leaf_tri, (int)sc.n_tris - 1, act, o, d, FLT_MAX, t);
    if (act) st_state<LAY == LAY_SEL>(&b.hit[pid], h);
}

/* ------------------------------------------------------------------ tail: all remaining bounces in one launch
 * Past the first few bounces a pass carries a fraction of a percent of its paths, but every bounce still costs two
 * launches (measured: ~2.3 us of GPU throughput each, whatever the grid size; profiles/README.md 4.6).  k_tail runs
 * shade(b), trace(b) for b = bounce0 .. path_len-1 inside one launch: each CTA owns a contiguous chunk of the
 * bounce-bounce0 queue and keeps wavefronting it on its own, with block barriers between the stages and a CTA-local
 * compaction (live paths stay inside the chunk's slot range of the ping-pong buffers, so the very same per-slot device
 * routines are used).  Per-path arithmetic is unchanged; the host picks bounce0 from an earlier pass's queue lengths
 * and any choice is correct. */
template <int LAY>
__global__ void __launch_bounds__(128) k_tail(SceneDev sc, const __grid_constant__ FrameParams fp, PassBuffers b, int bounce0) {
    __shared__ int s_next;
    const int count0 = b.counts[bounce0];
    const int per = (((count0 + (int)gridDim.x - 1) / (int)gridDim.x) + 31) & ~31;
    const int base = blockIdx.x * per;
    int n_cur = max(0, min(per, count0 - base));
    const int nl = fp.n_scene_lights + ((fp.tx_kind == 0) ? 0 : 8);
    const int n_nodes = (int)sc.n_tris - 1;
    const int lane = threadIdx.x & 31;
    const float4 *__restrict__ nodes = sc.nodes;
    for (int bounce = bounce0; bounce < fp.path_len && n_cur > 0; bounce++) {
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        /* ---- shade(bounce): one vertex per thread (the CTA-wide reflection queue of k_shade does not pay for a handful of paths) */
        for (int r0 = 0; r0 < n_cur; r0 += blockDim.x) {
            const int i = base + r0 + threadIdx.x;
            const bool valid = r0 + (int)threadIdx.x < n_cur;
            bool alive = false; int pid = -1; unsigned n_vert = 0, n_shadow = 0;
            float4 next_o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), next_d = next_o; float next_dist = 0.0f;
            if (valid) {
                VertexCtx v;
                if (shade_prologue(sc, b, bounce, i, v)) {
                    float cL = 0.0f, cB = 0.0f; int flags = 0;
                    float4 rec_d1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rec_d2 = rec_d1;
                    if (nl > 0) {
                        LightD l;
                        shade_pick_light(sc, fp, b, v, nl, l);
                        shade_light_sample(v, l, cL, rec_d1, flags);
                        shade_bsdf_light_sample(v, l, cB, rec_d2, flags);
                    }
                    b.sh_d1[i] = rec_d1; b.sh_d2[i] = rec_d2;
                    alive = shade_finish(fp, b, bounce, i, v, cL, cB, flags, next_o, next_d, next_dist);
                    n_vert = 1; n_shadow = (flags & 1) + ((flags >> 1) & 1);
                } else shade_miss(fp, b, i, v);
                pid = v.pid;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, alive);
            int off = 0;
            if (lane == 0 && mask) off = atomicAdd(&s_next, __popc(mask));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (alive) {
                const int slot = base + off + __popc(mask & ((1u << lane) - 1u));
                b.queue[(bounce + 1) & 1][slot] = pid;
                b.ray_o[(bounce + 1) & 1][slot] = next_o; b.ray_d[(bounce + 1) & 1][slot] = next_d; b.dist[(bounce + 1) & 1][slot] = next_dist;
            }
            unsigned vsum = __reduce_add_sync(0xffffffffu, n_vert), ssum = __reduce_add_sync(0xffffffffu, n_shadow);
            if (lane == 0 && (vsum | ssum)) { atomicAdd(&b.stats[0], (unsigned long long)vsum); atomicAdd(&b.stats[2], (unsigned long long)ssum); }
        }
        __syncthreads();
        const int n_next = (bounce + 1 < fp.path_len) ? s_next : 0;
        if (threadIdx.x == 0 && n_next > 0) atomicAdd(&b.counts[bounce + 1], n_next);     /* the next pass sizes itself from the true queue lengths */
        /* ---- trace(bounce): closest hits of bounce + 1, then the shadow rays of this bounce */
        for (int j0 = threadIdx.x & ~31; j0 < n_next; j0 += blockDim.x) {
            const bool act = j0 + lane < n_next;
            const int i = base + j0 + lane;
            float4 ro = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rd = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            if (act) { ro = b.ray_o[(bounce + 1) & 1][i]; rd = b.ray_d[(bounce + 1) & 1][i]; }
            float t;
            int h = traverse<false, LAY>(nodes, sc.leaf_tri, n_nodes, act, v3(ro.x, ro.y, ro.z), v3(rd.x, rd.y, rd.z), FLT_MAX, t);
            if (act) b.hit[i] = h;
        }
        for (int j0 = threadIdx.x & ~31; j0 < n_cur; j0 += blockDim.x) {
            const bool act = j0 + lane < n_cur;
            const int slot = base + j0 + lane;
            float4 ro = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(4)), rc = ro;
            if (act) { ro = b.sh_o[slot]; rc = b.sh_c[slot]; }
            const int flags = __float_as_int(ro.w);
            float L = 0.0f, B = 0.0f, t;
            V3 o = v3(ro.x, ro.y, ro.z);
            const bool need1 = act && !(flags & 4) && (flags & 1), need2 = act && !(flags & 4) && (flags & 2);
            if (__any_sync(0xffffffffu, need1)) {
                float4 d1 = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                if (need1) d1 = b.sh_d1[slot];
                int h = traverse<true, LAY>(nodes, sc.leaf_tri, n_nodes, need1, o, v3(d1.x, d1.y, d1.z), d1.w, t);
                if (need1 && h < 0) L = rc.x;
            }
            if (__any_sync(0xffffffffu, need2)) {
                float4 d2 = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
                if (need2) d2 = b.sh_d2[slot];
                int h = traverse<true, LAY>(nodes, sc.leaf_tri, n_nodes, need2, o, v3(d2.x, d2.y, d2.z), d2.w, t);
                if (need2 && h < 0) B = rc.y;
            }
            if (act) connect_finish(fp, b, bounce, slot, flags, L, B, rc.z, rc.w, rc.z);
        }
        __syncthreads();
        n_cur = n_next;
    }
}

/* ------------------------------------------------------------------ resolve + accumulate */
LYS_D V3 hue_to_rgb(float h) {                                       /* integrator.fut:139-148 */
    float hp = h * 6.0f;
    float x = 1.0f - lys_fabsf(fmodf(hp, 2.0f) - 1.0f);
    switch (trunc_u32(hp)) {
        case 0: return v3(1.0f, x, 0.0f); case 1: return v3(x, 1.0f, 0.0f); case 2: return v3(0.0f, 1.0f, x);
        case 3: return v3(0.0f, x, 1.0f); case 4: return v3(x, 0.0f, 1.0f); default: return v3(1.0f, 0.0f, x);
    }
}
LYS_D V3 resolve_pixel(const FrameParams &fp, const PassBuffers &b, int pid) {   /* visualize integrator.fut:150-168 */
    if (fp.render_mode == 1) {
        float bd = b.acc[pid].z;
        if (lys_isinff(bd)) return v3(0.0f, 0.0f, 0.0f);
        return hue_to_rgb(0.85f * (bd - 0.5f) / (10.0f - 0.5f));
    }
    V3 vis = fp.sensor_vis[b.chan[pid]];
    const float4 a = b.acc[pid];
    float s = a.x, z = a.y;
    float k = (float)fp.n_sensor;
    return v3(k * (vis.x == 1.0f ? s : z), k * (vis.y == 1.0f ? s : z), k * (vis.z == 1.0f ? s : z));
}
__global__ void __launch_bounds__(256) k_accumulate(const __grid_constant__ FrameParams fp, PassBuffers b, const float *__restrict__ img_old,
                                                    float *__restrict__ img_new, int merge, float n_frames, float out_scale) {
    int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= fp.n_local) return;
    int col, row; int ix = local_to_pixel(fp, pid, col, row);
    V3 c = resolve_pixel(fp, b, pid);
    if (merge) {                                                      /* sample_frame_accum :180-192 */
        V3 acc = v3(img_old[3ll * ix], img_old[3ll * ix + 1], img_old[3ll * ix + 2]);
        if (fp.render_mode == 1) c = (norm(acc) > 0.0f) ? acc : c;
        else c = ((n_frames - 1.0f) / n_frames) * acc + (1.0f / n_frames) * c;
    }
    if (out_scale != 1.0f) c = out_scale * c;                         /* pass-split frames: this rank's weight, folded into its last accumulate */
    img_new[3ll * ix] = c.x; img_new[3ll * ix + 1] = c.y; img_new[3ll * ix + 2] = c.z;
}

/* ------------------------------------------------------------------ point cloud (lib.fut:35-63) */
__global__ void __launch_bounds__(256) k_points_merge(const __grid_constant__ FrameParams fp, PassBuffers b, float4 *__restrict__ pos_int,
                                                      float *__restrict__ pdist, int first) {
    int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= fp.n_local) return;
    int col, row; int ix = local_to_pixel(fp, pid, col, row);
    const float4 acc = b.acc[pid];
    float bd = acc.z;
    float4 p2 = make_float4(-1.0f, -1.0f, -1.0f, 0.0f); float d2 = LYS_INF;     /* lib.fut:47 */
    if (!lys_isinff(bd)) {
        uint32_t r0 = fp.frame_rng ^ rng_split_hash((uint32_t)ix);
        V3 o, d; float w; int c;
        camera_sample(fp, col, row, r0, o, d, w, c);
        V3 pp = o + bd * d;                                                      /* to_cloud_points integrator.fut:122-126 */
        p2 = make_float4(pp.x, pp.y, pp.z, acc.w); d2 = bd;
    }
    if (first || !(pdist[ix] < d2)) { pos_int[ix] = p2; pdist[ix] = d2; }        /* merge lib.fut:48-51 */
}
__global__ void k_points_export(int n, const float4 *__restrict__ pos_int, float4 *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pos_int[i];
}

/* ------------------------------------------------------------------ render (lib.fut:187-196, matte argb.from_rgba) */
LYS_D uint32_t chan8(float x) { return lys_pin_argb_channel(x); }          /* matte argb.from_rgba (lys_pins.h) */
__global__ void __launch_bounds__(256) k_render(const float *__restrict__ img, int img_h, int img_w, int full_h, int full_w, int sub,
                                                int32_t *__restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= full_h * full_w) return;
    int i = k / full_w, j = k - i * full_w;
    int ii = i / sub, jj = j / sub;
    float r = 0.0f, g = 0.0f, bl = 0.0f;
    if (ii < img_h && jj < img_w) { const float *p = img + 3ll * ((long long)ii * img_w + jj); r = p[0]; g = p[1]; bl = p[2]; }
    out[k] = (int32_t)((chan8(1.0f) << 24) | (chan8(r) << 16) | (chan8(g) << 8) | chan8(bl));
}

/* ------------------------------------------------------------------ probes / tools */
template <int LAY>
__global__ void k_primary_probe(SceneDev sc, const __grid_constant__ FrameParams fp, PassBuffers b, int n, int *leaf, int *src, float *t) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = i < n;                                          /* no early return: traverse<> votes per warp */
    float4 ro = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rd = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    if (act) { ro = b.ray_o[0][i]; rd = b.ray_d[0][i]; }
    float th;
    int l = traverse<false, LAY>(sc.nodes, sc.leaf_tri, (int)sc.n_tris - 1, act, v3(ro.x, ro.y, ro.z), v3(rd.x, rd.y, rd.z), FLT_MAX, th);
    if (!act) return;
    const size_t ix = probe_index(fp, i);
    leaf[ix] = l;
    if (src) src[ix] = (l < 0) ? -1 : (int)__float_as_uint(sc.leaf_tri[4ll * l + 2].w);
    if (t) t[ix] = (l < 0) ? LYS_INF : th;
}
template <int LAY>
__global__ void k_trace_rays(SceneDev sc, const float *__restrict__ rays, const float *__restrict__ tmax, long long n,
                             int *out_leaf, float *out_t, int any) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = i < n;                                          /* no early return: traverse<> votes per warp */
    V3 o = v3(0.0f, 0.0f, 0.0f), d = v3(1.0f, 1.0f, 1.0f);
    if (act) { const float *r = rays + 6 * i; o = v3(r[0], r[1], r[2]); d = v3(r[3], r[4], r[5]); }
    float th; int l;
    if (any) l = traverse<true, LAY>(sc.nodes, sc.leaf_tri, (int)sc.n_tris - 1, act, o, d, act ? tmax[i] : 0.0f, th);
    else l = traverse<false, LAY>(sc.nodes, sc.leaf_tri, (int)sc.n_tris - 1, act, o, d, FLT_MAX, th);
    if (!act) return;
    out_leaf[i] = any ? (l >= 0 ? 1 : 0) : l;
    if (out_t) out_t[i] = (l < 0) ? LYS_INF : th;
}
__global__ void k_eval_math(int fn, const float *__restrict__ in, float *__restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = in[i], y;
    switch (fn) {
        case 0: y = det_sinf(x); break; case 1: y = det_cosf(x); break; case 2: y = det_expf(x); break;
        case 3: y = det_logf(x); break; case 4: y = det_pow5f(x); break; case 5: y = det_acosf(x); break;
        case 6: y = det_probitf(x); break; default: y = 0.0f;
    }
    out[i] = y;
}
__global__ void k_material_probe(const float *mat28, float wavelen, V3 wo, V3 wi, V3 n, uint32_t rng, float *out) {
    Onb onb = make_onb(n);
    Mat1 m = material_at(mat28, wavelen);
    float f, pdf;
    uber_eval(to_local(onb, wo), to_local(onb, wi), m, f, pdf);
    out[0] = f; out[1] = pdf;
    DirSample s = sample_bsdf(wo, onb, m, rng);
    out[2] = s.wi.x; out[3] = s.wi.y; out[4] = s.wi.z; out[5] = s.bsdf; out[6] = (float)s.kind; out[7] = s.pdf;
    out[8] = __uint_as_float(rng);
}

/* ------------------------------------------------------------------ host launchers */
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

/* persistent grids: SM count x resident CTAs per SM of each kernel (queried once per device).  The environment knobs
 * force what is otherwise chosen by scene size or by the previous pass (tests/test_gpu_parity.py::test_kernel_variants_bit_exact
 * runs each setting against the oracle): LYS_OCT_ONE_COPY (abi.cu), LYS_REFILL_MIN, LYS_TAIL_MAX, LYS_ADAPTIVE_GRIDS, LYS_SHADE_ORDER,
 * LYS_FUSE_GENERATE; LYS_PROFILE_TAIL keeps the fused tail under per-class timing. */
struct GridSizes { int trace[2] = {0, 0}, shade = 0, refill_min = LYS_REFILL_MIN_TRIS, adaptive = 1, sms = 148, tail_max = 8192, tail_items = 128, profile_tail = 0, order = 1, fuse_gen = 1; };
static GridSizes grid_sizes() {
    static GridSizes g[64];
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!g[dev].shade) {
        int sms = 148, bt[2] = {12, 16}, bs = 3;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bs, k_shade<LYS_SHADE_T>, LYS_SHADE_T, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bt[LAY_OCT], k_trace<LAY_OCT>, 128, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bt[LAY_SEL], k_trace<LAY_SEL>, 128, 0);
        for (int k = 0; k < 2; k++) g[dev].trace[k] = sms * (bt[k] > 0 ? bt[k] : 1);
        g[dev].shade = sms * (bs > 0 ? bs : 1);
        g[dev].sms = sms;
        const char *rfm = getenv("LYS_REFILL_MIN"); if (rfm) g[dev].refill_min = atoi(rfm);        /* tests: lane refill on small scenes too */
        const char *ord = getenv("LYS_SHADE_ORDER"); if (ord) g[dev].order = atoi(ord) ? 1 : 0;      /* 0: k_shade walks the slots in queue order */
        const char *fg = getenv("LYS_FUSE_GENERATE"); if (fg) g[dev].fuse_gen = atoi(fg) ? 1 : 0;    /* 0: k_generate and k_trace(-1) as two launches */
        const char *ptl = getenv("LYS_PROFILE_TAIL"); if (ptl) g[dev].profile_tail = atoi(ptl) ? 1 : 0;
        const char *tmx = getenv("LYS_TAIL_MAX"); if (tmx) g[dev].tail_max = atoi(tmx);          /* 0: no fused tail */
        const char *ad = getenv("LYS_ADAPTIVE_GRIDS"); if (ad) g[dev].adaptive = atoi(ad) ? 1 : 0;
    }
    return g[dev];
}
static int trace_layout(const SceneDev &sc) { return sc.oct_copies == 8 ? LAY_OCT : LAY_SEL; }
/* the two traversal variants a scene can select */
#define LYS_TRAV_DISPATCH(lay, CALL) do { if ((lay) == LAY_OCT) { CALL(LAY_OCT); } else { CALL(LAY_SEL); } } while (0)
static void launch_trace(const GridSizes &gs, int grid, const SceneDev &sc, const FrameParams &fp, PassBuffers &bufs, int bounce, cudaStream_t stream) {
    const int ordered = (gs.order && bounce >= 0) ? 1 : 0;      /* write the hits-first order of bounce + 1, walk the one of bounce */
    const int lay = trace_layout(sc);
    if (lay == LAY_SEL) k_trace<LAY_SEL><<<grid, 128, 0, stream>>>(sc, fp, bufs, bounce, ordered);
    else if (sc.n_tris >= gs.refill_min) k_trace<LAY_OCT_RF><<<grid, 128, 0, stream>>>(sc, fp, bufs, bounce, ordered);
    else k_trace<LAY_OCT><<<grid, 128, 0, stream>>>(sc, fp, bufs, bounce, ordered);
}
cudaError_t run_sample_pass(const SceneDev &sc, const FrameParams &fp, PassBuffers &bufs, cudaStream_t stream, uint64_t *launches, LaunchTimer *timer,
                            int *est_counts) {
    const int n = fp.n_local;
    if (n <= 0) return cudaSuccess;
    uint64_t nl = 0;
    LaunchTimer none; LaunchTimer &tm = timer ? *timer : none;
    const GridSizes gs = grid_sizes();
    const int g_trace = min(gs.trace[trace_layout(sc)], cdiv(2ll * n, 128)), g_shade = min(gs.shade, cdiv(n, LYS_SHADE_T));
    /* queue-length estimates: a snapshot of what an earlier pass left (the copy below may be updating it: harmless, the
     * numbers only size grids); valid if it is about the same sample grid */
    int est[LYS_MAX_PATH_LEN + 2];
    bool have_est = false;
    if (est_counts && gs.adaptive) {
        for (int k = 0; k <= LYS_MAX_PATH_LEN; k++) est[k] = est_counts[k];
        est[LYS_MAX_PATH_LEN + 1] = 0;
        have_est = est[0] == n;
        for (int k = 1; k <= LYS_MAX_PATH_LEN && have_est; k++) if (est[k] < 0 || est[k] > n) have_est = false;
    }
    const int g_min = max(1, gs.sms);                    /* never below one CTA per SM: a stale estimate costs at most ~10x on one pass */
    auto sized = [&](long long items, int threads, int g_full) {
        if (!have_est) return g_full;
        return (int)max((long long)g_min, min((long long)g_full, (2 * items + 4096 + threads - 1) / threads));
    };
    /* first bounce whose queue was short enough in the earlier pass: from there on one k_tail launch */
    int b_tail = fp.path_len;
    if (have_est && gs.tail_max > 0 && !(tm.on == 1 && !gs.profile_tail))      /* per-class timing (mode 1) wants every ray in the trace class */
        for (int k = 1; k < fp.path_len; k++) if (est[k] <= gs.tail_max) { b_tail = k; break; }
    if (gs.fuse_gen && tm.on != 1) {                     /* per-class timing (mode 1) keeps the two launches apart */
        tm.cur_bounce = -1;
        tm.begin(1, stream);
        const int g = cdiv(n, 128);
#define LYS_CALL(LAY) k_generate_trace<LAY><<<g, 128, 0, stream>>>(sc, fp, bufs)
        LYS_TRAV_DISPATCH(trace_layout(sc), LYS_CALL);
#undef LYS_CALL
        tm.end(stream); nl++;
    } else {
        tm.begin(0, stream); k_generate<<<cdiv(n, 256), 256, 0, stream>>>(fp, bufs); tm.end(stream); nl++;
        tm.cur_bounce = -1;
        tm.begin(1, stream);
        launch_trace(gs, g_trace, sc, fp, bufs, -1, stream);
        tm.end(stream); nl++;
    }
    for (int bnc = 0; bnc < b_tail; bnc++) {
        tm.cur_bounce = bnc;
        tm.begin(2, stream);
        k_shade<LYS_SHADE_T><<<sized(have_est ? est[bnc] : 0, LYS_SHADE_T, g_shade), LYS_SHADE_T, 0, stream>>>(sc, fp, bufs, bnc, (gs.order && bnc > 0) ? 1 : 0);   /* follow b.order (written by k_trace) */
        tm.end(stream); nl++;
        tm.begin(1, stream);
        launch_trace(gs, sized(have_est ? (long long)est[bnc] + est[bnc + 1] : 0, 128, g_trace), sc, fp, bufs, bnc, stream);
        tm.end(stream); nl++;
    }
    if (b_tail < fp.path_len) {
        const int g = (int)max((long long)gs.sms, min((long long)gs.sms * 4, (long long)((est[b_tail] + est[b_tail] / 4 + gs.tail_items - 1) / gs.tail_items)));   /* few, well filled CTAs (they stay resident for all the remaining bounces), never fewer than one per SM: a stale estimate must not serialise a long queue */
        tm.cur_bounce = b_tail;
        tm.begin(3, stream);
        const int lay = trace_layout(sc);
        if (lay == LAY_OCT) k_tail<LAY_OCT><<<g, 128, 0, stream>>>(sc, fp, bufs, b_tail);
        else k_tail<LAY_SEL><<<g, 128, 0, stream>>>(sc, fp, bufs, b_tail);
        tm.end(stream); nl++;
    }
    if (est_counts && gs.adaptive) {
        cudaError_t e = cudaMemcpyAsync(est_counts, bufs.counts, sizeof(int) * (LYS_MAX_PATH_LEN + 1), cudaMemcpyDeviceToHost, stream);
        if (e != cudaSuccess) return e;
    }
    if (launches) *launches += nl;
    return cudaGetLastError();
}
cudaError_t run_accumulate(const FrameParams &fp, const PassBuffers &bufs, const float *img_old, float *img_new, int merge,
                           float n_frames, cudaStream_t stream, uint64_t *launches, LaunchTimer *timer, float out_scale) {
    if (fp.n_local <= 0) return cudaSuccess;
    if (timer) timer->begin(4, stream);
    k_accumulate<<<cdiv(fp.n_local, 256), 256, 0, stream>>>(fp, bufs, img_old, img_new, merge, n_frames, out_scale);
    if (timer) timer->end(stream);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t run_points_merge(const FrameParams &fp, const PassBuffers &bufs, float4 *pos_int, float *dist, int first,
                             cudaStream_t stream, uint64_t *launches) {
    if (fp.n_local <= 0) return cudaSuccess;
    k_points_merge<<<cdiv(fp.n_local, 256), 256, 0, stream>>>(fp, bufs, pos_int, dist, first);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t run_points_export(const FrameParams &fp, const float4 *pos_int, float *out, cudaStream_t stream, uint64_t *launches) {
    int n = fp.gw * fp.gh;
    k_points_export<<<cdiv(n, 256), 256, 0, stream>>>(n, pos_int, reinterpret_cast<float4 *>(out));
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t run_render(const float *img, int img_h, int img_w, int full_h, int full_w, int subsampling, int32_t *out,
                       cudaStream_t stream, uint64_t *launches) {
    long long n = (long long)full_h * full_w;
    if (n <= 0) return cudaSuccess;
    k_render<<<cdiv(n, 256), 256, 0, stream>>>(img, img_h, img_w, full_h, full_w, subsampling, out);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t run_primary_probe(const SceneDev &sc, const FrameParams &fp, PassBuffers &bufs, int *leaf, int *src, float *t,
                              cudaStream_t stream, uint64_t *launches) {
    const int n = fp.n_local;
    if (n <= 0) return cudaSuccess;
    k_generate<<<cdiv(n, 256), 256, 0, stream>>>(fp, bufs);
    if (trace_layout(sc) == LAY_OCT) k_primary_probe<LAY_OCT><<<cdiv(n, 128), 128, 0, stream>>>(sc, fp, bufs, n, leaf, src, t);
    else k_primary_probe<LAY_SEL><<<cdiv(n, 128), 128, 0, stream>>>(sc, fp, bufs, n, leaf, src, t);
    if (launches) *launches += 2;
    return cudaGetLastError();
}
cudaError_t run_trace_rays(const SceneDev &sc, const float *rays, const float *tmax, int64_t n, int *out_leaf, float *out_t,
                           int any_hit, cudaStream_t stream, uint64_t *launches) {
    if (n <= 0) return cudaSuccess;
    if (trace_layout(sc) == LAY_OCT) k_trace_rays<LAY_OCT><<<cdiv(n, 128), 128, 0, stream>>>(sc, rays, tmax, n, out_leaf, out_t, any_hit);
    else k_trace_rays<LAY_SEL><<<cdiv(n, 128), 128, 0, stream>>>(sc, rays, tmax, n, out_leaf, out_t, any_hit);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t run_eval_math(int fn, const float *in, float *out, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    k_eval_math<<<cdiv(n, 256), 256, 0, stream>>>(fn, in, out, n);
    return cudaGetLastError();
}
cudaError_t run_material_probe(const float *mat28_dev, float wavelen, V3 wo, V3 wi, V3 n, uint32_t rng, float *out9_dev, cudaStream_t stream) {
    k_material_probe<<<1, 1, 0, stream>>>(mat28_dev, wavelen, wo, wi, n, rng, out9_dev);
    return cudaGetLastError();
}

} // namespace lys
