/* lys_oracle.h -- C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's Futhark
 * program (bryal/msc-futhark-ray-tracer, src/ *.fut files) used as the parity checker and
 * as the timed CPU baseline.  Nothing under csrc/ or the product library links,
 * includes or calls it; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, its
 * Futhark toolchain and its five third-party Futhark packages are absent from this
 * environment, so this restatement cannot be checked against outputs of the
 * reference itself.  It is auditable line-by-line against the .fut sources cited
 * in lys_oracle.cpp, and is pinned only by self-authored known-answer vectors
 * (tests/golden/) and by a second, independently written reading of the .fut text
 * (tests/test_oracle_kat.py, tests/test_oracle_pathtrace_kat.py: build, walks, BSDF,
 * camera and a complete scalar path tracer, all bit for bit).
 */
#ifndef LYS_ORACLE_H
#define LYS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_state orc_state;

/* Knobs (process-global).  path_len: reference constant 16 (integrator.fut:23).
 * refit_mode: 0 = reference-exact truncated Jacobi refit (bvh.fut:109-120),
 *             1 = converged boxes.  math_mode: 0 = lys_detmath.h, 1 = glibc libm. */
void orc_set_path_len(int n);
void orc_set_refit_mode(int mode);
void orc_set_math_mode(int mode);
void orc_set_threads(int n);          /* OpenMP threads for sample passes; 0 = all */
int  orc_get_threads(void);

/* lib.fut entries (same argument meaning as the futhark_entry_* ABI). */
orc_state *orc_init(int32_t seed, uint32_t h, uint32_t w, uint32_t cam_conf_id,
                    const float *tri_geoms, const uint32_t *tri_mats, int64_t n_tris,
                    const float *mat_data, int64_t n_mats,
                    float cam_pitch, float cam_yaw, const float *cam_origin);
orc_state *orc_resize(uint32_t h, uint32_t w, const orc_state *s);
orc_state *orc_key(int32_t e, int32_t key, const orc_state *s);
orc_state *orc_step(const orc_state *s);
void orc_render(const orc_state *s, int32_t *out /* [h][w] */);
void orc_sample_n_frames(const orc_state *s, uint32_t n, float *out /* [gh][gw][3] */);
orc_state *orc_sample_points_n(const orc_state *s, uint32_t spp, float *out /* [gh][gw][4] */);
orc_state *orc_advance_rng(const orc_state *s, uint32_t k);   /* k x advance_rng on the frame rng (rand.fut:11-12) */
void orc_free_state(orc_state *s);

/* state introspection */
void orc_state_dims(const orc_state *s, uint32_t *w, uint32_t *h, uint32_t *grid_w, uint32_t *grid_h);
void orc_state_image(const orc_state *s, float *out /* [img_h][img_w][3] */, uint32_t *img_h, uint32_t *img_w);
void orc_state_scalars(const orc_state *s, uint32_t *rng, uint32_t *n_frames, uint32_t *subsampling,
                       int32_t *mode, int32_t *render_mode, uint32_t *cam_conf_id,
                       float *cam /* pitch,yaw,ox,oy,oz,aperture,focal_dist */, float *ambience12);

/* BVH introspection (bvh.fut:76-121, radix_tree.fut:21-89).  Child encoding:
 * internal i -> i, leaf i -> ~i (= -i-1).  node_aabb: [n-1][6] = center xyz, half xyz. */
int64_t orc_bvh_size(const orc_state *s);
int64_t orc_n_lights(const orc_state *s);
void orc_bvh_get(const orc_state *s, float *bounds6, uint32_t *sorted_morton, int32_t *sorted_src_index,
                 int32_t *left, int32_t *right, int32_t *parent, float *node_aabb, float *leaf_aabb);
void orc_light_indices(const orc_state *s, int32_t *src_index);

/* Stand-alone pieces for known-answer tests. */
uint32_t orc_expand_bits(uint32_t x);
uint32_t orc_morton3d(float x, float y, float z);
uint32_t orc_hash(int32_t x);
uint32_t orc_rng_from_seed(int32_t seed);
uint32_t orc_rng_next(uint32_t s);
float    orc_rng_uniform(uint32_t s, float lo, float hi, uint32_t *s_out);
void orc_radix_tree(const uint32_t *keys, int64_t n, int32_t *left, int32_t *right, int32_t *parent);
float orc_spectrum_lookup(float v, const float *spectrum12);
int32_t orc_hit_aabb(float tmax, const float *ray6, const float *box6 /* center, half_dims */);             /* shapes.fut:114-135 */
int32_t orc_hit_triangle(float tmax, const float *ray6, const float *tri9, float *t_pos_normal7 /* or NULL */); /* shapes.fut:66-86 */
void orc_normalise(const float *v3, float *out3);                                                          /* vector `normalise` (lys_pins.h) */
int32_t orc_argb_from_rgba(float r, float g, float b, float a);                                           /* matte argb.from_rgba (lys_pins.h) */
void orc_eval_math(int fn, const float *in, float *out, int64_t n);
void orc_material_probe(const float *mat28, float wavelen, const float *wo, const float *wi, const float *normal,
                        uint32_t rng, float *out /* bsdf_f, bsdf_pdf, sample wi xyz, sample bsdf, pdf kind, pdf, rng_out */);

/* Per-pass probes on the state's NEXT pass (uses s->rng, does not modify s).
 * first_hit: sorted-leaf index of the primary-ray hit (-1 = miss), source triangle
 * index, and t.  radiance: [gh][gw][path_len] raw per-vertex radiance, distance same shape,
 * channel [gh][gw]. */
void orc_probe_primary(const orc_state *s, int32_t *leaf, int32_t *src_tri, float *t,
                       float *ray /* [gh][gw][6] or NULL */, float *wavelen /* or NULL */);
void orc_probe_pass(const orc_state *s, float *radiance, float *distance, int32_t *channel);
/* brute-force closest hit over all triangles (mk_fake_bvh semantics, bvh.fut:31-39) for rays [n][6] */
void orc_brute_force_hits(const orc_state *s, const float *rays, int64_t n, int32_t *src_tri, float *t);
void orc_closest_hits(const orc_state *s, const float *rays, int64_t n, int32_t *leaf, float *t);
/* per-ray work of the closest-hit walk: box tests and triangle tests (divergence studies) */
void orc_closest_hits_steps(const orc_state *s, const float *rays, int64_t n, int32_t *box_tests, int32_t *tri_tests);
void orc_any_hits(const orc_state *s, const float *rays, const float *tmax, int64_t n, int32_t *hit);
/* divergence studies (tools/simt_model.py): closest-hit rays of every path of the NEXT pass, and per-ray visit patterns */
void orc_probe_path_rays(const orc_state *s, float *rays /* [gh][gw][16][6] */, int32_t *n_rays /* [gh][gw] */);
void orc_closest_hits_pattern(const orc_state *s, const float *rays, int64_t n, int32_t max_steps, uint8_t *pattern, int32_t *len);

/* Work counters accumulated since the last reset (define the algorithmic-bytes figure). */
typedef struct {
    uint64_t paths, vertices, closest_rays, shadow_rays, node_visits, box_tests, tri_tests, loop_iters;
    uint64_t closest_box, closest_tri, shadow_box, shadow_tri;   /* box / triangle tests split by ray type */
} orc_counters;
void orc_counters_reset(void);
void orc_counters_get(orc_counters *out);

#ifdef __cplusplus
}
#endif
#endif
