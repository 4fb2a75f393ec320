/* lys_ext.h -- exports of libtracer that are NOT part of the Futhark-generated surface.
 *
 * They exist for three things the reference's API has no words for: multi-GPU pixel
 * partitioning, device-resident results for an NCCL reduce without a
 * host round trip, and introspection for parity tests / benchmarks (BVH arrays, per-pass
 * probes, build timing).  A host that only uses tracer.h never needs them.
 */
#ifndef LYS_EXT_H
#define LYS_EXT_H

#include "tracer.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- knobs (per context) ---------------------------------------------------------------- */
/* Maximum number of path vertices; the reference constant is 16 (integrator.fut:23); 1..16. */
int lys_context_set_path_len(struct futhark_context *ctx, int path_len);
/* 0 = reference-exact truncated Jacobi refit (bvh.fut:109-120), 1 = converged boxes,
 * 2 = reference-exact through the literal sweeps (slower; the overflow fallback, exposed for cross-checks). */
int lys_context_set_refit_mode(struct futhark_context *ctx, int mode);
/* Row-interleaved pixel partition for multi-GPU rendering: this context samples grid rows r with
 * r % world_size == rank; other pixels stay zero so that a sum-reduce over ranks is exact. */
int lys_context_set_partition(struct futhark_context *ctx, int rank, int world_size);
/* CUDA device ordinal and stream the context launches on (for event timing by the caller). */
int lys_context_device(struct futhark_context *ctx);
void *lys_context_stream(struct futhark_context *ctx);
/* number of kernel launches issued by this context so far */
uint64_t lys_context_launch_count(struct futhark_context *ctx);

/* Kernel-class device timing (CUDA events around every launch of the sample pass).  Off by default: the
 * extra event records perturb throughput, so it is used for a separate profiling step, never the timed run.
 * on = 1: one launch per stage and bounce, one pass at a time (exclusive times; every ray in the trace class);
 * on = 2: the production sequence (fused camera-ray launch, fused tail, passes pipelined over streams) with events around
 *         its launches: kernels of different passes overlap, so the class sums are SHARES of the step, not exclusive times.
 * Classes: 0 generate, 1 trace (closest hits of bounce b+1 + shadow rays of bounce b), 2 shade, 3 tail (k_tail: all bounces from the first sparse one on, in one launch), 4 accumulate. */
#define LYS_PROFILE_CLASSES 5
int lys_context_set_profiling(struct futhark_context *ctx, int on);
int lys_context_profile_get(struct futhark_context *ctx, float *ms /* [5] */, uint64_t *launches /* [5] */, int reset);
/* per-bounce device ms: ms36[0..17] = k_trace for bounce -1..16, ms36[18..35] = k_shade (index bounce + 1); call before a reset */
int lys_context_profile_detail(struct futhark_context *ctx, float *ms36);
/* A copy of `s` whose frame rng is advanced k steps (k sample passes ahead): pass-split multi-GPU rendering. */
int lys_state_advance_rng(struct futhark_context *ctx, struct futhark_opaque_state **out0, const struct futhark_opaque_state *s, uint32_t k);

/* ---- device-resident access ------------------------------------------------------------- */
/* Raw device pointers (valid until the array / state is freed). */
void *lys_device_ptr_f32_3d(struct futhark_context *ctx, struct futhark_f32_3d *arr);
void *lys_state_image_device_ptr(struct futhark_context *ctx, struct futhark_opaque_state *s,
                                 uint32_t *img_h, uint32_t *img_w);
/* Wrap device-resident data without a copy from the host (device-to-device copy). */
struct futhark_f32_3d *lys_new_f32_3d_from_device(struct futhark_context *ctx, const void *dev, int64_t d0, int64_t d1, int64_t d2);
struct futhark_u32_1d *lys_new_u32_1d_from_device(struct futhark_context *ctx, const void *dev, int64_t d0);

/* ---- state introspection ---------------------------------------------------------------- */
typedef struct {
    uint32_t dim_w, dim_h, subsampling, rng, img_h, img_w, n_frames, cam_conf_id;
    int32_t mode, render_mode;
    float cam_pitch, cam_yaw, cam_origin[3], aperture, focal_dist;
    float ambience[12];
    int64_t n_tris, n_mats, n_lights;
} lys_state_info;
int lys_state_info_get(struct futhark_context *ctx, const struct futhark_opaque_state *s, lys_state_info *out);
int lys_state_image(struct futhark_context *ctx, const struct futhark_opaque_state *s, float *out /* [img_h][img_w][3] */);

/* BVH arrays, same conventions as the oracle: child pointer internal i -> i, leaf i -> ~i;
 * node_aabb [n-1][6] and leaf_aabb [n][6] are (center xyz, half_dims xyz).  Any pointer may be NULL. */
int lys_state_bvh_get(struct futhark_context *ctx, const struct futhark_opaque_state *s,
                      float *bounds6, uint32_t *sorted_morton, int32_t *sorted_src_index,
                      int32_t *left, int32_t *right, int32_t *parent, float *node_aabb, float *leaf_aabb,
                      int32_t *node_height);
int lys_state_light_indices(struct futhark_context *ctx, const struct futhark_opaque_state *s, int32_t *src_index);

/* Rebuild the LBVH of the state's scene `reps` times from the device-resident triangle arrays and
 * report the mean device time in milliseconds (CUDA events around the build kernels only). */
int lys_state_bvh_rebuild_timed(struct futhark_context *ctx, const struct futhark_opaque_state *s, int reps, float *ms);

/* ---- per-pass probes on the state's NEXT pass (state is not modified) ------------------- */
/* primary rays: sorted-leaf index of the first hit (-1 = miss), source triangle index, t */
int lys_probe_primary(struct futhark_context *ctx, const struct futhark_opaque_state *s,
                      int32_t *leaf, int32_t *src_tri, float *t);
/* full pass: raw per-vertex radiance / cumulative distance [gh][gw][16], channel [gh][gw] */
int lys_probe_pass(struct futhark_context *ctx, const struct futhark_opaque_state *s,
                   float *radiance, float *distance, int32_t *channel);
/* arbitrary rays [n][6] (origin, unit direction) against the state's BVH */
int lys_trace_closest(struct futhark_context *ctx, const struct futhark_opaque_state *s,
                      const float *rays, int64_t n, int32_t *leaf, float *t);
int lys_trace_any(struct futhark_context *ctx, const struct futhark_opaque_state *s,
                  const float *rays, const float *tmax, int64_t n, int32_t *hit);
/* evaluate one lys_detmath.h function on the device: fn 0 sin, 1 cos, 2 exp, 3 log, 4 pow5, 5 acos, 6 probit */
int lys_eval_math(struct futhark_context *ctx, int fn, const float *in, float *out, int64_t n);
/* device evaluation of bsdf_f / bsdf_pdf / sample_dir for one interaction (out layout as orc_material_probe) */
int lys_material_probe(struct futhark_context *ctx, const float *mat28, float wavelen, const float *wo, const float *wi,
                       const float *normal, uint32_t rng, float *out9);

/* ---- batch rendering without leaving the device ---------------------------------------- */
/* Like futhark_entry_sample_n_frames but also reports per-call work counters measured on the device. */
typedef struct {
    uint64_t paths, vertices, closest_rays, shadow_rays, launches;
    float device_ms;
} lys_pass_stats;
int lys_sample_n_frames_stats(struct futhark_context *ctx, struct futhark_f32_3d **out0,
                              const struct futhark_opaque_state *s, uint32_t n, lys_pass_stats *stats);
/* The same with the image multiplied by `weight` inside the last accumulate kernel: the share of this rank's passes in a
 * pass-split multi-GPU frame, so that the per-rank images only need the one sum-reduce.  weight 1 = the
 * reference's result, bit for bit.  stats may be NULL. */
int lys_sample_n_frames_weighted(struct futhark_context *ctx, struct futhark_f32_3d **out0,
                                 const struct futhark_opaque_state *s, uint32_t n, float weight, lys_pass_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* LYS_EXT_H */
