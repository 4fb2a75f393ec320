/* tracer.h -- the drop-in boundary.
 *
 * This is the header the reference's build generates with
 *     futhark <backend> -o build/tracer --library src/lib.fut        (reference Makefile:61-63)
 * and that its hosts consume: demo-interactive/liblys.h:7 (`#include "tracer.h"`) and
 * demo-save/src/ffi.rs:1-75 (`#[link(name = "tracer", kind = "static")]`).  The generated
 * file is not shipped with the reference (.gitignore:7), so the surface below is
 * reconstructed from Futhark's `--library` conventions and pinned by the hosts' call sites,
 * cited per declaration.  The implementation behind it is hand-written sm_100a CUDA
 * (msc-futhark-ray-tracer_b200/csrc), built as libtracer.so / libtracer.a.
 *
 * Conventions (Futhark C API): entry points, `values` and `free` return 0 on success and
 * non-zero on failure, with the message available from futhark_context_get_error() (caller
 * frees).  Results are returned through leading out-pointers.  Inputs are never consumed;
 * every returned object is owned by the caller and released with the matching free function.
 * A context is used from one host thread at a time.
 */
#ifndef LYS_TRACER_H
#define LYS_TRACER_H

#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- context ------------------------------------------------------------------------- */
struct futhark_context_config;
struct futhark_context;

struct futhark_context_config *futhark_context_config_new(void);                 /* liblys.c:170, ffi.rs:11 */
void futhark_context_config_free(struct futhark_context_config *cfg);            /* liblys.c:307 */
/* "#k" or "k" selects CUDA device k; any other string selects the first device whose name
 * contains it (the cuda backend's -d syntax).                                     liblys.c:175 */
void futhark_context_config_set_device(struct futhark_context_config *cfg, const char *s);
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag);
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag);

struct futhark_context *futhark_context_new(struct futhark_context_config *cfg); /* liblys.c:189, ffi.rs:13 */
void futhark_context_free(struct futhark_context *ctx);                          /* liblys.c:306 */
int futhark_context_sync(struct futhark_context *ctx);
int futhark_context_clear_caches(struct futhark_context *ctx);
char *futhark_context_get_error(struct futhark_context *ctx);                    /* liblys.h:37 */

/* ---- arrays ---------------------------------------------------------------------------
 * futhark_new_* copies `data` from the host to device memory; futhark_values_* is a blocking
 * device-to-host copy; futhark_shape_* returns a pointer owned by the array. */
struct futhark_f32_1d;
struct futhark_f32_1d *futhark_new_f32_1d(struct futhark_context *ctx, const float *data, int64_t dim0); /* liblys.c:138, ffi.rs:21 */
int futhark_free_f32_1d(struct futhark_context *ctx, struct futhark_f32_1d *arr);
int futhark_values_f32_1d(struct futhark_context *ctx, struct futhark_f32_1d *arr, float *data);
const int64_t *futhark_shape_f32_1d(struct futhark_context *ctx, struct futhark_f32_1d *arr);

struct futhark_f32_2d;
struct futhark_f32_2d *futhark_new_f32_2d(struct futhark_context *ctx, const float *data, int64_t dim0, int64_t dim1); /* liblys.c:302, ffi.rs:27 */
int futhark_free_f32_2d(struct futhark_context *ctx, struct futhark_f32_2d *arr);
int futhark_values_f32_2d(struct futhark_context *ctx, struct futhark_f32_2d *arr, float *data);
const int64_t *futhark_shape_f32_2d(struct futhark_context *ctx, struct futhark_f32_2d *arr);

struct futhark_f32_3d;
struct futhark_f32_3d *futhark_new_f32_3d(struct futhark_context *ctx, const float *data, int64_t dim0, int64_t dim1, int64_t dim2); /* liblys.c:298, ffi.rs:34 */
int futhark_free_f32_3d(struct futhark_context *ctx, struct futhark_f32_3d *arr);
int futhark_values_f32_3d(struct futhark_context *ctx, struct futhark_f32_3d *arr, float *data);      /* ffi.rs:42, wrapper.rs:89 */
const int64_t *futhark_shape_f32_3d(struct futhark_context *ctx, struct futhark_f32_3d *arr);

struct futhark_u32_1d;
struct futhark_u32_1d *futhark_new_u32_1d(struct futhark_context *ctx, const uint32_t *data, int64_t dim0); /* liblys.c:300, ffi.rs:15 */
int futhark_free_u32_1d(struct futhark_context *ctx, struct futhark_u32_1d *arr);
int futhark_values_u32_1d(struct futhark_context *ctx, struct futhark_u32_1d *arr, uint32_t *data);
const int64_t *futhark_shape_u32_1d(struct futhark_context *ctx, struct futhark_u32_1d *arr);

struct futhark_i32_2d;
struct futhark_i32_2d *futhark_new_i32_2d(struct futhark_context *ctx, const int32_t *data, int64_t dim0, int64_t dim1);
int futhark_free_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);                      /* liblys.c:115 */
int futhark_values_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr, int32_t *data);    /* liblys.c:114 */
const int64_t *futhark_shape_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);

/* ---- opaque state (src/state.fut:8-19) --------------------------------------------------- */
struct futhark_opaque_state;
int futhark_free_opaque_state(struct futhark_context *ctx, struct futhark_opaque_state *obj);         /* liblys.c:40,96,110,157; ffi.rs:48 */

/* ---- entry points (src/lib.fut) ---------------------------------------------------------- */
/* lib.fut:76-106.  NOTE h before w.  Host call sites: liblys.c:139-144, wrapper.rs:52-65. */
int futhark_entry_init(struct futhark_context *ctx, struct futhark_opaque_state **out0,
                       const int32_t seed, const uint32_t h, const uint32_t w, const uint32_t cam_conf_id,
                       const struct futhark_f32_3d *tri_geoms, const struct futhark_u32_1d *tri_mats,
                       const struct futhark_f32_2d *mat_data,
                       const float cam_pitch, const float cam_yaw, const struct futhark_f32_1d *cam_origin);
/* lib.fut:108-109, liblys.c:39 */
int futhark_entry_resize(struct futhark_context *ctx, struct futhark_opaque_state **out0,
                         const uint32_t h, const uint32_t w, const struct futhark_opaque_state *s);
/* lib.fut:111-118, liblys.c:109 */
int futhark_entry_step(struct futhark_context *ctx, struct futhark_opaque_state **out0,
                       const struct futhark_opaque_state *s);
/* lib.fut:120-185, liblys.c:94-95.  e == 0 means key-down. */
int futhark_entry_key(struct futhark_context *ctx, struct futhark_opaque_state **out0,
                      const int32_t e, const int32_t key, const struct futhark_opaque_state *s);
/* lib.fut:187-196, liblys.c:113.  [h][w] packed ARGB. */
int futhark_entry_render(struct futhark_context *ctx, struct futhark_i32_2d **out0,
                         const struct futhark_opaque_state *s);
/* lib.fut:67-74 (host use commented out at demo-save/src/main.rs:37-41).  [h][w][3] f32. */
int futhark_entry_sample_n_frames(struct futhark_context *ctx, struct futhark_f32_3d **out0,
                                  const struct futhark_opaque_state *s, const uint32_t n);
/* lib.fut:35-63, wrapper.rs:79-85.  out1 = [h][w][4] f32 (x, y, z, intensity). */
int futhark_entry_sample_points_n(struct futhark_context *ctx, struct futhark_opaque_state **out0,
                                  struct futhark_f32_3d **out1,
                                  const struct futhark_opaque_state *s, const uint32_t samples_per_pixel);

/* ---- what else a `futhark cuda --library` header carries and a host can use ------------------
 * Not called by the reference's hosts (liblys.c, ffi.rs): the profiling report and zero-copy device access, with the
 * generated header's signatures.  The run-time-compilation and tuning setters of a generated header (nvrtc options, program /
 * PTX dumps, default group / tile sizes, named sizes) have no counterpart here -- nothing is compiled at run time and no grid
 * size is user-tunable -- and are not declared. */
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag);    /* kernel-class timing from the start */
void futhark_context_pause_profiling(struct futhark_context *ctx);
void futhark_context_unpause_profiling(struct futhark_context *ctx);
/* malloc'ed text (caller frees): kernel launches, device time per kernel class when profiling is on, pooled device memory */
char *futhark_context_report(struct futhark_context *ctx);

/* Zero-copy device access.  The generated CUDA header spells the pointer type CUdeviceptr (<cuda.h>); this typedef is
 * ABI-identical, so tracer.h does not need the CUDA headers.  futhark_new_raw_* COPIES dim0*... elements from device memory
 * at `data + offset` (offset in bytes) into a new array on the context's stream; futhark_values_raw_* returns the array's own
 * device memory, valid until the array is freed (work queued by the library is complete after futhark_context_sync). */
typedef unsigned long long futhark_deviceptr;
struct futhark_f32_1d *futhark_new_raw_f32_1d(struct futhark_context *ctx, const futhark_deviceptr data, int offset, int64_t dim0);
struct futhark_f32_2d *futhark_new_raw_f32_2d(struct futhark_context *ctx, const futhark_deviceptr data, int offset, int64_t dim0, int64_t dim1);
struct futhark_f32_3d *futhark_new_raw_f32_3d(struct futhark_context *ctx, const futhark_deviceptr data, int offset, int64_t dim0, int64_t dim1, int64_t dim2);
struct futhark_u32_1d *futhark_new_raw_u32_1d(struct futhark_context *ctx, const futhark_deviceptr data, int offset, int64_t dim0);
struct futhark_i32_2d *futhark_new_raw_i32_2d(struct futhark_context *ctx, const futhark_deviceptr data, int offset, int64_t dim0, int64_t dim1);
futhark_deviceptr futhark_values_raw_f32_1d(struct futhark_context *ctx, struct futhark_f32_1d *arr);
futhark_deviceptr futhark_values_raw_f32_2d(struct futhark_context *ctx, struct futhark_f32_2d *arr);
futhark_deviceptr futhark_values_raw_f32_3d(struct futhark_context *ctx, struct futhark_f32_3d *arr);
futhark_deviceptr futhark_values_raw_u32_1d(struct futhark_context *ctx, struct futhark_u32_1d *arr);
futhark_deviceptr futhark_values_raw_i32_2d(struct futhark_context *ctx, struct futhark_i32_2d *arr);

#ifdef __cplusplus
}
#endif
#endif /* LYS_TRACER_H */
