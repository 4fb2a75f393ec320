/* lys_wavefront.h -- per-pass parameters and work buffers of the wavefront path tracer. */
#pragma once
#include "lys_scene.h"
#include "lys_device.cuh"
#include <vector>

namespace lys {

/* Everything that is constant over one sample pass (one `sample_pixels` call, integrator.fut:103-116).
 * Camera vectors are derived on the host exactly as camera.fut:81-101 does per pixel. */
struct FrameParams {
    int gw, gh;                 /* sample grid (integrator.fut:175-176) */
    float fw, fh;
    int rank, world;            /* row-interleaved partition */
    int n_local;                /* pixels sampled by this rank */
    uint32_t frame_rng;         /* state.rng */
    V3 cam_origin, llc, horizontal, vertical, cam_u, cam_v;
    float lens_radius, offset_radius;
    int n_sensor;
    float sensor_mu[3], sensor_sigma[3];
    V3 sensor_vis[3];
    int tx_kind;                /* 0 none, 1 flash, 2 scanning (camera.fut:30-32) */
    float tx_radius, tx_theta;
    float tx_emission[12];
    float sector_x[9], sector_y[9];   /* rot_z (2*pi/8 * j) (1,0,0), host libm (shapes.fut:20-28) */
    int n_scene_lights;
    float ambience[12];
    int path_len;
    int render_mode;            /* 0 colour, 1 distance */
    float intensity_factor;     /* 1, or 1/spp for sample_points_n (lib.fut:39,42) */
};

/* Path id -> pixel.  Path ids are an implementation detail (everything a pixel's value depends on is keyed by the PIXEL index:
 * rng stream integrator.fut:109-114, framebuffer position), so they are laid out for the hardware: a warp's 32 consecutive
 * ids cover an 8 x 4 pixel tile instead of a 32 x 1 strip -- camera rays of a warp walk the same nodes, and the paths they
 * start stay close.  The local rows (this rank's rows of the interleaved partition) are cut into bands of 4; a band is
 * gw / 8 full tiles, then the gw % 8 leftover columns; an incomplete last band is walked row by row.  A bijection for any size. */
LYS_HDI void path_tile(int gw, int lrows, int pid, int &col, int &rl) {
    const int band_sz = 4 * gw;
    const int band = pid / band_sz, q = pid - band * band_sz;
    if (band * 4 + 4 > lrows) { const int r = q / gw; rl = band * 4 + r; col = q - r * gw; return; }
    const int full = gw >> 3;
    if (q < 32 * full) { const int t = q >> 5, k = q & 31; col = t * 8 + (k & 7); rl = band * 4 + (k >> 3); return; }
    const int rem = gw & 7, q2 = q - 32 * full, r = q2 / rem;
    rl = band * 4 + r; col = full * 8 + (q2 - r * rem);
}

/* Work buffers, sized for `cap` paths.  Path state is indexed by path id (= local pixel index),
 * queues and shadow records by queue slot. */
struct PassBuffers {
    int64_t cap = 0;
    /* rays live at QUEUE SLOTS (ping-pong per bounce): written coalesced at the compacted slot by k_shade, read
     * coalesced by k_trace / k_shade of the next bounce; everything else per path is indexed by path id */
    float4 *ray_o[2] = {nullptr, nullptr};    /* origin.xyz | wavelength */
    float4 *ray_d[2] = {nullptr, nullptr};    /* dir.xyz    | rng state (bits) */
    float *dist[2] = {nullptr, nullptr};      /* cumulative distance (integrator.fut:54), slot-indexed like the rays */
    float4 *acc = nullptr;      /* x: sum of vertex radiance (integrator.fut:164-168), y: sum of radiance * 0 (the other
                                   channels), z: distance of the nearest valid vertex (inf = none), w: its intensity */
    uint8_t *chan = nullptr;
    int *queue[2] = {nullptr, nullptr};
    int *hit = nullptr;         /* per slot: sorted-leaf index or -1 */
    int *order[2] = {nullptr, nullptr};   /* processing order (ping-pong per bounce): the slots of a bounce, hits first (from the front),
                                   misses last (from the back), written by k_trace with the hits; k_shade and the shadow-ray
                                   part of k_trace walk it, so their warps are all-hit or all-miss */
    int *split = nullptr;       /* [2][LYS_MAX_PATH_LEN + 1] hits / misses appended per bounce */
    float4 *sh_o = nullptr;     /* per slot: shadow origin.xyz | flags (bit0 ray1, bit1 ray2) */
    float4 *sh_d1 = nullptr;    /* dir1.xyz | tmax1 */
    float4 *sh_d2 = nullptr;    /* dir2.xyz | tmax2 */
    float4 *sh_c = nullptr;     /* cL, cB, emission (vertex 0), vertex distance */
    int *counts = nullptr;      /* [LYS_MAX_PATH_LEN + 1] active paths per bounce */
    unsigned long long *stats = nullptr;   /* [4] vertices, closest rays, shadow rays, paths */
    LightRec *tx_lights = nullptr;         /* [8] flash transmitter lights */
    /* probes (optional) */
    float *probe_rad = nullptr, *probe_dist = nullptr;   /* [cap][16] */
};

/* optional per-launch timing: events are recorded around each launch and resolved by the owner.
 * on = 1: per-class sequence (k_generate and k_trace(-1) as two launches, one launch per bounce, one pass at a time) so that
 *         every ray is in the trace class; on = 2: the PRODUCTION sequence (fused camera-ray launch, fused tail, passes
 *         pipelined over the slot streams) with events around its launches -- concurrent kernels share the GPU, so the
 *         per-class sums are used as SHARES of the step, not as exclusive times. */
struct LaunchTimer {
    int on = 0;
    std::vector<cudaEvent_t> ev0, ev1; std::vector<int> cls, bnc; size_t used = 0; int cur_bounce = 0;
    float ms[5] = {0, 0, 0, 0, 0}; uint64_t n[5] = {0, 0, 0, 0, 0};
    float detail[2][18] = {};          /* [0] trace, [1] shade; index = bounce + 1 */
    void begin(int c, cudaStream_t st) {
        if (!on) return;
        if (used == ev0.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ev0.push_back(a); ev1.push_back(b); cls.push_back(0); bnc.push_back(0); }
        cls[used] = c; bnc[used] = cur_bounce; cudaEventRecord(ev0[used], st);
    }
    void end(cudaStream_t st) { if (!on) return; cudaEventRecord(ev1[used], st); used++; }
    void resolve(cudaStream_t st) {
        if (!used) return;
        cudaStreamSynchronize(st);
        for (size_t i = 0; i < used; i++) { float t = 0; cudaEventElapsedTime(&t, ev0[i], ev1[i]); ms[cls[i]] += t; n[cls[i]]++;
            if ((cls[i] == 1 || cls[i] == 2) && bnc[i] >= -1 && bnc[i] < 17) detail[cls[i] - 1][bnc[i] + 1] += t; }
        used = 0;
    }
    void reset() { for (int i = 0; i < 5; i++) { ms[i] = 0; n[i] = 0; } for (int i = 0; i < 18; i++) { detail[0][i] = 0; detail[1][i] = 0; } used = 0; }
    ~LaunchTimer() { for (auto e : ev0) cudaEventDestroy(e); for (auto e : ev1) cudaEventDestroy(e); }
};

/* one sample pass: generate, (extend, shade, connect) x path_len.  Results stay in bufs.acc. */
/* est_counts (host, pinned, may be null): queue lengths per bounce of an earlier pass of the same frame; they only size the
 * persistent grids of the sparse late bounces (any grid size is correct: all kernels are grid-stride over device-side counts),
 * and the pass leaves its own counts there (async copy) for the next one */
cudaError_t run_sample_pass(const SceneDev &sc, const FrameParams &fp, PassBuffers &bufs, cudaStream_t stream, uint64_t *launches, LaunchTimer *timer = nullptr,
                            int *est_counts = nullptr);
/* resolve + merge into the image: mode 0 = replace (sample_frame), 1 = running average (sample_frame_accum) */
/* out_scale != 1: the merged pixel is multiplied by it (the weight of this rank's passes in a pass-split multi-GPU frame) */
cudaError_t run_accumulate(const FrameParams &fp, const PassBuffers &bufs, const float *img_old, float *img_new,
                           int merge, float n_frames, cudaStream_t stream, uint64_t *launches, LaunchTimer *timer = nullptr, float out_scale = 1.0f);
/* point cloud: resolve one pass into [gh][gw] (pos.xyz, distance, intensity) and merge (lib.fut:41-59) */
cudaError_t run_points_merge(const FrameParams &fp, const PassBuffers &bufs, float4 *pos_int, float *dist, int first,
                             cudaStream_t stream, uint64_t *launches);
cudaError_t run_points_export(const FrameParams &fp, const float4 *pos_int, float *out, cudaStream_t stream, uint64_t *launches);
/* render: upscale + pack ARGB (lib.fut:187-196) */
cudaError_t run_render(const float *img, int img_h, int img_w, int full_h, int full_w, int subsampling, int32_t *out,
                       cudaStream_t stream, uint64_t *launches);
/* probes / tools */
cudaError_t run_primary_probe(const SceneDev &sc, const FrameParams &fp, PassBuffers &bufs, int *leaf, int *src, float *t,
                              cudaStream_t stream, uint64_t *launches);
cudaError_t run_trace_rays(const SceneDev &sc, const float *rays, const float *tmax, int64_t n, int *out_leaf, float *out_t,
                           int any_hit, cudaStream_t stream, uint64_t *launches);
cudaError_t run_eval_math(int fn, const float *in, float *out, int64_t n, cudaStream_t stream);
cudaError_t run_material_probe(const float *mat28_dev, float wavelen, V3 wo, V3 wi, V3 n, uint32_t rng, float *out9_dev, cudaStream_t stream);

} // namespace lys
