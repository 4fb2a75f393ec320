/* abi.cu -- the futhark_* C ABI (include/tracer.h) and the lys_* extensions (include/lys_ext.h).
 *
 * Host-side restatement of the scalar parts of the reference program: src/lib.fut (entry points, camera
 * presets, key handling), src/state.fut (the opaque state record), the per-frame camera basis of
 * src/camera.fut:47-55,89-101 and the sky / flash spectra of src/spectrum.fut:64-91.  Everything that
 * touches pixels, rays or triangles is launched on the GPU (lbvh.cu, wavefront.cu); there is no CPU
 * fallback: without a CUDA device futhark_context_new() fails and every entry point returns an error.
 */
#include "../../include/tracer.h"
#include "../../include/lys_ext.h"
#include "lys_wavefront.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <deque>
#include <atomic>
#include <memory>
#include <string>
#include <vector>

using namespace lys;

/* ------------------------------------------------------------------ objects */
struct futhark_context_config { int device = 0; std::string device_name; int debugging = 0; int logging = 0; int profiling = 0; };   /* device = the #k of "#k name": k-th device whose name contains `device_name` */

struct futhark_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    int path_len = 16, refit_mode = 0, rank = 0, world = 1;
    uint64_t launches = 0;
    int logging = 0;
    std::multimap<size_t, void *> pool;          /* free device blocks by size */
    size_t pooled_bytes = 0, pool_cap = (size_t)1 << 30;   /* the pool never holds more than pool_cap bytes (set from the device size) */
    struct InitReadback { int lights[4]; float origin[3]; int crown_overflow; } *h_init = nullptr;   /* pinned: the one read-back of futhark_entry_init */
    PassBuffers bufs; BuildScratch scratch;
    float4 *pts_pos = nullptr; float *pts_dist = nullptr; int64_t pts_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr;
    LaunchTimer timer;
    /* pass pipelining: independent sample passes run on their own streams / buffer sets; only the accumulate
     * (or point-cloud merge) kernels are ordered, so one pass's latency-bound tail overlaps the next pass's head */
    struct PassSlot { cudaStream_t stream = nullptr; PassBuffers bufs; cudaEvent_t done = nullptr; };
    std::vector<PassSlot> slots;
    int pipeline = 12;                   /* 6 / 8 / 12 / 16 measured: 3305 / 3348 / 3381 / 3389 Mpaths/s on CornellBox (profiles/README.md 8.10) */
    int *h_counts = nullptr;             /* pinned: queue lengths of a recent pass (grid sizing only, see run_sample_pass) */
    uint64_t est_tag = 0;                /* what the estimates are about: scene, camera, path length */
    /* stepping loop of an interactive host (liblys.c:104-123: step, render, values, repeat): once the host steps the state
     * the previous step returned, the following steps are started AHEAD on the pass-slot streams (see futhark_entry_step),
     * so frame k's render + read-back overlap the passes of frames k+1 ..; states are immutable values, so a step computed
     * ahead is the step the host asks for next, bit for bit -- or it is dropped. */
    struct Ahead { uint64_t in_serial; struct futhark_opaque_state *out; };
    std::deque<Ahead> ahead;
    uint64_t last_step_out = 0;          /* serial of the state the latest futhark_entry_step returned */
    int ahead_depth = 3; uint64_t ahead_slot = 0;
};

namespace {

struct DevBlock {
    futhark_context *ctx; void *p; size_t bytes;
    DevBlock(futhark_context *c, void *q, size_t b) : ctx(c), p(q), bytes(b) {}
    ~DevBlock();
};
typedef std::shared_ptr<DevBlock> DevRef;

bool set_error(futhark_context *ctx, const std::string &msg) { ctx->error = msg; return false; }
bool cu_ok(futhark_context *ctx, cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    return set_error(ctx, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, call) do { if (!cu_ok(ctx, (call), #call)) return 1; } while (0)
#define CUB(ctx, call) do { if (!cu_ok(ctx, (call), #call)) return false; } while (0)

/* Block pool.  Every step / render / resize takes its image from here and gives it back when the state is freed, so the
 * steady state of a host loop costs no cudaMalloc / cudaFree.  Reuse is best fit within 12.5 % slack (a window dragged
 * through many sizes reuses near-size blocks instead of parking one block per size); the pool holds at most pool_cap bytes
 * (largest blocks are released first beyond that); when cudaMalloc fails the whole pool is released and the call retried
 * once, as Futhark's generated allocator does. */
void pool_release_all(futhark_context *ctx) {
    if (ctx->pool.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->pool) cudaFree(kv.second);
    ctx->pool.clear(); ctx->pooled_bytes = 0;
}
DevBlock::~DevBlock() {
    if (!p) return;
    ctx->pool.insert({bytes, p}); ctx->pooled_bytes += bytes;
    while (ctx->pooled_bytes > ctx->pool_cap && !ctx->pool.empty()) {
        auto last = std::prev(ctx->pool.end());
        ctx->pooled_bytes -= last->first;
        cudaFree(last->second);                 /* synchronises: nothing can still be using a block that is being dropped */
        ctx->pool.erase(last);
    }
}
DevRef dev_alloc(futhark_context *ctx, size_t bytes) {
    if (bytes == 0) bytes = 16;
    bytes = (bytes + 255) & ~(size_t)255;
    auto it = ctx->pool.lower_bound(bytes);
    if (it != ctx->pool.end() && it->first <= bytes + bytes / 8) {
        void *p = it->second; const size_t have = it->first;
        ctx->pool.erase(it); ctx->pooled_bytes -= have;
        return std::make_shared<DevBlock>(ctx, p, have);
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); pool_release_all(ctx); e = cudaMalloc(&p, bytes); }
    if (e != cudaSuccess) { cu_ok(ctx, e, "cudaMalloc"); return nullptr; }
    return std::make_shared<DevBlock>(ctx, p, bytes);
}
template <class T> bool raw_alloc(futhark_context *ctx, T *&p, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, sizeof(T) * (count ? count : 1));
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); pool_release_all(ctx); e = cudaMalloc(&q, sizeof(T) * (count ? count : 1)); }
    if (!cu_ok(ctx, e, "cudaMalloc")) return false;
    p = (T *)q; return true;
}
template <class T> void raw_free(T *&p) { if (p) cudaFree(p); p = nullptr; }

/* scene buffers come from (and go back to) the context's block pool: re-initialising a scene of the same size
 * (the e2e benchmark does it every step) costs no cudaMalloc / cudaFree (cudaFree synchronises the device) */
struct SceneHolder {
    SceneDev d;
    int refit_mode = 0;                  /* the mode the BVH was built with (a later rebuild must not change its semantics) */
    std::vector<DevRef> blocks;
    template <class T> bool take(futhark_context *ctx, T *&p, size_t count) {
        DevRef r = dev_alloc(ctx, sizeof(T) * (count ? count : 1));
        if (!r) return false;
        p = (T *)r->p; blocks.push_back(r); return true;
    }
};

/* ---- spectrum.fut (host) ---- */
struct Spectrum { float k[12]; };
float h_spectrum_lookup(float v, const Spectrum &s) { return spectrum_lookup12(v, s.k); }
Spectrum h_uniform_spectrum(float x) { Spectrum s; s.k[0] = 0.0f; s.k[1] = x; for (int i = 1; i < 6; i++) { s.k[2 * i] = -1.0f; s.k[2 * i + 1] = 0.0f; } return s; } /* :81-87 */
Spectrum h_blackbody(float T) {                                                   /* :64-72 */
    const float c = 299792458.0f, h = 6.62606957e-34f, kb = 1.3806488e-23f;
    const float nm[6] = {150.0f, 460.0f, 550.0f, 610.0f, 1000.0f, 2000.0f};
    Spectrum s;
    for (int i = 0; i < 6; i++) {
        float l = nm[i] * 1e-9f;
        float planck = (2 * h * c * c) / (powf(l, 5.0f) * (expf((h * c) / (l * kb * T)) - 1));
        s.k[2 * i] = l * 1e9f; s.k[2 * i + 1] = planck;
    }
    return s;
}
Spectrum h_blackbody_normalized(float T) {                                        /* :74-79 */
    Spectrum r = h_blackbody(T);
    float lambda_max = (2.8977721e-3f / T) * 1e9f;
    float mx = h_spectrum_lookup(lambda_max, r);
    for (int i = 0; i < 6; i++) r.k[2 * i + 1] = r.k[2 * i + 1] / mx;
    return r;
}
Spectrum h_scale(Spectrum s, float f) { for (int i = 0; i < 6; i++) s.k[2 * i + 1] = s.k[2 * i + 1] * f; return s; }

/* ---- camera presets (lib.fut:10-33) ---- */
struct CamConf {
    float aperture, focal_dist, offset_radius, fov;
    int n_sensor; float mu[3], sigma[3]; V3 vis[3];
    int tx_kind; float tx_radius, tx_theta; Spectrum tx_emission;
};
float h_from_deg(float d) { return d * LYS_PI / 180.0f; }                         /* linalg.fut:53 */
CamConf conf_visual() {
    CamConf c{}; c.aperture = 0; c.focal_dist = 1; c.offset_radius = 1; c.fov = h_from_deg(80);
    c.n_sensor = 3;
    c.mu[0] = 455; c.sigma[0] = 22; c.vis[0] = v3(0, 0, 1);
    c.mu[1] = 535; c.sigma[1] = 32; c.vis[1] = v3(0, 1, 0);
    c.mu[2] = 610; c.sigma[2] = 26; c.vis[2] = v3(1, 0, 0);
    c.tx_kind = 0; c.tx_radius = 0; c.tx_theta = 0; c.tx_emission = h_uniform_spectrum(0);
    return c;
}
CamConf conf_flash() { CamConf c = conf_visual(); c.tx_kind = 1; c.tx_radius = 0.05f; c.tx_emission = h_scale(h_blackbody_normalized(5500.0f), 1000.0f); return c; }
CamConf conf_lidar() {
    CamConf c{}; c.aperture = 0; c.focal_dist = 1; c.offset_radius = 0.01f; c.fov = h_from_deg(90);
    c.n_sensor = 1; c.mu[0] = 1550; c.sigma[0] = 10; c.vis[0] = v3(1, 0, 0);
    c.tx_kind = 2; c.tx_radius = 0.01f; c.tx_theta = h_from_deg(3); c.tx_emission = h_uniform_spectrum(1500);
    return c;
}
struct Camera { float pitch, yaw; V3 origin; CamConf conf; };
V3 h_cam_dir(const Camera &c) { return normalise(v3(sinf(c.yaw), sinf(c.pitch), -(cosf(c.yaw)))); }     /* camera.fut:47-49 */
V3 h_cam_right(const Camera &c) { return normalise(cross(h_cam_dir(c), v3(0, 1, 0))); }                /* :51-52 */
V3 h_cam_up(const Camera &c) { return normalise(cross(h_cam_right(c), h_cam_dir(c))); }                /* :54-55 */
Camera h_move(Camera cam, V3 m) {                                                                      /* :57-62 */
    V3 d = h_cam_dir(cam); d.y = 0;
    V3 fwd = normalise(d);
    cam.origin = ((cam.origin + (0.1f * m.z) * fwd) + (0.1f * m.x) * h_cam_right(cam)) + (0.1f * m.y) * v3(0, 1, 0);
    return cam;
}
Camera h_turn(Camera cam, float pitch, float yaw) {                                                    /* :64-66 */
    cam.pitch = lys_fmaxf(-0.5f * LYS_PI, lys_fminf(0.5f * LYS_PI, cam.pitch + pitch));
    cam.yaw = fmodf(cam.yaw + yaw, 2 * LYS_PI);
    return cam;
}

} // namespace

struct futhark_opaque_state {                          /* state.fut:8-19 */
    uint32_t dim_w, dim_h, subsampling;
    uint32_t rng;
    DevRef img; uint32_t img_h, img_w;
    uint32_t n_frames;
    Spectrum ambience;
    bool mode;
    int render_mode;
    uint32_t cam_conf_id;
    Camera cam;
    std::shared_ptr<SceneHolder> scene;
    /* a state produced on a pass-slot stream (futhark_entry_step in a stepping loop): `ready` is recorded behind the kernels
     * that write its image.  The context's stream is made to wait for it when the state is handed to the host (or dropped),
     * so every other entry point, all of which work on the context's stream, sees a finished image; the input image stays
     * referenced until then. */
    struct Sync { cudaEvent_t ready = nullptr; DevRef prev_img; ~Sync() { if (ready) cudaEventDestroy(ready); } };
    std::shared_ptr<Sync> sync;
    uint64_t serial = 0;                 /* identity of this value (pointers are reused by the allocator) */
};

template <class T, int R> struct fut_array { DevRef mem; int64_t shape[R]; int64_t count() const { int64_t c = 1; for (int i = 0; i < R; i++) c *= shape[i]; return c; } T *ptr() const { return (T *)mem->p; } };
struct futhark_f32_1d : fut_array<float, 1> {};
struct futhark_f32_2d : fut_array<float, 2> {};
struct futhark_f32_3d : fut_array<float, 3> {};
struct futhark_u32_1d : fut_array<uint32_t, 1> {};
struct futhark_i32_2d : fut_array<int32_t, 2> {};

namespace {

template <class A, class T> A *new_array(futhark_context *ctx, const T *data, const int64_t *shape, int rank, cudaMemcpyKind kind) {
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);          /* a host may hold contexts on several devices */
    A *a = new A();
    int64_t c = 1;
    for (int i = 0; i < rank; i++) { a->shape[i] = shape[i]; c *= shape[i]; }
    if (c < 0) { delete a; set_error(ctx, "negative array dimension"); return nullptr; }
    a->mem = dev_alloc(ctx, sizeof(T) * (size_t)c);
    if (!a->mem) { delete a; return nullptr; }
    if (c > 0 && data) {
        if (!cu_ok(ctx, cudaMemcpyAsync(a->mem->p, data, sizeof(T) * (size_t)c, kind, ctx->stream), "cudaMemcpy") ||
            !cu_ok(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize")) { delete a; return nullptr; }
    }
    return a;
}
template <class A, class T> int array_values(futhark_context *ctx, A *a, T *out) {
    if (!ctx || !a || !out) { if (ctx) set_error(ctx, "null argument"); return 1; }
    cudaSetDevice(ctx->device);
    CU(ctx, cudaMemcpyAsync(out, a->mem->p, sizeof(T) * (size_t)a->count(), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

void grid_dims(const futhark_opaque_state *s, uint32_t &gw, uint32_t &gh) {      /* integrator.fut:175-176 */
    gw = (s->dim_w + s->subsampling - 1) / s->subsampling;
    gh = (s->dim_h + s->subsampling - 1) / s->subsampling;
}

bool ensure_bufs(futhark_context *ctx, PassBuffers &b, int64_t n, bool probes) {
    if (b.cap < n) {
        raw_free(b.ray_o[0]); raw_free(b.ray_o[1]); raw_free(b.ray_d[0]); raw_free(b.ray_d[1]); raw_free(b.dist[0]); raw_free(b.dist[1]); raw_free(b.acc);
        raw_free(b.chan); raw_free(b.queue[0]); raw_free(b.queue[1]); raw_free(b.hit); raw_free(b.order[0]); raw_free(b.order[1]); raw_free(b.sh_o); raw_free(b.sh_d1); raw_free(b.sh_d2);
        raw_free(b.sh_c); raw_free(b.probe_rad); raw_free(b.probe_dist);
        b.cap = 0;
        size_t c = (size_t)n;
        if (!raw_alloc(ctx, b.ray_o[0], c) || !raw_alloc(ctx, b.ray_o[1], c) || !raw_alloc(ctx, b.ray_d[0], c) || !raw_alloc(ctx, b.ray_d[1], c) ||
            !raw_alloc(ctx, b.dist[0], c) || !raw_alloc(ctx, b.dist[1], c) || !raw_alloc(ctx, b.acc, c) || !raw_alloc(ctx, b.chan, c) ||
            !raw_alloc(ctx, b.queue[0], c) || !raw_alloc(ctx, b.queue[1], c) || !raw_alloc(ctx, b.hit, c) || !raw_alloc(ctx, b.order[0], c) || !raw_alloc(ctx, b.order[1], c) || !raw_alloc(ctx, b.sh_o, c) ||
            !raw_alloc(ctx, b.sh_d1, c) || !raw_alloc(ctx, b.sh_d2, c) || !raw_alloc(ctx, b.sh_c, c)) return false;
        b.cap = n;
    }
    if (!b.counts) {
        if (!raw_alloc(ctx, b.counts, LYS_MAX_PATH_LEN + 1) || !raw_alloc(ctx, b.stats, 4) || !raw_alloc(ctx, b.split, 2 * (LYS_MAX_PATH_LEN + 1))) return false;
        CUB(ctx, cudaMemsetAsync(b.stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
    }
    if (!ctx->bufs.tx_lights && !raw_alloc(ctx, ctx->bufs.tx_lights, 8)) return false;
    b.tx_lights = ctx->bufs.tx_lights;          /* one copy shared by all buffer sets */
    if (probes && !b.probe_rad) {
        if (!raw_alloc(ctx, b.probe_rad, (size_t)b.cap * 16) || !raw_alloc(ctx, b.probe_dist, (size_t)b.cap * 16)) return false;
    }
    return true;
}
bool ensure_pass_buffers(futhark_context *ctx, int64_t n, bool probes) { return ensure_bufs(ctx, ctx->bufs, n, probes); }
void free_bufs(PassBuffers &b, bool owns_tx) {
    raw_free(b.ray_o[0]); raw_free(b.ray_o[1]); raw_free(b.ray_d[0]); raw_free(b.ray_d[1]); raw_free(b.dist[0]); raw_free(b.dist[1]); raw_free(b.acc);
    raw_free(b.chan); raw_free(b.queue[0]); raw_free(b.queue[1]); raw_free(b.hit); raw_free(b.order[0]); raw_free(b.order[1]); raw_free(b.sh_o); raw_free(b.sh_d1); raw_free(b.sh_d2);
    raw_free(b.sh_c); raw_free(b.counts); raw_free(b.stats); raw_free(b.split); raw_free(b.probe_rad); raw_free(b.probe_dist);
    if (owns_tx) raw_free(b.tx_lights);
    b.cap = 0;
}
/* `want` pipeline slots with buffers for n paths each */
bool ensure_slots(futhark_context *ctx, int want, int64_t n) {
    while ((int)ctx->slots.size() < want) {
        futhark_context::PassSlot sl;
        if (!cu_ok(ctx, cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
            !cu_ok(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming), "cudaEventCreate")) return false;
        ctx->slots.push_back(sl);
    }
    for (int i = 0; i < want; i++) if (!ensure_bufs(ctx, ctx->slots[i].bufs, n, false)) return false;
    return true;
}
bool ensure_scratch(futhark_context *ctx, int64_t n) {
    BuildScratch &w = ctx->scratch;
    if (w.cap >= n) return true;
    raw_free(w.box_c); raw_free(w.box_h); raw_free(w.F); raw_free(w.chunk_lo); raw_free(w.chunk_hi); raw_free(w.keys[0]); raw_free(w.keys[1]); raw_free(w.vals[0]); raw_free(w.vals[1]);
    raw_free(w.rs_hist); raw_free(w.rs_status); raw_free(w.leaf_parent); raw_free(w.visits); raw_free(w.crown_box); raw_free(w.crown_cnt);
    if (w.crown_pairs) { cudaFree(w.crown_pairs); w.crown_pairs = nullptr; }
    w.cap = 0;
    size_t c = (size_t)n, tiles = (c + 4095) / 4096;
    if (!raw_alloc(ctx, w.chunk_lo, (c + 255) / 256) || !raw_alloc(ctx, w.chunk_hi, (c + 255) / 256) || !raw_alloc(ctx, w.box_c, c) || !raw_alloc(ctx, w.box_h, c) || !raw_alloc(ctx, w.F, 2 * c) || !raw_alloc(ctx, w.keys[0], c) ||
        !raw_alloc(ctx, w.keys[1], c) || !raw_alloc(ctx, w.vals[0], c) || !raw_alloc(ctx, w.vals[1], c) || !raw_alloc(ctx, w.rs_hist, 1024) ||
        !raw_alloc(ctx, w.rs_status, 4 * tiles * 256 + 4) || !raw_alloc(ctx, w.leaf_parent, c) || !raw_alloc(ctx, w.visits, c)) return false;
    w.crown_cap = (int)std::min<size_t>(2 * c + 1024, (size_t)1 << 30);
    { char *cp = nullptr; if (!raw_alloc(ctx, cp, (size_t)w.crown_cap * 12)) return false; w.crown_pairs = cp; }
    if (!raw_alloc(ctx, w.crown_box, (size_t)2 * w.crown_cap) || !raw_alloc(ctx, w.crown_cnt, 64)) return false;
    w.cap = n;
    return true;
}

/* the per-pass constants; camera vectors exactly as camera.fut:82-101 derives them */
bool make_frame_params(futhark_context *ctx, const futhark_opaque_state *s, uint32_t rng, float intensity_factor, FrameParams &fp) {
    uint32_t gw, gh; grid_dims(s, gw, gh);
    fp.gw = (int)gw; fp.gh = (int)gh; fp.fw = (float)gw; fp.fh = (float)gh;
    fp.rank = ctx->rank; fp.world = ctx->world;
    int rows = ((int)gh - ctx->rank + ctx->world - 1) / ctx->world;
    if (rows < 0) rows = 0;
    fp.n_local = rows * (int)gw;
    fp.frame_rng = rng;
    const Camera &cam = s->cam; const CamConf &cf = cam.conf;
    if (ctx->h_counts) {        /* queue-length estimates are only kept for the same scene seen by the same camera */
        uint64_t tag = 1469598103934665603ull;
        auto mix = [&tag](uint64_t x) { tag = (tag ^ x) * 1099511628211ull; };
        mix((uint64_t)(uintptr_t)s->scene.get()); mix((uint64_t)ctx->path_len); mix((uint64_t)s->cam_conf_id);
        uint32_t w[5]; memcpy(&w[0], &cam.origin.x, 4); memcpy(&w[1], &cam.origin.y, 4); memcpy(&w[2], &cam.origin.z, 4); memcpy(&w[3], &cam.pitch, 4); memcpy(&w[4], &cam.yaw, 4);
        for (int k = 0; k < 5; k++) mix(w[k]);
        if (tag != ctx->est_tag) { for (int k = 0; k <= LYS_MAX_PATH_LEN; k++) ctx->h_counts[k] = -1; ctx->est_tag = tag; }
    }
    float ratio = fp.fw / fp.fh;
    fp.lens_radius = cf.aperture / 2;
    float half_height = tanf(cf.fov / 2.0f);
    float half_width = ratio * half_height;
    V3 w = (-1.0f) * h_cam_dir(cam), u = h_cam_right(cam), v = h_cam_up(cam);
    float focus = cf.focal_dist;
    fp.cam_origin = cam.origin;
    fp.llc = ((cam.origin - (half_width * focus) * u) - (half_height * focus) * v) - focus * w;
    fp.horizontal = (2 * half_width * focus) * u;
    fp.vertical = (2 * half_height * focus) * v;
    fp.cam_u = u; fp.cam_v = v;
    fp.offset_radius = cf.offset_radius;
    fp.n_sensor = cf.n_sensor;
    for (int i = 0; i < 3; i++) { fp.sensor_mu[i] = cf.mu[i]; fp.sensor_sigma[i] = cf.sigma[i]; fp.sensor_vis[i] = cf.vis[i]; }
    fp.tx_kind = cf.tx_kind; fp.tx_radius = cf.tx_radius; fp.tx_theta = cf.tx_theta;
    memcpy(fp.tx_emission, cf.tx_emission.k, sizeof(fp.tx_emission));
    {   /* shapes.fut:18-28: a = 2*pi / n_sectors, angles a*i; rot_z b (1,0,0) */
        float a = 2 * LYS_PI / (float)8;
        for (int j = 0; j <= 8; j++) {
            float bj = a * (float)j;
            fp.sector_x[j] = 1.0f * cosf(bj) - 0.0f * sinf(bj);
            fp.sector_y[j] = 1.0f * sinf(bj) + 0.0f * cosf(bj);
        }
    }
    fp.n_scene_lights = (int)s->scene->d.n_lights;
    memcpy(fp.ambience, s->ambience.k, sizeof(fp.ambience));
    fp.path_len = ctx->path_len;
    fp.render_mode = s->render_mode;
    fp.intensity_factor = intensity_factor;
    if (cf.tx_kind == 1) {
        /* flash: disk c.origin (cam_dir c) radius 8 (camera.fut:116-118), identical for every ray */
        LightRec recs[8];
        V3 normal = h_cam_dir(cam);
        V3 c = cross(normal, v3(0, 1, 0));
        V3 right = (norm(c) == 0) ? v3(1, 0, 0) : normalise(c);
        V3 up = normalise(cross(right, normal));
        for (int k = 0; k < 8; k++) {
            V3 v0 = fp.sector_x[k] * right + fp.sector_y[k] * up;
            V3 v1 = fp.sector_x[k + 1] * right + fp.sector_y[k + 1] * up;
            V3 A = cam.origin, B = cam.origin + cf.tx_radius * v1, C = cam.origin + cf.tx_radius * v0;
            V3 e1 = B - A, e2 = C - A, nc = cross(e1, e2);
            float area = norm(nc) / 2.0f; V3 n = normalise(nc);
            LightRec &r = recs[k];
            r.a[0] = A.x; r.a[1] = A.y; r.a[2] = A.z; r.area = area;
            r.e1[0] = e1.x; r.e1[1] = e1.y; r.e1[2] = e1.z; r.inv_area = 1.0f / area;
            r.e2[0] = e2.x; r.e2[1] = e2.y; r.e2[2] = e2.z; r.theta = 0.0f;
            r.n[0] = n.x; r.n[1] = n.y; r.n[2] = n.z; r.kind = 0;
            memcpy(r.emission, cf.tx_emission.k, sizeof(r.emission));
            r.src_index = -1; r.pad[0] = r.pad[1] = r.pad[2] = 0;
        }
        CUB(ctx, cudaMemcpyAsync(ctx->bufs.tx_lights, recs, sizeof(recs), cudaMemcpyHostToDevice, ctx->stream));
        CUB(ctx, cudaStreamSynchronize(ctx->stream));     /* recs is a stack buffer */
    }
    return true;
}

uint32_t h_advance_rng(uint32_t s) { return lys_pin_lcg(s); }                      /* rand.fut:11-12 */
uint32_t h_rng_from_seed(int32_t seed) {                                            /* cpprandom rng_from_seed [seed] */
    return lys_pin_rng_from_seed(seed);
}

std::atomic<uint64_t> g_state_serial{1};
futhark_opaque_state *clone_state(const futhark_opaque_state *s) { futhark_opaque_state *r = new futhark_opaque_state(*s); r->serial = g_state_serial++; return r; }


/* sample_frame / sample_frame_accum (integrator.fut:172-192) into `img` */
bool sample_into(futhark_context *ctx, const futhark_opaque_state *s, uint32_t rng, const float *img_old, float *img_new, bool merge, float n_frames) {
    FrameParams fp;
    if (!ensure_pass_buffers(ctx, (int64_t)((s->dim_w + s->subsampling - 1) / s->subsampling) * ((s->dim_h + s->subsampling - 1) / s->subsampling), false)) return false;
    if (!make_frame_params(ctx, s, rng, 1.0f, fp)) return false;
    CUB(ctx, run_sample_pass(s->scene->d, fp, ctx->bufs, ctx->stream, &ctx->launches, &ctx->timer, ctx->h_counts));
    CUB(ctx, run_accumulate(fp, ctx->bufs, img_old, img_new, merge ? 1 : 0, n_frames, ctx->stream, &ctx->launches, &ctx->timer));
    if (ctx->timer.on && ctx->timer.used > 3000) ctx->timer.resolve(ctx->stream);
    return true;
}

/* drop the steps computed ahead; the context's stream waits for their kernels before their images return to the pool */
void drop_ahead(futhark_context *ctx) {
    for (auto &a : ctx->ahead) {
        if (a.out->sync && a.out->sync->ready) cudaStreamWaitEvent(ctx->stream, a.out->sync->ready, 0);
        delete a.out;
    }
    ctx->ahead.clear();
}

} // namespace

/* ================================================================== C ABI */
extern "C" {

struct futhark_context_config *futhark_context_config_new(void) { return new futhark_context_config(); }
void futhark_context_config_free(struct futhark_context_config *cfg) { delete cfg; }
void futhark_context_config_set_device(struct futhark_context_config *cfg, const char *s) {
    /* as the generated cuda backend parses it: an optional "#k" (k-th matching device), then a substring of the device
     * name; a bare number is a name substring, not an index */
    if (!cfg || !s) return;
    int k = 0;
    if (*s == '#') { s++; while (*s >= '0' && *s <= '9') k = k * 10 + (*s++ - '0'); while (*s == ' ' || *s == '\t') s++; }
    cfg->device = k; cfg->device_name = s;
}
void futhark_context_config_set_debugging(struct futhark_context_config *cfg, int flag) { if (cfg) cfg->debugging = flag; }
void futhark_context_config_set_logging(struct futhark_context_config *cfg, int flag) { if (cfg) cfg->logging = flag; }

struct futhark_context *futhark_context_new(struct futhark_context_config *cfg) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        fprintf(stderr, "libtracer: no CUDA device available (this library has no CPU path)\n");
        return nullptr;
    }
    int dev = -1;
    {
        int want = cfg ? cfg->device : 0, seen = 0;
        const char *name = cfg ? cfg->device_name.c_str() : "";
        for (int i = 0; i < count && dev < 0; i++) {
            cudaDeviceProp p;
            if (cudaGetDeviceProperties(&p, i) != cudaSuccess || !strstr(p.name, name)) continue;
            if (seen++ == want) dev = i;
        }
        if (dev < 0) { fprintf(stderr, "libtracer: no CUDA device #%d matching '%s'\n", want, name); return nullptr; }
    }
    if (cudaSetDevice(dev) != cudaSuccess) return nullptr;
    /* the sample pass alternates kernels with different local-memory footprints (traversal stacks, shading
     * temporaries); without this flag the driver may shrink / regrow the local-memory pool between launches */
    if (!getenv("LYS_NO_LMEM_FLAG")) {
        unsigned int flags = 0;
        if (cudaGetDeviceFlags(&flags) == cudaSuccess && !(flags & cudaDeviceLmemResizeToMax)) { cudaSetDeviceFlags(flags | cudaDeviceLmemResizeToMax); cudaGetLastError(); }
    }
    futhark_context *ctx = new futhark_context();
    ctx->device = dev; ctx->logging = cfg ? cfg->logging : 0;
    ctx->timer.on = (cfg && cfg->profiling) ? 1 : 0;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return nullptr; }
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    { const char *pe = getenv("LYS_PIPELINE"); if (pe) { int v = atoi(pe); if (v >= 1 && v <= 16) ctx->pipeline = v; } }
    { const char *ae = getenv("LYS_STEP_AHEAD"); if (ae) { int v = atoi(ae); if (v >= 0 && v <= 4) ctx->ahead_depth = v; } }   /* steps started ahead in a stepping loop (0: none) */
    if (cudaHostAlloc((void **)&ctx->h_init, sizeof(*ctx->h_init), cudaHostAllocDefault) != cudaSuccess) { cudaStreamDestroy(ctx->stream); delete ctx; return nullptr; }
    { size_t fr = 0, tot = 0; if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) ctx->pool_cap = std::max<size_t>((size_t)1 << 30, tot / 8); else cudaGetLastError(); }
    if (cudaHostAlloc((void **)&ctx->h_counts, sizeof(int) * (LYS_MAX_PATH_LEN + 1), cudaHostAllocDefault) == cudaSuccess) {
        for (int k = 0; k <= LYS_MAX_PATH_LEN; k++) ctx->h_counts[k] = -1;          /* no estimate yet */
    } else { ctx->h_counts = nullptr; cudaGetLastError(); }
    const char *pl = getenv("LYS_PATH_LEN");
    if (pl) { int v = atoi(pl); if (v >= 1 && v <= LYS_MAX_PATH_LEN) ctx->path_len = v; }
    return ctx;
}
void futhark_context_free(struct futhark_context *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    drop_ahead(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto &sl : ctx->slots) { cudaStreamSynchronize(sl.stream); free_bufs(sl.bufs, false); cudaStreamDestroy(sl.stream); cudaEventDestroy(sl.done); }
    ctx->slots.clear();
    free_bufs(ctx->bufs, true);
    BuildScratch &w = ctx->scratch;
    raw_free(w.box_c); raw_free(w.box_h); raw_free(w.F); raw_free(w.chunk_lo); raw_free(w.chunk_hi); raw_free(w.keys[0]); raw_free(w.keys[1]); raw_free(w.vals[0]); raw_free(w.vals[1]);
    raw_free(w.rs_hist); raw_free(w.rs_status); raw_free(w.leaf_parent); raw_free(w.visits); raw_free(w.crown_box); raw_free(w.crown_cnt);
    if (w.crown_pairs) cudaFree(w.crown_pairs);
    raw_free(ctx->pts_pos); raw_free(ctx->pts_dist);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->h_init) cudaFreeHost(ctx->h_init);
    for (auto &kv : ctx->pool) cudaFree(kv.second);
    ctx->pool.clear(); ctx->pooled_bytes = 0;
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
int futhark_context_sync(struct futhark_context *ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
int futhark_context_clear_caches(struct futhark_context *ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    pool_release_all(ctx);
    return 0;
}
/* profiling surface of a generated cuda-backend header (tracer.h) */
void futhark_context_config_set_profiling(struct futhark_context_config *cfg, int flag) { if (cfg) cfg->profiling = flag; }
void futhark_context_pause_profiling(struct futhark_context *ctx) { if (ctx) { ctx->timer.resolve(ctx->stream); ctx->timer.on = 0; } }
void futhark_context_unpause_profiling(struct futhark_context *ctx) { if (ctx) ctx->timer.on = 1; }
char *futhark_context_report(struct futhark_context *ctx) {
    if (!ctx) return nullptr;
    ctx->timer.resolve(ctx->stream);
    static const char *const cls[LYS_PROFILE_CLASSES] = {"generate", "trace", "shade", "tail", "accumulate"};
    const size_t pooled = ctx->pooled_bytes;
    std::string r = "libtracer (sm_100a): " + std::to_string((unsigned long long)ctx->launches) + " kernel launches, " +
                    std::to_string((unsigned long long)pooled) + " bytes of device memory pooled for reuse\n";
    char line[160];
    for (int i = 0; i < LYS_PROFILE_CLASSES; i++) {
        if (!ctx->timer.n[i]) continue;
        snprintf(line, sizeof line, "%-10s ran %8llu times; avg: %8.1fus; total: %10.1fus\n", cls[i], (unsigned long long)ctx->timer.n[i],
                 1e3 * ctx->timer.ms[i] / (double)ctx->timer.n[i], 1e3 * ctx->timer.ms[i]);
        r += line;
    }
    return strdup(r.c_str());
}
char *futhark_context_get_error(struct futhark_context *ctx) {
    if (!ctx || ctx->error.empty()) return nullptr;        /* NULL when there is nothing to report, as the generated code does */
    char *r = strdup(ctx->error.c_str());
    ctx->error.clear();
    return r;
}

/* ---- arrays ---- */
#define LYS_ARRAY_API(NAME, T, RANK, DIMS_DECL, DIMS_INIT)                                                            \
    struct futhark_##NAME *futhark_new_##NAME(struct futhark_context *ctx, const T *data, DIMS_DECL) {                \
        int64_t shape[RANK] = DIMS_INIT;                                                                              \
        return new_array<futhark_##NAME, T>(ctx, data, shape, RANK, cudaMemcpyHostToDevice);                          \
    }                                                                                                                 \
    int futhark_free_##NAME(struct futhark_context *ctx, struct futhark_##NAME *arr) { (void)ctx; delete arr; return 0; } \
    int futhark_values_##NAME(struct futhark_context *ctx, struct futhark_##NAME *arr, T *data) { return array_values(ctx, arr, data); } \
    const int64_t *futhark_shape_##NAME(struct futhark_context *ctx, struct futhark_##NAME *arr) { (void)ctx; return arr ? arr->shape : nullptr; } \
    struct futhark_##NAME *futhark_new_raw_##NAME(struct futhark_context *ctx, const futhark_deviceptr data, int offset, DIMS_DECL) { \
        int64_t shape[RANK] = DIMS_INIT;                                                                              \
        return new_array<futhark_##NAME, T>(ctx, (const T *)(uintptr_t)(data + (unsigned long long)(long long)offset), shape, RANK, cudaMemcpyDeviceToDevice); \
    }                                                                                                                 \
    futhark_deviceptr futhark_values_raw_##NAME(struct futhark_context *ctx, struct futhark_##NAME *arr) { (void)ctx; return arr ? (futhark_deviceptr)(uintptr_t)arr->mem->p : 0; }
#define LYS_D1 int64_t dim0
#define LYS_D2 int64_t dim0, int64_t dim1
#define LYS_D3 int64_t dim0, int64_t dim1, int64_t dim2
#define LYS_I1 {dim0}
#define LYS_I2 {dim0, dim1}
#define LYS_I3 {dim0, dim1, dim2}
LYS_ARRAY_API(f32_1d, float, 1, LYS_D1, LYS_I1)
LYS_ARRAY_API(f32_2d, float, 2, LYS_D2, LYS_I2)
LYS_ARRAY_API(f32_3d, float, 3, LYS_D3, LYS_I3)
LYS_ARRAY_API(u32_1d, uint32_t, 1, LYS_D1, LYS_I1)
LYS_ARRAY_API(i32_2d, int32_t, 2, LYS_D2, LYS_I2)

int futhark_free_opaque_state(struct futhark_context *ctx, struct futhark_opaque_state *obj) {
    if (ctx && obj && !ctx->ahead.empty() && ctx->ahead.front().in_serial == obj->serial) { cudaSetDevice(ctx->device); drop_ahead(ctx); }   /* nobody can ask for them any more */
    delete obj;
    return 0;
}

/* ---- init (lib.fut:76-106) ---- */
int futhark_entry_init(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t seed, const uint32_t h,
                       const uint32_t w, const uint32_t cam_conf_id, const struct futhark_f32_3d *tri_geoms,
                       const struct futhark_u32_1d *tri_mats, const struct futhark_f32_2d *mat_data, const float cam_pitch,
                       const float cam_yaw, const struct futhark_f32_1d *cam_origin) {
    if (!ctx) return 1;
    if (!out0 || !tri_geoms || !tri_mats || !mat_data || !cam_origin) { set_error(ctx, "init: null argument"); return 1; }
    cudaSetDevice(ctx->device);
    const int64_t n = tri_geoms->shape[0], m = mat_data->shape[0];
    if (tri_geoms->shape[1] != 3 || tri_geoms->shape[2] != 3) { set_error(ctx, "init: tri_geoms must be [n][3][3]"); return 1; }
    if (tri_mats->shape[0] != n) { set_error(ctx, "init: tri_mats must be [n] (same n as tri_geoms)"); return 1; }
    if (mat_data->shape[1] != 28) { set_error(ctx, "init: mat_data must be [m][28]"); return 1; }
    if (cam_origin->shape[0] != 3) { set_error(ctx, "init: cam_origin must be [3]"); return 1; }
    if (n < 2) { set_error(ctx, "init: at least 2 triangles are required (radix_tree.mk, radix_tree.fut:75)"); return 1; }
    if (n >= (1ll << 30)) { set_error(ctx, "init: too many triangles"); return 1; }
    if (m < 1) { set_error(ctx, "init: at least 1 material is required"); return 1; }

    auto holder = std::make_shared<SceneHolder>();
    SceneDev &sc = holder->d;
    sc.n_tris = n; sc.n_mats = m; sc.n_lights = 0;
    holder->refit_mode = ctx->refit_mode;
    size_t c = (size_t)n;
    SceneHolder &H = *holder;
    /* lights are found on the device (scene.fut:58-66); their number is only known after the read-back below, so the
     * arrays get a first capacity that covers every bundled scene and are regrown in the rare case it does not */
    int light_cap = (int)std::min<int64_t>(n, 4096);
    const size_t light_chunks = (c + 1023) / 1024;
    unsigned char *mat_flag = nullptr; int *light_chunk = nullptr, *light_info = nullptr;
    if (!H.take(ctx, sc.tris, 9 * c) || !H.take(ctx, sc.tri_mats, c) || !H.take(ctx, sc.mats, (size_t)m * 28) ||
        !H.take(ctx, sc.leaf_tri, 4 * c) || !H.take(ctx, sc.leaf_box, 2 * c) || !H.take(ctx, sc.leaf_frame, 3 * c) ||
        !H.take(ctx, sc.node_box, 2 * c) || !H.take(ctx, sc.left, c) || !H.take(ctx, sc.right, c) || !H.take(ctx, sc.parent, c) ||
        !H.take(ctx, sc.height, c) || !H.take(ctx, sc.morton, c) || !H.take(ctx, sc.sorted_idx, c) || !H.take(ctx, sc.bounds, 8) ||
        !H.take(ctx, sc.lights, (size_t)light_cap) || !H.take(ctx, sc.light_src, (size_t)light_cap) ||
        !H.take(ctx, mat_flag, (size_t)m) || !H.take(ctx, light_chunk, light_chunks) || !H.take(ctx, light_info, 4)) return 1;
    static const int one_copy = []() { const char *e = getenv("LYS_OCT_ONE_COPY"); return (e && atoi(e)) ? 1 : 0; }();   /* tests: the large-scene record arrays on a small scene */
    sc.oct_copies = (n - 1 <= LYS_OCT_MAX_NODES && !one_copy) ? 8 : 1;
    if (!H.take(ctx, sc.nodes, (size_t)sc.oct_copies * 2 * c)) return 1;
    CU(ctx, cudaMemcpyAsync(sc.tris, tri_geoms->mem->p, sizeof(float) * 9 * c, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(sc.tri_mats, tri_mats->mem->p, sizeof(uint32_t) * c, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(sc.mats, mat_data->mem->p, sizeof(float) * (size_t)m * 28, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(ctx, build_lights(sc, mat_flag, light_chunk, light_info, light_cap, ctx->stream, &ctx->launches));
    if (!ensure_scratch(ctx, n)) return 1;
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CU(ctx, build_lbvh(sc, ctx->scratch, ctx->refit_mode, ctx->stream, &ctx->launches));
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    /* the one read-back of init: number of lights + index check, camera origin, crown overflow flag of the refit */
    futhark_context::InitReadback *rb = ctx->h_init;
    rb->crown_overflow = 0;
    CU(ctx, cudaMemcpyAsync(rb->lights, light_info, sizeof(rb->lights), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(rb->origin, cam_origin->mem->p, sizeof(rb->origin), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->refit_mode == 0) CU(ctx, cudaMemcpyAsync(&rb->crown_overflow, ctx->scratch.crown_cnt + 63, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&sc.build_ms, ctx->ev0, ctx->ev1);
    if (rb->lights[1]) { set_error(ctx, "init: material index out of range"); return 1; }
    sc.n_lights = rb->lights[0];
    if (sc.n_lights > light_cap) {                     /* more lights than the first capacity: larger arrays, scatter + records again */
        light_cap = (int)sc.n_lights;
        if (!H.take(ctx, sc.lights, (size_t)light_cap) || !H.take(ctx, sc.light_src, (size_t)light_cap)) return 1;
        CU(ctx, rebuild_lights(sc, mat_flag, light_chunk, light_info, light_cap, ctx->stream, &ctx->launches));
    }
    if (rb->crown_overflow) CU(ctx, build_lbvh(sc, ctx->scratch, 2, ctx->stream, &ctx->launches));      /* pair buffer overflowed: literal sweeps (always exact) */
    const float org[3] = {rb->origin[0], rb->origin[1], rb->origin[2]};

    futhark_opaque_state *s = new futhark_opaque_state();
    s->serial = g_state_serial++;
    s->dim_w = w; s->dim_h = h; s->subsampling = 1;
    s->rng = h_rng_from_seed(seed);
    s->img_h = h; s->img_w = w;
    s->img = dev_alloc(ctx, sizeof(float) * 3 * (size_t)w * h);
    if (!s->img) { delete s; return 1; }
    CU(ctx, cudaMemsetAsync(s->img->p, 0, sizeof(float) * 3 * (size_t)w * h, ctx->stream));
    s->n_frames = 0; s->ambience = h_uniform_spectrum(0); s->mode = false;                 /* no_sky spectrum.fut:91 */
    if (cam_conf_id == 0) { s->render_mode = 0; s->cam.conf = conf_visual(); }
    else if (cam_conf_id == 1) { s->render_mode = 0; s->cam.conf = conf_flash(); }
    else { s->render_mode = 1; s->cam.conf = conf_lidar(); }
    s->cam_conf_id = cam_conf_id;
    s->cam.pitch = cam_pitch; s->cam.yaw = cam_yaw; s->cam.origin = v3(org[0], org[1], org[2]);
    s->scene = holder;
    *out0 = s;
    return 0;
}

int futhark_entry_resize(struct futhark_context *ctx, struct futhark_opaque_state **out0, const uint32_t h, const uint32_t w,
                         const struct futhark_opaque_state *s) {                 /* lib.fut:108-109 */
    if (!ctx) return 1;
    if (!out0 || !s) { set_error(ctx, "resize: null argument"); return 1; }
    futhark_opaque_state *r = clone_state(s);
    r->dim_w = w; r->dim_h = h; r->mode = false;
    *out0 = r;
    return 0;
}

int futhark_entry_key(struct futhark_context *ctx, struct futhark_opaque_state **out0, const int32_t e, const int32_t key,
                      const struct futhark_opaque_state *s) {                    /* lib.fut:120-185; key codes src/sdl.fut */
    if (!ctx) return 1;
    if (!out0 || !s) { set_error(ctx, "key: null argument"); return 1; }
    futhark_opaque_state *r = clone_state(s);
    if (e == 0) {
        switch (key) {
            case 0x32: r->subsampling = s->subsampling + 1; r->n_frames = 0; break;                               /* SDLK_2 */
            case 0x31: r->subsampling = (s->subsampling - 1u > 1u) ? s->subsampling - 1u : 1u; r->n_frames = 0; break; /* SDLK_1: u32.max 1 (sub-1) */
            case 0x77: r->cam = h_move(s->cam, v3(0, 0, 1)); r->n_frames = 0; break;                              /* w */
            case 0x61: r->cam = h_move(s->cam, v3(-1, 0, 0)); r->n_frames = 0; break;                             /* a */
            case 0x73: r->cam = h_move(s->cam, v3(0, 0, -1)); r->n_frames = 0; break;                             /* s */
            case 0x64: r->cam = h_move(s->cam, v3(1, 0, 0)); r->n_frames = 0; break;                              /* d */
            case 0x40000052: r->cam = h_turn(s->cam, -0.1f, 0.0f); r->n_frames = 0; break;                        /* UP */
            case 0x40000051: r->cam = h_turn(s->cam, 0.1f, 0.0f); r->n_frames = 0; break;                         /* DOWN */
            case 0x4000004F: r->cam = h_turn(s->cam, 0.0f, 0.1f); r->n_frames = 0; break;                         /* RIGHT */
            case 0x40000050: r->cam = h_turn(s->cam, 0.0f, -0.1f); r->n_frames = 0; break;                        /* LEFT */
            case 0x78: r->cam = h_move(s->cam, v3(0, 1, 0)); r->n_frames = 0; break;                              /* x */
            case 0x7A: r->cam = h_move(s->cam, v3(0, -1, 0)); r->n_frames = 0; break;                             /* z */
            case 0x20: r->mode = !s->mode; r->n_frames = 0; break;                                                /* SPACE */
            case 0x6E: r->mode = false; r->n_frames = 0; break;                                                   /* n */
            case 0x6D: r->mode = true; break;                                                                     /* m */
            case 0x69: r->cam.conf.aperture = lys_fminf(2.0f, s->cam.conf.aperture + 0.08f); break;               /* i */
            case 0x6B: r->cam.conf.aperture = lys_fmaxf(0.0f, s->cam.conf.aperture - 0.08f); break;               /* k */
            case 0x6F: r->cam.conf.focal_dist = s->cam.conf.focal_dist * 1.14f; break;                            /* o */
            case 0x6C: r->cam.conf.focal_dist = lys_fmaxf(0.1f, s->cam.conf.focal_dist / 1.14f); break;           /* l */
            case 0x74:                                                                                            /* t */
                if (s->cam_conf_id == 0) { r->cam.conf = conf_flash(); r->cam_conf_id = 1; r->render_mode = 0; }
                else if (s->cam_conf_id == 1) { r->cam.conf = conf_lidar(); r->cam_conf_id = 2; r->render_mode = 1; }
                else { r->cam.conf = conf_visual(); r->cam_conf_id = 0; r->render_mode = 0; }
                r->n_frames = 0; break;
            case 0x70: r->ambience = (s->ambience.k[1] == 0) ? h_scale(h_blackbody_normalized(17000.0f), 5.0f) : h_uniform_spectrum(0); break; /* p */
            default: break;
        }
    }
    *out0 = r;
    return 0;
}

/* ---- step (lib.fut:111-118) ----
 * step_on_slot: one step computed on a pass-slot stream instead of the context's stream.  The slot stream first waits for
 * everything issued on the context's stream so far (the block pool is ordered on that stream), and its accumulate kernel for
 * the input state's own `ready` event. */
static futhark_opaque_state *step_on_slot(struct futhark_context *ctx, const struct futhark_opaque_state *s) {
    uint32_t gw, gh; grid_dims(s, gw, gh);
    const bool accum = s->mode && s->n_frames > 0;
    if (accum && (s->img_w != gw || s->img_h != gh)) return nullptr;
    const int S = ctx->ahead_depth + 1;
    if (!ensure_slots(ctx, S, (int64_t)gw * gh)) return nullptr;
    futhark_context::PassSlot &sl = ctx->slots[ctx->ahead_slot++ % (uint64_t)S];
    std::unique_ptr<futhark_opaque_state> r(clone_state(s));
    r->img = dev_alloc(ctx, sizeof(float) * 3 * (size_t)gw * gh);
    if (!r->img) return nullptr;
    r->img_w = gw; r->img_h = gh;
    r->sync = std::make_shared<futhark_opaque_state::Sync>();
    r->sync->prev_img = s->img;
    FrameParams fp;
    if (!make_frame_params(ctx, s, s->rng, 1.0f, fp)) return nullptr;
    if (!cu_ok(ctx, cudaEventCreateWithFlags(&r->sync->ready, cudaEventDisableTiming), "cudaEventCreate") ||
        !cu_ok(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream), "cudaEventRecord") ||
        !cu_ok(ctx, cudaStreamWaitEvent(sl.stream, ctx->ev_fork, 0), "cudaStreamWaitEvent")) return nullptr;
    /* the pass reads the scene and the state's scalars only; the input IMAGE is needed by the accumulate kernel, so the wait
     * for the state that produces it goes between the two: the passes of consecutive steps overlap, as in sample_n_frames */
    const bool ok = cu_ok(ctx, run_sample_pass(s->scene->d, fp, sl.bufs, sl.stream, &ctx->launches, nullptr, ctx->h_counts), "sample pass") &&
                    (!s->sync || cu_ok(ctx, cudaStreamWaitEvent(sl.stream, s->sync->ready, 0), "cudaStreamWaitEvent")) &&
                    cu_ok(ctx, run_accumulate(fp, sl.bufs, (const float *)s->img->p, (float *)r->img->p, accum ? 1 : 0, (float)s->n_frames, sl.stream, &ctx->launches), "accumulate");
    /* whatever was launched writes r->img: the event goes behind it in any case, and the caller orders the context's stream
     * after it before the image can go back to the pool */
    cudaEventRecord(r->sync->ready, sl.stream);
    if (!ok) { cudaStreamWaitEvent(ctx->stream, r->sync->ready, 0); return nullptr; }
    r->rng = h_advance_rng(s->rng);                                                /* integrator.fut:116 */
    r->n_frames = accum ? s->n_frames + 1 : 1;
    return r.release();
}
int futhark_entry_step(struct futhark_context *ctx, struct futhark_opaque_state **out0, const struct futhark_opaque_state *s) { /* lib.fut:111-118 */
    if (!ctx) return 1;
    if (!out0 || !s) { set_error(ctx, "step: null argument"); return 1; }
    cudaSetDevice(ctx->device);
    uint32_t gw, gh; grid_dims(s, gw, gh);
    bool accum = s->mode && s->n_frames > 0;
    if (accum && (s->img_w != gw || s->img_h != gh)) { set_error(ctx, "step: accumulated image shape does not match the sample grid"); return 1; }
    /* the host steps the state the previous step returned: a stepping loop.  Not with a row partition (the caller reduces
     * the image itself), per-launch timing, or the flash preset (its light records are re-uploaded per pass). */
    const bool loop = ctx->ahead_depth > 0 && s->serial == ctx->last_step_out && ctx->world == 1 && !ctx->timer.on && s->cam.conf.tx_kind != 1;
    futhark_opaque_state *r = nullptr;
    if (!ctx->ahead.empty() && ctx->ahead.front().in_serial == s->serial) {         /* computed ahead */
        r = ctx->ahead.front().out;
        ctx->ahead.pop_front();
    } else {
        drop_ahead(ctx);
        if (loop) r = step_on_slot(ctx, s);
        if (!r) {                                                                   /* on the context's stream */
            r = clone_state(s);
            size_t bytes = sizeof(float) * 3 * (size_t)gw * gh;
            r->img = dev_alloc(ctx, bytes);
            if (!r->img) { delete r; return 1; }
            if (ctx->world > 1 && cudaMemsetAsync(r->img->p, 0, bytes, ctx->stream) != cudaSuccess) { delete r; set_error(ctx, "memset failed"); return 1; }
            r->img_w = gw; r->img_h = gh;
            r->sync.reset();
            if (!sample_into(ctx, s, s->rng, (const float *)s->img->p, (float *)r->img->p, accum, (float)s->n_frames)) { delete r; return 1; }
            r->rng = h_advance_rng(s->rng);                                         /* integrator.fut:116 */
            r->n_frames = accum ? s->n_frames + 1 : 1;
        }
    }
    if (r->sync && r->sync->ready) CU(ctx, cudaStreamWaitEvent(ctx->stream, r->sync->ready, 0));   /* every later call sees the finished image */
    ctx->last_step_out = r->serial;
    if (loop) {                                                                     /* start the next steps now */
        while ((int)ctx->ahead.size() < ctx->ahead_depth) {
            const futhark_opaque_state *base = ctx->ahead.empty() ? r : ctx->ahead.back().out;
            futhark_opaque_state *nx = step_on_slot(ctx, base);
            if (!nx) { ctx->error.clear(); break; }                                  /* a step that cannot run ahead is reported when the host asks for it */
            ctx->ahead.push_back({base->serial, nx});
        }
    }
    *out0 = r;
    return 0;
}

int futhark_entry_render(struct futhark_context *ctx, struct futhark_i32_2d **out0, const struct futhark_opaque_state *s) { /* lib.fut:187-196 */
    if (!ctx) return 1;
    if (!out0 || !s) { set_error(ctx, "render: null argument"); return 1; }
    cudaSetDevice(ctx->device);
    int64_t shape[2] = {(int64_t)s->dim_h, (int64_t)s->dim_w};
    futhark_i32_2d *a = new_array<futhark_i32_2d, int32_t>(ctx, nullptr, shape, 2, cudaMemcpyDeviceToDevice);
    if (!a) return 1;
    if (!cu_ok(ctx, run_render((const float *)s->img->p, (int)s->img_h, (int)s->img_w, (int)s->dim_h, (int)s->dim_w, (int)s->subsampling,
                               a->ptr(), ctx->stream, &ctx->launches), "render")) { delete a; return 1; }
    *out0 = a;
    return 0;
}

/* sample_n_frames (lib.fut:67-74); out_scale is applied by the last accumulate (1 = the reference's result) */
static int sample_n_frames_impl(struct futhark_context *ctx, struct futhark_f32_3d **out0, const struct futhark_opaque_state *s, uint32_t n,
                                float out_scale, lys_pass_stats *stats) {
    if (!ctx) return 1;
    if (!out0 || !s) { set_error(ctx, "sample_n_frames: null argument"); return 1; }
    cudaSetDevice(ctx->device);
    uint32_t gw, gh; grid_dims(s, gw, gh);
    int64_t shape[3] = {(int64_t)gh, (int64_t)gw, 3};
    std::unique_ptr<futhark_f32_3d> a(new_array<futhark_f32_3d, float>(ctx, nullptr, shape, 3, cudaMemcpyDeviceToDevice));      /* freed on every early return */
    if (!a) return 1;
    uint64_t l0 = ctx->launches;
    if (ctx->world > 1) CU(ctx, cudaMemsetAsync(a->ptr(), 0, sizeof(float) * 3 * (size_t)gw * gh, ctx->stream));
    const uint32_t passes = n < 1 ? 1 : n;                               /* the first sample_frame always runs (lib.fut:68) */
    const int S = ctx->timer.on == 1 ? 1 : (int)std::min<uint32_t>((uint32_t)ctx->pipeline, passes);
    if (!ensure_slots(ctx, S, (int64_t)gw * gh)) return 1;
    FrameParams fp;
    if (!make_frame_params(ctx, s, s->rng, 1.0f, fp)) return 1;          /* uploads the flash lights once */
    for (int i = 0; i < S; i++) CU(ctx, cudaMemsetAsync(ctx->slots[i].bufs.stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int i = 0; i < S; i++) CU(ctx, cudaStreamWaitEvent(ctx->slots[i].stream, ctx->ev_fork, 0));
    uint32_t rng = s->rng;
    cudaEvent_t prev = nullptr;
    int rc = 0;
    for (uint32_t k = 0; k < passes && !rc; k++) {
        futhark_context::PassSlot &sl = ctx->slots[k % S];
        fp.frame_rng = rng;
        if (!cu_ok(ctx, run_sample_pass(s->scene->d, fp, sl.bufs, sl.stream, &ctx->launches, &ctx->timer, ctx->h_counts), "sample pass")) { rc = 1; break; }
        if (prev && !cu_ok(ctx, cudaStreamWaitEvent(sl.stream, prev, 0), "cudaStreamWaitEvent")) { rc = 1; break; }      /* running average is order dependent */
        if (!cu_ok(ctx, run_accumulate(fp, sl.bufs, a->ptr(), a->ptr(), k > 0 ? 1 : 0, (float)k, sl.stream, &ctx->launches, &ctx->timer,
                                       k + 1 == passes ? out_scale : 1.0f), "accumulate") ||
            !cu_ok(ctx, cudaEventRecord(sl.done, sl.stream), "cudaEventRecord")) { rc = 1; break; }
        prev = sl.done;
        rng = h_advance_rng(rng);
    }
    /* join every slot stream back into the context's stream, also after an error: nothing may stay forked */
    for (int i = 0; i < S; i++) if (cudaEventRecord(ctx->slots[i].done, ctx->slots[i].stream) == cudaSuccess) cudaStreamWaitEvent(ctx->stream, ctx->slots[i].done, 0);
    if (rc) { cudaStreamSynchronize(ctx->stream); return 1; }             /* `a` goes back to the pool only after its writers have drained */
    CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    if (ctx->timer.on == 1 || ctx->timer.used > 3000) ctx->timer.resolve(ctx->stream);
    if (stats) {
        unsigned long long hs[4] = {0, 0, 0, 0}, one[4];
        for (int i = 0; i < S; i++) {
            CU(ctx, cudaMemcpyAsync(one, ctx->slots[i].bufs.stats, sizeof(one), cudaMemcpyDeviceToHost, ctx->stream));
            CU(ctx, cudaStreamSynchronize(ctx->stream));
            for (int q = 0; q < 4; q++) hs[q] += one[q];
        }
        stats->paths = (uint64_t)fp.n_local * passes; stats->vertices = hs[0]; stats->shadow_rays = hs[2];
        stats->closest_rays = 0; stats->launches = ctx->launches - l0;
        cudaEventElapsedTime(&stats->device_ms, ctx->ev0, ctx->ev1);
    }
    *out0 = a.release();
    return 0;
}
int lys_sample_n_frames_stats(struct futhark_context *ctx, struct futhark_f32_3d **out0, const struct futhark_opaque_state *s, uint32_t n,
                              lys_pass_stats *stats) {
    return sample_n_frames_impl(ctx, out0, s, n, 1.0f, stats);
}
int lys_sample_n_frames_weighted(struct futhark_context *ctx, struct futhark_f32_3d **out0, const struct futhark_opaque_state *s, uint32_t n,
                                 float weight, lys_pass_stats *stats) {
    return sample_n_frames_impl(ctx, out0, s, n, weight, stats);
}
int futhark_entry_sample_n_frames(struct futhark_context *ctx, struct futhark_f32_3d **out0, const struct futhark_opaque_state *s, const uint32_t n) {
    return sample_n_frames_impl(ctx, out0, s, n, 1.0f, nullptr);
}

int futhark_entry_sample_points_n(struct futhark_context *ctx, struct futhark_opaque_state **out0, struct futhark_f32_3d **out1,
                                  const struct futhark_opaque_state *s, const uint32_t spp) {       /* lib.fut:35-63 */
    if (!ctx) return 1;
    if (!out0 || !out1 || !s) { set_error(ctx, "sample_points_n: null argument"); return 1; }
    cudaSetDevice(ctx->device);
    uint32_t gw, gh; grid_dims(s, gw, gh);
    int64_t np = (int64_t)gw * gh;
    int64_t shape[3] = {(int64_t)gh, (int64_t)gw, 4};
    std::unique_ptr<futhark_f32_3d> a(new_array<futhark_f32_3d, float>(ctx, nullptr, shape, 3, cudaMemcpyDeviceToDevice));      /* freed on every early return */
    if (!a) return 1;
    if (ctx->pts_cap < np) {
        raw_free(ctx->pts_pos); raw_free(ctx->pts_dist); ctx->pts_cap = 0;
        if (!raw_alloc(ctx, ctx->pts_pos, (size_t)np) || !raw_alloc(ctx, ctx->pts_dist, (size_t)np)) return 1;
        ctx->pts_cap = np;
    }
    if (ctx->world > 1) { CU(ctx, cudaMemsetAsync(ctx->pts_pos, 0, sizeof(float4) * (size_t)np, ctx->stream)); }
    float factor = 1 / (float)spp;                                                 /* lib.fut:39 */
    uint32_t rng = s->rng;
    uint32_t passes = spp < 1 ? 1 : spp;                                           /* the first pass always runs (lib.fut:52) */
    const int S = (int)std::min<uint32_t>((uint32_t)ctx->pipeline, passes);
    if (!ensure_slots(ctx, S, np)) return 1;
    FrameParams fp;
    if (!make_frame_params(ctx, s, rng, factor, fp)) return 1;
    fp.render_mode = 1;
    CU(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int i = 0; i < S; i++) CU(ctx, cudaStreamWaitEvent(ctx->slots[i].stream, ctx->ev_fork, 0));
    cudaEvent_t prev = nullptr;
    int rc = 0;
    for (uint32_t k = 0; k < passes; k++) {
        futhark_context::PassSlot &sl = ctx->slots[k % S];
        fp.frame_rng = rng;
        if (!cu_ok(ctx, run_sample_pass(s->scene->d, fp, sl.bufs, sl.stream, &ctx->launches, nullptr, ctx->h_counts), "sample pass") ||
            (prev && !cu_ok(ctx, cudaStreamWaitEvent(sl.stream, prev, 0), "cudaStreamWaitEvent")) ||                   /* merge keeps the earlier point on ties (lib.fut:51) */
            !cu_ok(ctx, run_points_merge(fp, sl.bufs, ctx->pts_pos, ctx->pts_dist, k == 0 ? 1 : 0, sl.stream, &ctx->launches), "points merge") ||
            !cu_ok(ctx, cudaEventRecord(sl.done, sl.stream), "cudaEventRecord")) { rc = 1; break; }
        prev = sl.done;
        rng = h_advance_rng(rng);
    }
    for (int i = 0; i < S; i++) if (cudaEventRecord(ctx->slots[i].done, ctx->slots[i].stream) == cudaSuccess) cudaStreamWaitEvent(ctx->stream, ctx->slots[i].done, 0);   /* nothing stays forked */
    if (rc) { cudaStreamSynchronize(ctx->stream); return 1; }
    if (!cu_ok(ctx, run_points_export(fp, ctx->pts_pos, a->ptr(), ctx->stream, &ctx->launches), "points export")) return 1;
    futhark_opaque_state *r = clone_state(s);
    r->rng = rng;
    *out0 = r; *out1 = a.release();
    return 0;
}

/* ================================================================== extensions */
int lys_context_set_path_len(struct futhark_context *ctx, int path_len) {
    if (!ctx || path_len < 1 || path_len > LYS_MAX_PATH_LEN) { if (ctx) set_error(ctx, "path_len must be in 1..16"); return 1; }
    ctx->path_len = path_len; return 0;
}
int lys_context_set_refit_mode(struct futhark_context *ctx, int mode) { if (!ctx || mode < 0 || mode > 2) return 1; ctx->refit_mode = mode; return 0; }
int lys_context_set_partition(struct futhark_context *ctx, int rank, int world_size) {
    if (!ctx || world_size < 1 || rank < 0 || rank >= world_size) { if (ctx) set_error(ctx, "bad partition"); return 1; }
    ctx->rank = rank; ctx->world = world_size; return 0;
}
int lys_context_set_profiling(struct futhark_context *ctx, int on) { if (!ctx) return 1; ctx->timer.resolve(ctx->stream); ctx->timer.on = (on == 2) ? 2 : (on != 0); return 0; }
int lys_context_profile_get(struct futhark_context *ctx, float *ms, uint64_t *launches, int reset) {
    if (!ctx) return 1;
    ctx->timer.resolve(ctx->stream);
    for (int i = 0; i < LYS_PROFILE_CLASSES; i++) { if (ms) ms[i] = ctx->timer.ms[i]; if (launches) launches[i] = ctx->timer.n[i]; }
    if (reset) ctx->timer.reset();
    return 0;
}
int lys_context_profile_detail(struct futhark_context *ctx, float *ms36) {
    if (!ctx || !ms36) return 1;
    ctx->timer.resolve(ctx->stream);
    memcpy(ms36, ctx->timer.detail, sizeof(float) * 36);
    return 0;
}
int lys_state_advance_rng(struct futhark_context *ctx, struct futhark_opaque_state **out0, const struct futhark_opaque_state *s, uint32_t k) {
    if (!ctx || !out0 || !s) return 1;
    futhark_opaque_state *r = clone_state(s);
    for (uint32_t i = 0; i < k; i++) r->rng = h_advance_rng(r->rng);
    *out0 = r;
    return 0;
}
int lys_context_device(struct futhark_context *ctx) { return ctx ? ctx->device : -1; }
void *lys_context_stream(struct futhark_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t lys_context_launch_count(struct futhark_context *ctx) { return ctx ? ctx->launches : 0; }

void *lys_device_ptr_f32_3d(struct futhark_context *ctx, struct futhark_f32_3d *arr) { (void)ctx; return arr ? arr->mem->p : nullptr; }
void *lys_state_image_device_ptr(struct futhark_context *ctx, struct futhark_opaque_state *s, uint32_t *img_h, uint32_t *img_w) {
    (void)ctx; if (!s) return nullptr;
    if (img_h) *img_h = s->img_h;
    if (img_w) *img_w = s->img_w;
    return s->img->p;
}
struct futhark_f32_3d *lys_new_f32_3d_from_device(struct futhark_context *ctx, const void *dev, int64_t d0, int64_t d1, int64_t d2) {
    int64_t shape[3] = {d0, d1, d2};
    return new_array<futhark_f32_3d, float>(ctx, (const float *)dev, shape, 3, cudaMemcpyDeviceToDevice);
}
struct futhark_u32_1d *lys_new_u32_1d_from_device(struct futhark_context *ctx, const void *dev, int64_t d0) {
    int64_t shape[1] = {d0};
    return new_array<futhark_u32_1d, uint32_t>(ctx, (const uint32_t *)dev, shape, 1, cudaMemcpyDeviceToDevice);
}

int lys_state_info_get(struct futhark_context *ctx, const struct futhark_opaque_state *s, lys_state_info *o) {
    if (!ctx || !s || !o) return 1;
    o->dim_w = s->dim_w; o->dim_h = s->dim_h; o->subsampling = s->subsampling; o->rng = s->rng; o->img_h = s->img_h; o->img_w = s->img_w;
    o->n_frames = s->n_frames; o->cam_conf_id = s->cam_conf_id; o->mode = s->mode ? 1 : 0; o->render_mode = s->render_mode;
    o->cam_pitch = s->cam.pitch; o->cam_yaw = s->cam.yaw; o->cam_origin[0] = s->cam.origin.x; o->cam_origin[1] = s->cam.origin.y; o->cam_origin[2] = s->cam.origin.z;
    o->aperture = s->cam.conf.aperture; o->focal_dist = s->cam.conf.focal_dist;
    memcpy(o->ambience, s->ambience.k, sizeof(o->ambience));
    o->n_tris = s->scene->d.n_tris; o->n_mats = s->scene->d.n_mats; o->n_lights = s->scene->d.n_lights;
    return 0;
}
int lys_state_image(struct futhark_context *ctx, const struct futhark_opaque_state *s, float *out) {
    if (!ctx || !s || !out) return 1;
    cudaSetDevice(ctx->device);
    CU(ctx, cudaMemcpyAsync(out, s->img->p, sizeof(float) * 3 * (size_t)s->img_h * s->img_w, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
int lys_state_bvh_get(struct futhark_context *ctx, const struct futhark_opaque_state *s, float *bounds6, uint32_t *sorted_morton,
                      int32_t *sorted_src_index, int32_t *left, int32_t *right, int32_t *parent, float *node_aabb, float *leaf_aabb,
                      int32_t *node_height) {
    if (!ctx || !s) return 1;
    cudaSetDevice(ctx->device);
    const SceneDev &d = s->scene->d;
    size_t n = (size_t)d.n_tris, nn = n - 1;
    auto get = [&](void *dst, const void *src, size_t bytes) -> bool {
        return !dst || cu_ok(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpy");
    };
    std::vector<float4> nb, lb;
    if (node_aabb) nb.resize(2 * nn);
    if (leaf_aabb) lb.resize(2 * n);
    if (!get(bounds6, d.bounds, 24) || !get(sorted_morton, d.morton, 4 * n) || !get(sorted_src_index, d.sorted_idx, 4 * n) ||
        !get(left, d.left, 4 * nn) || !get(right, d.right, 4 * nn) || !get(parent, d.parent, 4 * nn) || !get(node_height, d.height, 4 * nn) ||
        !get(node_aabb ? nb.data() : nullptr, d.node_box, 32 * nn) || !get(leaf_aabb ? lb.data() : nullptr, d.leaf_box, 32 * n)) return 1;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (node_aabb) for (size_t i = 0; i < nn; i++) { float *p = node_aabb + 6 * i; p[0] = nb[2 * i].x; p[1] = nb[2 * i].y; p[2] = nb[2 * i].z; p[3] = nb[2 * i + 1].x; p[4] = nb[2 * i + 1].y; p[5] = nb[2 * i + 1].z; }
    if (leaf_aabb) for (size_t i = 0; i < n; i++) { float *p = leaf_aabb + 6 * i; p[0] = lb[2 * i].x; p[1] = lb[2 * i].y; p[2] = lb[2 * i].z; p[3] = lb[2 * i + 1].x; p[4] = lb[2 * i + 1].y; p[5] = lb[2 * i + 1].z; }
    return 0;
}
int lys_state_light_indices(struct futhark_context *ctx, const struct futhark_opaque_state *s, int32_t *src_index) {
    if (!ctx || !s || !src_index) return 1;
    cudaSetDevice(ctx->device);
    const SceneDev &d = s->scene->d;
    if (d.n_lights > 0) { CU(ctx, cudaMemcpyAsync(src_index, d.light_src, sizeof(int) * (size_t)d.n_lights, cudaMemcpyDeviceToHost, ctx->stream)); CU(ctx, cudaStreamSynchronize(ctx->stream)); }
    return 0;
}
int lys_state_bvh_rebuild_timed(struct futhark_context *ctx, const struct futhark_opaque_state *s, int reps, float *ms) {
    if (!ctx || !s || reps < 1) return 1;
    cudaSetDevice(ctx->device);
    SceneDev &d = s->scene->d;
    const int mode = s->scene->refit_mode;          /* the mode the scene was built with: a timing helper must not change results */
    if (!ensure_scratch(ctx, d.n_tris)) return 1;
    float total = 0.0f;
    for (int r = 0; r < reps; r++) {
        CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        CU(ctx, build_lbvh(d, ctx->scratch, mode, ctx->stream, &ctx->launches));
        CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        float t = 0.0f; cudaEventElapsedTime(&t, ctx->ev0, ctx->ev1); total += t;
    }
    if (mode == 0) {                                /* outside the timed region: same overflow fallback as init, so the state keeps exact boxes */
        int ovf = 0;
        CU(ctx, cudaMemcpyAsync(&ovf, ctx->scratch.crown_cnt + 63, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        if (ovf) { CU(ctx, build_lbvh(d, ctx->scratch, 2, ctx->stream, &ctx->launches)); CU(ctx, cudaStreamSynchronize(ctx->stream)); }
    }
    if (ms) *ms = total / (float)reps;
    return 0;
}

int lys_probe_primary(struct futhark_context *ctx, const struct futhark_opaque_state *s, int32_t *leaf, int32_t *src_tri, float *t) {
    if (!ctx || !s || !leaf) return 1;
    cudaSetDevice(ctx->device);
    uint32_t gw, gh; grid_dims(s, gw, gh);
    size_t np = (size_t)gw * gh;
    if (!ensure_pass_buffers(ctx, (int64_t)np, false)) return 1;
    FrameParams fp; if (!make_frame_params(ctx, s, s->rng, 1.0f, fp)) return 1;
    if (ctx->world != 1) { set_error(ctx, "probe requires world_size 1"); return 1; }
    DevRef dl = dev_alloc(ctx, 4 * np), ds = dev_alloc(ctx, 4 * np), dt = dev_alloc(ctx, 4 * np);
    if (!dl || !ds || !dt) return 1;
    CU(ctx, run_primary_probe(s->scene->d, fp, ctx->bufs, (int *)dl->p, (int *)ds->p, (float *)dt->p, ctx->stream, &ctx->launches));
    CU(ctx, cudaMemcpyAsync(leaf, dl->p, 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
    if (src_tri) CU(ctx, cudaMemcpyAsync(src_tri, ds->p, 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
    if (t) CU(ctx, cudaMemcpyAsync(t, dt->p, 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
int lys_probe_pass(struct futhark_context *ctx, const struct futhark_opaque_state *s, float *radiance, float *distance, int32_t *channel) {
    if (!ctx || !s) return 1;
    cudaSetDevice(ctx->device);
    uint32_t gw, gh; grid_dims(s, gw, gh);
    size_t np = (size_t)gw * gh;
    if (ctx->world != 1) { set_error(ctx, "probe requires world_size 1"); return 1; }
    if (!ensure_pass_buffers(ctx, (int64_t)np, true)) return 1;
    FrameParams fp; if (!make_frame_params(ctx, s, s->rng, 1.0f, fp)) return 1;
    PassBuffers b = ctx->bufs;
    CU(ctx, run_sample_pass(s->scene->d, fp, b, ctx->stream, &ctx->launches));
    if (radiance) CU(ctx, cudaMemcpyAsync(radiance, b.probe_rad, 4 * 16 * np, cudaMemcpyDeviceToHost, ctx->stream));
    if (distance) CU(ctx, cudaMemcpyAsync(distance, b.probe_dist, 4 * 16 * np, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint8_t> ch;
    if (channel) { ch.resize(np); CU(ctx, cudaMemcpyAsync(ch.data(), b.chan, np, cudaMemcpyDeviceToHost, ctx->stream)); }
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (channel) for (size_t i = 0; i < np; i++) { int col, rl; path_tile((int)gw, (int)gh, (int)i, col, rl); channel[(size_t)rl * gw + col] = ch[i]; }   /* path order -> pixel order */
    /* probes off again for the timed paths */
    raw_free(ctx->bufs.probe_rad); raw_free(ctx->bufs.probe_dist);
    return 0;
}
static int trace_common(struct futhark_context *ctx, const struct futhark_opaque_state *s, const float *rays, const float *tmax, int64_t n,
                        int32_t *out_i, float *out_t, int any) {
    if (!ctx || !s || !rays || !out_i || n < 0) return 1;
    cudaSetDevice(ctx->device);
    if (n == 0) return 0;
    DevRef dr = dev_alloc(ctx, 24 * (size_t)n), dm = dev_alloc(ctx, 4 * (size_t)n), di = dev_alloc(ctx, 4 * (size_t)n), dt = dev_alloc(ctx, 4 * (size_t)n);
    if (!dr || !dm || !di || !dt) return 1;
    CU(ctx, cudaMemcpyAsync(dr->p, rays, 24 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (tmax) CU(ctx, cudaMemcpyAsync(dm->p, tmax, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, run_trace_rays(s->scene->d, (const float *)dr->p, (const float *)dm->p, n, (int *)di->p, (float *)dt->p, any, ctx->stream, &ctx->launches));
    CU(ctx, cudaMemcpyAsync(out_i, di->p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_t) CU(ctx, cudaMemcpyAsync(out_t, dt->p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
int lys_trace_closest(struct futhark_context *ctx, const struct futhark_opaque_state *s, const float *rays, int64_t n, int32_t *leaf, float *t) {
    return trace_common(ctx, s, rays, nullptr, n, leaf, t, 0);
}
int lys_trace_any(struct futhark_context *ctx, const struct futhark_opaque_state *s, const float *rays, const float *tmax, int64_t n, int32_t *hit) {
    if (!tmax) return 1;
    return trace_common(ctx, s, rays, tmax, n, hit, nullptr, 1);
}
int lys_eval_math(struct futhark_context *ctx, int fn, const float *in, float *out, int64_t n) {
    if (!ctx || !in || !out || n < 0) return 1;
    cudaSetDevice(ctx->device);
    if (n == 0) return 0;
    DevRef di = dev_alloc(ctx, 4 * (size_t)n), dout = dev_alloc(ctx, 4 * (size_t)n);
    if (!di || !dout) return 1;
    CU(ctx, cudaMemcpyAsync(di->p, in, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, run_eval_math(fn, (const float *)di->p, (float *)dout->p, n, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out, dout->p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}
int lys_material_probe(struct futhark_context *ctx, const float *mat28, float wavelen, const float *wo, const float *wi, const float *normal,
                       uint32_t rng, float *out9) {
    if (!ctx || !mat28 || !wo || !wi || !normal || !out9) return 1;
    cudaSetDevice(ctx->device);
    DevRef dm = dev_alloc(ctx, 28 * 4), dout = dev_alloc(ctx, 9 * 4);
    if (!dm || !dout) return 1;
    CU(ctx, cudaMemcpyAsync(dm->p, mat28, 28 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, run_material_probe((const float *)dm->p, wavelen, v3(wo[0], wo[1], wo[2]), v3(wi[0], wi[1], wi[2]), v3(normal[0], normal[1], normal[2]), rng,
                               (float *)dout->p, ctx->stream));
    CU(ctx, cudaMemcpyAsync(out9, dout->p, 9 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

} // extern "C"
