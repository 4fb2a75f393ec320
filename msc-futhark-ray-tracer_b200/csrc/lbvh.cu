/* lbvh.cu -- LBVH build on sm_100a.
 *
 * Replaces `obj_bvh.build` (reference src/bvh.fut:86-121) and `radix_tree.mk`
 * (src/radix_tree.fut:21-89):
 *   k_tri_boxes      bvh.fut:87          per-triangle AABB (center, half) + exact corner union per 256-triangle chunk
 *   k_bounds_fold    bvh.fut:88-90       scene bounds = LEFT FOLD of containing_aabb, reproduced
 *                                        exactly by a speculative block-skip fold (see below)
 *   k_morton         bvh.fut:91-94       30-bit Morton code of the normalised box centre
 *   k_hist/k_onesweep bvh.fut:95-97      stable LSD radix sort of (key, index): onesweep, 8-bit digits,
 *                                        decoupled look-back, 4 passes
 *   k_gather_leaves  bvh.fut:95 (unzip3) sorted triangles / boxes
 *   k_karras         radix_tree.fut:31-88 internal nodes + parent pointers
 *   k_refit          bvh.fut:105-120     converged boxes F and node heights, atomic bottom-up
 *   k_crown_fixup    bvh.fut:109,118-120 the reference stops after floor(log2 n)+2 Jacobi sweeps from
 *                                        zero boxes; nodes higher than that keep truncated boxes,
 *                                        recomputed here exactly
 *   k_pack_nodes                         traversal layout: 2 x float4 per node (min|left, max|right)
 */
#include "lys_scene.h"
#include "lys_device.cuh"
#include <cstdio>

namespace lys {

/* ------------------------------------------------------------------ boxes + per-chunk exact unions */
#define BOX_CHUNK 256        /* triangles per chunk == threads per block of k_tri_boxes */
__device__ __forceinline__ float warp_min(float v) { for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
__device__ __forceinline__ float warp_max(float v) { for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
/* One triangle per thread: its AABB (bvh.fut:87), and per block the exact min/max over the block of the
 * corners (center - half, center + half) exactly as containing_aabb derives them (shapes.fut:97-98);
 * a NaN corner poisons the chunk union so that the fold never skips the chunk. */
__global__ void __launch_bounds__(BOX_CHUNK) k_tri_boxes(const float *__restrict__ tris, int n, float4 *__restrict__ box_c,
                                                         float4 *__restrict__ box_h, float4 *__restrict__ chunk_lo, float4 *__restrict__ chunk_hi) {
    __shared__ float red[6][BOX_CHUNK / 32];
    __shared__ int poison;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) poison = 0;
    __syncthreads();
    float lo[3] = {LYS_INF, LYS_INF, LYS_INF}, hi[3] = {-LYS_INF, -LYS_INF, -LYS_INF};
    if (i < n) {
        const float *t = tris + 9ll * i;
        Box b = triangle_box(v3(t[0], t[1], t[2]), v3(t[3], t[4], t[5]), v3(t[6], t[7], t[8]));
        box_c[i] = make_float4(b.c.x, b.c.y, b.c.z, 0.0f);
        box_h[i] = make_float4(b.h.x, b.h.y, b.h.z, 0.0f);
        V3 l = b.c - b.h, h = b.c + b.h;
        lo[0] = l.x; lo[1] = l.y; lo[2] = l.z; hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
        if (l.x != l.x || l.y != l.y || l.z != l.z || h.x != h.x || h.y != h.y || h.z != h.z) poison = 1;
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float a = warp_min(lo[k]), b = warp_max(hi[k]);
        if (lane == 0) { red[k][warp] = a; red[3 + k][warp] = b; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float r[6];
        for (int k = 0; k < 3; k++) {
            float a = red[k][0], b = red[3 + k][0];
            for (int w = 1; w < BOX_CHUNK / 32; w++) { a = fminf(a, red[k][w]); b = fmaxf(b, red[3 + k][w]); }
            r[k] = a; r[3 + k] = b;
        }
        if (poison) { float q = lys_u2f(0x7fc00000u); for (int k = 0; k < 6; k++) r[k] = q; }
        chunk_lo[blockIdx.x] = make_float4(r[0], r[1], r[2], 0.0f);
        chunk_hi[blockIdx.x] = make_float4(r[3], r[4], r[5], 0.0f);
    }
}

/* ------------------------------------------------------------------ exact left fold of the scene bounds
 * S_k = containing_aabb(S_{k-1}, box_k) (bvh.fut:88-90) is not associative in f32: center/half are
 * re-derived from the corners at every step, so a tree reduction does not reproduce the reference's
 * sequential (`c` backend) result.  It is emulated exactly:
 *   - an element leaves S unchanged iff contain(S, box) == S bit-for-bit;
 *   - if S is a fixed point of the re-derivation (contain(S, S) == S) and a whole chunk's exact corner
 *     union lies inside S's corners, every element of the chunk leaves S unchanged -> the chunk is
 *     skipped after a 6-comparison test on its precomputed union (32 chunks per warp step);
 *   - any other chunk is staged in shared memory and folded by warp 0 in order, 32 elements per step:
 *     the lanes test their element against the current S, the first lane whose result differs
 *     commits it, and the step restarts behind that element.
 * State changes are rare (a few hundred per million triangles), so the walk is O(#chunks/32) warp
 * steps plus the dirty chunks.  The result equals the sequential fold for any input. */
#define FOLD_THREADS BOX_CHUNK
__device__ __forceinline__ Box shfl_box(Box b, int src) {
    Box r;
    r.c.x = __shfl_sync(0xffffffffu, b.c.x, src); r.c.y = __shfl_sync(0xffffffffu, b.c.y, src); r.c.z = __shfl_sync(0xffffffffu, b.c.z, src);
    r.h.x = __shfl_sync(0xffffffffu, b.h.x, src); r.h.y = __shfl_sync(0xffffffffu, b.h.y, src); r.h.z = __shfl_sync(0xffffffffu, b.h.z, src);
    return r;
}
__global__ void __launch_bounds__(FOLD_THREADS, 1)
k_bounds_fold(const float4 *__restrict__ box_c, const float4 *__restrict__ box_h, const float4 *__restrict__ chunk_lo,
              const float4 *__restrict__ chunk_hi, int n, float *__restrict__ bounds_out /* 6 */) {
    __shared__ float4 sc[BOX_CHUNK], sh[BOX_CHUNK];
    __shared__ Box S_sh;
    __shared__ int next_chunk;       /* first chunk (>= cursor) that must be folded element-wise, or n_chunks */
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_chunks = (n + BOX_CHUNK - 1) / BOX_CHUNK;
    Box S; S.c = v3(0.0f, 0.0f, 0.0f); S.h = v3(-LYS_INF, -LYS_INF, -LYS_INF);            /* bvh.fut:88-89 */
    int cursor = 0;
    while (cursor < n_chunks) {
        /* warp 0: skip clean chunks, 32 at a time (all threads hold the same S) */
        if (warp == 0) {
            bool stable = box_bits_equal(contain(S, S), S);
            V3 slo = S.c - S.h, shi = S.c + S.h;
            int c = cursor;
            if (stable) {
                while (c < n_chunks) {
                    int j = c + lane;
                    bool dirty = false;
                    if (j < n_chunks) {
                        float4 l = __ldg(chunk_lo + j), h = __ldg(chunk_hi + j);
                        dirty = !(l.x >= slo.x && l.y >= slo.y && l.z >= slo.z && h.x <= shi.x && h.y <= shi.y && h.z <= shi.z);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, dirty);
                    if (m) { c += __ffs(m) - 1; break; }
                    c += 32;
                }
                if (c > n_chunks) c = n_chunks;
            }
            if (lane == 0) next_chunk = c;
        }
        __syncthreads();
        cursor = next_chunk;
        if (cursor >= n_chunks) break;
        /* stage the dirty chunk */
        const int base = cursor * BOX_CHUNK;
        const int limit = min(BOX_CHUNK, n - base);
        if (tid < limit) { sc[tid] = box_c[base + tid]; sh[tid] = box_h[base + tid]; }
        __syncthreads();
        if (warp == 0) {
            Box s = S;
            int p = 0;
            while (p < limit) {
                int e = p + lane;
                Box ns = s; bool ch = false;
                if (e < limit) {
                    float4 c4 = sc[e], h4 = sh[e];
                    Box b; b.c = v3(c4.x, c4.y, c4.z); b.h = v3(h4.x, h4.y, h4.z);
                    ns = contain(s, b);
                    ch = !box_bits_equal(ns, s);
                }
                unsigned m = __ballot_sync(0xffffffffu, ch);
                if (!m) { p += 32; continue; }
                int k = __ffs(m) - 1;
                s = shfl_box(ns, k);
                p += k + 1;
            }
            if (lane == 0) S_sh = s;
        }
        __syncthreads();
        S = S_sh;
        cursor++;
        __syncthreads();
    }
    if (tid == 0) {
        bounds_out[0] = S.c.x; bounds_out[1] = S.c.y; bounds_out[2] = S.c.z;
        bounds_out[3] = S.h.x; bounds_out[4] = S.h.y; bounds_out[5] = S.h.z;
    }
}

/* ------------------------------------------------------------------ Morton keys */
__global__ void k_morton(const float4 *__restrict__ box_c, int n, const float *__restrict__ bounds,
                         uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 bc = v3(bounds[0], bounds[1], bounds[2]), bh = v3(bounds[3], bounds[4], bounds[5]);
    V3 bmin = bc - bh;                      /* aabb_min_corner shapes.fut:88-89 */
    V3 bdim = 2.0f * bh;                    /* aabb_dimensions shapes.fut:94 */
    float4 c = box_c[i];
    V3 p = v3(c.x, c.y, c.z) - bmin;        /* normalise_position bvh.fut:91-92 */
    V3 q = v3(p.x / bdim.x, p.y / bdim.y, p.z / bdim.z);
    keys[i] = morton30(q);
    vals[i] = (uint32_t)i;
}

/* ------------------------------------------------------------------ onesweep radix sort (stable, LSD, 8-bit digits) */
#define RS_RADIX 256
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)      /* 4096 keys per tile */
#define RS_WARPS (RS_THREADS / 32)
#define RS_FLAG_LOCAL 0x40000000u
#define RS_FLAG_INCL 0x80000000u
#define RS_VALUE_MASK 0x3fffffffu

/* digit histograms of all four passes in one read of the keys */
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t *__restrict__ keys, int n, uint32_t *__restrict__ ghist /* [4][256] */) {
    __shared__ uint32_t h[4][RS_RADIX];
    for (int i = threadIdx.x; i < 4 * RS_RADIX; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t k = keys[i];
        atomicAdd(&h[0][k & 255u], 1u); atomicAdd(&h[1][(k >> 8) & 255u], 1u);
        atomicAdd(&h[2][(k >> 16) & 255u], 1u); atomicAdd(&h[3][k >> 24], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * RS_RADIX; i += blockDim.x) {
        uint32_t v = (&h[0][0])[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}
/* exclusive scan of each pass's 256 digit counts -> global digit base offsets */
__global__ void k_rs_scan(uint32_t *__restrict__ ghist) {
    __shared__ uint32_t s[RS_RADIX];
    int p = blockIdx.x, d = threadIdx.x;
    uint32_t v = ghist[p * RS_RADIX + d];
    s[d] = v; __syncthreads();
    for (int off = 1; off < RS_RADIX; off <<= 1) {
        uint32_t t = (d >= off) ? s[d - off] : 0u; __syncthreads();
        s[d] += t; __syncthreads();
    }
    ghist[p * RS_RADIX + d] = s[d] - v;
}
/* One digit pass.  Tiles are claimed through an atomic ticket so that every tile a look-back waits on
 * has already started (forward progress).  status[tile][digit]: 2 flag bits + 30-bit count. */
__global__ void __launch_bounds__(RS_THREADS)
k_rs_onesweep(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
              uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int n, int shift,
              const uint32_t *__restrict__ gbase /* [256] */, uint32_t *status /* [tiles][256] */, uint32_t *ticket) {
    __shared__ uint32_t wcnt[RS_WARPS][RS_RADIX];       /* per-warp digit counts -> per-warp digit bases */
    __shared__ uint32_t tile_off[RS_RADIX];             /* exclusive offset of a digit inside the sorted tile */
    __shared__ uint32_t out_base[RS_RADIX];             /* global position of the tile's first element of a digit */
    __shared__ uint32_t skeys[RS_TILE], svals[RS_TILE];
    __shared__ uint32_t tile_id_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) tile_id_s = atomicAdd(ticket, 1u);
    for (int i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = tile_id_s;
    const long long tile_base = (long long)tile * RS_TILE;

    /* 1. load (warp w owns elements [w*ITEMS*32, (w+1)*ITEMS*32) of the tile, round r = 32 consecutive keys) and rank */
    uint32_t key[RS_ITEMS], val[RS_ITEMS]; uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        long long g = tile_base + warp * (RS_ITEMS * 32) + r * 32 + lane;
        bool ok = g < n;
        key[r] = ok ? keys_in[g] : 0xffffffffu;
        val[r] = ok ? vals_in[g] : 0u;
        uint32_t d = ok ? ((key[r] >> shift) & 255u) : 0xffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t below = __popc(peers & ((1u << lane) - 1u));
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && lane == leader) { old = wcnt[warp][d]; wcnt[warp][d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + below;                          /* stable rank among this warp's keys of digit d */
        __syncwarp();
    }
    __syncthreads();
    /* 2. per digit: scan over warps, tile count; publish the local count, then look back */
    {
        const int d = tid;                              /* RS_THREADS == RS_RADIX */
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { uint32_t c = wcnt[w][d]; wcnt[w][d] = run; run += c; }
        const uint32_t count = run;
        volatile uint32_t *st = status;
        st[(size_t)tile * RS_RADIX + d] = RS_FLAG_LOCAL | count;
        /* exclusive scan of the tile's digit counts (for the shared-memory reorder) */
        tile_off[d] = count; __syncthreads();
        for (int off = 1; off < RS_RADIX; off <<= 1) {
            uint32_t t = (d >= off) ? tile_off[d - off] : 0u; __syncthreads();
            tile_off[d] += t; __syncthreads();
        }
        uint32_t incl_off = tile_off[d]; __syncthreads();
        tile_off[d] = incl_off - count;
        /* decoupled look-back */
        uint32_t excl = 0;
        for (long long t = (long long)tile - 1; t >= 0; t--) {
            uint32_t s;
            do { s = st[(size_t)t * RS_RADIX + d]; } while ((s & (RS_FLAG_LOCAL | RS_FLAG_INCL)) == 0u);
            excl += s & RS_VALUE_MASK;
            if (s & RS_FLAG_INCL) break;
        }
        st[(size_t)tile * RS_RADIX + d] = RS_FLAG_INCL | (excl + count);
        out_base[d] = gbase[d] + excl;
    }
    __syncthreads();
    /* 3. reorder through shared memory, then write runs of equal digits */
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        long long g = tile_base + warp * (RS_ITEMS * 32) + r * 32 + lane;
        if (g < n) {
            uint32_t d = (key[r] >> shift) & 255u;
            uint32_t p = tile_off[d] + wcnt[warp][d] + rank[r];
            skeys[p] = key[r]; svals[p] = val[r];
        }
    }
    __syncthreads();
    const int tile_n = (int)min((long long)RS_TILE, (long long)n - tile_base);
    for (int p = tid; p < tile_n; p += RS_THREADS) {
        uint32_t k = skeys[p];
        uint32_t d = (k >> shift) & 255u;
        uint32_t dst = out_base[d] + ((uint32_t)p - tile_off[d]);
        keys_out[dst] = k; vals_out[dst] = svals[p];
    }
}

/* ------------------------------------------------------------------ sorted leaves */
__global__ void k_gather_leaves(const float *__restrict__ tris, const uint32_t *__restrict__ tri_mats,
                                const float4 *__restrict__ box_c, const float4 *__restrict__ box_h,
                                const uint32_t *__restrict__ sorted_idx, int n,
                                float4 *__restrict__ leaf_tri /* [n][3] */, float4 *__restrict__ leaf_box /* [n][2] */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = sorted_idx[i];
    const float *t = tris + 9ll * s;
    V3 a = v3(t[0], t[1], t[2]), b = v3(t[3], t[4], t[5]), c = v3(t[6], t[7], t[8]);
    V3 e1 = b - a, e2 = c - a;                                          /* shapes.fut:69-70 */
    leaf_tri[3ll * i + 0] = make_float4(a.x, a.y, a.z, __uint_as_float(tri_mats[s]));
    leaf_tri[3ll * i + 1] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(s));
    leaf_tri[3ll * i + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
    leaf_box[2ll * i + 0] = box_c[s];
    leaf_box[2ll * i + 1] = box_h[s];
}

/* ------------------------------------------------------------------ Karras radix tree */
__device__ __forceinline__ int karras_delta(const uint32_t *__restrict__ L, int n, int i, int j) {  /* radix_tree.fut:22-29 */
    if (j < 0 || j >= n) return -1;
    uint32_t a = L[i], b = L[j];
    return (a == b) ? 32 + __clz((uint32_t)i ^ (uint32_t)j) : __clz(a ^ b);
}
__global__ void k_karras(const uint32_t *__restrict__ L, int n, int *__restrict__ left, int *__restrict__ right,
                         int *__restrict__ parent, int *__restrict__ leaf_parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int dd = karras_delta(L, n, i, i + 1) - karras_delta(L, n, i, i - 1);
    int d = (dd > 0) - (dd < 0);
    int dmin = karras_delta(L, n, i, i - d);
    int lmax = 2;
    while (karras_delta(L, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (karras_delta(L, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = karras_delta(L, n, i, j);
    int s = 0;
    for (int q = 1; q <= l; q *= 2) {
        int t = (l + q * 2 - 1) / (q * 2);
        if (karras_delta(L, n, i, i + (s + t) * d) > dnode) s += t;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lp, rp;
    if (lo == gamma) { lp = ~gamma; leaf_parent[gamma] = i; } else { lp = gamma; parent[gamma] = i; }
    if (hi == gamma + 1) { rp = ~(gamma + 1); leaf_parent[gamma + 1] = i; } else { rp = gamma + 1; parent[gamma + 1] = i; }
    left[i] = lp; right[i] = rp;
    if (i == 0) parent[0] = -1;
}

/* ------------------------------------------------------------------ bottom-up refit: converged boxes F and heights */
__device__ __forceinline__ Box load_box_cg(const float4 *p) {
    float4 c = __ldcg(p), h = __ldcg(p + 1);
    Box b; b.c = v3(c.x, c.y, c.z); b.h = v3(h.x, h.y, h.z); return b;
}
__global__ void k_refit(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ parent,
                        const int *__restrict__ leaf_parent, const float4 *leaf_box, float4 *node_box /* [n-1][2] F */,
                        int *height, unsigned int *visits, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cur = leaf_parent[i];
    while (cur >= 0) {
        unsigned int old = atomicAdd(&visits[cur], 1u);
        if (old == 0u) return;                       /* the sibling subtree is not finished yet */
        __threadfence();
        int l = left[cur], r = right[cur];
        Box bl = (l < 0) ? load_box_cg(leaf_box + 2ll * (~l)) : load_box_cg(node_box + 2ll * l);
        Box br = (r < 0) ? load_box_cg(leaf_box + 2ll * (~r)) : load_box_cg(node_box + 2ll * r);
        int hl = (l < 0) ? 0 : __ldcg(height + l), hr = (r < 0) ? 0 : __ldcg(height + r);
        Box b = contain(bl, br);                     /* bvh.fut:114-117: always (left, right) */
        node_box[2ll * cur + 0] = make_float4(b.c.x, b.c.y, b.c.z, 0.0f);
        node_box[2ll * cur + 1] = make_float4(b.h.x, b.h.y, b.h.z, 0.0f);
        height[cur] = 1 + max(hl, hr);
        __threadfence();
        cur = parent[cur];
    }
}

/* ------------------------------------------------------------------ truncated-Jacobi fix-up
 * After `depth` sweeps from zero boxes a node of height <= depth holds its converged box F; a taller
 * node holds A_depth(v) with A_k(v) = contain(A_{k-1}(left), A_{k-1}(right)), A_0 = {0,0}, leaves exact. */
struct CrownFrame { int node, k, stage; Box lbox; };
__device__ Box crown_eval(int root, int depth, const int *left, const int *right, const float4 *leaf_box,
                          const float4 *F, const int *height) {
    CrownFrame st[40];                 /* depth <= 33 for n < 2^31 */
    int sp = 0;
    st[0].node = root; st[0].k = depth; st[0].stage = 0;
    Box ret; ret.c = v3(0.0f, 0.0f, 0.0f); ret.h = v3(0.0f, 0.0f, 0.0f);
    bool have_ret = false;
    while (true) {
        CrownFrame &f = st[sp];
        if (have_ret) {
            have_ret = false;
            if (f.stage == 0) { f.lbox = ret; f.stage = 1; }
            else {
                ret = contain(f.lbox, ret);            /* bvh.fut:115-116: (left, right) */
                if (sp == 0) return ret;
                sp--; have_ret = true;
                continue;
            }
        }
        int child = (f.stage == 0) ? left[f.node] : right[f.node];
        int ck = f.k - 1;
        if (child < 0) {
            float4 c = leaf_box[2ll * (~child)], h = leaf_box[2ll * (~child) + 1];
            ret.c = v3(c.x, c.y, c.z); ret.h = v3(h.x, h.y, h.z); have_ret = true;
        } else if (height[child] <= ck) {
            float4 c = F[2ll * child], h = F[2ll * child + 1];
            ret.c = v3(c.x, c.y, c.z); ret.h = v3(h.x, h.y, h.z); have_ret = true;
        } else if (ck == 0) {
            ret.c = v3(0.0f, 0.0f, 0.0f); ret.h = v3(0.0f, 0.0f, 0.0f); have_ret = true;   /* bvh.fut:105-107 */
        } else {
            sp++;
            st[sp].node = child; st[sp].k = ck; st[sp].stage = 0;
        }
    }
}
__global__ void k_crown_fixup(const int *__restrict__ left, const int *__restrict__ right, const float4 *__restrict__ leaf_box,
                              const float4 *__restrict__ F, const int *__restrict__ height, int n_nodes, int depth,
                              float4 *__restrict__ A /* [n-1][2] */, int converged) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    float4 c = F[2ll * i], h = F[2ll * i + 1];
    if (!converged && height[i] > depth) {
        Box b = crown_eval(i, depth, left, right, leaf_box, F, height);
        c = make_float4(b.c.x, b.c.y, b.c.z, 0.0f); h = make_float4(b.h.x, b.h.y, b.h.z, 0.0f);
    }
    A[2ll * i] = c; A[2ll * i + 1] = h;
}

/* traversal layout: node i -> (min.xyz | left), (max.xyz | right); min/max as hit_aabb derives them (shapes.fut:120) */
__global__ void k_pack_nodes(const float4 *__restrict__ A, const int *__restrict__ left, const int *__restrict__ right,
                             int n_nodes, float4 *__restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    float4 c = A[2ll * i], h = A[2ll * i + 1];
    V3 mn = v3(c.x, c.y, c.z) - v3(h.x, h.y, h.z), mx = v3(c.x, c.y, c.z) + v3(h.x, h.y, h.z);
    nodes[2ll * i + 0] = make_float4(mn.x, mn.y, mn.z, __int_as_float(left[i]));
    nodes[2ll * i + 1] = make_float4(mx.x, mx.y, mx.z, __int_as_float(right[i]));
}

/* ------------------------------------------------------------------ host driver */
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

cudaError_t build_lbvh(SceneDev &sc, BuildScratch &ws, int refit_mode, cudaStream_t stream, uint64_t *launches) {
    const int n = (int)sc.n_tris;
    const int n_nodes = n - 1;
    const int T = 256;
    uint64_t nl = 0;
    k_tri_boxes<<<cdiv(n, BOX_CHUNK), BOX_CHUNK, 0, stream>>>(sc.tris, n, ws.box_c, ws.box_h, ws.chunk_lo, ws.chunk_hi); nl++;
    k_bounds_fold<<<1, FOLD_THREADS, 0, stream>>>(ws.box_c, ws.box_h, ws.chunk_lo, ws.chunk_hi, n, sc.bounds); nl++;
    k_morton<<<cdiv(n, T), T, 0, stream>>>(ws.box_c, n, sc.bounds, ws.keys[0], ws.vals[0]); nl++;
    /* radix sort */
    const int tiles = cdiv(n, RS_TILE);
    cudaMemsetAsync(ws.rs_hist, 0, 4 * RS_RADIX * sizeof(uint32_t), stream);
    cudaMemsetAsync(ws.rs_status, 0, (size_t)4 * tiles * RS_RADIX * sizeof(uint32_t) + 4 * sizeof(uint32_t), stream);
    k_rs_hist<<<min(cdiv(n, RS_THREADS * 8), 148 * 8), RS_THREADS, 0, stream>>>(ws.keys[0], n, ws.rs_hist); nl++;
    k_rs_scan<<<4, RS_RADIX, 0, stream>>>(ws.rs_hist); nl++;
    int cur = 0;
    for (int p = 0; p < 4; p++) {
        uint32_t *status = ws.rs_status + (size_t)p * tiles * RS_RADIX;
        uint32_t *ticket = ws.rs_status + (size_t)4 * tiles * RS_RADIX + p;
        k_rs_onesweep<<<tiles, RS_THREADS, 0, stream>>>(ws.keys[cur], ws.vals[cur], ws.keys[cur ^ 1], ws.vals[cur ^ 1], n, 8 * p,
                                                        ws.rs_hist + p * RS_RADIX, status, ticket); nl++;
        cur ^= 1;
    }
    /* 4 passes -> result back in buffer 0 */
    cudaMemcpyAsync(sc.morton, ws.keys[cur], (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(sc.sorted_idx, ws.vals[cur], (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    k_gather_leaves<<<cdiv(n, T), T, 0, stream>>>(sc.tris, sc.tri_mats, ws.box_c, ws.box_h, sc.sorted_idx, n, sc.leaf_tri, sc.leaf_box); nl++;
    k_karras<<<cdiv(n_nodes, T), T, 0, stream>>>(sc.morton, n, sc.left, sc.right, sc.parent, ws.leaf_parent); nl++;
    cudaMemsetAsync(ws.visits, 0, (size_t)n_nodes * sizeof(unsigned int), stream);
    k_refit<<<cdiv(n, T), T, 0, stream>>>(sc.left, sc.right, sc.parent, ws.leaf_parent, sc.leaf_box, ws.F, sc.height, ws.visits, n); nl++;
    int depth = (int)(log2f((float)n)) + 2;                                 /* bvh.fut:109 */
    k_crown_fixup<<<cdiv(n_nodes, T), T, 0, stream>>>(sc.left, sc.right, sc.leaf_box, ws.F, sc.height, n_nodes, depth, sc.node_box, refit_mode); nl++;
    k_pack_nodes<<<cdiv(n_nodes, T), T, 0, stream>>>(sc.node_box, sc.left, sc.right, n_nodes, sc.nodes); nl++;
    if (launches) *launches += nl;
    return cudaGetLastError();
}

} // namespace lys
