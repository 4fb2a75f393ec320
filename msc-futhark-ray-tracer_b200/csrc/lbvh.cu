/* lbvh.cu -- LBVH build on sm_100a.
 *
 * Replaces `obj_bvh.build` (reference src/bvh.fut:86-121) and `radix_tree.mk`
 * (src/radix_tree.fut:21-89):
 *   k_tri_boxes      bvh.fut:87          per-triangle AABB (center, half) + exact corner union per 256-triangle chunk
 *   k_bounds_fold    bvh.fut:88-90       scene bounds = LEFT FOLD of containing_aabb, reproduced
 *                                        exactly by a per-axis speculative chunk-skipping fold (see below)
 *   k_morton         bvh.fut:91-94       30-bit Morton code of the normalised box centre
 *   k_hist/k_onesweep bvh.fut:95-97      stable LSD radix sort of (key, index): onesweep, 8-bit digits,
 *                                        decoupled look-back, 4 passes
 *   k_gather_leaves  bvh.fut:95 (unzip3) sorted triangles / boxes
 *   k_karras         radix_tree.fut:31-88 internal nodes + parent pointers
 *   k_refit          bvh.fut:105-120     converged boxes F and node heights, atomic bottom-up
 *   k_crown_*        bvh.fut:109,118-120 the reference stops after floor(log2 n)+2 Jacobi sweeps from
 *                                        zero boxes; nodes higher than that keep truncated boxes,
 *                                        recomputed here exactly through level-synchronous worklists
 *   k_pack_records                       traversal layout: 2 x float4 per node (its box, left child, escape link), per octant
 */
#include "lys_scene.h"
#include "lys_device.cuh"
#include <cstdio>
#include <cstdlib>

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
#define FOLD_THREADS 1024
#define FOLD_GRANULE 1024      /* elements tested per dirty step = 4 chunks */
#define FOLD_MAX_SUPER 1024    /* super-chunk (32 chunks) unions kept in shared memory */
/* containing_aabb (shapes.fut:96-101) acts on each axis independently, so the fold is three independent scalar
 * automata: CTA `axis` folds (center[axis], half[axis]).  Splitting them also splits the state changes
 * (185 / 2 / 151 on the 1M-triangle scene instead of 335 jointly), and the three CTAs run concurrently. */
struct Box1 { float c, h; };
__device__ __forceinline__ Box1 contain1(Box1 a, Box1 b) {
    float mn = lys_fminf(a.c - a.h, b.c - b.h);
    float mx = lys_fmaxf(a.c + a.h, b.c + b.h);
    Box1 r; r.c = 0.5f * (mn + mx); r.h = mx - r.c; return r;
}
__device__ __forceinline__ bool box1_equal(Box1 a, Box1 b) { return lys_f2u(a.c) == lys_f2u(b.c) && lys_f2u(a.h) == lys_f2u(b.h); }
__device__ __forceinline__ float f4_axis(const float4 &v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }
/* dynamic shared memory: the first `n_chunk_sm` chunk unions of this axis (lo then hi) */
__global__ void __launch_bounds__(FOLD_THREADS, 1)
k_bounds_fold(const float4 *__restrict__ box_c, const float4 *__restrict__ box_h, const float4 *__restrict__ chunk_lo,
              const float4 *__restrict__ chunk_hi, int n, int n_chunk_sm, float *__restrict__ bounds_out /* 6 */) {
    extern __shared__ float chunk_sm[];
    __shared__ float sup_lo[FOLD_MAX_SUPER], sup_hi[FOLD_MAX_SUPER];
    __shared__ Box1 S_sh;
    __shared__ int next_chunk, first_changed[2];
    const int axis = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_chunks = (n + BOX_CHUNK - 1) / BOX_CHUNK;
    const int n_super = (n_chunks + 31) / 32;
    const int n_super_sm = min(n_super, FOLD_MAX_SUPER);
    float *clo_sm = chunk_sm, *chi_sm = chunk_sm + n_chunk_sm;
    for (int j = tid; j < n_chunk_sm; j += FOLD_THREADS) { clo_sm[j] = f4_axis(__ldg(chunk_lo + j), axis); chi_sm[j] = f4_axis(__ldg(chunk_hi + j), axis); }
    if (tid < 2) first_changed[tid] = 0x7fffffff;
    __syncthreads();
    /* prologue: unions of 32 chunk unions (8192 triangles); NaN poisons */
    for (int sidx = warp; sidx < n_super_sm; sidx += FOLD_THREADS / 32) {
        int j = sidx * 32 + lane;
        float l = LYS_INF, h = -LYS_INF;
        if (j < n_chunks) {
            if (j < n_chunk_sm) { l = clo_sm[j]; h = chi_sm[j]; } else { l = f4_axis(__ldg(chunk_lo + j), axis); h = f4_axis(__ldg(chunk_hi + j), axis); }
        }
        bool bad = (l != l) || (h != h);
        l = warp_min(l); h = warp_max(h);
        if (__ballot_sync(0xffffffffu, bad)) { l = lys_u2f(0x7fc00000u); h = l; }
        if (lane == 0) { sup_lo[sidx] = l; sup_hi[sidx] = h; }
    }
    Box1 S; S.c = 0.0f; S.h = -LYS_INF;                                                     /* bvh.fut:88-89 */
    __syncthreads();
    int cursor = 0, round = 0;
    int pre_base = -1; float pre_c = 0.0f, pre_h = 0.0f;                                    /* prefetched element of the next granule */
    while (cursor < n_chunks) {
        /* warp 0: find the first chunk >= cursor that S does not provably absorb */
        if (warp == 0) {
            int c = cursor;
            if (box1_equal(contain1(S, S), S)) {
                const float slo = S.c - S.h, shi = S.c + S.h;
                while (c < n_chunks) {
                    if ((c & 31) == 0 && (c >> 5) < n_super_sm) {          /* aligned + indexed: try whole super-chunks first */
                        int s0 = c >> 5;
                        bool found = false;
                        while (s0 < n_super_sm) {
                            int sj = s0 + lane;
                            bool dirty = (sj < n_super_sm) && !(sup_lo[sj] >= slo && sup_hi[sj] <= shi);
                            unsigned m = __ballot_sync(0xffffffffu, dirty);
                            if (m) { s0 += __ffs(m) - 1; found = true; break; }
                            s0 += 32;
                        }
                        c = min(s0, n_super_sm) << 5;
                        if (!found && n_super_sm == n_super) { c = n_chunks; break; }
                        if (c >= n_chunks) { c = n_chunks; break; }
                    }
                    /* chunk level inside the current (dirty or unindexed) super-chunk */
                    int j = (c & ~31) + lane;
                    bool dirty = false;
                    if (j >= c && j < n_chunks) {
                        float l, h;
                        if (j < n_chunk_sm) { l = clo_sm[j]; h = chi_sm[j]; } else { l = f4_axis(__ldg(chunk_lo + j), axis); h = f4_axis(__ldg(chunk_hi + j), axis); }
                        dirty = !(l >= slo && h <= shi);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, dirty);
                    if (m) { c = (c & ~31) + __ffs(m) - 1; break; }
                    c = (c & ~31) + 32;
                }
                if (c > n_chunks) c = n_chunks;
            }
            if (lane == 0) next_chunk = c;
        }
        __syncthreads();
        cursor = next_chunk;
        if (cursor >= n_chunks) break;
        /* fold a granule of FOLD_GRANULE elements starting at the dirty chunk, all threads testing in parallel */
        const int base = cursor * BOX_CHUNK;
        const int limit = min(FOLD_GRANULE, n - base);
        Box1 mine; mine.c = 0.0f; mine.h = 0.0f;
        if (tid < limit) {
            if (pre_base == base) { mine.c = pre_c; mine.h = pre_h; }
            else { mine.c = f4_axis(box_c[base + tid], axis); mine.h = f4_axis(box_h[base + tid], axis); }
        }
        pre_base = base + FOLD_GRANULE;                                    /* dirty granules come in runs */
        if (pre_base + tid < n) { pre_c = f4_axis(box_c[pre_base + tid], axis); pre_h = f4_axis(box_h[pre_base + tid], axis); }
        int start = 0;
        while (true) {
            /* two barriers per round; the `first_changed` slot of the NEXT round is reset between them */
            Box1 ns = S; bool ch = false;
            if (tid >= start && tid < limit) { ns = contain1(S, mine); ch = !box1_equal(ns, S); }
            unsigned m = __ballot_sync(0xffffffffu, ch);
            if (m && lane == __ffs(m) - 1) atomicMin(&first_changed[round & 1], tid);
            __syncthreads();
            int f = first_changed[round & 1];
            if (tid == 0) first_changed[(round + 1) & 1] = 0x7fffffff;
            round++;
            if (f == 0x7fffffff) break;
            if (tid == f) S_sh = ns;
            __syncthreads();
            S = S_sh;
            start = f + 1;
        }
        cursor += (limit + BOX_CHUNK - 1) / BOX_CHUNK;
    }
    if (tid == 0) { bounds_out[axis] = S.c; bounds_out[3 + axis] = S.h; }
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
                                float4 *__restrict__ leaf_tri /* [n][4] */, float4 *__restrict__ leaf_box /* [n][2] */, float4 *__restrict__ leaf_frame /* [n][3] */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = sorted_idx[i];
    const float *t = tris + 9ll * s;
    V3 a = v3(t[0], t[1], t[2]), b = v3(t[3], t[4], t[5]), c = v3(t[6], t[7], t[8]);
    V3 e1 = b - a, e2 = c - a;                                          /* shapes.fut:69-70 */
    V3 nc = cross(e1, e2);                                              /* shapes.fut:71: ray independent, stored once */
    leaf_tri[4ll * i + 0] = make_float4(a.x, a.y, a.z, __uint_as_float(tri_mats[s]));
    leaf_tri[4ll * i + 1] = make_float4(nc.x, nc.y, nc.z, __int_as_float((int)0x80000000));   /* .w: escape link (k_pack_records) */
    leaf_tri[4ll * i + 2] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(s));
    leaf_tri[4ll * i + 3] = make_float4(e2.x, e2.y, e2.z, 0.0f);
    leaf_box[2ll * i + 0] = box_c[s];
    leaf_box[2ll * i + 1] = box_h[s];
    /* the shading frame of a hit on this triangle depends on the triangle only: unit normal (shapes.fut:84, never flipped) and
     * mk_orthonormal_basis of it (material.fut:374-379), computed here once with the operations k_shade would repeat per vertex */
    const V3 nn = normalise(nc);
    const Onb f = make_onb(nn);
    leaf_frame[3ll * i + 0] = make_float4(nn.x, nn.y, nn.z, 0.0f);
    leaf_frame[3ll * i + 1] = make_float4(f.b.x, f.b.y, f.b.z, 0.0f);
    leaf_frame[3ll * i + 2] = make_float4(f.t.x, f.t.y, f.t.z, 0.0f);
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
 * The reference runs `depth` = floor(log2 n)+2 Jacobi sweeps from zero boxes (bvh.fut:105-120).  After them a
 * node of height <= depth holds its converged box F; a taller ("crown") node holds A_depth(v) with
 *     A_k(v) = contain(A_{k-1}(left), A_{k-1}(right)),  A_0(internal) = {0,0},  A_k(leaf) = leaf box,
 * and A_k(u) = F(u) whenever height(u) <= k.  The (node, k) pairs that are NOT resolved by those rules are
 * discovered top-down from the crown nodes, one level per launch (k_crown_expand), then evaluated bottom-up
 * (k_crown_eval).  Level j holds the pairs with k = depth - j; level sizes live on the device.
 * If the pair buffer overflows (adversarially deep trees) the literal Jacobi sweeps are run instead. */
struct CrownPair { int node; int child[2]; };    /* child slot (absolute) or -1 = resolved by rule */
/* cnt[0..31] = pairs per level, cnt[32..62] = first slot of each level (cnt[32] = 0), cnt[63] = overflow flag.
 * Every block derives its level range with three L2 loads by one thread (data written by the previous level). */
__device__ __forceinline__ void crown_level_range(int *cnt, int level, int cap, int &base, int &count, int &next_base) {
    __shared__ int range[3];
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile int *vc = cnt;
        int b = vc[32 + level], c = vc[level];
        range[0] = b; range[1] = min(c, max(cap - b, 0)); range[2] = b + c;
        if (level + 1 < 31) vc[32 + level + 1] = b + c;          /* same value from every block */
    }
    __syncthreads();
    base = range[0]; count = range[1]; next_base = range[2];
}
__global__ void k_crown_find(const float4 *__restrict__ F, const int *__restrict__ height, int n_nodes, int depth,
                             float4 *__restrict__ A, CrownPair *pairs, int *cnt, int cap, int *overflow) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    A[2ll * i] = F[2ll * i]; A[2ll * i + 1] = F[2ll * i + 1];
    if (height[i] > depth) {
        int slot = atomicAdd(&cnt[0], 1);
        if (slot < cap) { pairs[slot].node = i; pairs[slot].child[0] = -1; pairs[slot].child[1] = -1; } else *overflow = 1;
    }
}
__device__ __forceinline__ void crown_expand_level(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ height,
                               CrownPair *pairs, int *cnt, int level, int depth, int cap, int *overflow) {
    int base, count, next_base; crown_level_range(cnt, level, cap, base, count, next_base);
    const int k = depth - level;                 /* pairs of this level carry k; children carry k - 1 */
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    /* warp-uniform trip count: slots of the next level are claimed with ONE atomic per warp (ballot + popcount) */
    for (int p0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); p0 < count; p0 += gridDim.x * blockDim.x) {
        const int p = p0 + lane;
        const bool valid = p < count;
        int node = 0, cl = -1, cr = -1;
        bool needl = false, needr = false;
        if (valid) {
            node = __ldcg(&pairs[base + p].node);
            cl = left[node]; cr = right[node];
            needl = cl >= 0 && k - 1 >= 1 && height[cl] > k - 1;
            needr = cr >= 0 && k - 1 >= 1 && height[cr] > k - 1;
        }
        const unsigned ml = __ballot_sync(0xffffffffu, needl), mr = __ballot_sync(0xffffffffu, needr);
        int wbase = 0;
        if (lane == 0 && (ml | mr)) wbase = atomicAdd(&cnt[level + 1], __popc(ml) + __popc(mr));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        int sl = -1, sr = -1;
        if (needl) { sl = next_base + wbase + __popc(ml & lt); if (sl < cap) { pairs[sl].node = cl; pairs[sl].child[0] = -1; pairs[sl].child[1] = -1; } else { *overflow = 1; sl = -1; } }
        if (needr) { sr = next_base + wbase + __popc(ml) + __popc(mr & lt); if (sr < cap) { pairs[sr].node = cr; pairs[sr].child[0] = -1; pairs[sr].child[1] = -1; } else { *overflow = 1; sr = -1; } }
        if (valid) { pairs[base + p].child[0] = sl; pairs[base + p].child[1] = sr; }
    }
}
__device__ __forceinline__ Box crown_child_box(int c, int slot, int ck, const float4 *leaf_box, const float4 *F, const float4 *pbox) {
    const float4 *src;
    if (c < 0) src = leaf_box + 2ll * (~c);
    else if (slot >= 0) { float4 a = __ldcg(pbox + 2ll * slot), b = __ldcg(pbox + 2ll * slot + 1); Box r; r.c = v3(a.x, a.y, a.z); r.h = v3(b.x, b.y, b.z); return r; }
    else if (ck == 0) { Box z; z.c = v3(0.0f, 0.0f, 0.0f); z.h = v3(0.0f, 0.0f, 0.0f); return z; }   /* bvh.fut:105-107 */
    else src = F + 2ll * c;                          /* height(c) <= ck: converged */
    float4 a = src[0], b = src[1];
    Box r; r.c = v3(a.x, a.y, a.z); r.h = v3(b.x, b.y, b.z); return r;
}
__device__ __forceinline__ void crown_eval_level(const int *__restrict__ left, const int *__restrict__ right, const float4 *__restrict__ leaf_box,
                             const float4 *__restrict__ F, const CrownPair *pairs, float4 *pbox, int *cnt,
                             int level, int depth, int cap, float4 *__restrict__ A) {
    int base, count, next_base; crown_level_range(cnt, level, cap, base, count, next_base); (void)next_base;
    const int k = depth - level;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < count; p += gridDim.x * blockDim.x) {
        CrownPair pr; pr.node = __ldcg(&pairs[base + p].node); pr.child[0] = __ldcg(&pairs[base + p].child[0]); pr.child[1] = __ldcg(&pairs[base + p].child[1]);
        int l = left[pr.node], r = right[pr.node];
        Box b = contain(crown_child_box(l, pr.child[0], k - 1, leaf_box, F, pbox), crown_child_box(r, pr.child[1], k - 1, leaf_box, F, pbox));
        float4 c4 = make_float4(b.c.x, b.c.y, b.c.z, 0.0f), h4 = make_float4(b.h.x, b.h.y, b.h.z, 0.0f);
        pbox[2ll * (base + p)] = c4; pbox[2ll * (base + p) + 1] = h4;
        if (level == 0) { A[2ll * pr.node] = c4; A[2ll * pr.node + 1] = h4; }
    }
}
__global__ void k_crown_expand(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ height,
                               CrownPair *pairs, int *cnt, int level, int depth, int cap, int *overflow) {
    crown_expand_level(left, right, height, pairs, cnt, level, depth, cap, overflow);
}
__global__ void k_crown_eval(const int *__restrict__ left, const int *__restrict__ right, const float4 *__restrict__ leaf_box,
                             const float4 *__restrict__ F, const CrownPair *pairs, float4 *pbox, int *cnt,
                             int level, int depth, int cap, float4 *__restrict__ A) {
    crown_eval_level(left, right, leaf_box, F, pairs, pbox, cnt, level, depth, cap, A);
}
/* The top levels are small (67, 92, 133, ... pairs on the 1M-triangle scene) and a launch per level costs ~6 us of
 * pure latency, so one CTA walks them with block barriers in between; loads of data written earlier in the same
 * kernel go through L2 (ld.cg / volatile; __syncthreads orders them within the CTA).  Only the last CROWN_TAIL_LEVELS (large) levels get their own launches. */
#define CROWN_TAIL_LEVELS 8
__global__ void __launch_bounds__(1024) k_crown_top_expand(const int *__restrict__ left, const int *__restrict__ right, const int *__restrict__ height,
                                                           CrownPair *pairs, int *cnt, int n_levels, int depth, int cap, int *overflow) {
    for (int lv = 0; lv < n_levels; lv++) { crown_expand_level(left, right, height, pairs, cnt, lv, depth, cap, overflow); __syncthreads(); }
}
__global__ void __launch_bounds__(1024) k_crown_top_eval(const int *__restrict__ left, const int *__restrict__ right, const float4 *__restrict__ leaf_box,
                                                         const float4 *__restrict__ F, const CrownPair *pairs, float4 *pbox, int *cnt,
                                                         int first_level, int depth, int cap, float4 *__restrict__ A) {
    for (int lv = first_level; lv >= 0; lv--) { crown_eval_level(left, right, leaf_box, F, pairs, pbox, cnt, lv, depth, cap, A); __syncthreads(); }
}
/* the reference's own sweep (bvh.fut:114-120), ping-pong; used when the pair buffer overflows */
__global__ void k_jacobi_sweep(const int *__restrict__ left, const int *__restrict__ right, const float4 *__restrict__ leaf_box,
                               const float4 *__restrict__ in, float4 *__restrict__ out, int n_nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    int l = left[i], r = right[i];
    const float4 *pl = (l < 0) ? leaf_box + 2ll * (~l) : in + 2ll * l;
    const float4 *pr = (r < 0) ? leaf_box + 2ll * (~r) : in + 2ll * r;
    float4 a = pl[0], b = pl[1], c = pr[0], d = pr[1];
    Box bl, br; bl.c = v3(a.x, a.y, a.z); bl.h = v3(b.x, b.y, b.z); br.c = v3(c.x, c.y, c.z); br.h = v3(d.x, d.y, d.z);
    Box o = contain(bl, br);
    out[2ll * i] = make_float4(o.c.x, o.c.y, o.c.z, 0.0f); out[2ll * i + 1] = make_float4(o.h.x, o.h.y, o.h.z, 0.0f);
}
__global__ void k_copy_f4(const float4 *__restrict__ in, float4 *__restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

/* Traversal records (lys_scene.h): node i -> (near.xyz | left child) (far.xyz | escape link), once per ray-direction octant
 * (copies = 8: near / far picked per axis as hit_aabb's swap would, shapes.fut:124-126) or once as (min | max) (copies = 1);
 * min / max = center -+ half as hit_aabb derives them per visit (shapes.fut:120).
 * Escape links: the reference's walk always goes left first (bvh.fut:126-142), whatever the ray, so the node visited after a
 * failed box test or a finished leaf is a property of the TREE -- the right child of the nearest ancestor-or-self that is a
 * left child, or the end marker on the right spine -- and is stored instead of being kept on a per-ray stack (a threaded tree;
 * the right child itself is the escape link of the left child, nobody else needs it).  Thread x < n_nodes: internal node x;
 * thread n_nodes + j: leaf j (link in leaf_tri[4j+1].w).  The climb is short: half of all nodes are left children. */
__global__ void k_pack_records(const float4 *__restrict__ A, const int *__restrict__ left, const int *__restrict__ right,
                               const int *__restrict__ parent, const int *__restrict__ leaf_parent, int n,
                               float4 *__restrict__ nodes, int copies, float4 *__restrict__ leaf_tri) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_nodes = n - 1;
    if (x >= n_nodes + n) return;
    const bool leaf = x >= n_nodes;
    int self = leaf ? ~(x - n_nodes) : x;               /* child encoding: internal i -> i, leaf j -> ~j */
    int p = leaf ? leaf_parent[x - n_nodes] : parent[x];
    while (p >= 0 && right[p] == self) { self = p; p = parent[p]; }      /* climb while we are a right child */
    const float link = __int_as_float((p < 0) ? (int)0x80000000 : right[p]);
    if (leaf) { leaf_tri[4ll * (x - n_nodes) + 1].w = link; return; }
    float4 c = A[2ll * x], h = A[2ll * x + 1];
    V3 mn = v3(c.x, c.y, c.z) - v3(h.x, h.y, h.z), mx = v3(c.x, c.y, c.z) + v3(h.x, h.y, h.z);
    const float l = __int_as_float(left[x]);
    for (int o = 0; o < copies; o++) {                  /* copies == 1: o = 0 = no swap */
        float4 *q = nodes + 2ll * ((long long)o * n_nodes + x);
        q[0] = make_float4((o & 4) ? mx.x : mn.x, (o & 2) ? mx.y : mn.y, (o & 1) ? mx.z : mn.z, l);
        q[1] = make_float4((o & 4) ? mn.x : mx.x, (o & 2) ? mn.y : mx.y, (o & 1) ? mn.z : mx.z, link);
    }
}


/* ------------------------------------------------------------------ host driver */
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

/* per-device launch facts (SM count, one-time function attributes) */
struct BuildDevice { int sms = 0; bool fold_attr = false; };
static BuildDevice &build_device() {
    static BuildDevice d[64];
    int dev = 0; cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!d[dev].sms) { int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); d[dev].sms = sms > 0 ? sms : 148; }
    return d[dev];
}

cudaError_t build_lbvh(SceneDev &sc, BuildScratch &ws, int refit_mode, cudaStream_t stream, uint64_t *launches) {
    BuildDevice &dev = build_device();
    const int n = (int)sc.n_tris;
    const int n_nodes = n - 1;
    const int T = 256;
    uint64_t nl = 0;
    k_tri_boxes<<<cdiv(n, BOX_CHUNK), BOX_CHUNK, 0, stream>>>(sc.tris, n, ws.box_c, ws.box_h, ws.chunk_lo, ws.chunk_hi); nl++;
    {
        const int max_chunk_sm = 24576;                                    /* 24576 x 8 B = 192 KB of dynamic shared memory (6.3 M triangles) */
        if (!dev.fold_attr) {                                              /* a function attribute is per device: a host may hold contexts on several */
            cudaError_t e = cudaFuncSetAttribute(k_bounds_fold, cudaFuncAttributeMaxDynamicSharedMemorySize, max_chunk_sm * 8);
            if (e != cudaSuccess) return e;
            dev.fold_attr = true;
        }
        const int n_chunk_sm = min(cdiv(n, BOX_CHUNK), max_chunk_sm);
        k_bounds_fold<<<3, FOLD_THREADS, (size_t)n_chunk_sm * 8, stream>>>(ws.box_c, ws.box_h, ws.chunk_lo, ws.chunk_hi, n, n_chunk_sm, sc.bounds); nl++;
    }
    k_morton<<<cdiv(n, T), T, 0, stream>>>(ws.box_c, n, sc.bounds, ws.keys[0], ws.vals[0]); nl++;
    /* radix sort */
    const int tiles = cdiv(n, RS_TILE);
    cudaMemsetAsync(ws.rs_hist, 0, 4 * RS_RADIX * sizeof(uint32_t), stream);
    cudaMemsetAsync(ws.rs_status, 0, (size_t)4 * tiles * RS_RADIX * sizeof(uint32_t) + 4 * sizeof(uint32_t), stream);
    k_rs_hist<<<min(cdiv(n, RS_THREADS * 8), dev.sms * 8), RS_THREADS, 0, stream>>>(ws.keys[0], n, ws.rs_hist); nl++;
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
    k_gather_leaves<<<cdiv(n, T), T, 0, stream>>>(sc.tris, sc.tri_mats, ws.box_c, ws.box_h, sc.sorted_idx, n, sc.leaf_tri, sc.leaf_box, sc.leaf_frame); nl++;
    k_karras<<<cdiv(n_nodes, T), T, 0, stream>>>(sc.morton, n, sc.left, sc.right, sc.parent, ws.leaf_parent); nl++;
    cudaMemsetAsync(ws.visits, 0, (size_t)n_nodes * sizeof(unsigned int), stream);
    k_refit<<<cdiv(n, T), T, 0, stream>>>(sc.left, sc.right, sc.parent, ws.leaf_parent, sc.leaf_box, ws.F, sc.height, ws.visits, n); nl++;
    const int depth = (int)(log2f((float)n)) + 2;                           /* bvh.fut:109 */
    if (refit_mode == 1) {                                                  /* converged boxes */
        k_copy_f4<<<cdiv(2ll * n_nodes, T), T, 0, stream>>>(ws.F, sc.node_box, 2ll * n_nodes); nl++;
    } else if (refit_mode == 2) {                                           /* literal Jacobi sweeps (fallback / cross-check) */
        cudaMemsetAsync(ws.F, 0, (size_t)2 * n_nodes * sizeof(float4), stream);     /* zero boxes, bvh.fut:105-107 */
        float4 *a = ws.F, *b = sc.node_box;
        for (int it = 0; it < depth; it++) { k_jacobi_sweep<<<cdiv(n_nodes, T), T, 0, stream>>>(sc.left, sc.right, sc.leaf_box, a, b, n_nodes); nl++; float4 *t = a; a = b; b = t; }
        if (a != sc.node_box) { k_copy_f4<<<cdiv(2ll * n_nodes, T), T, 0, stream>>>(a, sc.node_box, 2ll * n_nodes); nl++; }
    } else {
        cudaMemsetAsync(ws.crown_cnt, 0, 64 * sizeof(int), stream);
        k_crown_find<<<cdiv(n_nodes, T), T, 0, stream>>>(ws.F, sc.height, n_nodes, depth, sc.node_box, (CrownPair *)ws.crown_pairs, ws.crown_cnt, ws.crown_cap, ws.crown_cnt + 63); nl++;
        /* expand levels 0 .. depth-2 (level lv creates level lv+1); eval levels depth-1 .. 0 */
        const int G = dev.sms * 4;
        const int n_top = max(0, (depth - 1) - CROWN_TAIL_LEVELS);            /* expand levels handled by the single CTA */
        if (n_top > 0) { k_crown_top_expand<<<1, 1024, 0, stream>>>(sc.left, sc.right, sc.height, (CrownPair *)ws.crown_pairs, ws.crown_cnt, n_top, depth, ws.crown_cap, ws.crown_cnt + 63); nl++; }
        for (int lv = n_top; lv < depth - 1; lv++) { k_crown_expand<<<G, T, 0, stream>>>(sc.left, sc.right, sc.height, (CrownPair *)ws.crown_pairs, ws.crown_cnt, lv, depth, ws.crown_cap, ws.crown_cnt + 63); nl++; }
        for (int lv = depth - 1; lv > n_top; lv--) { k_crown_eval<<<G, T, 0, stream>>>(sc.left, sc.right, sc.leaf_box, ws.F, (const CrownPair *)ws.crown_pairs, ws.crown_box, ws.crown_cnt, lv, depth, ws.crown_cap, sc.node_box); nl++; }
        k_crown_top_eval<<<1, 1024, 0, stream>>>(sc.left, sc.right, sc.leaf_box, ws.F, (const CrownPair *)ws.crown_pairs, ws.crown_box, ws.crown_cnt, min(n_top, depth - 1), depth, ws.crown_cap, sc.node_box); nl++;
    }
    k_pack_records<<<cdiv(n_nodes + n, T), T, 0, stream>>>(sc.node_box, sc.left, sc.right, sc.parent, ws.leaf_parent, n, sc.nodes, sc.oct_copies, sc.leaf_tri); nl++;
    if (launches) *launches += nl;
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ lights (scene.fut:58-66), on the device
 * get_lights keeps the triangles whose material has a non-zero emission spectrum (a knot with wavelength >= 0 and value > 0,
 * scene.fut:59-60), IN INPUT ORDER: direct.fut:116-119 picks light `rand % n_lights` by position.  Four small launches
 * (flag per material, count per 1024-triangle chunk + index range check, scan of the chunk counts by one CTA, ordered
 * scatter) and the record kernel; the host learns n_lights with the single read-back that ends futhark_entry_init. */
#define LIGHT_CHUNK 1024
__global__ void k_mat_emissive(const float *__restrict__ mats, int m, unsigned char *__restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float *e = mats + 28ll * i + 16;
    bool on = false, dark = true;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (e[2 * k] >= 0.0f && e[2 * k + 1] > 0.0f) on = true;
        if (__float_as_uint(e[2 * k + 1]) != 0u) dark = false;
    }
    /* bit 0: emissive (scene.fut:59-60).  bit 1: every emission VALUE is +0.0 -- whatever the knots' wavelengths, spectrum_lookup
     * (spectrum.fut:30-49) then returns +0.0 for any wavelength (0, a knot value, or +0 + (+0 - +0) * t), so k_shade skips the
     * lookup of the vertex-0 emission term (integrator.fut:51-53) for such materials */
    flag[i] = (on ? 1 : 0) | (dark ? 2 : 0);
}
__device__ __forceinline__ bool tri_is_light(const uint32_t *__restrict__ tri_mats, const unsigned char *__restrict__ mat_flag, int i, int n, int m, bool &bad) {
    if (i >= n) return false;
    const uint32_t k = tri_mats[i];
    if (k >= (uint32_t)m) { bad = true; return false; }
    return (mat_flag[k] & 1) != 0;
}
__global__ void __launch_bounds__(LIGHT_CHUNK) k_light_count(const uint32_t *__restrict__ tri_mats, const unsigned char *__restrict__ mat_flag, int n, int m,
                                                              int *__restrict__ chunk_cnt, int *__restrict__ info) {
    bool bad = false;
    const bool on = tri_is_light(tri_mats, mat_flag, blockIdx.x * LIGHT_CHUNK + threadIdx.x, n, m, bad);
    const int c = __syncthreads_count(on);
    if (__syncthreads_or(bad) && threadIdx.x == 0) info[1] = 1;             /* material index out of range */
    if (threadIdx.x == 0) chunk_cnt[blockIdx.x] = c;
}
/* exclusive scan of the chunk counts in place (one CTA, 1024 chunks per round); info[0] = number of lights */
__global__ void __launch_bounds__(1024) k_light_scan(int *__restrict__ chunk_cnt, int n_chunks, int *__restrict__ info) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_chunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_chunks ? chunk_cnt[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_sync(0xffffffffu, x, max(lane - o, 0)); if (lane >= o) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) { int w = wsum[lane], z = w; for (int o = 1; o < 32; o <<= 1) { int y = __shfl_sync(0xffffffffu, z, max(lane - o, 0)); if (lane >= o) z += y; } wsum[lane] = z - w; }
        __syncthreads();
        const int excl = carry + wsum[warp] + x - v;
        if (i < n_chunks) chunk_cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) info[0] = carry;
}
__global__ void __launch_bounds__(LIGHT_CHUNK) k_light_scatter(const uint32_t *__restrict__ tri_mats, const unsigned char *__restrict__ mat_flag, int n, int m,
                                                                const int *__restrict__ chunk_off, int cap, int *__restrict__ light_src) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * LIGHT_CHUNK + threadIdx.x;
    bool bad = false;
    const bool on = tri_is_light(tri_mats, mat_flag, i, n, m, bad);
    const unsigned mask = __ballot_sync(0xffffffffu, on);
    if (lane == 0) wsum[warp] = __popc(mask);
    __syncthreads();
    if (warp == 0) { int w = wsum[lane], z = w; for (int o = 1; o < 32; o <<= 1) { int y = __shfl_sync(0xffffffffu, z, max(lane - o, 0)); if (lane >= o) z += y; } wsum[lane] = z - w; }
    __syncthreads();
    if (on) { const int slot = chunk_off[blockIdx.x] + wsum[warp] + __popc(mask & ((1u << lane) - 1u)); if (slot < cap) light_src[slot] = i; }
}
__global__ void k_build_lights(const float *__restrict__ tris, const uint32_t *__restrict__ tri_mats, const float *__restrict__ mats,
                               const int *__restrict__ src, const int *__restrict__ info, int cap, LightRec *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(info[0], cap)) return;
    int s = src[i];
    const float *t = tris + 9ll * s;
    V3 a = v3(t[0], t[1], t[2]), bb = v3(t[3], t[4], t[5]), c = v3(t[6], t[7], t[8]);
    V3 e1 = bb - a, e2 = c - a;
    V3 nc = cross(e1, e2);
    float area = norm(nc) / 2.0f;                       /* direct.fut:17-20,37 */
    V3 n = normalise(nc);                               /* triangle_normal shapes.fut:59-62 */
    LightRec r;
    r.a[0] = a.x; r.a[1] = a.y; r.a[2] = a.z; r.area = area;
    r.e1[0] = e1.x; r.e1[1] = e1.y; r.e1[2] = e1.z; r.inv_area = 1.0f / area;
    r.e2[0] = e2.x; r.e2[1] = e2.y; r.e2[2] = e2.z; r.theta = 0.0f;
    r.n[0] = n.x; r.n[1] = n.y; r.n[2] = n.z; r.kind = 0;
    const float *em = mats + 28ll * tri_mats[s] + 16;
    for (int k = 0; k < 12; k++) r.emission[k] = em[k];
    r.src_index = s; r.pad[0] = r.pad[1] = r.pad[2] = 0;
    out[i] = r;
}
cudaError_t build_lights(SceneDev &sc, unsigned char *mat_flag, int *chunk_cnt, int *info, int cap, cudaStream_t stream, uint64_t *launches) {
    const int n = (int)sc.n_tris, m = (int)sc.n_mats;
    const int chunks = cdiv(n, LIGHT_CHUNK);
    cudaMemsetAsync(info, 0, 4 * sizeof(int), stream);
    k_mat_emissive<<<cdiv(m, 128), 128, 0, stream>>>(sc.mats, m, mat_flag);
    sc.mat_flag = mat_flag;
    k_light_count<<<chunks, LIGHT_CHUNK, 0, stream>>>(sc.tri_mats, mat_flag, n, m, chunk_cnt, info);
    k_light_scan<<<1, 1024, 0, stream>>>(chunk_cnt, chunks, info);
    k_light_scatter<<<chunks, LIGHT_CHUNK, 0, stream>>>(sc.tri_mats, mat_flag, n, m, chunk_cnt, cap, sc.light_src);
    k_build_lights<<<cdiv(cap, 128), 128, 0, stream>>>(sc.tris, sc.tri_mats, sc.mats, sc.light_src, info, cap, sc.lights);
    if (launches) *launches += 5;
    return cudaGetLastError();
}
/* after the host has learnt that there are more lights than `old cap`: scatter + records again into larger arrays */
cudaError_t rebuild_lights(SceneDev &sc, const unsigned char *mat_flag, const int *chunk_off, const int *info, int cap, cudaStream_t stream, uint64_t *launches) {
    const int n = (int)sc.n_tris, m = (int)sc.n_mats;
    k_light_scatter<<<cdiv(n, LIGHT_CHUNK), LIGHT_CHUNK, 0, stream>>>(sc.tri_mats, mat_flag, n, m, chunk_off, cap, sc.light_src);
    k_build_lights<<<cdiv(cap, 128), 128, 0, stream>>>(sc.tris, sc.tri_mats, sc.mats, sc.light_src, info, cap, sc.lights);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

} // namespace lys
