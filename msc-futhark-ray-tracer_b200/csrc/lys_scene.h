/* lys_scene.h -- device-resident scene (HBM layout) shared by the build, the wavefront kernels and the ABI.
 *
 * All pointers are device pointers.  Layout (n triangles, m materials, L lights):
 *   tris      [n][9]  f32   input triangles as uploaded (src/scene.fut:26-35)
 *   tri_mats  [n]     u32
 *   mats      [m][28] f32   material rows (src/scene.fut:37-53)
 *   leaf_tri  [n][4]  float4  sorted leaves, 64 B: (a.xyz | mat_ix), (e1 x e2 | escape link) = the plane test's sector,
 *                             then (e1.xyz | source index), (e2.xyz | 0) for the barycentric test.  Escape link: the node the
 *                             left-first walk visits after this leaf (lbvh.cu: k_pack_records)
 *   leaf_box  [n][2]  float4  sorted leaf boxes (center | half_dims)
 *   leaf_frame [n][3] float4  shading frame of a hit on the leaf: unit normal, then b and t of mk_orthonormal_basis (material.fut:374-379)
 *   nodes     [copies][n-1][2] float4  traversal records, 32 B = one sector: (near.xyz | left child) (far.xyz | escape link) of
 *                             node i.  copies = 8 (scenes up to LYS_OCT_MAX_NODES nodes): one copy per ray-direction octant (bit
 *                             2/1/0 = 1/dir.x, .y, .z < 0) with near / far already picked per axis as hit_aabb's swap would
 *                             (shapes.fut:124-126), so the box test needs no select; copies = 1: (min | max), box test with
 *                             selects.  Corners are derived from (center, half) with the subtraction / addition hit_aabb does
 *                             per visit (shapes.fut:120).  Escape link: the node the left-first walk (bvh.fut:126-142) visits
 *                             after this node's box fails -- a property of the tree, so the walk needs no stack
 *                             (lbvh.cu: k_pack_records, wavefront.cu: traverse)
 *   node_box  [n-1][2] float4 node boxes as the reference stores them (center | half_dims)
 *   left/right/parent/height [n-1] i32; child encoding: internal i -> i, leaf i -> ~i
 *   morton, sorted_idx [n] u32; bounds [6] f32 (center, half_dims)
 *   lights    [L] LightRec
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lys {

#define LYS_OCT_MAX_NODES (1 << 16)      /* 8 x 2 MB of octant records at most: stays L2 resident */

struct LightRec {            /* 32 floats = 128 B, float4-aligned */
    float a[3]; float area;                /* vertex a, triangle area (direct.fut:17-20) */
    float e1[3]; float inv_area;           /* b - a, 1/area (direct.fut:22,42) */
    float e2[3]; float theta;              /* c - a, frustum half-angle (light.fut:5) */
    float n[3]; int kind;                  /* triangle_normal (shapes.fut:59-62); 0 diffuse, 1 frustum */
    float emission[12];                    /* spectrum knots */
    int src_index; int pad[3];
};

struct SceneDev {
    int64_t n_tris = 0, n_mats = 0, n_lights = 0;
    float *tris = nullptr; uint32_t *tri_mats = nullptr; float *mats = nullptr;
    float4 *leaf_tri = nullptr, *leaf_box = nullptr, *leaf_frame = nullptr, *nodes = nullptr, *node_box = nullptr;
    int oct_copies = 8;                    /* copies in `nodes`: 8 (one per ray-direction octant) or 1 (scenes above LYS_OCT_MAX_NODES: box test with selects) */
    int *left = nullptr, *right = nullptr, *parent = nullptr, *height = nullptr;
    uint32_t *morton = nullptr, *sorted_idx = nullptr;
    float *bounds = nullptr;
    LightRec *lights = nullptr;
    int *light_src = nullptr;
    const unsigned char *mat_flag = nullptr;   /* [m] bit 0: emissive material, bit 1: all emission values are +0.0 (lbvh.cu: k_mat_emissive) */
    float build_ms = 0.0f;
};

struct BuildScratch {
    int64_t cap = 0;
    float4 *box_c = nullptr, *box_h = nullptr, *F = nullptr, *chunk_lo = nullptr, *chunk_hi = nullptr;
    uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr};
    uint32_t *rs_hist = nullptr, *rs_status = nullptr;
    int *leaf_parent = nullptr; unsigned int *visits = nullptr;
    void *crown_pairs = nullptr; float4 *crown_box = nullptr; int *crown_cnt = nullptr; int crown_cap = 0;   /* cnt[0..62] level sizes, cnt[63] overflow flag */
};

/* refit_mode: 0 reference-exact (truncated Jacobi via crown worklists), 1 converged, 2 reference-exact via literal sweeps */
cudaError_t build_lbvh(SceneDev &sc, BuildScratch &ws, int refit_mode, cudaStream_t stream, uint64_t *launches);
/* lights (scene.fut:58-66) on the device: emissive triangles in input order -> sc.light_src / sc.lights (capacity `cap`).
 * info[0] = number of lights found (may exceed cap: then call rebuild_lights with larger arrays), info[1] = 1 if a material
 * index is out of range.  mat_flag [n_mats] bytes and chunk_cnt [ceil(n / 1024)] ints are scratch that rebuild_lights reuses. */
cudaError_t build_lights(SceneDev &sc, unsigned char *mat_flag, int *chunk_cnt, int *info, int cap, cudaStream_t stream, uint64_t *launches);
cudaError_t rebuild_lights(SceneDev &sc, const unsigned char *mat_flag, const int *chunk_off, const int *info, int cap, cudaStream_t stream, uint64_t *launches);

} // namespace lys
