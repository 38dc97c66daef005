// rr_internal.h -- data layout in HBM shared by the builder, the render
// kernels and the C-ABI host layer.  
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rr_api.h"

#define RR_EPSILON 1e-6f              // reference src/Trace.cl:6
#define RR_TAU 6.28318530717958647692f // reference src/Trace.cl:5
#define RR_MAX_DEPTH 64               // deepest binary hierarchy accepted (the reference's BVHStackSize, src/Trace.cl:2)
#define RR_STACK_MAX 200               // traversal stack entries at most: a 4-wide node pushes up to 3 per level
#define RR_MAX_INVISIBLE_PASSES 256u   // pass-throughs of Invisible surfaces per path before it is ended
#define RR_MAX_BOUNCES 0x7fffffu       // the bounce counter shares a 32-bit slot word with the 9-bit pass counter
#define RR_QUEUE_EPOCH_SHIFT 48        // queue counter word: frame epoch (16 bits) << 48 | work items (pixels in tile-major order) handed out
#define RR_DIRECT_MAX 4               // segments this small are tested without a hierarchy
#define RR_LEAF_MAX 4                 // primitives per leaf of a hierarchy at most (the leaf reference keeps count - 1 in 2 bits)
#ifndef RR_LEAF_DEFAULT
#define RR_LEAF_DEFAULT 2             // what rr_upload_scene builds with
#endif
#ifndef RR_TOP_CLUSTER_DEFAULT
#define RR_TOP_CLUSTER_DEFAULT 512    // Karras subtrees of at most this many primitives are the clusters of the SAH-ordered top (0: off)
#endif
#define RR_TOP_MIN_CLUSTERS 32        // ... built only for segments of at least this many times the cluster size
#ifndef RR_LEAF_DEFAULT_SPHERES
#define RR_LEAF_DEFAULT_SPHERES 1     // ... for the sphere hierarchy (one sphere per leaf measured best: C2 +6.6 %, C4 +2.8 % over two)
#endif
#define RR_MAX_PRIMS 0x1ffffff0ull    // ... and the first sorted slot in the 29 bits above them
#define RR_TILE_W 8                   // default work tile: 8 x 4 pixels = one warp
#define RR_TILE_H 4
#define RR_NODE_QUADS 8   // float4 per traversal node (4-wide node, 128 bytes)
#ifndef RR_POOL
#define RR_POOL 96        // path slots per warp of the render kernel (rr_render.cu)
#endif
#ifndef RR_NT
#define RR_NT 128         // threads per CTA of the render kernel
#endif
#ifndef RR_MIN_CTAS
#define RR_MIN_CTAS 5     // resident CTAs per SM the render kernel is compiled for
#endif
#ifndef RR_TOP_STAGE
#define RR_TOP_STAGE 0    // > 0: this many nodes of the top of the largest hierarchy are staged in shared memory (A/B switch)
#endif
#ifndef RR_PIXEL_QUEUE
#define RR_PIXEL_QUEUE 1  // 1: the counter hands out PIXELS (tile-major order), a warp takes exactly as many as it has free slots;
#endif                    // 0: a warp pops whole tiles and consumes them as slots free up (round 1; A/B switch)
#ifndef RR_ADAPTIVE_KEEP
#define RR_ADAPTIVE_KEEP 1  // node-step runs that start with few lanes keep stepping until a quarter of them has finished (A/B switch)
#endif
#ifndef RR_TILE_ORDER
#define RR_TILE_ORDER 0   // order in which the queue hands out the tiles of a launch: 0 row-major, 1 reversed, 2 scattered (A/B switch)
#endif
#ifndef RR_STREAM_NRM
#define RR_STREAM_NRM 0   // 1: vertex normals are loaded with the evict-first policy (LDG.E.EF): fetched once per accepted hit, rarely reused
#endif
#ifndef RR_STREAM_GEOM
#define RR_STREAM_GEOM 0  // 1: ... and the triangle positions too (A/B switch)
#endif
#ifndef RR_KEEP_NUM
#define RR_KEEP_NUM 3       // ... threshold = RR_KEEP_NUM / 4 of the lanes the run started with
#endif
#define RR_TOP_TAG 0x40000000  // node references at or above this value address the staged copy
#define RR_POOL_WORDS 28  // 32-bit words of one slot in shared memory
#define RR_COLD_WORDS 25  // ... and in the per-warp global scratch

namespace rr {

// One LBVH over a segmented primitive array.  Index spaces:
//   prim  : index in the uploaded array (original order)
//   slot  : position after the sort; segment s owns slots [sfirst, sfirst+count)
//   inner : inner node i of segment s lives at global index sfirst + i (count-1 used)
struct Lbvh {
  uint64_t n = 0;          // primitives covered by segments
  uint64_t n_nodes = 0;    // node slots of the traversal array: n Karras slots + the nodes of the SAH-ordered tops
  uint32_t* seg_root = nullptr;  // [n_segs] root node of every segment (its Karras root, or the root of its SAH top)
  uint32_t n_segs = 0;
  // build products (kept for rr_bvh_read and packing)
  uint64_t* codes = nullptr;   // [n] sorted keys
  uint32_t* order = nullptr;   // [n] prim at slot
  int32_t* left = nullptr;     // [n]
  int32_t* right = nullptr;    // [n]
  int32_t* parent = nullptr;   // [n] parent of inner node
  float* bounds = nullptr;     // [n*6] inner boxes
  float* prim_box = nullptr;   // [n_prims_total*6] original order
  float* seg_box = nullptr;    // [n_segs*6]
  uint32_t* seg_first = nullptr;  // [n_segs] first prim (original array)
  uint32_t* seg_count = nullptr;  // [n_segs]
  uint32_t* seg_sfirst = nullptr; // [n_segs] first slot
  // traversal arrays
  float4* nodes = nullptr;     // [n*RR_NODE_QUADS] 4-wide traversal nodes: delta-inflated child boxes + refs
  uint32_t max_depth = 0;    // deepest leaf of the binary hierarchy (root = 1)
  uint32_t wide_levels = 0;  // levels of the 4-wide hierarchy
};

// Per-mesh record read by the kernel: ten float4 (LDG.128 each).  When the scene has spheres one
// pseudo-mesh is appended at index n_meshes (world space, identity transform, RR_MF_SPHERES).
struct DMesh {
  float4 ri0, ri1, ri2;  // rows of transpose(makeRotation) (reference src/Trace.cl:452-454); .w = pos.x / pos.y / pos.z
  float4 r0, r1, r2;     // rows of makeRotation; r0.w = scale, r1.w = 1/scale (exact, valid with RR_MF_POW2)
  float4 bmin;           // local-space root box, delta-inflated; .w = bits(first slot / inner-node base)
  float4 bmax;           //                                     ; .w = bits(primitive count)
  float4 wmin;           // world-space box (conservative);       .w = bits(flags)
  float4 wmax;           //                                     ; .w = bits(material index)
};
#define RR_MF_SKIP 1u      // scale <= EPSILON (reference src/Trace.cl:448) or no primitives
#define RR_MF_CULL 2u      // cullBackface (reference src/Trace.cl:460-462)
#define RR_MF_POW2 4u      // scale is a power of two: x / scale == x * (1/scale) bit for bit
#define RR_MF_UNIT 8u      // scale == 1: the division is the identity
#define RR_MF_SPHERES 16u  // the sphere set (extension)
#define RR_MF_TYPE_SHIFT 8

struct DMaterial {
  int32_t type;
  float ior;
  float emissionStrength;
  float reflectiveness;
  float color[3];
  float specularProbability;
  float emissionColor[3];
  float pad;
};

static_assert(sizeof(DMaterial) == 48, "the shade phase reads a material as three float4: (type, ior, emissionStrength, reflectiveness) (color, specularProbability) (emissionColor, pad)");

struct DCamera {
  float pos[3];
  float pitch, yaw, roll, fov, aspect;
};

struct Counters {
  unsigned long long rays, box_tests, tri_tests, sphere_tests, tiles;
  unsigned long long stack_overflows;  // node steps that could not push their far children (must stay 0: the stacks are sized per scene)
  unsigned long long queue_errors;     // tile pops that met a counter of another frame (rr_queue_reset during a frame)
  unsigned long long phase_runs[5], phase_lanes[5];  // scheduler statistics (instrumented kernel only)
  // drain of the persistent slot pool (instrumented kernel only): per warp, the time between its first failed tile pop
  // (queue empty: no new pixels) and its exit, in ns of %globaltimer
  unsigned long long tail_ns_sum, tail_ns_max, tail_warps;
};

// Warp scheduler knobs of k_render.  Phases: 0 pixel, 1 shade, 2 setup, 3 traverse, 4 leaf.
struct Tuning {
  uint32_t weight[5];   // a phase runs when weight * ready lanes is the largest
  uint32_t trav_keep;   // the traversal loop keeps stepping while at least this many lanes can step
  uint32_t speculate;   // 1: a lane with a postponed leaf keeps traversing
  uint32_t ctas_per_sm; // persistent CTAs per SM (0 = default)
};

// Scene features (k_render<..., FEAT>)
#define RR_FEAT_SPHERES 1     // the sphere set
#define RR_FEAT_MATERIALS 2   // Checker / Glassy / Invisible materials (Solid and OneSided are always compiled in)
#define RR_FEAT_TLAS 4        // more than 32 meshes: the implicit top level
#define RR_FEAT_ALL 7

#define RR_TRAV_KEEP_DEFAULT 20  // default of Tuning::trav_keep (a compile-time constant in the untuned kernels)
#define RR_TLAS_MAX_LEVELS 14  // 4^13 chunks of 32 meshes: more than the 31-bit mesh index allows

// Everything a render kernel needs (passed by value).
struct RenderParams {
  // scene
  const DMesh* meshes;
  int32_t n_meshes;
  const DMaterial* materials;
  const float4* nodes;       // RR_NODE_QUADS float4 per node: triangle hierarchies, then the sphere hierarchy
  const float4* tri_geom;    // 3 x float4 per slot: (A, primId) (B-A) (C-A)
  const float4* tri_nrm;     // 3 x float4 per slot: nA nB nC
  // spheres (one segment, world space)
  int32_t n_spheres;
  int32_t last_mesh;         // n_meshes - 1, or n_meshes when the sphere pseudo-mesh exists
  // more than 32 meshes: implicit tree of world boxes (lo, hi pairs) over the Morton-ordered meshes; nullptr otherwise
  const float4* tlas_blocks; // one box per block of 8 meshes
  const float4* tlas;        // level 0: one box per chunk of 32 meshes; level l: one box per 4^l chunks
  const uint32_t* tlas_levels; // number of levels in `tlas`, then the index of the first box of every level
  const float4* top_nodes;   // RR_TOP_STAGE: the staged nodes (root first; the root's references point at the staged children)
  uint32_t top_count;        //               how many (0: nothing staged)
  int32_t top_root;          //               node index of the hierarchy root they belong to
  const float4* sph_geom;    // (center, radius) per slot
  const uint32_t* sph_order; // slot -> sphere index
  Tuning tune;
  // frame
  DCamera cam;
  uint32_t width, height, spp, max_bounces;
  int32_t frame_index;
  uint32_t tile_w, tile_h, tiles_x, tiles_y;
  uint32_t tile_begin, tile_stride;  // static partition: this rank renders tile_begin + k*tile_stride ...
  uint32_t tile_pixels, queue_items; // RR_PIXEL_QUEUE: tile_w * tile_h, and the work items of this launch (its tiles x tile_pixels, < 2^32)
  uint32_t queue_tiles, tile_mul;    // tiles of this launch; RR_TILE_ORDER 2: multiplier of the tile permutation (coprime to queue_tiles)
  const uint32_t* tile_order;        // nullptr, or the launch's tiles in the order the queue hands them out (rr_set_tile_order)
  uint32_t* cost;                    // instrumented kernel: path segments per pixel (rr_render_cost), or nullptr
  uint2* stack;                      // traversal stacks, stack_entries * RR_POOL entries per warp (scratch)
  uint32_t stack_entries;            // 3 per level of the deepest 4-wide hierarchy + slack
  uint32_t* cold;                    // cold slot words, RR_COLD_WORDS * RR_POOL per warp (scratch)
  uint32_t stack_warps;              // warps the scratch was sized for
  uint32_t pool_use;                 // path slots per warp that take pixels (<= RR_POOL; the others stay idle for this frame)
  unsigned long long* queue;         // tile counter (may live in a peer GPU's memory): frame epoch << 48 | next tile
  uint32_t queue_epoch;              // epoch this launch expects in the counter (0 for a context-local queue)
  uint8_t* frame;                    // RGBA8 (may live in a peer GPU's memory)
  float* radiance;                   // optional, local
  Counters* counters;
  int32_t* hit_mesh;                 // primary-hit outputs (primary kernel only)
  int32_t* hit_prim;
  float* hit_dst;
};

// ---- device memory for scene data (rr_api.cu) -------------------------------
// cudaMalloc / cudaFree of the large scene arrays cost hundreds of milliseconds per upload (the driver maps and
// unmaps the memory each time).  Scene-sized blocks therefore go through a small per-device cache: dev_free() parks
// a block, dev_malloc() reuses a parked block of a fitting size, dev_trim() gives parked blocks back to the driver.
cudaError_t dev_malloc_bytes(void** p, size_t bytes);
void dev_free(void* p);
void dev_trim(int ordinal);
template <class T>
inline cudaError_t dev_malloc(T** p, size_t bytes) { return dev_malloc_bytes(reinterpret_cast<void**>(p), bytes); }

// ---- builder (rr_lbvh.cu) -------------------------------------------------
// boxes: prim boxes [n_total*6] on the device, segments on the HOST (first,count per segment,
// sorted by first, non-overlapping).  Builds everything in `out` on `stream`.
// ref_offset is added to the inner-node references of the packed traversal nodes (the sphere
// hierarchy is stored behind the triangle hierarchies in one array).
cudaError_t lbvh_build(Lbvh& out, const float* d_prim_box, uint64_t n_total, const uint32_t* h_seg_first,
                       const uint32_t* h_seg_count, uint32_t n_segs, int32_t ref_offset, uint32_t leaf_max, uint32_t top_cluster,
                       cudaStream_t stream);
// Conservative slack added to every box a ray is tested against: 2^-18 of the largest |coordinate|
// of the segment box (about 60 ulp), so that the closest hit does not depend on the traversal order.
__host__ __device__ inline float box_delta(const float* seg_box6) {
  float m = 0.0f;
  for (int k = 0; k < 6; ++k) {
    float a = seg_box6[k] < 0.0f ? -seg_box6[k] : seg_box6[k];
    if (a > m && a < 3.0e38f) m = a;
  }
  return m * 3.814697265625e-06f;
}
// The per-ray part of the slack: 2^-18 of the largest |origin coordinate| (rr_render.cu RaySlack / make_slack).
__host__ __device__ inline float ray_slack(float ox, float oy, float oz) {
  const float ax = ox < 0.0f ? -ox : ox, ay = oy < 0.0f ? -oy : oy, az = oz < 0.0f ? -oz : oz;
  float m = ax > ay ? ax : ay;
  m = m > az ? m : az;
  return (m < 3.0e38f ? m : 0.0f) * 3.814697265625e-06f;
}
void lbvh_free(Lbvh& b);

cudaError_t launch_tri_boxes(const rr_triangle* d_tris, uint64_t n, float* d_box, cudaStream_t s);
cudaError_t launch_sphere_boxes(const rr_sphere* d_sph, uint64_t n, float* d_box, cudaStream_t s);
cudaError_t launch_pack_tris(const rr_triangle* d_tris, const uint32_t* d_order, uint64_t n, float4* geom, float4* nrm,
                             cudaStream_t s);
cudaError_t launch_pack_spheres(const rr_sphere* d_sph, const uint32_t* d_order, uint64_t n, float4* geom,
                                cudaStream_t s);

// ---- render (rr_render.cu) --------------------------------------------------
// slack: the instantiation with the per-ray culling slack (rr_render.cu RaySlack; chosen per frame by rr_api.cu frame_needs_slack)
// feat: RR_FEAT_* bits of the uploaded scene; the kernel instantiation without the code of absent features is launched
cudaError_t launch_render(const RenderParams& p, bool count_tests, bool slack, int feat, int sm_count, cudaStream_t s);
cudaError_t launch_primary(const RenderParams& p, bool slack, int sm_count, cudaStream_t s);
void default_tuning(Tuning& t);
size_t render_stack_bytes_per_warp(uint32_t stack_entries);
size_t render_cold_bytes_per_warp();
int render_max_warps_per_sm();
cudaError_t launch_math_probe(int fn, const float* x, const float* y, float* out, uint64_t n, cudaStream_t s);
cudaError_t launch_rng_probe(uint32_t pixel, int32_t frame, uint32_t* out_u32, float* out_f32, cudaStream_t s);

}  // namespace rr
