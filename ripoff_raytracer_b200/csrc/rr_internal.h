// rr_internal.h -- data layout in HBM shared by the builder, the render
// kernels and the C-ABI host layer.  
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rr_api.h"

#define RR_EPSILON 1e-6f              // reference src/Trace.cl:6
#define RR_TAU 6.28318530717958647692f // reference src/Trace.cl:5
#define RR_STACK 64                   // reference src/Trace.cl:2 (BVHStackSize)
#define RR_MAX_INVISIBLE_PASSES 256u   // pass-throughs of Invisible surfaces per path before it is ended
#define RR_DIRECT_MAX 4               // segments this small are tested without a hierarchy
#define RR_TILE_W 8                   // default work tile: 8 x 4 pixels = one warp
#define RR_TILE_H 4

namespace rr {

// One LBVH over a segmented primitive array.  Index spaces:
//   prim  : index in the uploaded array (original order)
//   slot  : position after the sort; segment s owns slots [sfirst, sfirst+count)
//   inner : inner node i of segment s lives at global index sfirst + i (count-1 used)
struct Lbvh {
  uint64_t n = 0;          // primitives covered by segments
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
  float4* nodes = nullptr;     // [n*4] per inner node: child boxes + refs
  uint32_t max_depth = 0;
};

// Per-mesh record read by the kernel (all floats/ints, 16-byte multiples).
struct DMesh {
  float Rinv[9];  // rows of transpose(makeRotation)   reference src/Trace.cl:452-454
  float R[9];     // rows of makeRotation
  float pos[3];
  float scale;
  float bmin[3];  // local-space root box
  float bmax[3];
  uint32_t sfirst;  // first slot / inner-node base
  uint32_t count;   // triangles
  int32_t cull;     // cullBackface (reference src/Trace.cl:460-462)
  int32_t type;     // material type
  int32_t skip;     // scale <= EPSILON (reference src/Trace.cl:448)
  int32_t material; // index into the material table
  int32_t pad[2];
};

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

struct DCamera {
  float pos[3];
  float pitch, yaw, roll, fov, aspect;
};

struct Counters {
  unsigned long long rays, rays_reused, box_tests, tri_tests, sphere_tests, tiles;
};

// Everything a render kernel needs (passed by value).
struct RenderParams {
  // scene
  const DMesh* meshes;
  int32_t n_meshes;
  const DMaterial* materials;
  const float4* tri_nodes;   // 4 x float4 per inner node
  const float4* tri_geom;    // 3 x float4 per slot: (A, primId) (B-A) (C-A)
  const float4* tri_nrm;     // 3 x float4 per slot: nA nB nC
  // spheres (one segment, world space)
  int32_t n_spheres;
  const float4* sph_nodes;
  const float4* sph_geom;    // (center, radius) per slot
  const uint32_t* sph_order; // slot -> sphere index
  float sph_bmin[3], sph_bmax[3];
  // frame
  DCamera cam;
  uint32_t width, height, spp, max_bounces;
  int32_t frame_index;
  uint32_t tile_w, tile_h, tiles_x, tiles_y;
  uint32_t tile_begin, tile_stride;  // static partition: this rank renders tile_begin + k*tile_stride ...
  unsigned long long* queue;         // tile counter (may live in a peer GPU's memory)
  uint8_t* frame;                    // RGBA8 (may live in a peer GPU's memory)
  float* radiance;                   // optional, local
  Counters* counters;
  int32_t* hit_mesh;                 // primary-hit outputs (primary kernel only)
  int32_t* hit_prim;
  float* hit_dst;
};

// ---- builder (rr_lbvh.cu) -------------------------------------------------
// boxes: prim boxes [n_total*6] on the device, segments on the HOST (first,count per segment,
// sorted by first, non-overlapping).  Builds everything in `out` on `stream`.
cudaError_t lbvh_build(Lbvh& out, const float* d_prim_box, uint64_t n_total, const uint32_t* h_seg_first,
                       const uint32_t* h_seg_count, uint32_t n_segs, cudaStream_t stream);
void lbvh_free(Lbvh& b);

cudaError_t launch_tri_boxes(const rr_triangle* d_tris, uint64_t n, float* d_box, cudaStream_t s);
cudaError_t launch_sphere_boxes(const rr_sphere* d_sph, uint64_t n, float* d_box, cudaStream_t s);
cudaError_t launch_pack_tris(const rr_triangle* d_tris, const uint32_t* d_order, uint64_t n, float4* geom, float4* nrm,
                             cudaStream_t s);
cudaError_t launch_pack_spheres(const rr_sphere* d_sph, const uint32_t* d_order, uint64_t n, float4* geom,
                                cudaStream_t s);

// ---- render (rr_render.cu) --------------------------------------------------
cudaError_t launch_render(const RenderParams& p, bool count_tests, int sm_count, cudaStream_t s);
cudaError_t launch_primary(const RenderParams& p, cudaStream_t s);
cudaError_t launch_math_probe(int fn, const float* x, const float* y, float* out, uint64_t n, cudaStream_t s);
cudaError_t launch_rng_probe(uint32_t pixel, int32_t frame, uint32_t* out_u32, float* out_f32, cudaStream_t s);

}  // namespace rr
