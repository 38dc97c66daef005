// rr_api.cu -- the thin C-ABI host layer over the CUDA kernels (include/rr_api.h).
//
// Replaces the OpenCL context / queue / program / buffer plumbing of the
// reference's host driver (src/image.hpp:30-278) and the device selection of
// src/main.cpp:54-157.  No JIT: the kernels ship as sm_100a SASS.  No CPU
// fallback: every entry point that needs a device fails with RR_ERR_NO_DEVICE /
// RR_ERR_CUDA when there is none.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "rr_internal.h"
#include "rr_math.cuh"

namespace rr {

static thread_local std::string g_last_error;

static int fail(int status, const std::string& detail) {
  g_last_error = detail;
  return status;
}
static int cuda_fail(cudaError_t e, const char* where) {
  return fail(e == cudaErrorMemoryAllocation ? RR_ERR_OUT_OF_MEMORY : RR_ERR_CUDA,
              std::string(where) + ": " + cudaGetErrorString(e));
}
#define RR_CUDA(x)                                 \
  do {                                             \
    cudaError_t e_ = (x);                          \
    if (e_ != cudaSuccess) return cuda_fail(e_, #x); \
  } while (0)

// ---- scene memory cache (see rr_internal.h) ---------------------------------------
namespace {
struct CacheBlock { void* p; size_t bytes; int ordinal; };
std::mutex g_cache_mu;
std::vector<CacheBlock> g_parked;                           // free blocks kept for reuse
std::unordered_map<void*, std::pair<size_t, int>> g_live;   // blocks handed out: size, device
}  // namespace

cudaError_t dev_malloc_bytes(void** p, size_t bytes) {
  *p = nullptr;
  bytes = (std::max<size_t>(bytes, 1) + 511) & ~size_t(511);
  int ordinal = 0;
  cudaError_t e = cudaGetDevice(&ordinal);
  if (e != cudaSuccess) return e;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    int best = -1;
    for (int i = 0; i < (int)g_parked.size(); ++i) {  // smallest parked block that fits without wasting more than half
      const CacheBlock& b = g_parked[i];
      if (b.ordinal != ordinal || b.bytes < bytes || b.bytes > 2 * bytes) continue;
      if (best < 0 || b.bytes < g_parked[best].bytes) best = i;
    }
    if (best >= 0) {
      *p = g_parked[best].p;
      g_live[*p] = {g_parked[best].bytes, ordinal};
      g_parked.erase(g_parked.begin() + best);
      return cudaSuccess;
    }
  }
  e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the parked blocks back and try once more
    cudaGetLastError();
    dev_trim(ordinal);
    e = cudaMalloc(p, bytes);
  }
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  g_live[*p] = {bytes, ordinal};
  return cudaSuccess;
}

void dev_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_cache_mu);
  auto it = g_live.find(p);
  if (it == g_live.end()) { cudaFree(p); return; }  // not ours
  g_parked.push_back({p, it->second.first, it->second.second});
  g_live.erase(it);
}

void dev_trim(int ordinal) {
  std::vector<void*> drop;
  {
    std::lock_guard<std::mutex> lock(g_cache_mu);
    for (size_t i = 0; i < g_parked.size();) {
      if (g_parked[i].ordinal == ordinal) { drop.push_back(g_parked[i].p); g_parked.erase(g_parked.begin() + i); }
      else ++i;
    }
  }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ordinal);
  for (void* q : drop) cudaFree(q);
  cudaSetDevice(prev);
}

struct Device {
  int ordinal = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // scene
  rr_triangle* tris = nullptr;
  rr_sphere* spheres = nullptr;
  float* tri_box = nullptr;
  float* sph_box = nullptr;
  Lbvh tb, sb;
  float4 *tri_geom = nullptr, *tri_nrm = nullptr, *sph_geom = nullptr;
  float4* nodes = nullptr;  // triangle hierarchies, then the sphere hierarchy
  float4* top_nodes = nullptr;  // RR_TOP_STAGE: copies of the top nodes of the largest hierarchy
  uint32_t top_count = 0;
  int32_t top_root = -1;
  DMesh* meshes = nullptr;
  DMaterial* materials = nullptr;
  // inputs of k_prepare_meshes, kept so that rr_update_meshes can re-pose the scene without a rebuild
  rr_mesh* meshes_in = nullptr;
  uint32_t *mesh_seg = nullptr, *mesh_pos = nullptr;
  std::vector<uint64_t> entry_count;  // host: primitives of every mesh (+ the sphere set), for the visiting order
  // host copies for frame_needs_slack(): the uploaded MeshInfo array, the largest |coordinate| of every mesh's local
  // box and of the sphere set's box
  std::vector<rr_mesh> h_meshes;
  std::vector<float> h_mesh_absmax;
  float h_sph_absmax = 0.0f;
  int h_sphere_materials = 0;  // RR_FEAT_MATERIALS if a sphere has a Checker / Glassy / Invisible material
  float4* tlas_blocks = nullptr;      // > 32 meshes: the implicit top-level tree, one allocation: block boxes ...
  float4* tlas = nullptr;             // ... then the chunk level and the levels above it (RenderParams::tlas)
  uint32_t* tlas_levels = nullptr;    // device: level count, first box of every level (RenderParams::tlas_levels)
  float build_ms = 0.0f;
  // progressive mode: per-pixel integer sums of the 8-bit frames, three planes (R, G, B) of width*height u32
  uint32_t* accum = nullptr;
  size_t accum_capacity = 0;  // pixels allocated
  uint32_t accum_w = 0, accum_h = 0, accum_frames = 0;
  float accum_ms = 0.0f;      // device time of the last k_accum_add
  // frame
  uint8_t* frame = nullptr;
  size_t frame_bytes = 0;
  float* radiance = nullptr;
  size_t radiance_bytes = 0;
  uint2* stack = nullptr;               // traversal stacks of the render kernel (scratch)
  uint32_t* cold = nullptr;             // cold slot words of the render kernel (scratch)
  uint32_t stack_warps = 0;
  uint32_t stack_entries = 0;           // per slot: 3 per level of the deepest wide hierarchy + slack
  unsigned long long* queue = nullptr;  // local tile counter
  uint32_t* tile_order = nullptr;       // rr_set_tile_order: the launch's tiles in hand-out order (nullptr: row-major)
  uint32_t tile_order_n = 0;
  Counters* counters = nullptr;
  // shared (multi-process) attachments.  The exporter's shared frame is an allocation of its own (never d.frame, which
  // ensure_frame may free and move); shared_bytes bounds what rr_render_shared may write through either mapping.
  unsigned long long* shared_queue = nullptr;
  uint8_t* shared_frame = nullptr;
  size_t shared_bytes = 0;
  uint32_t shared_w = 0, shared_h = 0;
  uint32_t shared_epoch = 0;  // frames rendered through the shared queue since the export / import
  bool shared_imported = false;
  // progress polling (rr_render_progress): a stream of its own, so that the 8-byte copy overtakes the running kernel
  cudaStream_t poll_stream = nullptr;
  unsigned long long* poll_host = nullptr;  // pinned
};

}  // namespace rr

struct rr_ctx {
  std::vector<rr::Device> dev;
  size_t n_tris = 0, n_meshes = 0, n_spheres = 0;
  bool has_scene = false;
  bool peer_ok = false;  // devices 1.. can address device 0's memory
  rr::Tuning tune;
  uint32_t pool_use = 0;  // rr_set_tuning value 8: path slots per warp that take pixels (0 = RR_POOL)
  std::atomic<uint64_t> progress_total{0};  // tiles of the frame being rendered (0: none)
  std::atomic<const unsigned long long*> progress_queue{nullptr};  // the counter its warps pop
  std::atomic<uint32_t> progress_items_per_tile{1};  // the counter counts pixels (RR_PIXEL_QUEUE): this many per tile
};

namespace rr {

// ---- per-mesh records, computed on the device so that the rotation uses the
// same sin/cos as the kernel would (reference src/Trace.cl:452-454 rebuilds it per ray).
__global__ void k_prepare_meshes(const rr_mesh* __restrict__ meshes, const uint32_t* __restrict__ mesh_seg, int n_meshes,
                                 const float* __restrict__ seg_box, const uint32_t* __restrict__ seg_sfirst,
                                 const uint32_t* __restrict__ seg_root, const uint32_t* __restrict__ seg_count, const rr_sphere* __restrict__ spheres,
                                 int n_spheres, const float* __restrict__ sph_seg_box, uint32_t sph_node_base,
                                 const uint32_t* __restrict__ mesh_pos, DMesh* __restrict__ out,
                                 DMaterial* __restrict__ mats) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_meshes + n_spheres) return;
  const rr_material* src;
  if (i < n_meshes) {
    const rr_mesh& m = meshes[i];
    src = &m.material;
    DMesh d;
    const float cx = cos_c(m.pitch), sx = sin_c(m.pitch);
    const float cy = cos_c(m.yaw), sy = sin_c(m.yaw);
    const float cz = cos_c(m.roll), sz = sin_c(m.roll);
    // makeRotation, src/Trace.cl:90-100
    float R[9];
    R[0] = cy * cz; R[1] = cy * sz; R[2] = -sy;
    R[3] = cz * sy * sx - cx * sz; R[4] = cx * cz + sx * sy * sz; R[5] = cy * sx;
    R[6] = sx * sz + cx * cz * sy; R[7] = cx * sy * sz - cz * sx; R[8] = cx * cy;
    const float px = m.pos.s[0], py = m.pos.s[1], pz = m.pos.s[2];
    // transpose_mat, src/Trace.cl:109-116 (rows of the inverse), position in .w
    d.ri0 = make_float4(R[0], R[3], R[6], px);
    d.ri1 = make_float4(R[1], R[4], R[7], py);
    d.ri2 = make_float4(R[2], R[5], R[8], pz);
    const uint32_t sbits = __float_as_uint(m.scale);
    const uint32_t sexp = (sbits >> 23) & 0xffu;
    const bool pow2 = m.scale > 0.0f && (sbits & 0x007fffffu) == 0u && sexp > 64u && sexp < 190u;
    d.r0 = make_float4(R[0], R[1], R[2], m.scale);
    d.r1 = make_float4(R[3], R[4], R[5], pow2 ? 1.0f / m.scale : 0.0f);
    d.r2 = make_float4(R[6], R[7], R[8], 0.0f);
    const uint32_t s = mesh_seg[i];
    float sb[6];
    for (int k = 0; k < 6; ++k) sb[k] = seg_box[6 * s + k];
    const float delta = box_delta(sb);
    const uint32_t count = seg_count[s];
    const int t = m.material.type;
    uint32_t flags = (uint32_t)t << RR_MF_TYPE_SHIFT;
    if (t != RR_MATERIAL_GLASSY && t != RR_MATERIAL_INVISIBLE && t != RR_MATERIAL_ONESIDED) flags |= RR_MF_CULL;  // :460-462
    if (m.scale <= RR_EPSILON || count == 0) flags |= RR_MF_SKIP;                                             // :448
    if (pow2) flags |= RR_MF_POW2;
    if (m.scale == 1.0f) flags |= RR_MF_UNIT;
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k) { lo[k] = sb[k] - delta; hi[k] = sb[3 + k] + delta; }
    // .w: the root node of the segment's hierarchy, or (count <= RR_DIRECT_MAX: no hierarchy) its first sorted slot
    d.bmin = make_float4(lo[0], lo[1], lo[2], __uint_as_float(count > RR_DIRECT_MAX ? seg_root[s] : seg_sfirst[s]));
    d.bmax = make_float4(hi[0], hi[1], hi[2], __uint_as_float(count));
    // world-space box of the (inflated) local root box: LocalToWorldHit of its 8 corners, then a generous slack
    float wlo[3] = {INFINITY, INFINITY, INFINITY}, whi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (!(flags & RR_MF_SKIP)) {
      for (int c = 0; c < 8; ++c) {
        const float lx = ((c & 1) ? hi[0] : lo[0]) * m.scale, ly = ((c & 2) ? hi[1] : lo[1]) * m.scale,
                    lz = ((c & 4) ? hi[2] : lo[2]) * m.scale;
        const float w[3] = {R[0] * lx + R[1] * ly + R[2] * lz + px, R[3] * lx + R[4] * ly + R[5] * lz + py,
                            R[6] * lx + R[7] * ly + R[8] * lz + pz};
        for (int k = 0; k < 3; ++k) { wlo[k] = fminf(wlo[k], w[k]); whi[k] = fmaxf(whi[k], w[k]); }
      }
      float mabs = 0.0f;
      for (int k = 0; k < 3; ++k) mabs = fmaxf(mabs, fmaxf(fabsf(wlo[k]), fabsf(whi[k])));
      const float e = mabs * 1.220703125e-4f + 1.0e-6f;  // 2^-13 relative
      for (int k = 0; k < 3; ++k) { wlo[k] -= e; whi[k] += e; }
    }
    d.wmin = make_float4(wlo[0], wlo[1], wlo[2], __uint_as_float(flags));
    d.wmax = make_float4(whi[0], whi[1], whi[2], __int_as_float(i));
    out[mesh_pos[i]] = d;  // the kernel visits the meshes in mesh_pos order (largest first)
  } else {
    src = &spheres[i - n_meshes].material;
    if (i == n_meshes) {  // the sphere set as a world-space pseudo-mesh
      DMesh d;
      d.ri0 = make_float4(1, 0, 0, 0); d.ri1 = make_float4(0, 1, 0, 0); d.ri2 = make_float4(0, 0, 1, 0);
      d.r0 = make_float4(1, 0, 0, 1); d.r1 = make_float4(0, 1, 0, 1); d.r2 = make_float4(0, 0, 1, 0);
      float sb[6];
      for (int k = 0; k < 6; ++k) sb[k] = sph_seg_box[k];
      const float delta = box_delta(sb);
      d.bmin = make_float4(sb[0] - delta, sb[1] - delta, sb[2] - delta, __uint_as_float(sph_node_base));
      d.bmax = make_float4(sb[3] + delta, sb[4] + delta, sb[5] + delta, __uint_as_float((uint32_t)n_spheres));
      d.wmin = make_float4(d.bmin.x, d.bmin.y, d.bmin.z, __uint_as_float(RR_MF_SPHERES | RR_MF_UNIT | RR_MF_POW2));
      d.wmax = make_float4(d.bmax.x, d.bmax.y, d.bmax.z, __int_as_float(n_meshes));
      out[mesh_pos[n_meshes]] = d;
    }
  }
  DMaterial mm;
  mm.type = src->type;
  mm.ior = src->ior;
  mm.emissionStrength = src->emissionStrength;
  mm.reflectiveness = src->reflectiveness;
  mm.specularProbability = src->specularProbability;
  for (int k = 0; k < 3; ++k) { mm.color[k] = src->color.s[k]; mm.emissionColor[k] = src->emissionColor.s[k]; }
  mm.pad = 0.0f;
  mats[i] = mm;
}

// Triangle assembly on the device (replaces the host loop of src/readobj.hpp:313-343): corner indices -> the 96-byte
// Triangle records, six coalesced 16-byte stores per triangle.  HBM-bound: 24 B of indices + 72 B of gathered vertex
// data read, 96 B written per triangle.
__global__ void k_gather_tris(const float* __restrict__ pos, const float* __restrict__ nrm, const uint32_t* __restrict__ corners,
                              uint64_t n, float4* __restrict__ tris) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* c = corners + 6 * i;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* v = pos + 3 * (uint64_t)__ldg(c + k);
    tris[6 * i + k] = make_float4(__ldg(v), __ldg(v + 1), __ldg(v + 2), 0.0f);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* v = nrm + 3 * (uint64_t)__ldg(c + 3 + k);
    tris[6 * i + 3 + k] = make_float4(__ldg(v), __ldg(v + 1), __ldg(v + 2), 0.0f);
  }
}

// Progressive mode: the host loop of src/main.cpp:575-582 as one byte-streaming kernel.  Four pixels per thread:
// the new frame's RGBA8 quad (16 B) is added to the three u32 sum planes (16 B each, read + write) and the quad
// is overwritten IN PLACE with sum / frames (integer division, alpha 255).  HBM-bound: 4 + 12 + 12 + 4 = 32 B per pixel.
__global__ void __launch_bounds__(256) k_accum_add(uint32_t* __restrict__ frame, uint32_t* __restrict__ sum_r,
                                                   uint32_t* __restrict__ sum_g, uint32_t* __restrict__ sum_b,
                                                   size_t n_pixels, uint32_t frames) {
  const size_t n_quads = n_pixels / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += stride) {
    const uint4 px = reinterpret_cast<const uint4*>(frame)[q];
    uint4 r = reinterpret_cast<uint4*>(sum_r)[q], g = reinterpret_cast<uint4*>(sum_g)[q], b = reinterpret_cast<uint4*>(sum_b)[q];
    r.x += px.x & 0xffu; g.x += (px.x >> 8) & 0xffu; b.x += (px.x >> 16) & 0xffu;
    r.y += px.y & 0xffu; g.y += (px.y >> 8) & 0xffu; b.y += (px.y >> 16) & 0xffu;
    r.z += px.z & 0xffu; g.z += (px.z >> 8) & 0xffu; b.z += (px.z >> 16) & 0xffu;
    r.w += px.w & 0xffu; g.w += (px.w >> 8) & 0xffu; b.w += (px.w >> 16) & 0xffu;
    reinterpret_cast<uint4*>(sum_r)[q] = r;
    reinterpret_cast<uint4*>(sum_g)[q] = g;
    reinterpret_cast<uint4*>(sum_b)[q] = b;
    uint4 avg;
    avg.x = (r.x / frames) | ((g.x / frames) << 8) | ((b.x / frames) << 16) | 0xff000000u;
    avg.y = (r.y / frames) | ((g.y / frames) << 8) | ((b.y / frames) << 16) | 0xff000000u;
    avg.z = (r.z / frames) | ((g.z / frames) << 8) | ((b.z / frames) << 16) | 0xff000000u;
    avg.w = (r.w / frames) | ((g.w / frames) << 8) | ((b.w / frames) << 16) | 0xff000000u;
    reinterpret_cast<uint4*>(frame)[q] = avg;
  }
  // ragged tail (n_pixels % 4 pixels)
  const size_t i = n_quads * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pixels) {
    const uint32_t px = frame[i];
    const uint32_t r = sum_r[i] + (px & 0xffu), g = sum_g[i] + ((px >> 8) & 0xffu), b = sum_b[i] + ((px >> 16) & 0xffu);
    sum_r[i] = r; sum_g[i] = g; sum_b[i] = b;
    frame[i] = (r / frames) | ((g / frames) << 8) | ((b / frames) << 16) | 0xff000000u;
  }
}

// ---- measured roofline denominators (rr_probe_peak; bench.py) --------------------------------------------------
// MEASURED_PEAKS.json holds an HBM copy and a bf16 GEMM figure; the two roofs this path is read against -- FP32 FMA
// issue and L2 bandwidth -- are measured here, on the device the bench runs on, by two plain kernels.
__global__ void __launch_bounds__(256) k_peak_fma(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.0f, a2 = a0 + 2.0f, a3 = a0 + 3.0f, a4 = a0 + 4.0f, a5 = a0 + 5.0f, a6 = a0 + 6.0f, a7 = a0 + 7.0f;
  const float m = 0.999f, c = 1e-3f;
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {  // 8 independent FFMA chains per thread; unrolled so that the loop counter and branch are < 2 % of the issue slots
    a0 = __fmaf_rn(a0, m, c); a1 = __fmaf_rn(a1, m, c); a2 = __fmaf_rn(a2, m, c); a3 = __fmaf_rn(a3, m, c);
    a4 = __fmaf_rn(a4, m, c); a5 = __fmaf_rn(a5, m, c); a6 = __fmaf_rn(a6, m, c); a7 = __fmaf_rn(a7, m, c);
  }
  const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 12345.678f) out[0] = r;  // keeps the chains alive
}
__global__ void __launch_bounds__(256) k_peak_l2(const uint4* __restrict__ buf, size_t n16, int passes, unsigned* out) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int pass = 0; pass < passes; ++pass)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
      const uint4 v = __ldcg(buf + i);  // L2 only: the L1 would not hold 32 MB anyway
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) out[0] = acc;
}

struct IndexedInput {  // host arrays of rr_upload_scene_indexed
  const float* positions = nullptr;
  size_t n_positions = 0;
  const float* normals = nullptr;
  size_t n_normals = 0;
  const uint32_t* corners = nullptr;
};

static void free_scene(Device& d) {
  cudaSetDevice(d.ordinal);
  dev_free(d.tris); dev_free(d.spheres); dev_free(d.tri_box); dev_free(d.sph_box);
  dev_free(d.tri_geom); dev_free(d.tri_nrm); dev_free(d.sph_geom); dev_free(d.meshes); dev_free(d.materials);
  dev_free(d.top_nodes); d.top_nodes = nullptr; d.top_count = 0; d.top_root = -1;
  dev_free(d.nodes); dev_free(d.meshes_in); dev_free(d.mesh_seg); dev_free(d.mesh_pos); dev_free(d.tlas_blocks); dev_free(d.tlas_levels);
  d.meshes_in = nullptr; d.mesh_seg = nullptr; d.mesh_pos = nullptr; d.tlas_blocks = nullptr; d.tlas = nullptr; d.tlas_levels = nullptr;
  d.tris = nullptr; d.spheres = nullptr; d.tri_box = nullptr; d.sph_box = nullptr;
  d.tri_geom = d.tri_nrm = d.sph_geom = nullptr; d.meshes = nullptr; d.materials = nullptr; d.nodes = nullptr;
  lbvh_free(d.tb);
  lbvh_free(d.sb);
}

struct SegPlan {
  std::vector<uint32_t> first, count;  // sorted, non-overlapping, unique
  std::vector<uint32_t> mesh_seg;      // mesh -> segment
};

// Argument checks shared by the three upload entry points.
static int check_upload_counts(size_t n_tris, size_t n_meshes, size_t n_spheres) {
  if (n_tris >= RR_MAX_PRIMS || n_spheres >= RR_MAX_PRIMS) return fail(RR_ERR_INVALID_ARGUMENT, "too many primitives (a leaf reference keeps the slot in 29 bits)");
  if (n_meshes >= 0x1fffff00ull) return fail(RR_ERR_INVALID_ARGUMENT, "too many meshes (the slot word keeps mesh + 1 in 29 bits)");
  return RR_OK;
}

static int plan_segments(const rr_mesh_range* ranges, size_t n_meshes, size_t n_tris, SegPlan& plan) {
  struct R { uint64_t first, count; };
  std::vector<R> uniq;
  for (size_t i = 0; i < n_meshes; ++i) {
    if (ranges[i].firstTriangle > n_tris || ranges[i].numTriangles > n_tris - ranges[i].firstTriangle)
      return fail(RR_ERR_BAD_MESH_RANGE, "mesh " + std::to_string(i) + ": triangle range outside the array");
    uniq.push_back({ranges[i].firstTriangle, ranges[i].numTriangles});
  }
  std::sort(uniq.begin(), uniq.end(), [](const R& a, const R& b) { return a.first != b.first ? a.first < b.first : a.count < b.count; });
  uniq.erase(std::unique(uniq.begin(), uniq.end(), [](const R& a, const R& b) { return a.first == b.first && a.count == b.count; }), uniq.end());
  for (size_t k = 1; k < uniq.size(); ++k)
    if (uniq[k - 1].first + uniq[k - 1].count > uniq[k].first)
      return fail(RR_ERR_BAD_MESH_RANGE, "mesh triangle ranges overlap without being identical");
  plan.first.clear(); plan.count.clear();
  for (auto& r : uniq) { plan.first.push_back((uint32_t)r.first); plan.count.push_back((uint32_t)r.count); }
  plan.mesh_seg.resize(n_meshes);
  for (size_t i = 0; i < n_meshes; ++i) {  // uniq is sorted by (first, count): binary search
    const R want{ranges[i].firstTriangle, ranges[i].numTriangles};
    const auto it = std::lower_bound(uniq.begin(), uniq.end(), want,
                                     [](const R& a, const R& b) { return a.first != b.first ? a.first < b.first : a.count < b.count; });
    plan.mesh_seg[i] = (uint32_t)(it - uniq.begin());
  }
  return RR_OK;
}

// World boxes of the implicit top-level tree (RenderParams::tlas_blocks / tlas): output box w is the union of `fan`
// consecutive input boxes -- mesh world boxes (skipped meshes excluded) for the blocks, boxes of the level below
// otherwise.  An empty union is stored as a NaN box, which no ray enters.  A few thousand boxes at most per launch.
__global__ void k_tlas_boxes(const DMesh* __restrict__ meshes, const float4* __restrict__ in_boxes, int n_in, int fan,
                             float4* __restrict__ out, int n_out) {
  const int w = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (w >= n_out) return;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = fan * w; i < min(fan * w + fan, n_in); ++i) {
    float4 a, b;
    bool use;
    if (meshes) {
      a = meshes[i].wmin; b = meshes[i].wmax;
      use = !(__float_as_uint(a.w) & RR_MF_SKIP);
    } else {
      a = in_boxes[2 * i]; b = in_boxes[2 * i + 1];
      use = a.x == a.x;  // not an empty (NaN) box
    }
    if (use) {
      lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
      hi[0] = fmaxf(hi[0], b.x); hi[1] = fmaxf(hi[1], b.y); hi[2] = fmaxf(hi[2], b.z);
    }
  }
  const bool empty = lo[0] > hi[0];
  const float q = __int_as_float(0x7fc00000);
  out[2 * w] = empty ? make_float4(q, q, q, 0.0f) : make_float4(lo[0], lo[1], lo[2], 0.0f);
  out[2 * w + 1] = empty ? make_float4(q, q, q, 0.0f) : make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// Tight world boxes of ROTATED meshes.  The world box k_prepare_meshes derives from the eight corners of the local root
// box is exact for an axis-aligned mesh but up to sqrt(3) too wide per axis for a rotated one, and every ray that
// enters it pays a mesh entry (transform, root test, finish).  Here the vertices themselves go through
// LocalToWorldHit (src/Trace.cl:139-147): grid.x = mesh, grid.y = 1024-triangle chunk of its range, block-reduced
// min / max into ordered-int atomics.  The hierarchy and the result arithmetic are untouched: the box only culls.
__device__ __forceinline__ int wb_f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float wb_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_world_box_init(int* __restrict__ box_ord, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * n) box_ord[i] = (i % 6) < 3 ? wb_f2ord(INFINITY) : wb_f2ord(-INFINITY);
}

__global__ void __launch_bounds__(256) k_world_box_accum(const rr_triangle* __restrict__ tris, const rr_mesh* __restrict__ meshes,
                                                         const uint32_t* __restrict__ mesh_seg, const uint32_t* __restrict__ seg_first,
                                                         const uint32_t* __restrict__ seg_count, const uint32_t* __restrict__ mesh_pos,
                                                         const DMesh* __restrict__ dm, int* __restrict__ box_ord) {
  const uint32_t i = blockIdx.x;
  const DMesh& d = dm[mesh_pos[i]];
  const float4 r0 = d.r0, r1 = d.r1, r2 = d.r2;
  if (__float_as_uint(d.wmin.w) & RR_MF_SKIP) return;
  if (r0.x == 1.0f && r0.y == 0.0f && r0.z == 0.0f && r1.x == 0.0f && r1.y == 1.0f && r1.z == 0.0f && r2.x == 0.0f && r2.y == 0.0f &&
      r2.z == 1.0f)
    return;  // axis-aligned: the corner box is exact
  const float scale = meshes[i].scale;
  const float px = d.ri0.w, py = d.ri1.w, pz = d.ri2.w;
  const uint32_t first = seg_first[mesh_seg[i]], count = seg_count[mesh_seg[i]];
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint32_t base = blockIdx.y * 1024u; base < count; base += gridDim.y * 1024u) {
    for (uint32_t t = base + threadIdx.x; t < min(base + 1024u, count); t += 256u) {
      const rr_triangle& tr = tris[first + t];
      const rr_float3* v[3] = {&tr.posA, &tr.posB, &tr.posC};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float lx = v[k]->s[0] * scale, ly = v[k]->s[1] * scale, lz = v[k]->s[2] * scale;
        const float w[3] = {r0.x * lx + r0.y * ly + r0.z * lz + px, r1.x * lx + r1.y * ly + r1.z * lz + py,
                            r2.x * lx + r2.y * ly + r2.z * lz + pz};
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], w[a]); hi[a] = fmaxf(hi[a], w[a]); }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
  if ((threadIdx.x & 31) == 0 && lo[0] <= hi[0]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(box_ord + 6 * i + a, wb_f2ord(lo[a]));
      atomicMax(box_ord + 6 * i + 3 + a, wb_f2ord(hi[a]));
    }
  }
}

// The tight box (with the same slack as the corner box) replaces the corner box where it is smaller.
__global__ void k_world_box_apply(const int* __restrict__ box_ord, const uint32_t* __restrict__ mesh_pos, int n_meshes, DMesh* __restrict__ dm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_meshes) return;
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) { lo[a] = wb_ord2f(box_ord[6 * i + a]); hi[a] = wb_ord2f(box_ord[6 * i + 3 + a]); }
  if (!(lo[0] <= hi[0])) return;  // not a rotated mesh (or no triangles)
  float mabs = 0.0f;
  for (int a = 0; a < 3; ++a) mabs = fmaxf(mabs, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
  const float e = mabs * 1.220703125e-4f + 1.0e-6f;  // 2^-13 relative, as in k_prepare_meshes
  DMesh& d = dm[mesh_pos[i]];
  d.wmin.x = fmaxf(d.wmin.x, lo[0] - e); d.wmin.y = fmaxf(d.wmin.y, lo[1] - e); d.wmin.z = fmaxf(d.wmin.z, lo[2] - e);
  d.wmax.x = fminf(d.wmax.x, hi[0] + e); d.wmax.y = fminf(d.wmax.y, hi[1] + e); d.wmax.z = fminf(d.wmax.z, hi[2] + e);
}

static inline uint64_t spread21(uint64_t v) {  // 21 bits -> every third bit
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

// Per-mesh records and the material table from the MeshInfo array on the device (upload and rr_update_meshes), in
// the order the kernel visits them.  Up to 32 entries (meshes + the sphere set): one chunk, most primitives first (a
// hit there prunes the small ones by their world box).  More: Morton order of the world-box centres, so that the
// 32-mesh chunks are compact and the top level (k_tlas_boxes) can cull them; inside a chunk, most primitives first.
// Equal world distances are resolved by the original mesh index in the kernel: the order never changes a result.
static int prepare_meshes(Device& d, size_t n_meshes, size_t n_spheres) {
  const size_t n_entries = n_meshes + (n_spheres ? 1 : 0);
  if (n_meshes + n_spheres == 0) return RR_OK;
  cudaStream_t st = d.stream;
  const int n = (int)(n_meshes + n_spheres);
  auto launch = [&]() -> cudaError_t {
    k_prepare_meshes<<<(n + 127) / 128, 128, 0, st>>>(d.meshes_in, d.mesh_seg, (int)n_meshes, d.tb.seg_box, d.tb.seg_sfirst, d.tb.seg_root,
                                                      d.tb.seg_count, d.spheres, (int)n_spheres, d.sb.seg_box, (uint32_t)d.tb.n_nodes,
                                                      d.mesh_pos, d.meshes, d.materials);
    return cudaGetLastError();
  };
  // tight world boxes of the rotated meshes: accumulated once (they do not depend on the visiting order), applied
  // after every k_prepare_meshes launch
  struct Tmp { int* p = nullptr; ~Tmp() { dev_free(p); } } box_ord;
  bool have_tight = false;
  auto tighten = [&]() -> cudaError_t {
    if (n_meshes == 0 || getenv("RR_NO_TIGHT_BOXES")) return cudaSuccess;
    if (!have_tight) {
      cudaError_t e = dev_malloc(&box_ord.p, n_meshes * 6 * sizeof(int));
      if (e != cudaSuccess) return e;
      k_world_box_init<<<(unsigned)((6 * n_meshes + 255) / 256), 256, 0, st>>>(box_ord.p, (int)n_meshes);
      uint64_t max_count = 0;
      for (size_t k = 0; k < n_meshes; ++k) max_count = std::max(max_count, d.entry_count[k]);
      const dim3 grid((unsigned)n_meshes, (unsigned)std::min<uint64_t>(std::max<uint64_t>((max_count + 1023) / 1024, 1), 65535));
      k_world_box_accum<<<grid, 256, 0, st>>>(d.tris, d.meshes_in, d.mesh_seg, d.tb.seg_first, d.tb.seg_count, d.mesh_pos, d.meshes,
                                              box_ord.p);
      have_tight = true;
    }
    k_world_box_apply<<<(unsigned)((n_meshes + 127) / 128), 128, 0, st>>>(box_ord.p, d.mesh_pos, (int)n_meshes, d.meshes);
    return cudaGetLastError();
  };
  std::vector<uint32_t> pos(n_meshes + 1, 0);
  std::vector<uint32_t> order(n_entries);  // order[k] = entry at position k
  for (size_t k = 0; k < n_entries; ++k) order[k] = (uint32_t)k;
  auto by_count = [&](uint32_t a, uint32_t b) { return d.entry_count[a] > d.entry_count[b]; };
  if (n_entries <= 32) {
    std::stable_sort(order.begin(), order.end(), by_count);
  } else {
    // pass 1 in upload order, to learn the world boxes
    for (size_t k = 0; k <= n_meshes; ++k) pos[k] = (uint32_t)k;
    RR_CUDA(cudaMemcpyAsync(d.mesh_pos, pos.data(), (n_meshes + 1) * 4, cudaMemcpyHostToDevice, st));
    RR_CUDA(launch());
    RR_CUDA(tighten());
    std::vector<float> wb(n_entries * 8);
    RR_CUDA(cudaMemcpy2DAsync(wb.data(), 32, reinterpret_cast<const char*>(d.meshes) + offsetof(DMesh, wmin), sizeof(DMesh), 32,
                              n_entries, cudaMemcpyDeviceToHost, st));
    RR_CUDA(cudaStreamSynchronize(st));
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    std::vector<char> skip(n_entries);
    for (size_t k = 0; k < n_entries; ++k) {
      uint32_t flags;
      memcpy(&flags, &wb[8 * k + 3], 4);
      skip[k] = (flags & RR_MF_SKIP) != 0;
      if (skip[k]) continue;
      for (int a = 0; a < 3; ++a) {
        const float c = 0.5f * (wb[8 * k + a] + wb[8 * k + 4 + a]);
        if (c == c && std::fabs(c) < 3.0e38f) { lo[a] = std::min(lo[a], c); hi[a] = std::max(hi[a], c); }
      }
    }
    std::vector<uint64_t> key(n_entries, ~0ull);  // skipped meshes go last
    for (size_t k = 0; k < n_entries; ++k) {
      if (skip[k]) continue;
      uint64_t code = 0;
      for (int a = 0; a < 3; ++a) {
        const float c = 0.5f * (wb[8 * k + a] + wb[8 * k + 4 + a]);
        const float ext = hi[a] - lo[a];
        double u = (ext > 0.0f && c == c) ? ((double)c - lo[a]) / ext : 0.0;
        u = std::min(std::max(u, 0.0), 1.0);
        code |= spread21((uint64_t)(u * 2097151.0)) << (2 - a);
      }
      key[k] = code;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    for (size_t c = 0; c < n_entries; c += 8) std::stable_sort(order.begin() + c, order.begin() + std::min(c + 8, n_entries), by_count);
  }
  for (size_t k = 0; k < n_entries; ++k) pos[order[k]] = (uint32_t)k;
  RR_CUDA(cudaMemcpyAsync(d.mesh_pos, pos.data(), (n_meshes + 1) * 4, cudaMemcpyHostToDevice, st));
  RR_CUDA(launch());
  RR_CUDA(tighten());
  uint32_t level_info[1 + RR_TLAS_MAX_LEVELS] = {0};
  if (n_entries > 32) {
    // blocks of 8 meshes, chunks of 4 blocks, then 4 boxes per box until one box is left
    const int n_blocks = (int)((n_entries + 7) / 8);
    std::vector<int> count;  // boxes per level of RenderParams::tlas
    for (int n = (n_blocks + 3) / 4;; n = (n + 3) / 4) {
      count.push_back(n);
      if (n == 1 || (int)count.size() == RR_TLAS_MAX_LEVELS) break;
    }
    size_t total = (size_t)n_blocks;
    for (int c : count) total += (size_t)c;
    if (!d.tlas_blocks) RR_CUDA(dev_malloc(&d.tlas_blocks, total * 2 * sizeof(float4)));
    d.tlas = d.tlas_blocks + 2 * (size_t)n_blocks;
    k_tlas_boxes<<<(n_blocks + 127) / 128, 128, 0, st>>>(d.meshes, nullptr, (int)n_entries, 8, d.tlas_blocks, n_blocks);
    RR_CUDA(cudaGetLastError());
    const float4* in = d.tlas_blocks;
    int n_in = n_blocks;
    uint32_t off = 0;
    for (size_t l = 0; l < count.size(); ++l) {
      float4* out = d.tlas + 2 * (size_t)off;
      k_tlas_boxes<<<(count[l] + 127) / 128, 128, 0, st>>>(nullptr, in, n_in, 4, out, count[l]);
      RR_CUDA(cudaGetLastError());
      level_info[1 + l] = off;
      off += (uint32_t)count[l];
      in = out;
      n_in = count[l];
    }
    level_info[0] = (uint32_t)count.size();
    if (!d.tlas_levels) RR_CUDA(dev_malloc(&d.tlas_levels, sizeof(level_info)));
    RR_CUDA(cudaMemcpyAsync(d.tlas_levels, level_info, sizeof(level_info), cudaMemcpyHostToDevice, st));
  }
  RR_CUDA(cudaStreamSynchronize(st));  // `pos` is pageable host memory read by the copies above
  return RR_OK;
}

static void upload_mark(cudaStream_t st, const char* what) {  // RR_BUILD_TIMING=1: see rr_lbvh.cu build_mark
  static const bool on = getenv("RR_BUILD_TIMING") != nullptr;
  static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  if (!on) return;
  cudaStreamSynchronize(st);
  const auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[upload] %-40s %8.1f us\n", what, std::chrono::duration<double, std::micro>(now - last).count());
  last = now;
}

static int upload_device(Device& d, const rr_triangle* tris, const IndexedInput* indexed, size_t n_tris, const rr_mesh* meshes,
                         size_t n_meshes, const SegPlan& plan, const rr_sphere* spheres, size_t n_spheres) {
  RR_CUDA(cudaSetDevice(d.ordinal));
  upload_mark(d.stream, "(since the last upload)");
  free_scene(d);
  cudaStream_t st = d.stream;
  RR_CUDA(dev_malloc(&d.tris, std::max<size_t>(n_tris, 1) * sizeof(rr_triangle)));
  RR_CUDA(dev_malloc(&d.tri_box, std::max<size_t>(n_tris, 1) * 24));
  RR_CUDA(dev_malloc(&d.spheres, std::max<size_t>(n_spheres, 1) * sizeof(rr_sphere)));
  RR_CUDA(dev_malloc(&d.sph_box, std::max<size_t>(n_spheres, 1) * 24));
  struct Staged {  // the raw OBJ arrays on the device, only until the triangles are assembled
    float *pos = nullptr, *nrm = nullptr;
    uint32_t* corners = nullptr;
    ~Staged() { dev_free(pos); dev_free(nrm); dev_free(corners); }
  } staged;
  if (n_tris && !indexed) {
    RR_CUDA(cudaMemcpyAsync(d.tris, tris, n_tris * sizeof(rr_triangle), cudaMemcpyHostToDevice, st));
  } else if (n_tris) {
    RR_CUDA(dev_malloc(&staged.pos, indexed->n_positions * 12));
    RR_CUDA(dev_malloc(&staged.nrm, indexed->n_normals * 12));
    RR_CUDA(dev_malloc(&staged.corners, n_tris * 24));
    RR_CUDA(cudaMemcpyAsync(staged.pos, indexed->positions, indexed->n_positions * 12, cudaMemcpyHostToDevice, st));
    RR_CUDA(cudaMemcpyAsync(staged.nrm, indexed->normals, indexed->n_normals * 12, cudaMemcpyHostToDevice, st));
    RR_CUDA(cudaMemcpyAsync(staged.corners, indexed->corners, n_tris * 24, cudaMemcpyHostToDevice, st));
    k_gather_tris<<<(unsigned)((n_tris + 255) / 256), 256, 0, st>>>(staged.pos, staged.nrm, staged.corners, n_tris,
                                                                    reinterpret_cast<float4*>(d.tris));
    RR_CUDA(cudaGetLastError());
  }
  if (n_spheres) RR_CUDA(cudaMemcpyAsync(d.spheres, spheres, n_spheres * sizeof(rr_sphere), cudaMemcpyHostToDevice, st));
  upload_mark(st, "(free, alloc) + H2D of the scene");
  RR_CUDA(cudaEventRecord(d.ev0, st));
  RR_CUDA(launch_tri_boxes(d.tris, n_tris, d.tri_box, st));
  // primitives per leaf (1 .. RR_LEAF_MAX); RR_LEAF_MAX_PRIMS in the environment overrides the default for A/B runs
  uint32_t leaf_max = RR_LEAF_DEFAULT;
  if (const char* e = getenv("RR_LEAF_MAX_PRIMS")) leaf_max = (uint32_t)std::max(1, atoi(e));
  uint32_t top_cluster = RR_TOP_CLUSTER_DEFAULT;  // RR_TOP_CLUSTER in the environment overrides it (0: plain Karras top) for A/B runs
  if (const char* e = getenv("RR_TOP_CLUSTER")) top_cluster = (uint32_t)std::max(0, atoi(e));
  RR_CUDA(lbvh_build(d.tb, d.tri_box, n_tris, plan.first.data(), plan.count.data(), (uint32_t)plan.first.size(), 0, leaf_max, top_cluster, st));
  upload_mark(st, "triangle LBVH");
  RR_CUDA(dev_malloc(&d.tri_geom, std::max<uint64_t>(d.tb.n, 1) * 48));
  RR_CUDA(dev_malloc(&d.tri_nrm, std::max<uint64_t>(d.tb.n, 1) * 48));
  RR_CUDA(launch_pack_tris(d.tris, d.tb.order, d.tb.n, d.tri_geom, d.tri_nrm, st));
  RR_CUDA(launch_sphere_boxes(d.spheres, n_spheres, d.sph_box, st));
  uint32_t sf = 0, sc = (uint32_t)n_spheres;
  uint32_t sphere_leaf_max = RR_LEAF_DEFAULT_SPHERES;
  if (const char* e = getenv("RR_LEAF_MAX_SPHERES")) sphere_leaf_max = (uint32_t)std::max(1, atoi(e));
  RR_CUDA(lbvh_build(d.sb, d.sph_box, n_spheres, &sf, &sc, n_spheres ? 1u : 0u, (int32_t)d.tb.n_nodes, sphere_leaf_max, 0, st));
  RR_CUDA(dev_malloc(&d.sph_geom, std::max<size_t>(n_spheres, 1) * 16));
  RR_CUDA(launch_pack_spheres(d.spheres, d.sb.order, d.sb.n, d.sph_geom, st));
  upload_mark(st, "pack triangles, sphere LBVH");
  // one node array: triangle hierarchies at [0, tb.n_nodes) (Karras slots, then the SAH tops), the sphere hierarchy behind them
  const size_t node_bytes = RR_NODE_QUADS * sizeof(float4);
  RR_CUDA(dev_malloc(&d.nodes, std::max<uint64_t>(d.tb.n_nodes + d.sb.n_nodes, 1) * node_bytes));
  if (d.tb.n_nodes) RR_CUDA(cudaMemcpyAsync(d.nodes, d.tb.nodes, d.tb.n_nodes * node_bytes, cudaMemcpyDeviceToDevice, st));
  if (d.sb.n_nodes) RR_CUDA(cudaMemcpyAsync(d.nodes + RR_NODE_QUADS * d.tb.n_nodes, d.sb.nodes, d.sb.n_nodes * node_bytes, cudaMemcpyDeviceToDevice, st));
#if RR_TOP_STAGE
  {  // the root of the largest triangle hierarchy and its inner children, copied out for the kernel to stage
    size_t big = 0;
    uint64_t sfirst = 0, big_sfirst = 0;
    for (size_t k = 0; k < plan.count.size(); ++k) {
      if (plan.count[k] > plan.count[big]) { big = k; big_sfirst = sfirst; }
      sfirst += plan.count[k];
    }
    if (!plan.count.empty() && plan.count[big] > RR_DIRECT_MAX) {
      RR_CUDA(cudaStreamSynchronize(st));
      uint32_t root = 0;  // the Karras root of the segment, or the root of its SAH-ordered top
      RR_CUDA(cudaMemcpy(&root, d.tb.seg_root + big, 4, cudaMemcpyDeviceToHost));
      big_sfirst = root;
      std::vector<float4> top(RR_NODE_QUADS * (size_t)RR_TOP_STAGE);
      RR_CUDA(cudaMemcpy(top.data(), d.nodes + RR_NODE_QUADS * big_sfirst, 128, cudaMemcpyDeviceToHost));
      uint32_t n_top = 1;
      int32_t refs[4], cnt;
      memcpy(refs, &top[6], 16);
      memcpy(&cnt, &top[7].x, 4);
      for (int k = 0; k < 4 && n_top < (uint32_t)RR_TOP_STAGE; ++k) {
        if (k >= cnt || refs[k] < 0) continue;  // unused child slot, or a leaf
        RR_CUDA(cudaMemcpy(top.data() + RR_NODE_QUADS * (size_t)n_top, d.nodes + RR_NODE_QUADS * (size_t)refs[k], 128, cudaMemcpyDeviceToHost));
        refs[k] = RR_TOP_TAG + (int32_t)n_top;
        ++n_top;
      }
      memcpy(&top[6], refs, 16);
      RR_CUDA(dev_malloc(&d.top_nodes, sizeof(float4) * top.size()));
      RR_CUDA(cudaMemcpy(d.top_nodes, top.data(), sizeof(float4) * top.size(), cudaMemcpyHostToDevice));
      d.top_count = n_top;
      d.top_root = (int32_t)big_sfirst;
    }
  }
#endif
  // mesh + material tables (the inputs stay on the device for rr_update_meshes; free_scene releases them)
  rr_mesh*& d_meshes_in = d.meshes_in;
  uint32_t *&d_mesh_seg = d.mesh_seg, *&d_mesh_pos = d.mesh_pos;
  d.entry_count.assign(n_meshes + (n_spheres ? 1 : 0), 0);
  for (size_t i = 0; i < n_meshes; ++i) d.entry_count[i] = plan.count[plan.mesh_seg[i]];
  if (n_spheres) d.entry_count[n_meshes] = n_spheres;
  RR_CUDA(dev_malloc(&d_mesh_pos, (n_meshes + 1) * 4));
  RR_CUDA(dev_malloc(&d_meshes_in, std::max<size_t>(n_meshes, 1) * sizeof(rr_mesh)));
  RR_CUDA(dev_malloc(&d_mesh_seg, std::max<size_t>(n_meshes, 1) * 4));
  RR_CUDA(dev_malloc(&d.meshes, (n_meshes + 1) * sizeof(DMesh)));
  RR_CUDA(dev_malloc(&d.materials, std::max<size_t>(n_meshes + n_spheres, 1) * sizeof(DMaterial)));
  if (n_meshes) {
    RR_CUDA(cudaMemcpyAsync(d_meshes_in, meshes, n_meshes * sizeof(rr_mesh), cudaMemcpyHostToDevice, st));
    RR_CUDA(cudaMemcpyAsync(d_mesh_seg, plan.mesh_seg.data(), n_meshes * 4, cudaMemcpyHostToDevice, st));
  }
  {
    const int rc = prepare_meshes(d, n_meshes, n_spheres);
    if (rc) return rc;
  }
  RR_CUDA(cudaEventRecord(d.ev1, st));
  RR_CUDA(cudaStreamSynchronize(st));
  {  // what frame_needs_slack() looks at
    auto absmax6 = [](const float* b) {
      float m = 0.0f;
      for (int k = 0; k < 6; ++k) { const float a = std::fabs(b[k]); if (a > m && a < 3.0e38f) m = a; }
      return m;
    };
    d.h_meshes.assign(meshes, meshes + n_meshes);
    d.h_mesh_absmax.assign(n_meshes, 0.0f);
    std::vector<float> sb(6 * std::max<size_t>(plan.first.size(), 1));
    if (!plan.first.empty()) RR_CUDA(cudaMemcpy(sb.data(), d.tb.seg_box, plan.first.size() * 24, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n_meshes; ++i) d.h_mesh_absmax[i] = plan.count[plan.mesh_seg[i]] ? absmax6(&sb[6 * plan.mesh_seg[i]]) : 0.0f;
    d.h_sph_absmax = 0.0f;
    d.h_sphere_materials = 0;
    for (size_t i = 0; i < n_spheres; ++i) {
      const int t = spheres[i].material.type;
      if (t == RR_MATERIAL_CHECKER || t == RR_MATERIAL_GLASSY || t == RR_MATERIAL_INVISIBLE) d.h_sphere_materials = RR_FEAT_MATERIALS;
    }
    if (n_spheres) {
      float b6[6];
      RR_CUDA(cudaMemcpy(b6, d.sb.seg_box, 24, cudaMemcpyDeviceToHost));
      d.h_sph_absmax = absmax6(b6);
    }
  }
  RR_CUDA(cudaEventElapsedTime(&d.build_ms, d.ev0, d.ev1));
  upload_mark(st, "node array, meshes, top level");
  dev_free(d.tb.nodes); d.tb.nodes = nullptr;  // copied into d.nodes
  dev_free(d.sb.nodes); d.sb.nodes = nullptr;
  static_assert(3 * RR_MAX_DEPTH + 4 <= RR_STACK_MAX, "stack pointer must fit the slot word");
  {  // traversal stacks sized for THIS scene: a wide node pushes at most 3 children per level
    const uint32_t need = 3u * std::max(d.tb.wide_levels, d.sb.wide_levels) + 4u;
    if (need > RR_STACK_MAX)
      return fail(RR_ERR_BVH_DEPTH, "hierarchy of " + std::to_string(std::max(d.tb.wide_levels, d.sb.wide_levels)) +
                                        " wide levels exceeds the traversal stack (" + std::to_string(RR_STACK_MAX) + " entries)");
    if (need > d.stack_entries || !d.stack) {
      cudaFree(d.cold);  // one block: cold slot words of every warp, then the stacks
      d.cold = nullptr; d.stack = nullptr;
      d.stack_entries = 0;
      const size_t cold_bytes = (size_t)d.stack_warps * render_cold_bytes_per_warp();
      RR_CUDA(cudaMalloc(&d.cold, cold_bytes + (size_t)d.stack_warps * render_stack_bytes_per_warp(need)));
      d.stack = reinterpret_cast<uint2*>(reinterpret_cast<char*>(d.cold) + cold_bytes);
      d.stack_entries = need;
    }
  }
  if (d.tb.max_depth > RR_MAX_DEPTH || d.sb.max_depth > RR_MAX_DEPTH)
    return fail(RR_ERR_BVH_DEPTH, "LBVH depth " + std::to_string(std::max(d.tb.max_depth, d.sb.max_depth)) +
                                      " exceeds the traversal stack (" + std::to_string(RR_MAX_DEPTH) + " levels)");
  return RR_OK;
}

static int ensure_frame(Device& d, uint32_t W, uint32_t H, bool want_radiance) {
  RR_CUDA(cudaSetDevice(d.ordinal));
  const size_t need = (size_t)W * H * 4;
  if (need > d.frame_bytes) {
    cudaFree(d.frame);
    d.frame = nullptr; d.frame_bytes = 0;
    RR_CUDA(cudaMalloc(&d.frame, need));
    d.frame_bytes = need;
  }
  const size_t rneed = (size_t)W * H * 12;
  if (want_radiance && rneed > d.radiance_bytes) {
    cudaFree(d.radiance);
    d.radiance = nullptr; d.radiance_bytes = 0;
    RR_CUDA(cudaMalloc(&d.radiance, rneed));
    d.radiance_bytes = rneed;
  }
  return RR_OK;
}

// Does this frame need the kernel instantiation with the per-ray culling slack (rr_render.cu RaySlack)?  Every box is
// inflated at build time by box_delta = 2^-18 (64 ulp) of the largest |coordinate| A of its mesh, which covers the
// rounding error of slab and triangle tests while the mesh-local ray origin stays within 16 A.  A ray starts at the
// camera or at a hit point, so |origin| <= O = max(|cam|, radius of the scene), and in mesh m's space
// |local origin| <= (O + |pos_m|) / |scale_m|.  The world-space tests (mesh world boxes, 2^-13 relative slack of their
// own) tolerate a larger ratio.  If any mesh exceeds its bound the frame runs with the per-ray term (a few % slower).
static bool frame_needs_slack(const rr_ctx* ctx, const Device& d, const rr_camera* cam) {
  if (ctx->tune.speculate & 16u) return true;   // tuning bits 4 / 5: force either instantiation (tests, A/B)
  if (ctx->tune.speculate & 32u) return false;
  auto len3 = [](const float* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
  float radius = 1.7320508f * d.h_sph_absmax;  // of everything a path can hit
  for (size_t i = 0; i < d.h_meshes.size(); ++i) {
    const rr_mesh& m = d.h_meshes[i];
    if (!(m.scale > RR_EPSILON) || d.h_mesh_absmax[i] == 0.0f) continue;
    radius = std::max(radius, len3(m.pos.s) + m.scale * 1.7320508f * d.h_mesh_absmax[i]);
  }
  const float O = std::max(len3(cam->position.s), radius);
  if (!(O < 3.0e38f)) return true;
  if (d.h_sph_absmax > 0.0f && O > 16.0f * d.h_sph_absmax) return true;
  for (size_t i = 0; i < d.h_meshes.size(); ++i) {
    const rr_mesh& m = d.h_meshes[i];
    const float A = d.h_mesh_absmax[i];
    if (!(m.scale > RR_EPSILON) || A == 0.0f) continue;
    const float P = len3(m.pos.s);
    if ((O + P) / m.scale > 16.0f * A) return true;               // mesh-local tests: node boxes, root box
    if (O > 256.0f * std::max(P, 0.5f * m.scale * A)) return true;  // world-space tests: the mesh's world box
  }
  return false;
}

// RR_FEAT_* bits of the scene as it is now (materials can change through rr_update_meshes).  Tuning bit 6 (64) forces
// the instantiation with everything (A/B).
static int scene_features(const rr_ctx* ctx, const Device& d) {
  if (ctx->tune.speculate & 64u) return RR_FEAT_ALL;
  int f = d.h_sphere_materials;
  if (ctx->n_spheres) f |= RR_FEAT_SPHERES;
  if (ctx->n_meshes + (ctx->n_spheres ? 1 : 0) > 32) f |= RR_FEAT_TLAS;
  for (const rr_mesh& m : d.h_meshes) {
    const int t = m.material.type;
    if (t == RR_MATERIAL_CHECKER || t == RR_MATERIAL_GLASSY || t == RR_MATERIAL_INVISIBLE) { f |= RR_FEAT_MATERIALS; break; }
  }
  return f;
}

// Tiles of a launch and the multiplier of the scattered tile order (RR_TILE_ORDER 2): about 0.618 n, made coprime to n
static void set_queue_tiles(RenderParams& p, uint32_t n) {
  p.queue_tiles = std::max<uint32_t>(n, 1u);
  uint32_t m = (uint32_t)((double)p.queue_tiles * 0.6180339887498949) | 1u;
  auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
  while (gcd(m, p.queue_tiles) != 1u) m += 2u;
  p.tile_mul = m % p.queue_tiles;
  if (p.queue_tiles == 1u) p.tile_mul = 0u;
}

static void fill_params(const rr_ctx* ctx, const Device& d, const rr_camera* cam, uint32_t W, uint32_t H, uint32_t spp,
                        uint32_t bounces, int32_t frame_index, uint32_t tile_size, RenderParams& p) {
  memset(&p, 0, sizeof(p));
  p.meshes = d.meshes;
  p.n_meshes = (int32_t)ctx->n_meshes;
  p.materials = d.materials;
  p.nodes = d.nodes;
  p.tri_geom = d.tri_geom;
  p.tri_nrm = d.tri_nrm;
  p.n_spheres = (int32_t)ctx->n_spheres;
  p.last_mesh = (int32_t)ctx->n_meshes - 1 + (ctx->n_spheres ? 1 : 0);
  if (p.last_mesh >= 32 && !(ctx->tune.speculate & 4u)) {  // tuning bit 2: linear mesh scan (A/B)
    p.tlas_blocks = d.tlas_blocks;
    p.tlas = d.tlas;
    p.tlas_levels = d.tlas_levels;
  }
  p.sph_geom = d.sph_geom;
  p.sph_order = d.sb.order;
  p.top_nodes = d.top_nodes;
  p.top_count = d.top_count;
  p.top_root = d.top_root;
  p.tune = ctx->tune;
  for (int k = 0; k < 3; ++k) p.cam.pos[k] = cam->position.s[k];
  p.cam.pitch = cam->pitch; p.cam.yaw = cam->yaw; p.cam.roll = cam->roll; p.cam.fov = cam->fov; p.cam.aspect = cam->aspectRatio;
  p.width = W; p.height = H; p.spp = spp; p.max_bounces = bounces; p.frame_index = frame_index;
  // Work tiles: 8 x 4 pixels by default.  A caller's tile size (the reference's TILE_SIZE, src/settings.hpp:48, bounds
  // the length of one OpenCL launch) is honoured up to 32 x 32: a tile is the unit ONE warp pops from the queue, and
  // the image does not depend on the tiling (src/image.hpp:228: the per-tile seed term is 0).
  p.tile_w = tile_size ? std::min<uint32_t>(tile_size, 32u) : RR_TILE_W;
  p.tile_h = tile_size ? std::min<uint32_t>(tile_size, 32u) : RR_TILE_H;
  p.tiles_x = (W + p.tile_w - 1) / p.tile_w;
  p.tiles_y = (H + p.tile_h - 1) / p.tile_h;
  p.tile_begin = 0;
  p.tile_stride = 1;
  p.tile_pixels = p.tile_w * p.tile_h;
  p.queue_items = p.tiles_x * p.tiles_y * p.tile_pixels;  // checked against 2^32 by render_frame
  set_queue_tiles(p, p.tiles_x * p.tiles_y);
  if (d.tile_order) {  // a caller-supplied hand-out order: the launch renders exactly the tiles of the table
    p.tile_order = d.tile_order;
    p.queue_items = d.tile_order_n * p.tile_pixels;
    set_queue_tiles(p, d.tile_order_n);
  }
  p.stack = d.stack;
  p.stack_entries = d.stack_entries;
  p.cold = d.cold;
  p.stack_warps = d.stack_warps;
  p.pool_use = ctx->pool_use ? std::min<uint32_t>(ctx->pool_use, RR_POOL) : RR_POOL;
  p.queue = d.queue;
  p.frame = d.frame;
  p.radiance = nullptr;
  p.counters = d.counters;
}

// the queue counter hands out 32-bit work items (pixels in tile-major order, ragged border tiles padded)
static bool too_many_items(const RenderParams& p) { return (uint64_t)p.tiles_x * p.tiles_y * p.tile_pixels > 0xffffffffull; }

static int check_render_args(const rr_ctx* ctx, const rr_camera* cam, uint32_t W, uint32_t H, uint32_t spp, uint32_t bounces = 0) {
  if (!ctx || !cam) return fail(RR_ERR_INVALID_ARGUMENT, "null context or camera");
  if (bounces > RR_MAX_BOUNCES) return fail(RR_ERR_INVALID_ARGUMENT, "max_bounces above 8388607 (the bounce counter has 23 bits)");
  if (!ctx->has_scene) return fail(RR_ERR_NO_SCENE, "rr_upload_scene has not been called");
  if (W == 0 || H == 0 || spp == 0) return fail(RR_ERR_INVALID_ARGUMENT, "width, height and spp must be positive");
  if ((uint64_t)W * H > 0x7fffffffull) return fail(RR_ERR_INVALID_ARGUMENT, "image too large (pixel index must fit 31 bits)");
  return RR_OK;
}

static void read_stats(const Counters& c, uint64_t samples, float ms, float build_ms, rr_stats* out) {
  if (!out) return;
  out->samples = samples;
  out->rays = c.rays;
  out->stack_overflows = c.stack_overflows;
  out->box_tests = c.box_tests;
  out->tri_tests = c.tri_tests;
  out->sphere_tests = c.sphere_tests;
  out->tiles = c.tiles;
  out->render_ms = ms;
  out->build_ms = build_ms;
  for (int k = 0; k < 5; ++k) { out->phase_runs[k] = c.phase_runs[k]; out->phase_lanes[k] = c.phase_lanes[k]; }
  out->tail_avg_ms = c.tail_warps ? (float)((double)c.tail_ns_sum / (double)c.tail_warps * 1e-6) : 0.0f;
  out->tail_max_ms = (float)((double)c.tail_ns_max * 1e-6);
}

// Renders one frame on all devices of the context.  Device 0 owns the queue and
// the frame; peers (in-process multi-GPU) address them directly over NVLink.
static int render_frame(rr_ctx* ctx, const rr_camera* cam, uint32_t W, uint32_t H, uint32_t spp, uint32_t bounces,
                        int32_t frame_index, uint32_t tile_size, bool want_radiance, bool count_tests, int mode,
                        uint32_t rank, uint32_t world, rr_stats* stats_out) {
  // mode 0: local queue, 1: shared (imported/exported) queue + frame, 2: static stride partition
  int rc = check_render_args(ctx, cam, W, H, spp, bounces);
  if (rc) return rc;
  if (ctx->tune.speculate & 8u) count_tests = true;  // tuning bit 3: the instrumented kernel for every kind of frame (tools/tail_model.py)
  const size_t nd = ctx->dev.size();
  if (want_radiance && nd > 1) return fail(RR_ERR_UNSUPPORTED, "radiance output needs a single-device context");
  Device& d0 = ctx->dev[0];
  for (size_t k = 0; k < nd; ++k) {
    rc = ensure_frame(ctx->dev[k], W, H, want_radiance);
    if (rc) return rc;
  }
  RR_CUDA(cudaSetDevice(d0.ordinal));
  if (mode != 1) {
    RR_CUDA(cudaMemsetAsync(d0.queue, 0, sizeof(unsigned long long), d0.stream));
    if (mode == 2) RR_CUDA(cudaMemsetAsync(d0.frame, 0, (size_t)W * H * 4, d0.stream));
  } else {
    if (!d0.shared_queue) return fail(RR_ERR_INVALID_ARGUMENT, "rr_render_shared needs rr_queue_export or rr_queue_import first");
    if ((size_t)W * H * 4 > d0.shared_bytes)
      return fail(RR_ERR_QUEUE, "frame of " + std::to_string(W) + "x" + std::to_string(H) + " exceeds the shared frame (" +
                                    std::to_string(d0.shared_w) + "x" + std::to_string(d0.shared_h) + ")");
    d0.shared_epoch++;  // the epoch rr_queue_reset has put into the counter for this frame
  }
  RR_CUDA(cudaStreamSynchronize(d0.stream));
  {
    RenderParams q;
    fill_params(ctx, d0, cam, W, H, spp, bounces, frame_index, tile_size, q);
    if (too_many_items(q)) return fail(RR_ERR_INVALID_ARGUMENT, "image too large (its tiles hold more than 2^32 - 1 work items)");
    ctx->progress_queue.store(mode == 1 ? d0.shared_queue : d0.queue);
    ctx->progress_items_per_tile.store(RR_PIXEL_QUEUE ? q.tile_pixels : 1u);
    ctx->progress_total.store((uint64_t)q.tiles_x * q.tiles_y);
  }
  for (size_t k = 0; k < nd; ++k) {
    Device& d = ctx->dev[k];
    RR_CUDA(cudaSetDevice(d.ordinal));
    RR_CUDA(cudaMemsetAsync(d.counters, 0, sizeof(Counters), d.stream));
    RenderParams p;
    fill_params(ctx, d, cam, W, H, spp, bounces, frame_index, tile_size, p);
    if (mode == 1) {
      p.queue = d.shared_queue;
      p.frame = d.shared_frame;
      p.queue_epoch = d0.shared_epoch & 0xffffu;
    } else {
      p.queue = d0.queue;  // peers pop device 0's counter
      p.frame = d0.frame;
      if (mode == 2 && !p.tile_order) {  // this rank's tiles: rank, rank + world, ... (a tile-order table already names the tiles)
        p.tile_begin = rank; p.tile_stride = world;
        const uint64_t tiles = (uint64_t)p.tiles_x * p.tiles_y;
        p.queue_items = (uint32_t)((tiles > rank ? (tiles - rank + world - 1) / world : 0) * p.tile_pixels);
        set_queue_tiles(p, p.queue_items / p.tile_pixels);
      }
    }
    if (want_radiance) p.radiance = d.radiance;
#ifdef RR_L2_PIN_EXPERIMENT
    {  // A/B only: one array of the scene in a persisting-L2 access window (RR_L2_PIN = nodes | geom | nrm | cold)
      const char* what = getenv("RR_L2_PIN");
      if (what) {
        const char* mb = getenv("RR_L2_PIN_MB");
        const char* ratio = getenv("RR_L2_PIN_RATIO");
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, d.ordinal);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, d.ordinal);
        size_t carve = mb ? (size_t)atoi(mb) << 20 : (size_t)max_persist;
        carve = std::min(carve, (size_t)max_persist);
        RR_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
        void* base = nullptr; size_t bytes = 0;
        if (!strcmp(what, "nodes")) { base = d.nodes; bytes = (d.tb.n_nodes + d.sb.n_nodes) * 128; }
        else if (!strcmp(what, "geom")) { base = d.tri_geom; bytes = d.tb.n * 48; }
        else if (!strcmp(what, "nrm")) { base = d.tri_nrm; bytes = d.tb.n * 48; }
        else if (!strcmp(what, "cold")) { base = d.cold; bytes = (size_t)d.stack_warps * render_cold_bytes_per_warp(); }
        bytes = std::min(bytes, (size_t)max_window);
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.base_ptr = base;
        av.accessPolicyWindow.num_bytes = bytes;
        av.accessPolicyWindow.hitRatio = ratio ? (float)atof(ratio) : 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        RR_CUDA(cudaStreamSetAttribute(d.stream, cudaStreamAttributeAccessPolicyWindow, &av));
        static bool said = false;
        if (!said) { fprintf(stderr, "L2 pin: %s, %zu MB window, carve-out %zu MB (max %d MB, max window %d MB), ratio %.2f\n", what, bytes >> 20, carve >> 20, max_persist >> 20, max_window >> 20, av.accessPolicyWindow.hitRatio); said = true; }
      }
    }
#endif
    RR_CUDA(cudaEventRecord(d.ev0, d.stream));
    RR_CUDA(launch_render(p, count_tests, frame_needs_slack(ctx, d, cam), scene_features(ctx, d), d.sm_count, d.stream));
    RR_CUDA(cudaEventRecord(d.ev1, d.stream));
  }
  Counters total;
  memset(&total, 0, sizeof(total));
  float ms_max = 0.0f;
  for (size_t k = 0; k < nd; ++k) {
    Device& d = ctx->dev[k];
    RR_CUDA(cudaSetDevice(d.ordinal));
    RR_CUDA(cudaStreamSynchronize(d.stream));
    float ms = 0.0f;
    RR_CUDA(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
    ms_max = std::max(ms_max, ms);
    Counters c;
    RR_CUDA(cudaMemcpy(&c, d.counters, sizeof(c), cudaMemcpyDeviceToHost));
    total.rays += c.rays; total.stack_overflows += c.stack_overflows; total.queue_errors += c.queue_errors; total.box_tests += c.box_tests;
    total.tri_tests += c.tri_tests; total.sphere_tests += c.sphere_tests; total.tiles += c.tiles;
    for (int q = 0; q < 5; ++q) { total.phase_runs[q] += c.phase_runs[q]; total.phase_lanes[q] += c.phase_lanes[q]; }
    total.tail_ns_sum += c.tail_ns_sum; total.tail_warps += c.tail_warps; total.tail_ns_max = std::max(total.tail_ns_max, c.tail_ns_max);
  }
  ctx->progress_total.store(0);
  read_stats(total, (uint64_t)W * H * spp, ms_max, d0.build_ms, stats_out);
  if (total.stack_overflows)
    return fail(RR_ERR_BVH_DEPTH, std::to_string(total.stack_overflows) + " node steps ran out of traversal stack (" +
                                      std::to_string(d0.stack_entries) + " entries per path): the frame is incomplete");
  if (total.queue_errors)
    return fail(RR_ERR_QUEUE, "the shared tile counter changed epoch during the frame: rr_queue_reset must run between two barriers, "
                              "after every rank has returned from rr_render_shared (include/rr_api.h)");
  return RR_OK;
}

}  // namespace rr

using namespace rr;

extern "C" {

const char* rr_error_string(int status) {
  switch (status) {
    case RR_OK: return "success";
    case RR_ERR_INVALID_ARGUMENT: return "invalid argument";
    case RR_ERR_NO_DEVICE: return "no usable CUDA device";
    case RR_ERR_CUDA: return "CUDA runtime error";
    case RR_ERR_OUT_OF_MEMORY: return "out of device memory";
    case RR_ERR_NO_SCENE: return "no scene uploaded";
    case RR_ERR_BAD_MESH_RANGE: return "bad mesh triangle range";
    case RR_ERR_BVH_DEPTH: return "BVH deeper than the traversal stack";
    case RR_ERR_IO: return "I/O error";
    case RR_ERR_UNSUPPORTED: return "unsupported";
    case RR_ERR_QUEUE: return "shared tile queue misuse";
    default: return "unknown status";
  }
}
const char* rr_last_error(void) { return g_last_error.c_str(); }
int rr_version(void) { return 400; }

int rr_device_count(int* out) {
  if (!out) return fail(RR_ERR_INVALID_ARGUMENT, "null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *out = 0; cudaGetLastError(); return fail(RR_ERR_NO_DEVICE, cudaGetErrorString(e)); }
  *out = n;
  return RR_OK;
}

int rr_device_info(int ordinal, char* name, size_t name_len, int* sm_count, uint64_t* mem_bytes) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, ordinal);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(RR_ERR_NO_DEVICE, cudaGetErrorString(e)); }
  if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (mem_bytes) *mem_bytes = (uint64_t)prop.totalGlobalMem;
  return RR_OK;
}

int rr_create(const int* cuda_ordinals, int n, rr_ctx** out) {
  if (!out) return fail(RR_ERR_INVALID_ARGUMENT, "null output");
  *out = nullptr;
  int count = 0;
  int rc = rr_device_count(&count);
  if (rc) return rc;
  if (count == 0) return fail(RR_ERR_NO_DEVICE, "cudaGetDeviceCount() == 0");
  int def = 0;
  if (!cuda_ordinals || n <= 0) { cuda_ordinals = &def; n = 1; }
  rr_ctx* ctx = new rr_ctx();
  default_tuning(ctx->tune);
  ctx->dev.resize(n);
  for (int k = 0; k < n; ++k) {
    Device& d = ctx->dev[k];
    d.ordinal = cuda_ordinals[k];
    if (d.ordinal < 0 || d.ordinal >= count) { delete ctx; return fail(RR_ERR_NO_DEVICE, "device ordinal out of range"); }
    cudaError_t e = cudaSetDevice(d.ordinal);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, d.ordinal);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&d.ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&d.ev1);
    if (e == cudaSuccess) e = cudaMalloc(&d.queue, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&d.counters, sizeof(Counters));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.poll_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost(&d.poll_host, sizeof(unsigned long long));
    if (e == cudaSuccess) {
      d.stack_warps = (uint32_t)(prop.multiProcessorCount * render_max_warps_per_sm());
      // the scratch block itself is sized at upload time (its stack part depends on the scene)
    }
    if (e != cudaSuccess) { rr_destroy(ctx); return cuda_fail(e, "rr_create"); }
    d.sm_count = prop.multiProcessorCount;
  }
  ctx->peer_ok = true;
  for (int k = 1; k < n; ++k) {  // peers must reach device 0 (queue + frame live there)
    int can = 0;
    cudaDeviceCanAccessPeer(&can, ctx->dev[k].ordinal, ctx->dev[0].ordinal);
    if (!can) { ctx->peer_ok = false; break; }
    cudaSetDevice(ctx->dev[k].ordinal);
    cudaError_t e = cudaDeviceEnablePeerAccess(ctx->dev[0].ordinal, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ctx->peer_ok = false;
    cudaGetLastError();
  }
  if (n > 1 && !ctx->peer_ok) { rr_destroy(ctx); return fail(RR_ERR_UNSUPPORTED, "multi-device context needs peer access to device 0"); }
  *out = ctx;
  return RR_OK;
}

void rr_destroy(rr_ctx* ctx) {
  if (!ctx) return;
  for (Device& d : ctx->dev) {
    if (d.ordinal < 0) continue;
    cudaSetDevice(d.ordinal);
    if (d.stream) cudaStreamSynchronize(d.stream);
    free_scene(d);
    if (d.shared_imported) {
      if (d.shared_queue) cudaIpcCloseMemHandle(d.shared_queue);
      if (d.shared_frame) cudaIpcCloseMemHandle(d.shared_frame);
    } else if (d.shared_frame) {
      cudaFree(d.shared_frame);
    }
    if (d.poll_stream) cudaStreamDestroy(d.poll_stream);
    if (d.poll_host) cudaFreeHost(d.poll_host);
    cudaFree(d.frame); cudaFree(d.radiance); cudaFree(d.accum); cudaFree(d.queue); cudaFree(d.tile_order); cudaFree(d.counters); cudaFree(d.cold);
    dev_trim(d.ordinal);
    if (d.ev0) cudaEventDestroy(d.ev0);
    if (d.ev1) cudaEventDestroy(d.ev1);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  cudaGetLastError();
  delete ctx;
}

int rr_upload_scene(rr_ctx* ctx, const rr_triangle* tris, size_t n_tris, const rr_mesh* meshes,
                    const rr_mesh_range* ranges, size_t n_meshes, const rr_sphere* spheres, size_t n_spheres) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  if ((n_tris && !tris) || (n_meshes && (!meshes || !ranges)) || (n_spheres && !spheres))
    return fail(RR_ERR_INVALID_ARGUMENT, "null array with non-zero count");
  int rc = check_upload_counts(n_tris, n_meshes, n_spheres);
  if (rc) return rc;
  SegPlan plan;
  rc = plan_segments(ranges, n_meshes, n_tris, plan);
  if (rc) return rc;
  ctx->has_scene = false;
  for (Device& d : ctx->dev) {
    rc = upload_device(d, tris, nullptr, n_tris, meshes, n_meshes, plan, spheres, n_spheres);
    if (rc) return rc;
  }
  ctx->n_tris = n_tris; ctx->n_meshes = n_meshes; ctx->n_spheres = n_spheres;
  ctx->has_scene = true;
  return RR_OK;
}

int rr_upload_scene_indexed(rr_ctx* ctx, const float* positions, size_t n_positions, const float* normals,
                            size_t n_normals, const uint32_t* corners, size_t n_tris, const rr_mesh* meshes,
                            const rr_mesh_range* ranges, size_t n_meshes, const rr_sphere* spheres, size_t n_spheres) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  if ((n_tris && (!positions || !normals || !corners)) || (n_meshes && (!meshes || !ranges)) || (n_spheres && !spheres))
    return fail(RR_ERR_INVALID_ARGUMENT, "null array with non-zero count");
  if (n_positions >= 0xffffffffull || n_normals >= 0xffffffffull) return fail(RR_ERR_INVALID_ARGUMENT, "too many vertices (32-bit indices)");
  int rc = check_upload_counts(n_tris, n_meshes, n_spheres);
  if (rc) return rc;
  for (size_t i = 0; i < n_tris; ++i)  // the gather kernel trusts the indices
    for (int k = 0; k < 6; ++k)
      if (corners[6 * i + k] >= (k < 3 ? n_positions : n_normals))
        return fail(RR_ERR_INVALID_ARGUMENT, "triangle " + std::to_string(i) + ": corner index out of range");
  SegPlan plan;
  rc = plan_segments(ranges, n_meshes, n_tris, plan);
  if (rc) return rc;
  IndexedInput in;
  in.positions = positions; in.n_positions = n_positions; in.normals = normals; in.n_normals = n_normals; in.corners = corners;
  ctx->has_scene = false;
  for (Device& d : ctx->dev) {
    rc = upload_device(d, nullptr, &in, n_tris, meshes, n_meshes, plan, spheres, n_spheres);
    if (rc) return rc;
  }
  ctx->n_tris = n_tris; ctx->n_meshes = n_meshes; ctx->n_spheres = n_spheres;
  ctx->has_scene = true;
  return RR_OK;
}

int rr_upload_scene_ref(rr_ctx* ctx, const rr_triangle* tris, size_t n_tris, const rr_mesh* meshes, size_t n_meshes,
                        const rr_ref_node* nodes, size_t n_nodes) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  if ((n_meshes && (!nodes || !meshes)) || (n_tris && !tris)) return fail(RR_ERR_INVALID_ARGUMENT, "null array with non-zero count");
  std::vector<rr_mesh_range> ranges(n_meshes);
  std::vector<uint64_t> stack;
  for (size_t i = 0; i < n_meshes; ++i) {
    uint64_t lo = UINT64_MAX, hi = 0;
    stack.assign(1, meshes[i].nodeIdx);
    size_t visited = 0;
    while (!stack.empty()) {
      uint64_t ni = stack.back();
      stack.pop_back();
      if (ni >= n_nodes || ++visited > n_nodes) return fail(RR_ERR_BAD_MESH_RANGE, "mesh " + std::to_string(i) + ": node index outside nodeList");
      const rr_ref_node& nd = nodes[ni];
      if (nd.numTriangles > 0) {  // leaf (an unsplit OBJ root keeps childIndex != 0 but is still a leaf)
        lo = std::min(lo, nd.firstTriangleIdx);
        hi = std::max(hi, nd.firstTriangleIdx + nd.numTriangles);
      } else if (nd.childIndex != 0) {
        stack.push_back(nd.childIndex);
        stack.push_back(nd.childIndex + 1);
      }
    }
    ranges[i].firstTriangle = lo == UINT64_MAX ? 0 : lo;
    ranges[i].numTriangles = lo == UINT64_MAX ? 0 : hi - lo;
  }
  return rr_upload_scene(ctx, tris, n_tris, meshes, ranges.data(), n_meshes, nullptr, 0);
}

int rr_render_ex(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                 uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, uint8_t* rgba_out,
                 float* radiance_out, rr_stats* stats_out, int count_tests) {
  int rc = render_frame(ctx, cam, width, height, spp, max_bounces, frame_index, tile_size, radiance_out != nullptr,
                        count_tests != 0, 0, 0, 1, stats_out);
  if (rc) return rc;
  Device& d0 = ctx->dev[0];
  RR_CUDA(cudaSetDevice(d0.ordinal));
  if (rgba_out) RR_CUDA(cudaMemcpy(rgba_out, d0.frame, (size_t)width * height * 4, cudaMemcpyDeviceToHost));
  if (radiance_out) RR_CUDA(cudaMemcpy(radiance_out, d0.radiance, (size_t)width * height * 12, cudaMemcpyDeviceToHost));
  return RR_OK;
}

int rr_render(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_bounces,
              int32_t frame_index, uint32_t tile_size, uint8_t* rgba_out) {
  if (!rgba_out) return fail(RR_ERR_INVALID_ARGUMENT, "null output image");
  return rr_render_ex(ctx, cam, width, height, spp, max_bounces, frame_index, tile_size, rgba_out, nullptr, nullptr, 0);
}

int rr_set_tuning(rr_ctx* ctx, const uint32_t* values, size_t n) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  if (!values) { default_tuning(ctx->tune); ctx->pool_use = 0; return RR_OK; }
  if (n > 8) ctx->pool_use = values[8];  // slots per warp in use (0 = all); not a scheduler knob: every instantiation honours it
  uint32_t* dst[8] = {&ctx->tune.weight[0], &ctx->tune.weight[1], &ctx->tune.weight[2], &ctx->tune.weight[3],
                      &ctx->tune.weight[4], &ctx->tune.trav_keep, &ctx->tune.speculate, &ctx->tune.ctas_per_sm};
  for (size_t k = 0; k < n && k < 8; ++k) *dst[k] = values[k];
  if (ctx->tune.trav_keep == 0) ctx->tune.trav_keep = 1;
  for (int k = 0; k < 5; ++k) if (ctx->tune.weight[k] == 0) ctx->tune.weight[k] = 1;
  return RR_OK;
}

int rr_render_device(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                     uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, rr_stats* stats_out) {
  return render_frame(ctx, cam, width, height, spp, max_bounces, frame_index, tile_size, false, false, 0, 0, 1, stats_out);
}

int rr_render_progress(rr_ctx* ctx, uint64_t* tiles_popped, uint64_t* tiles_total) {
  if (!ctx || !tiles_popped || !tiles_total) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  Device& d = ctx->dev[0];
  const uint64_t total = ctx->progress_total.load();
  *tiles_total = total;
  *tiles_popped = 0;
  const unsigned long long* src = ctx->progress_queue.load();
  if (!total || !src) return RR_OK;  // nothing is being rendered
  RR_CUDA(cudaSetDevice(d.ordinal));
  RR_CUDA(cudaMemcpyAsync(d.poll_host, src, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.poll_stream));
  RR_CUDA(cudaStreamSynchronize(d.poll_stream));
  // (every warp's last, failing pop counts too, hence the clamp)
  const uint64_t popped = (*d.poll_host & ((1ull << RR_QUEUE_EPOCH_SHIFT) - 1ull)) / std::max<uint32_t>(ctx->progress_items_per_tile.load(), 1u);
  *tiles_popped = popped < total ? popped : total;
  return RR_OK;
}

int rr_read_frame(rr_ctx* ctx, uint8_t* rgba_out, size_t bytes) {
  if (!ctx || !rgba_out) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  Device& d0 = ctx->dev[0];
  const bool from_shared = d0.shared_frame && !d0.shared_imported;  // the exporting rank reads the gathered frame
  const uint8_t* src = from_shared ? d0.shared_frame : d0.frame;
  if (!src || bytes > (from_shared ? d0.shared_bytes : d0.frame_bytes)) return fail(RR_ERR_INVALID_ARGUMENT, "no frame of that size has been rendered");
  RR_CUDA(cudaSetDevice(d0.ordinal));
  RR_CUDA(cudaMemcpy(rgba_out, src, bytes, cudaMemcpyDeviceToHost));
  return RR_OK;
}

int rr_update_meshes(rr_ctx* ctx, const rr_mesh* meshes, size_t n_meshes) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  if (!ctx->has_scene) return fail(RR_ERR_NO_SCENE, "rr_upload_scene has not been called");
  if (n_meshes != ctx->n_meshes) return fail(RR_ERR_INVALID_ARGUMENT, "mesh count differs from the uploaded scene (" + std::to_string(ctx->n_meshes) + ")");
  if (n_meshes && !meshes) return fail(RR_ERR_INVALID_ARGUMENT, "null mesh array");
  for (Device& d : ctx->dev) {
    RR_CUDA(cudaSetDevice(d.ordinal));
    if (n_meshes) RR_CUDA(cudaMemcpyAsync(d.meshes_in, meshes, n_meshes * sizeof(rr_mesh), cudaMemcpyHostToDevice, d.stream));
    d.h_meshes.assign(meshes, meshes + n_meshes);
    const int rc = prepare_meshes(d, n_meshes, ctx->n_spheres);  // synchronises: the caller's array may go away
    if (rc) return rc;
  }
  return RR_OK;
}

int rr_accum_reset(rr_ctx* ctx, uint32_t width, uint32_t height) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  if (width == 0 || height == 0 || (uint64_t)width * height > 0x7fffffffull) return fail(RR_ERR_INVALID_ARGUMENT, "bad image size");
  Device& d = ctx->dev[0];
  RR_CUDA(cudaSetDevice(d.ordinal));
  const size_t n = (size_t)width * height, padded = (n + 3) & ~size_t(3);  // planes stay 16-byte aligned
  if (padded > d.accum_capacity) {
    cudaFree(d.accum);
    d.accum = nullptr; d.accum_capacity = 0;
    RR_CUDA(cudaMalloc(&d.accum, 3 * padded * sizeof(uint32_t)));
    d.accum_capacity = padded;
  }
  RR_CUDA(cudaMemsetAsync(d.accum, 0, 3 * d.accum_capacity * sizeof(uint32_t), d.stream));
  RR_CUDA(cudaStreamSynchronize(d.stream));
  d.accum_w = width; d.accum_h = height; d.accum_frames = 0;
  return RR_OK;
}

int rr_accum_last_ms(rr_ctx* ctx, float* ms_out) {
  if (!ctx || !ms_out) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  *ms_out = ctx->dev[0].accum_ms;
  return RR_OK;
}

int rr_accum_frame_count(rr_ctx* ctx, uint32_t* frames_out) {
  if (!ctx || !frames_out) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  *frames_out = ctx->dev[0].accum_frames;
  return RR_OK;
}

int rr_accum_add_frame(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                       uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, uint8_t* rgba_avg_out,
                       rr_stats* stats_out) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  Device& d = ctx->dev[0];
  if (!d.accum || d.accum_w != width || d.accum_h != height)
    return fail(RR_ERR_INVALID_ARGUMENT, "rr_accum_reset with this image size has not been called");
  if (d.accum_frames == 0x00ffffffu) return fail(RR_ERR_UNSUPPORTED, "too many frames for the 32-bit sums");
  int rc = render_frame(ctx, cam, width, height, spp, max_bounces, frame_index, tile_size, false, false, 0, 0, 1, stats_out);
  if (rc) return rc;
  RR_CUDA(cudaSetDevice(d.ordinal));
  const size_t n = (size_t)width * height;
  d.accum_frames++;
  uint32_t* planes = d.accum;
  const int grid = (int)std::min<size_t>((n / 4 + 255) / 256 + 1, (size_t)d.sm_count * 8);
  RR_CUDA(cudaEventRecord(d.ev0, d.stream));
  k_accum_add<<<grid, 256, 0, d.stream>>>(reinterpret_cast<uint32_t*>(d.frame), planes, planes + d.accum_capacity,
                                          planes + 2 * d.accum_capacity, n, d.accum_frames);
  RR_CUDA(cudaGetLastError());
  RR_CUDA(cudaEventRecord(d.ev1, d.stream));
  if (rgba_avg_out) RR_CUDA(cudaMemcpyAsync(rgba_avg_out, d.frame, n * 4, cudaMemcpyDeviceToHost, d.stream));
  RR_CUDA(cudaStreamSynchronize(d.stream));
  RR_CUDA(cudaEventElapsedTime(&d.accum_ms, d.ev0, d.ev1));
  return RR_OK;
}

int rr_render_progressive(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                          uint32_t max_bounces, int32_t first_frame_index, uint32_t n_frames, uint32_t tile_size,
                          uint8_t* rgba_out, rr_stats* stats_out) {
  if (!rgba_out) return fail(RR_ERR_INVALID_ARGUMENT, "null output image");
  if (n_frames == 0) return fail(RR_ERR_INVALID_ARGUMENT, "n_frames must be positive");
  int rc = check_render_args(ctx, cam, width, height, spp);
  if (rc) return rc;
  rc = rr_accum_reset(ctx, width, height);
  if (rc) return rc;
  rr_stats total;
  memset(&total, 0, sizeof(total));
  for (uint32_t k = 0; k < n_frames; ++k) {
    rr_stats st;
    rc = rr_accum_add_frame(ctx, cam, width, height, spp, max_bounces, (int32_t)((uint32_t)first_frame_index + k), tile_size,
                            k + 1 == n_frames ? rgba_out : nullptr, &st);
    if (rc) return rc;
    total.samples += st.samples; total.rays += st.rays; total.stack_overflows += st.stack_overflows; total.box_tests += st.box_tests;
    total.tri_tests += st.tri_tests; total.sphere_tests += st.sphere_tests; total.tiles += st.tiles;
    total.render_ms += st.render_ms; total.build_ms = st.build_ms;
  }
  if (stats_out) *stats_out = total;
  return RR_OK;
}

int rr_set_tile_order(rr_ctx* ctx, const uint32_t* tiles, uint32_t n_tiles) {
  if (!ctx || (n_tiles && !tiles)) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  for (Device& d : ctx->dev) {
    RR_CUDA(cudaSetDevice(d.ordinal));
    RR_CUDA(cudaStreamSynchronize(d.stream));
    cudaFree(d.tile_order);
    d.tile_order = nullptr; d.tile_order_n = 0;
    if (n_tiles) {
      RR_CUDA(cudaMalloc(&d.tile_order, (size_t)n_tiles * 4));
      RR_CUDA(cudaMemcpy(d.tile_order, tiles, (size_t)n_tiles * 4, cudaMemcpyHostToDevice));
      d.tile_order_n = n_tiles;
    }
  }
  return RR_OK;
}

int rr_render_cost(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_bounces,
                   uint32_t* segments_out) {
  int rc = check_render_args(ctx, cam, width, height, spp, max_bounces);
  if (rc) return rc;
  if (!segments_out) return fail(RR_ERR_INVALID_ARGUMENT, "null output");
  Device& d = ctx->dev[0];
  RR_CUDA(cudaSetDevice(d.ordinal));
  rc = ensure_frame(d, width, height, false);
  if (rc) return rc;
  RenderParams p;
  fill_params(ctx, d, cam, width, height, spp, max_bounces, 0, 0, p);
  if (too_many_items(p)) return fail(RR_ERR_INVALID_ARGUMENT, "image too large (its tiles hold more than 2^32 - 1 work items)");
  const size_t n = (size_t)width * height;
  uint32_t* dc = nullptr;
  cudaError_t e = cudaMalloc(&dc, n * 4);
  p.cost = dc;
  if (e == cudaSuccess) e = cudaMemsetAsync(dc, 0, n * 4, d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.queue, 0, sizeof(unsigned long long), d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.counters, 0, sizeof(Counters), d.stream);
  if (e == cudaSuccess) e = launch_render(p, true, frame_needs_slack(ctx, d, cam), RR_FEAT_ALL, d.sm_count, d.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
  if (e == cudaSuccess) e = cudaMemcpy(segments_out, dc, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(dc);
  RR_CUDA(e);
  return RR_OK;
}

int rr_primary_hits(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, int32_t* mesh_out,
                    int32_t* prim_out, float* dst_out) {
  int rc = check_render_args(ctx, cam, width, height, 1);
  if (rc) return rc;
  Device& d = ctx->dev[0];
  RR_CUDA(cudaSetDevice(d.ordinal));
  const size_t n = (size_t)width * height;
  RenderParams p;
  fill_params(ctx, d, cam, width, height, 1, 1, 0, 0, p);
  if (too_many_items(p)) return fail(RR_ERR_INVALID_ARGUMENT, "image too large (its tiles hold more than 2^32 - 1 work items)");
  int32_t *dm = nullptr, *dp = nullptr;
  float* dd = nullptr;
  cudaError_t e = cudaMalloc(&dm, n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&dp, n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&dd, n * 4);
  p.hit_mesh = dm; p.hit_prim = dp; p.hit_dst = dd;
  if (e == cudaSuccess) e = cudaMemsetAsync(d.queue, 0, sizeof(unsigned long long), d.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d.counters, 0, sizeof(Counters), d.stream);
  if (e == cudaSuccess) e = launch_primary(p, frame_needs_slack(ctx, d, cam), d.sm_count, d.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
  if (e == cudaSuccess && mesh_out) e = cudaMemcpy(mesh_out, dm, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && prim_out) e = cudaMemcpy(prim_out, dp, n * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && dst_out) e = cudaMemcpy(dst_out, dd, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(dm); cudaFree(dp); cudaFree(dd);
  if (e != cudaSuccess) return cuda_fail(e, "rr_primary_hits");
  return RR_OK;
}

int rr_bvh_size(rr_ctx* ctx, int which, uint64_t* n_prims) {
  if (!ctx || !n_prims) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  if (!ctx->has_scene) return fail(RR_ERR_NO_SCENE, "no scene");
  *n_prims = which ? ctx->dev[0].sb.n : ctx->dev[0].tb.n;
  return RR_OK;
}

int rr_bvh_read(rr_ctx* ctx, int which, uint64_t* codes, uint32_t* order, int32_t* left, int32_t* right,
                int32_t* parent, float* bounds) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  if (!ctx->has_scene) return fail(RR_ERR_NO_SCENE, "no scene");
  Device& d = ctx->dev[0];
  RR_CUDA(cudaSetDevice(d.ordinal));
  const Lbvh& b = which ? d.sb : d.tb;
  const size_t n = b.n;
  if (!n) return RR_OK;
  if (codes) RR_CUDA(cudaMemcpy(codes, b.codes, n * 8, cudaMemcpyDeviceToHost));
  if (order) RR_CUDA(cudaMemcpy(order, b.order, n * 4, cudaMemcpyDeviceToHost));
  if (left) RR_CUDA(cudaMemcpy(left, b.left, n * 4, cudaMemcpyDeviceToHost));
  if (right) RR_CUDA(cudaMemcpy(right, b.right, n * 4, cudaMemcpyDeviceToHost));
  if (parent) RR_CUDA(cudaMemcpy(parent, b.parent, n * 4, cudaMemcpyDeviceToHost));
  if (bounds) RR_CUDA(cudaMemcpy(bounds, b.bounds, n * 24, cudaMemcpyDeviceToHost));
  return RR_OK;
}

// ---- multi-process tile queue -------------------------------------------------
int rr_queue_export(rr_ctx* ctx, uint32_t width, uint32_t height, uint8_t* queue_handle, uint8_t* frame_handle) {
  if (!ctx || !queue_handle || !frame_handle) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == RR_IPC_HANDLE_BYTES, "IPC handle size");
  if (width == 0 || height == 0 || (uint64_t)width * height > 0x7fffffffull) return fail(RR_ERR_INVALID_ARGUMENT, "bad image size");
  Device& d = ctx->dev[0];
  RR_CUDA(cudaSetDevice(d.ordinal));
  if (d.shared_imported) return fail(RR_ERR_INVALID_ARGUMENT, "this context has imported a queue; export from a context of its own");
  const size_t bytes = (size_t)width * height * 4;
  if (d.shared_frame && bytes != d.shared_bytes) {  // a new size: the old handles die with the old allocation
    RR_CUDA(cudaStreamSynchronize(d.stream));
    cudaFree(d.shared_frame);
    d.shared_frame = nullptr; d.shared_bytes = 0;
  }
  if (!d.shared_frame) {
    RR_CUDA(cudaMalloc(&d.shared_frame, bytes));
    d.shared_bytes = bytes;
  }
  d.shared_w = width; d.shared_h = height;
  cudaIpcMemHandle_t hq, hf;
  RR_CUDA(cudaIpcGetMemHandle(&hq, d.queue));
  RR_CUDA(cudaIpcGetMemHandle(&hf, d.shared_frame));
  memcpy(queue_handle, &hq, sizeof(hq));
  memcpy(frame_handle, &hf, sizeof(hf));
  d.shared_queue = d.queue;
  d.shared_imported = false;
  d.shared_epoch = 0;
  RR_CUDA(cudaMemset(d.shared_queue, 0, sizeof(unsigned long long)));
  return RR_OK;
}

int rr_queue_import(rr_ctx* ctx, uint32_t width, uint32_t height, const uint8_t* queue_handle,
                    const uint8_t* frame_handle) {
  if (!ctx || !queue_handle || !frame_handle) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  Device& d = ctx->dev[0];
  if (width == 0 || height == 0 || (uint64_t)width * height > 0x7fffffffull) return fail(RR_ERR_INVALID_ARGUMENT, "bad image size");
  RR_CUDA(cudaSetDevice(d.ordinal));
  if (d.shared_imported) {
    if (d.shared_queue) cudaIpcCloseMemHandle(d.shared_queue);
    if (d.shared_frame) cudaIpcCloseMemHandle(d.shared_frame);
    d.shared_queue = nullptr; d.shared_frame = nullptr; d.shared_imported = false; d.shared_bytes = 0;
  } else if (d.shared_frame) {
    return fail(RR_ERR_INVALID_ARGUMENT, "this context has exported a queue; import into a context of its own");
  }
  cudaIpcMemHandle_t hq, hf;
  memcpy(&hq, queue_handle, sizeof(hq));
  memcpy(&hf, frame_handle, sizeof(hf));
  void *pq = nullptr, *pf = nullptr;
  RR_CUDA(cudaIpcOpenMemHandle(&pq, hq, cudaIpcMemLazyEnablePeerAccess));
  RR_CUDA(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
  d.shared_queue = (unsigned long long*)pq;
  d.shared_frame = (uint8_t*)pf;
  d.shared_bytes = (size_t)width * height * 4;  // as declared by the caller: must be the exported size
  d.shared_w = width; d.shared_h = height;
  d.shared_epoch = 0;
  d.shared_imported = true;
  return RR_OK;
}

int rr_queue_reset(rr_ctx* ctx) {
  if (!ctx) return fail(RR_ERR_INVALID_ARGUMENT, "null context");
  Device& d = ctx->dev[0];
  if (!d.shared_queue || d.shared_imported) return fail(RR_ERR_INVALID_ARGUMENT, "only the exporting rank resets the queue");
  RR_CUDA(cudaSetDevice(d.ordinal));
  // the epoch of the frame that follows: this rank's next rr_render_shared is its (shared_epoch + 1)-th
  const unsigned long long word = (unsigned long long)((d.shared_epoch + 1u) & 0xffffu) << RR_QUEUE_EPOCH_SHIFT;
  RR_CUDA(cudaMemcpy(d.shared_queue, &word, sizeof(word), cudaMemcpyHostToDevice));
  return RR_OK;
}

int rr_render_shared(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                     uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, rr_stats* stats_out) {
  return render_frame(ctx, cam, width, height, spp, max_bounces, frame_index, tile_size, false, false, 1, 0, 1, stats_out);
}

int rr_render_strided(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                      uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, uint32_t rank, uint32_t world,
                      rr_stats* stats_out) {
  if (world == 0 || rank >= world) return fail(RR_ERR_INVALID_ARGUMENT, "rank/world");
  return render_frame(ctx, cam, width, height, spp, max_bounces, frame_index, tile_size, false, false, 2, rank, world, stats_out);
}

int rr_frame_device_ptr(rr_ctx* ctx, uint64_t* ptr_out, uint64_t* bytes_out) {
  if (!ctx || !ptr_out) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  *ptr_out = (uint64_t)(uintptr_t)ctx->dev[0].frame;
  if (bytes_out) *bytes_out = ctx->dev[0].frame_bytes;
  return RR_OK;
}

// ---- probes used by the parity tests (include/rr_api.h "Test hooks") ------------
int rr_probe_math(int fn, const float* x, const float* y, float* out, uint64_t n) {
  if (!x || !out) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  float *dx = nullptr, *dy = nullptr, *dout = nullptr;
  cudaError_t e = cudaMalloc(&dx, n * 4 + 4);
  if (e == cudaSuccess) e = cudaMalloc(&dy, n * 4 + 4);
  if (e == cudaSuccess) e = cudaMalloc(&dout, n * 4 + 4);
  if (e == cudaSuccess) e = cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dy, y ? y : x, n * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = launch_math_probe(fn, dx, dy, dout, n, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dy); cudaFree(dout);
  if (e != cudaSuccess) return cuda_fail(e, "rr_probe_math");
  return RR_OK;
}

int rr_probe_rng(uint32_t pixel, int32_t frame, uint32_t* out_u32_8, float* out_f32_9) {
  if (!out_u32_8 || !out_f32_9) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  uint32_t* du = nullptr;
  float* df = nullptr;
  cudaError_t e = cudaMalloc(&du, 8 * 4);
  if (e == cudaSuccess) e = cudaMalloc(&df, 9 * 4);
  if (e == cudaSuccess) e = launch_rng_probe(pixel, frame, du, df, 0);
  if (e == cudaSuccess) e = cudaMemcpy(out_u32_8, du, 32, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_f32_9, df, 36, cudaMemcpyDeviceToHost);
  cudaFree(du); cudaFree(df);
  if (e != cudaSuccess) return cuda_fail(e, "rr_probe_rng");
  return RR_OK;
}

int rr_probe_peak(int what, float* value_out) {
  if (!value_out) return fail(RR_ERR_INVALID_ARGUMENT, "null argument");
  *value_out = 0.0f;
  int dev = 0;
  RR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  RR_CUDA(cudaGetDeviceProperties(&prop, dev));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  void* buf = nullptr;
  unsigned* out = nullptr;
  cudaError_t e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaMalloc(&out, 16);
  float best_ms = 1e30f;
  double work = 0.0;
  if (what == 0) {  // FP32 FMA: 8 CTAs of 256 threads per SM, 8 chains per thread
    const int iters = 1 << 14, grid = prop.multiProcessorCount * 8;
    work = (double)grid * 256 * iters * 8 * 2;  // flop
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) {
      cudaEventRecord(e0, 0);
      k_peak_fma<<<grid, 256>>>(reinterpret_cast<float*>(out), iters);
      cudaEventRecord(e1, 0);
      e = cudaEventSynchronize(e1);
      float ms = 0.0f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
      if (k > 0 && ms < best_ms) best_ms = ms;
    }
    if (e == cudaSuccess) *value_out = (float)(work / (best_ms * 1e-3) / 1e12);  // TFLOP/s
  } else {  // L2 read bandwidth: a 32 MB buffer read 64 times with 16-byte loads
    const size_t bytes = 32u << 20;
    const int passes = 64;
    if (e == cudaSuccess) e = cudaMalloc(&buf, bytes);
    if (e == cudaSuccess) e = cudaMemset(buf, 1, bytes);
    work = (double)bytes * passes;
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) {
      cudaEventRecord(e0, 0);
      k_peak_l2<<<prop.multiProcessorCount * 8, 256>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, passes, out);
      cudaEventRecord(e1, 0);
      e = cudaEventSynchronize(e1);
      float ms = 0.0f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
      if (k > 0 && ms < best_ms) best_ms = ms;
    }
    if (e == cudaSuccess) *value_out = (float)(work / (best_ms * 1e-3) / 1e9);  // GB/s
  }
  cudaFree(buf); cudaFree(out);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e != cudaSuccess) return cuda_fail(e, "rr_probe_peak");
  return RR_OK;
}

}  // extern "C"
