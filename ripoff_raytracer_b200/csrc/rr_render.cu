// rr_render.cu -- the per-pixel path-tracing kernel (sm_100a).
//
// B200-native replacement of the reference's `raytrace` OpenCL kernel
// (src/Trace.cl:623-653) and everything it calls.  Each device function names
// the reference lines whose arithmetic it reproduces.  Everything that decides
// a RESULT (triangle test, instance transforms, shading, RNG, tonemap) is kept
// in the reference's operation order and this file is compiled with
// -fmad=false, so those results are bit-identical to the CPU oracle
// (DESIGN.md section 3).  Ray/box tests only CULL: they run in FMA form with an
// approximate reciprocal against delta-inflated boxes (rr_internal.h box_delta),
// which keeps them conservative, so the closest hit is the same whatever order
// the hierarchy is walked in.
//
// Structure -- a warp-synchronous state machine:
//   * persistent warps take pixels from one 64-bit atomic counter (the reference's
//     mutex-guarded std::queue of tiles, src/image.hpp:286-314): the counter numbers
//     the pixels tile by tile (8x4 tiles) and a warp adds exactly as many as it has
//     free path slots, one atomic per refill;
//   * per pixel the spp samples run serially in the lane because the RNG state
//     is carried across samples (src/Trace.cl:632,639-642);
//   * every lane is in one of five phases (pixel, shade, mesh setup, node step,
//     leaf test); each round the warp votes and runs the phase with the most
//     ready lanes, so the hot node-step loop executes with most lanes active
//     instead of each lane walking its own ray while 31 others wait;
//   * a node step is one 128-byte fetch (a 4-wide node: the binary LBVH node
//     collapsed with its grandchildren) and four slab tests; a slot may postpone
//     one leaf and keep walking (speculative traversal);
//   * every warp owns a pool of 96 path slots (one pixel each), three times as many as lanes, so that the phase
//     that runs finds ~30 ready slots; the hot words of a slot live in shared memory, the words only the shade /
//     pixel phases touch and the traversal stack below its top in an L2-resident per-warp scratch.  96 registers,
//     5 CTAs of 4 warps per SM.
#include <math.h>

#include "rr_internal.h"
#include "rr_math.cuh"

namespace rr {

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __noinline__ V3 div3(V3 a, float s);
__device__ __forceinline__ V3 operator/(V3 a, float s) { return div3(a, s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Code size matters more than call overhead here: the SM's instruction cache holds 32 KB (about 2 000
// instructions) and every warp walks through all phases of the kernel every few microseconds.  The IEEE
// square root / division sequences and the software log / cos / powr are therefore kept out of line, one
// copy each.
// fast_normalize / normalize of the numerics contract
__device__ __noinline__ V3 normalize(V3 a) {
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
__device__ __noinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __noinline__ V3 div3(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ V3 ld3(const float* p) { return mk(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ V3 xyz(float4 q) { return mk(q.x, q.y, q.z); }

// ---- RNG: reference src/Trace.cl:158-217 (u32 arithmetic, exact) ----------
__device__ __forceinline__ float map_u32(uint32_t s) { return (float)(s + 1u) * (1.0f / 4294967296.0f); }
__device__ __forceinline__ float random_value(uint32_t& state) {
  state = state * 747796405u + 2891336453u;
  uint32_t result = ((state >> ((state >> 28) + 4u)) ^ state) * 277803737u;
  result = (result >> 22) ^ result;
  return map_u32(result);
}
__device__ __forceinline__ uint32_t make_seed(uint32_t pixelIndex, int32_t frameIndex, uint32_t rayIdx) {
  uint32_t s = pixelIndex * 1664525u + (uint32_t)frameIndex * 1013904223u;
  s ^= (rayIdx + 0x9E3779B9u);
  s = s * 22695477u + 1u;
  return s;
}
__device__ __forceinline__ float rand01(uint32_t& state) {
  state = state * 747796405u + 2891336453u;
  uint32_t z = state;
  z = (z ^ (z >> 16)) * 0x7feb352du;
  z = (z ^ (z >> 15)) * 0x846ca68bu;
  z = z ^ (z >> 16);
  return map_u32(z);
}
// src/Trace.cl:179-187 (out of line; the state goes in and out by value so that it stays in a register)
struct NormalDraw { float v; uint32_t state; };
__device__ __noinline__ NormalDraw random_normal_draw(uint32_t state) {
  float u1 = random_value(state);
  float u2 = random_value(state);
  u1 = fmaxf(u1, RR_EPSILON);
  float r = sqrtf(-2.0f * log_c(u1));
  float theta = RR_TAU * u2;
  NormalDraw d;
  d.v = r * cos_c(theta);
  d.state = state;
  return d;
}
__device__ __forceinline__ float random_normal(uint32_t& state) {
  const NormalDraw d = random_normal_draw(state);
  state = d.state;
  return d.v;
}
__device__ __forceinline__ bool finite_f(float x) { return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u; }
// src/Trace.cl:189-200
__device__ __forceinline__ V3 random_direction(uint32_t& state) {
  float x = random_normal(state);
  float y = random_normal(state);
  float z = random_normal(state);
  V3 v = normalize(mk(x, y, z));
  if (!finite_f(v.x) || !finite_f(v.y) || !finite_f(v.z)) v = mk(0.0f, 1.0f, 0.0f);
  return v;
}


// ---- intersection ------------------------------------------------------------
// Culling slab test (replaces src/Trace.cl:259-274 for traversal decisions): t = b*inv - o*inv in one
// FFMA (finite reciprocal, see rcp_approx).
// 1/x for the slab tests only.  |x| is clamped to 1e-18 so that the reciprocal stays finite: with an
// infinite reciprocal b*inv - o*inv is inf - inf = NaN on one plane only and the slab would reject
// rays that run inside it.
__device__ __forceinline__ float rcp_approx(float x) {
  const float xs = fabsf(x) < 1.0e-18f ? copysignf(1.0e-18f, x) : x;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(xs));
  return r;
}
// Per-ray part of the conservative culling slack.  The rounding error of a slab test (and of the triangle test's
// `origin - A`, src/Trace.cl:283) grows with the ray ORIGIN, not with the box.  As long as the (mesh-local) origin stays
// within 16x the largest box coordinate, the build-time box_delta (64 ulp of that coordinate) covers it and the kernel
// runs without any per-ray term (SLACK = false: nothing below costs an instruction or a register).  For a camera far
// outside, or a mesh posed with a tiny scale (origin / scale), the host (rr_api.cu frame_needs_slack) launches the
// SLACK = true instantiation: every slab is additionally widened by ray_slack(origin) = 2^-18 of the largest |origin
// coordinate| -- in the parametric form of the test ek = slack * |1/d| per axis (oracle/rr_oracle.c ray_slack states the
// same bound; tests: |origin| / extent up to 10^4, tests/test_gpu_parity.py::test_far_origin_is_bit_exact).
template <bool SLACK> struct RaySlack;
template <> struct RaySlack<true> {
  V3 e;
  __device__ __forceinline__ float x() const { return e.x; }
  __device__ __forceinline__ float y() const { return e.y; }
  __device__ __forceinline__ float z() const { return e.z; }
};
template <> struct RaySlack<false> {
  __device__ __forceinline__ float x() const { return 0.0f; }
  __device__ __forceinline__ float y() const { return 0.0f; }
  __device__ __forceinline__ float z() const { return 0.0f; }
};
__device__ __forceinline__ void make_slack(RaySlack<true>& r, const V3& o, const V3& inv) {
  const float s = ray_slack(o.x, o.y, o.z);
  r.e = mk(s * fabsf(inv.x), s * fabsf(inv.y), s * fabsf(inv.z));
}
__device__ __forceinline__ void make_slack(RaySlack<false>&, const V3&, const V3&) {}

template <bool SLACK>
__device__ __forceinline__ bool box_cull(float lox, float loy, float loz, float hix, float hiy, float hiz, const V3& inv,
                                         const V3& noi, const RaySlack<SLACK>& ek, float tbest, float& tn) {
  const float t0x = __fmaf_rn(lox, inv.x, noi.x), t1x = __fmaf_rn(hix, inv.x, noi.x);
  const float t0y = __fmaf_rn(loy, inv.y, noi.y), t1y = __fmaf_rn(hiy, inv.y, noi.y);
  const float t0z = __fmaf_rn(loz, inv.z, noi.z), t1z = __fmaf_rn(hiz, inv.z, noi.z);
  float tf;
  if (SLACK) {
    tn = fmaxf(fmaxf(fminf(t0x, t1x) - ek.x(), fminf(t0y, t1y) - ek.y()), fminf(t0z, t1z) - ek.z());
    tf = fminf(fminf(fmaxf(t0x, t1x) + ek.x(), fmaxf(t0y, t1y) + ek.y()), fmaxf(t0z, t1z) + ek.z());
  } else {
    tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  }
  return tf >= fmaxf(tn, 0.0f) && tn <= tbest;
}

constexpr int32_t NO_PRIM = 0x7fffffff;
// flags kept with the mesh number in the slot word W_M, so that the leaf phase needs no look-up in the mesh table
constexpr uint32_t WM_BACK = 1u << 31, WM_SPHERES = 1u << 30, WM_CULL = 1u << 29, WM_MESH = WM_CULL - 1u;

struct SceneHit {  // HitInfo of src/Trace.cl:67-74 as the shade phase sees it
  bool did;
  float dst;
  V3 point, normal;
  bool back;
  int32_t material;  // index into the material table
};

// ---- shading -------------------------------------------------------------------
__device__ __forceinline__ V3 lerp3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }               // :84
__device__ __forceinline__ V3 reflect3(V3 inDir, V3 n) { return inDir - (2 * dot(inDir, n)) * n; }        // :234-236
__device__ __forceinline__ V3 refract3(V3 inDir, V3 n, float iorA, float iorB) {                          // :219-232
  float ratio = iorA / iorB;
  float cosIn = -dot(inDir, n);
  float sinSqr = ratio * ratio * (1 - cosIn * cosIn);
  if (sinSqr > 1) return mk(0.0f, 0.0f, 0.0f);
  return ratio * inDir + (ratio * cosIn - sqrtf(1 - sinSqr)) * n;
}
__device__ __forceinline__ float reflectance(V3 inDir, V3 n, float iorA, float iorB) {                    // :401-432
  float ratio = iorA / iorB;
  float cosIn = -dot(inDir, n);
  if (cosIn <= 0) return 1;
  float sinSqr = ratio * ratio * (1 - cosIn * cosIn);
  if (sinSqr >= 1) return 1;
  float cosR = sqrtf(1 - sinSqr);
  float denPerp = iorA * cosIn + iorB * cosR;
  float denPar = iorA * cosIn + iorB * cosR;  // same expression twice, as in the reference (:414-416)
  if (fminf(denPerp, denPar) < RR_EPSILON) return 1;
  float rPerp = (iorA * cosIn - iorB * cosR) / denPerp;
  rPerp *= rPerp;
  float rPar = (iorB * cosIn - iorA * cosR) / denPar;
  rPar *= rPar;
  return (rPerp + rPar) / 2;
}

// One bounce of Trace() after the closest hit is known (src/Trace.cl:497-591).
// Returns false when the path ends.  `bounce` is advanced as the reference does.
// MATERIALS = false: the scene has only Solid and OneSided materials (RR_FEAT_MATERIALS); the other branches are left out.
template <bool MATERIALS>
__device__ __forceinline__ bool shade(const RenderParams& p, const SceneHit& hit, V3& origin, V3& dir, V3& throughput,
                                      V3& incoming, uint32_t& bounce, uint32_t& passes, uint32_t& rng) {
  if (!hit.did) return false;
  const DMaterial* M = p.materials + hit.material;
  // the 48-byte material record in three 16-byte loads, all requested before the first use
  const float4 m0 = __ldg(reinterpret_cast<const float4*>(M)), m1 = __ldg(reinterpret_cast<const float4*>(M) + 1),
               m2 = __ldg(reinterpret_cast<const float4*>(M) + 2);
  const int32_t type = __float_as_int(m0.x);
  if (MATERIALS && type == RR_MATERIAL_INVISIBLE) {
    // `continue` without counting a bounce (:502-506).  Guard: when hit.point + dir*1e-6 rounds back
    // to hit.point the reference loops forever; the path is ended after RR_MAX_INVISIBLE_PASSES.
    if (++passes > RR_MAX_INVISIBLE_PASSES) return false;
    origin = hit.point + dir * RR_EPSILON;
    return true;
  }
  V3 color = mk(m1.x, m1.y, m1.z);
  const V3 emissionColor = mk(m2.x, m2.y, m2.z);
  float emissionStrength = m0.z;
  const float specProb = m1.w;
  const float reflectiveness = m0.w;
  if (MATERIALS && type == RR_MATERIAL_CHECKER) {  // :509-533
    const float size = emissionStrength;
    const int xi = (int)floorf(hit.point.x / size);
    const int zi = (int)floorf(hit.point.z / size);
    const bool isEven = (((uint32_t)xi + (uint32_t)zi) & 1u) == 0u;
    color = isEven ? color : emissionColor;
    emissionStrength = 0.0f;
  }
  if ((MATERIALS && type == RR_MATERIAL_CHECKER) || type == RR_MATERIAL_SOLID) {  // :525-532, :559-567
    const bool isSpec = specProb >= random_value(rng);
    const V3 diffuseDir = normalize(hit.normal + random_direction(rng));
    const V3 specularDir = reflect3(dir, hit.normal);
    dir = normalize(lerp3(diffuseDir, specularDir, reflectiveness * (isSpec ? 1.0f : 0.0f)));
  } else if (MATERIALS && type == RR_MATERIAL_GLASSY) {  // :534-558
    const float ior = m0.y;
    const float iorCur = hit.back ? ior : 1.0f;
    const float iorNext = hit.back ? 1.0f : ior;
    const V3 reflectDir = reflect3(dir, hit.normal);
    const V3 refractDir = refract3(dir, hit.normal, iorCur, iorNext);
    const float reflectWeight = reflectance(dir, hit.normal, iorCur, iorNext);
    const float refractWeight = 1.0f - reflectWeight;
    const bool willReflect = rand01(rng) < reflectWeight;
    dir = willReflect ? reflectDir : refractDir;
    throughput = throughput * (willReflect ? reflectWeight : refractWeight);
  }
  // OneSided front face: direction unchanged (tinted pass-through costing one bounce)
  incoming = incoming + throughput * (emissionColor * emissionStrength);  // :575-576
  origin = hit.point + dir * RR_EPSILON;                                   // :579-580
  throughput = throughput * color;                                         // :582
  const float pmax = fmaxf(throughput.x, fmaxf(throughput.y, throughput.z));
  if (bounce > 3) {  // :585-590
    const float q = fmaxf(0.05f, 1.0f - pmax);
    if (rand01(rng) < q) return false;
    throughput = throughput / (1.0f - q);
  }
  bounce++;
  return true;
}

// src/Trace.cl:596-621 + the uv of :634-635
__device__ __noinline__ V3 primary_dir(const DCamera cam, uint32_t x, uint32_t y, uint32_t W, uint32_t H) {
  const float u = (float)x / (float)W;
  const float v = (float)(1.0f - (float)y / (float)H);
  float ndc0 = u * 2.0f - 1.0f;
  const float ndc1 = v * 2.0f - 1.0f;
  ndc0 *= cam.aspect;
  const float scale = tan_c((cam.fov * 0.5f) * 0.017453292519943295f);
  const V3 dc = normalize(mk(ndc0 * scale, ndc1 * scale, 1.0f));
  const float cx = cos_c(cam.pitch), sx = sin_c(cam.pitch);
  const float cy = cos_c(cam.yaw), sy = sin_c(cam.yaw);
  const float cz = cos_c(cam.roll), sz = sin_c(cam.roll);
  const V3 r0 = mk(cy * cz, cz * sy * sx - cx * sz, sx * sz + cx * cz * sy);
  const V3 r1 = mk(cy * sz, cx * cz + sx * sy * sz, cx * sy * sz - cz * sx);
  const V3 r2 = mk(-sy, cy * sx, cx * cy);
  return normalize(mk(dot(r0, dc), dot(r1, dc), dot(r2, dc)));
}

// src/Trace.cl:643-652 (+ host alpha = 255, src/image.hpp:271)
__device__ __noinline__ float gamma_c(float c) { return powr_c(fminf(fmaxf(c, 0.0f), 1.0f), 1.0f / 2.2f); }
__device__ __noinline__ uint32_t tonemap_rgba(V3 c) {
  const float r = gamma_c(c.x);
  const float g = gamma_c(c.y);
  const float b = gamma_c(c.z);
  const uint32_t R = (uint32_t)(unsigned char)(r * 255.0f), G = (uint32_t)(unsigned char)(g * 255.0f),
                 B = (uint32_t)(unsigned char)(b * 255.0f);
  return R | (G << 8) | (B << 16) | (255u << 24);
}


// ---- tile queue ----------------------------------------------------------------
// Tiles are numbered row-major.  With a shared counter (possibly in a peer
// GPU's memory, hence the system-scope atomic) every warp of every GPU pops the
// next tile; with a static partition rank r renders tiles r, r+world, ...
// The counter word is (frame epoch << 48 | tiles popped).  A pop that meets another epoch (rr_queue_reset ran while
// this launch was still popping, or a rank is a frame behind) takes nothing and is counted: rr_render_shared then
// fails with RR_ERR_QUEUE instead of rendering a tile of the wrong frame.
__device__ __forceinline__ bool pop_tile(const RenderParams& p, uint32_t& tile) {
  unsigned long long t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd_system(p.queue, 1ull);
  t = __shfl_sync(0xffffffffu, t, 0);
  if ((uint32_t)(t >> RR_QUEUE_EPOCH_SHIFT) != p.queue_epoch) {  // (counted at once: no register is kept for an event that must not happen)
    if ((threadIdx.x & 31) == 0) atomicAdd(&p.counters->queue_errors, 1ull);
    return false;
  }
  t &= (1ull << RR_QUEUE_EPOCH_SHIFT) - 1ull;
  t = (unsigned long long)p.tile_begin + t * p.tile_stride;
  tile = (uint32_t)t;
  return t < (unsigned long long)p.tiles_x * p.tiles_y;
}

constexpr int NT = RR_NT;        // threads per CTA (warps are independent: no block-level barrier in the loop)
constexpr int WARPS = NT / 32;
constexpr int POOL = RR_POOL;    // path slots per warp (rr_internal.h)
constexpr int ROUNDS = POOL / 32;
enum { PH_PIXEL = 0, PH_SHADE = 1, PH_SETUP = 2, PH_TRAV = 3, PH_LEAF = 4 };
// Vote key of a slot: one byte per phase, so that ONE warp reduction (REDUX) over the keys counts the
// ready slots of every phase.  A slot that needs a pixel has key 0 and W_PIX == PIX_NEED.
constexpr uint32_t K_T = 1u, K_L = 1u << 8, K_S = 1u << 16, K_H = 1u << 24;
constexpr int32_t PIX_NEED = -1, PIX_IDLE = -2;
// Words of one slot.  Hot words live in shared memory (word w of slot s at pool[w * POOL + s]); the words only
// the shade / pixel phases touch live in a per-warp global scratch with the same layout (cold[w * POOL + s]).
enum {
  W_KEY = 0, W_PIX,
  W_OX, W_OY, W_OZ, W_DX, W_DY, W_DZ,           // world ray
  W_BDST, W_BMAT,                               // closest hit so far: distance, material | back << 31 (mesh = min(material, n_meshes))
  W_CAND, W_M,                                  // candidate meshes of the chunk; (current mesh + 1, 0 = new ray) | WM_* flags
  W_LOX, W_LOY, W_LOZ, W_LDX, W_LDY, W_LDZ, W_LIX, W_LIY, W_LIZ,  // mesh-local ray and its reciprocal direction
  W_LT, W_LPRIM,                                // closest hit inside the current mesh (its normal: C_LNX)
  W_CUR, W_SPC, W_PSLOT,                        // traversal: node ref, stack pointer | postponed count << 8, postponed slot
  W_TOPN, W_TOPD,                               // the top entry of the traversal stack (the rest is in global scratch)
  NW
};
enum {
  C_RNG = 0, C_SAMPLE, C_BOUNCE,                // bounce | passes << 23
  C_BPRIM, C_BPX, C_BPY, C_BPZ, C_BNX, C_BNY, C_BNZ,  // primitive, point and normal of the closest hit
  C_LNX, C_LNY, C_LNZ,                          // normal of the closest hit inside the current mesh
  C_THR, C_THR1, C_THR2, C_INC, C_INC1, C_INC2, C_ACC, C_ACC1, C_ACC2, C_PD, C_PD1, C_PD2,
  NC
};
static_assert(NW == RR_POOL_WORDS && NC == RR_COLD_WORDS, "rr_internal.h RR_POOL_WORDS / RR_COLD_WORDS");
// Node references.  In the packed nodes a leaf is -((first slot << 2 | count - 1) + 2): up to RR_LEAF_MAX consecutive
// sorted slots (rr_lbvh.cu k_pack_wide), and "inner node or pop" is cur >= -1.
constexpr int32_t REF_POP = -1;                   // take the next entry of the stack
constexpr int32_t REF_END = (int32_t)0x80000000;  // traversal of this mesh finished
__device__ __forceinline__ bool ref_is_leaf(int32_t r) { return r < REF_POP && r != REF_END; }
// The leaf word (first slot << 2 | count - 1) of a leaf reference.  The node-step loop only carries it (pend_slot) with
// a "one leaf is postponed" flag (pend_cnt = 1); the leaf phase takes it apart.
__device__ __forceinline__ uint32_t ref_leaf_word(int32_t r) { return (uint32_t)(-r) - 2u; }

// World-box tests of the meshes [base, base + 32): bit k set = the ray enters mesh base + k's box before `tmax`.
// Every lane walks the whole chunk, so the loop is convergent.  Out of line: one copy serves shade, pixel and setup.
// With a top level (`blocks` != nullptr, more than 32 meshes) the chunk is four blocks of eight meshes and a block
// whose box the ray misses is skipped.
template <bool SLACK, bool TLAS>
__device__ __noinline__ uint32_t scan_meshes_fn(const DMesh* __restrict__ meshes, const float4* __restrict__ blocks, int32_t base,
                                                int32_t last_mesh, V3 winv, V3 wnoi, RaySlack<SLACK> wek, float tmax, unsigned* tests) {
  uint32_t mask = 0;
  const int32_t end = min(base + 32, last_mesh + 1);
  for (int32_t k0 = base; k0 < end; k0 += 8) {
    float tn;
    if (TLAS && blocks) {
      const float4 blo = __ldg(blocks + 2 * (k0 >> 3)), bhi = __ldg(blocks + 2 * (k0 >> 3) + 1);
      if (tests) ++*tests;
      if (!box_cull(blo.x, blo.y, blo.z, bhi.x, bhi.y, bhi.z, winv, wnoi, wek, tmax, tn)) continue;
    }
    const int32_t e = min(k0 + 8, end);
    if (tests) *tests += (unsigned)(e - k0);
    for (int32_t k = k0; k < e; ++k) {
      const DMesh* M = meshes + k;
      const float4 wlo = __ldg(&M->wmin), whi = __ldg(&M->wmax);
      const bool hit = box_cull(wlo.x, wlo.y, wlo.z, whi.x, whi.y, whi.z, winv, wnoi, wek, tmax, tn);
      if (hit && !(__float_as_uint(wlo.w) & RR_MF_SKIP)) mask |= 1u << (k - base);
    }
  }
  return mask;
}

// Top level over the meshes (replaces the linear mesh loop of src/Trace.cl:444-482 for scenes
// with many instances).  The meshes are stored in Morton order of their world boxes and carry an implicit tree of
// world boxes (rr_api.cu prepare_meshes): blocks of 8 meshes, chunks of 32, then every level groups 4 boxes of the
// level below (128, 512, 2048 ... meshes).  The walk needs no stack and no per-ray state: from chunk `c` onwards, at
// every level whose span starts at c (top level first) a box the ray misses skips its whole span; returns the first
// chunk the ray enters before `tmax`, or a value > last_chunk.  Out of line: scenes with at most 32 meshes never come here.
template <bool SLACK>
__device__ __noinline__ int32_t next_chunk_fn(const float4* __restrict__ tlas, const uint32_t* __restrict__ lv, int32_t c,
                                              int32_t last_chunk, V3 winv, V3 wnoi, RaySlack<SLACK> wek, float tmax, unsigned* tests) {
  const int levels = (int)__ldg(lv);  // lv: level count, then the first box of every level
  while (c <= last_chunk) {
    bool skipped = false;
    float tn;
    for (int l = levels - 1; l >= 0; --l) {  // level l: boxes over 4^l chunks (l = 0: the chunks themselves)
      const int32_t span = 1 << (2 * l);
      if (c & (span - 1)) continue;
      const float4* b = tlas + 2 * ((size_t)__ldg(lv + 1 + l) + (size_t)(c >> (2 * l)));
      const float4 lo = __ldg(b), hi = __ldg(b + 1);
      if (tests) ++*tests;
      if (!box_cull(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, winv, wnoi, wek, tmax, tn)) { c += span; skipped = true; break; }
    }
    if (!skipped) break;
  }
  return c;
}

#if RR_TOP_STAGE
__device__ __forceinline__ uint32_t smem_addr(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
#endif

// FEAT: the scene features this instantiation contains code for (RR_FEAT_*, rr_internal.h).  The kernel's cost is
// dominated by the code a warp streams through, so a scene without spheres, without Checker /
// Glassy / Invisible materials and with at most 32 meshes runs an instantiation that is 8 KB (15 %) smaller.
template <bool COUNT, bool PRIMARY, bool SLACK, int FEAT, bool TUNED>
__global__ void __launch_bounds__(NT, RR_MIN_CTAS) k_render(const RenderParams p) {
  constexpr bool F_SPHERES = (FEAT & RR_FEAT_SPHERES) != 0, F_MATERIALS = (FEAT & RR_FEAT_MATERIALS) != 0, F_TLAS = (FEAT & RR_FEAT_TLAS) != 0;
  extern __shared__ uint32_t pool_all[];
#if RR_TOP_STAGE
  // Top of the largest hierarchy staged in shared memory (north_star: "shared-memory or TMA staging of the BVH top
  // levels"): ONE bulk async copy (TMA, cp.async.bulk -> UBLKCP) per CTA brings the root and its inner children in,
  // completion through an mbarrier.  A/B switch: the production build has RR_TOP_STAGE = 0 (profiles/README.md).
  const float4* const top_s = reinterpret_cast<const float4*>(pool_all + WARPS * (NW * POOL + 32 + 8));
  __shared__ __align__(8) unsigned long long top_bar;
  if (p.top_count) {
    const uint32_t bar = smem_addr(&top_bar), dst = smem_addr(top_s), bytes = p.top_count * 128u;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(p.top_nodes), "r"(bytes), "r"(bar)
                   : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(bar) : "memory");
  }
#endif
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu;
  const unsigned lanes_below = (1u << lane) - 1u;
  uint32_t* const pool = pool_all + warp * (NW * POOL + 32 + 8);
  uint32_t* const sel = pool + NW * POOL;  // slot picked for each lane in the current phase
  uint32_t* const tstate = sel + 32;       // the warp's current tile (only the pixel phase touches it): x0, y0, w, next, pixels
  uint2* const stack = p.stack + (size_t)(blockIdx.x * WARPS + warp) * ((size_t)p.stack_entries * POOL);  // [depth][slot]
  uint32_t* const cold = p.cold + (size_t)(blockIdx.x * WARPS + warp) * (NC * POOL);
#define CW(w, s) cold[(w) * POOL + (s)]
#define CF(w, s) __uint_as_float(cold[(w) * POOL + (s)])
#define CSF(w, s, v) cold[(w) * POOL + (s)] = __float_as_uint(v)
#define CLD3(w, s) mk(CF(w, s), CF((w) + 1, s), CF((w) + 2, s))
#define CST3(w, s, v) do { CSF(w, s, (v).x); CSF((w) + 1, s, (v).y); CSF((w) + 2, s, (v).z); } while (0)
#define PW(w, s) pool[(w) * POOL + (s)]
#define PF(w, s) __uint_as_float(pool[(w) * POOL + (s)])
#define PSF(w, s, v) pool[(w) * POOL + (s)] = __float_as_uint(v)
#define PLD3(w, s) mk(PF(w, s), PF((w) + 1, s), PF((w) + 2, s))
#define PST3(w, s, v) do { PSF(w, s, (v).x); PSF((w) + 1, s, (v).y); PSF((w) + 2, s, (v).z); } while (0)
  const V3 cam_pos = mk(p.cam.pos[0], p.cam.pos[1], p.cam.pos[2]);
  // warp-uniform state
  bool more_pixels = true;  // the tile queue or the warp's current tile still holds pixels
  if (lane < 8) tstate[lane] = lane == 2 ? 1u : 0u;
  uint32_t n_need = p.pool_use;  // slots waiting for a pixel
  // statistics
  uint32_t n_rays = 0, n_tiles = 0;  // per lane / per warp: far below 2^32 even for an 8K, 1024-spp frame on one GPU
  unsigned c_box = 0, c_tri = 0, c_sph = 0;
  unsigned ph_runs[5] = {0, 0, 0, 0, 0}, ph_lanes[5] = {0, 0, 0, 0, 0};
  unsigned long long t_empty = 0;  // COUNT: %globaltimer when this warp found the tile queue empty

  // Scheduler knobs: compile-time constants (the defaults of default_tuning) unless rr_set_tuning changed them
  // (TUNED: +1.5 % C4, +3.5 % C2 for not carrying them, profiles/ab_r02_fixed_tuning.jsonl)
  const uint32_t wP = TUNED ? p.tune.weight[PH_PIXEL] : 1u, wH = TUNED ? p.tune.weight[PH_SHADE] : 1u,
                 wS = TUNED ? p.tune.weight[PH_SETUP] : 1u, wT = TUNED ? p.tune.weight[PH_TRAV] : 1u,
                 wL = TUNED ? p.tune.weight[PH_LEAF] : 1u;
  const uint32_t trav_keep = TUNED ? p.tune.trav_keep : (uint32_t)RR_TRAV_KEEP_DEFAULT;
  const bool speculate = TUNED ? (p.tune.speculate & 1u) != 0 : true;

#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    PW(W_KEY, lane + 32 * r) = 0u;
    PW(W_PIX, lane + 32 * r) = lane + 32 * r < p.pool_use ? (uint32_t)PIX_NEED : (uint32_t)PIX_IDLE;
  }
  __syncwarp();

  // ---- per-slot working set of a phase (loaded from / stored to the pool by the phase) ----
  int s = -1;  // the slot this lane works on
  int32_t pix = 0;
  uint32_t rng = 0, sample = 0, bounce = 0, passes = 0;
  V3 origin = cam_pos, dir = mk(0, 0, 1);
  float best_dst = INFINITY;
  int32_t best_mat = 0, best_mesh = 0x7fffffff, best_prim = -1;
  bool best_back = false;
  uint32_t cand = 0;  // candidate meshes of the current 32-mesh chunk (bit k = mesh (m & ~31) + k) not yet visited
  int m = 0;
  uint32_t mflags = 0;
  V3 lo = origin, ld = dir, linv = dir, lnoi = dir;
  float lt = INFINITY;
  int32_t lprim = NO_PRIM;
  bool lback = false;
  int32_t cur = REF_END;
  int sp = 0;
  uint32_t pend_slot = 0, pend_cnt = 0;

  // Up to 32 slots with `ready` set are handed to the lanes (lane j gets the j-th ready slot); returns how many.
  auto select = [&](const bool (&ready)[ROUNDS]) -> int {
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {  // static indexing only: a runtime-indexed ready[] would live in local memory
      const unsigned b = __ballot_sync(full, ready[r]);
      const int rank = base + __popc(b & lanes_below);
      if (ready[r] && rank < 32) sel[rank] = lane + 32u * r;
      base += __popc(b);
    }
    __syncwarp();
    const int n = min(base, 32);
    s = (int)lane < n ? (int)sel[lane] : -1;
    return n;
  };
  auto trav_key = [&]() -> uint32_t {
    if (cur == REF_END && pend_cnt == 0) return K_S;  // mesh done: the setup phase finishes it and enters the next one
    return ((cur >= REF_POP && (speculate || pend_cnt == 0)) ? K_T : 0u) | (pend_cnt ? K_L : 0u);
  };

  auto scan_meshes = [&](int32_t base, const V3& winv, const V3& wnoi, const RaySlack<SLACK>& wek, float tmax) -> uint32_t {
    unsigned tests = 0;
    const uint32_t mask = scan_meshes_fn<SLACK, F_TLAS>(p.meshes, p.tlas_blocks, base, p.last_mesh, winv, wnoi, wek, tmax, COUNT ? &tests : nullptr);
    if (COUNT) c_box += tests;
    return mask;
  };
  // The mesh just traversed has a closest hit (local space): LocalToWorldHit and the keep-min of
  // src/Trace.cl:465-481.  Meshes are visited in our own order, so equal distances are resolved by the
  // original mesh index, which is what the reference's in-order strict `<` does.
  auto finish_mesh = [&]() {
    if (lprim == NO_PRIM) return;
    const DMesh* M = p.meshes + m;
    const int32_t mesh_index = __float_as_int(__ldg(&M->wmax.w));
    if (F_SPHERES && (mflags & RR_MF_SPHERES)) {
      const int32_t mat = mesh_index + lprim;
      const int32_t type = __ldg(&p.materials[mat].type);
      if (!(type == RR_MATERIAL_ONESIDED && lback) && (lt < best_dst || (lt == best_dst && mesh_index < best_mesh))) {
        best_dst = lt; best_mat = mat; best_back = lback; best_mesh = mesh_index; best_prim = lprim;
        const V3 wp = origin + dir * lt;
        CST3(C_BPX, s, wp);
        CW(C_BNX, s) = CW(C_LNX, s); CW(C_BNY, s) = CW(C_LNY, s); CW(C_BNZ, s) = CW(C_LNZ, s);
      }
    } else {
      const int32_t type = (int32_t)((mflags >> RR_MF_TYPE_SHIFT) & 0xffu);
      if (!(type == RR_MATERIAL_ONESIDED && lback)) {
        // LocalToWorldHit, src/Trace.cl:139-156
        const float4 r0 = __ldg(&M->r0), r1 = __ldg(&M->r1), r2 = __ldg(&M->r2);
        const V3 pos = mk(__ldg(&M->ri0.w), __ldg(&M->ri1.w), __ldg(&M->ri2.w));
        const V3 lp = (lo + ld * lt) * r0.w;
        const V3 wp = mk(dot(xyz(r0), lp), dot(xyz(r1), lp), dot(xyz(r2), lp)) + pos;
        const V3 ln = CLD3(C_LNX, s);
        const V3 wn = normalize(mk(dot(xyz(r0), ln), dot(xyz(r1), ln), dot(xyz(r2), ln)));
        const float wd = length(wp - origin);
        if (wd < best_dst || (wd == best_dst && mesh_index < best_mesh)) {
          best_dst = wd; best_mat = mesh_index; best_back = lback; best_mesh = mesh_index; best_prim = lprim;
          CST3(C_BPX, s, wp);
          CST3(C_BNX, s, wn);
        }
      }
    }
    lprim = NO_PRIM;
  };
  // Enters the next candidate mesh (src/Trace.cl:444-463): world box against the closest hit so far,
  // WorldToLocalRay, root box.  Returns the slot's new key (traversal started, or K_H: no candidate left).
  // Two stages per trip.  (1) The cheap, divergent part: lanes look for their next candidate whose world box still lies in
  // front of the closest hit so far (two loads and a box test per candidate; lanes differ in trips).  (2) The expensive
  // part -- the mesh record, WorldToLocalRay with its normalisation, the root box -- runs ONCE for all lanes that found
  // a candidate (structured control flow: every lane leaves stage 1 through the same exit, so they reconverge before
  // stage 2), instead of once per trip of the search with whichever lanes got that far.
  auto enter_next_mesh = [&](const V3& winv, const V3& wnoi, const RaySlack<SLACK>& wek) -> uint32_t {
    uint32_t key = K_H;
    bool searching = true;
    do {
      bool found = false;
      while (searching && !found) {
        if (cand == 0) {
          int32_t base = (m & ~31) + 32;
          if (F_TLAS && p.tlas && base <= p.last_mesh) {  // skip the chunks (and groups of chunks) the ray does not enter before the closest hit so far
            unsigned tests = 0;
            base = next_chunk_fn<SLACK>(p.tlas, p.tlas_levels, base >> 5, p.last_mesh >> 5, winv, wnoi, wek, best_dst, COUNT ? &tests : nullptr) << 5;
            if (COUNT) c_box += tests;
          }
          if (base > p.last_mesh) {
            searching = false;  // no candidate left: K_H
          } else {
            m = base;
            cand = scan_meshes(base, winv, wnoi, wek, best_dst);
          }
        } else {
          const int k = __ffs((int)cand) - 1;
          cand &= cand - 1u;
          m = (m & ~31) + k;
          found = true;
          if (best_dst < INFINITY) {  // a hit exists: the box may now lie behind it
            const DMesh* M = p.meshes + m;
            const float4 wlo = __ldg(&M->wmin), whi = __ldg(&M->wmax);
            float tn;
            if (COUNT) c_box++;
            found = box_cull(wlo.x, wlo.y, wlo.z, whi.x, whi.y, whi.z, winv, wnoi, wek, best_dst, tn);
          }
        }
      }
      if (found) {
        const DMesh* M = p.meshes + m;
        const float4 i0 = __ldg(&M->ri0), i1 = __ldg(&M->ri1), i2 = __ldg(&M->ri2);
        const float4 blo = __ldg(&M->bmin), bhi = __ldg(&M->bmax);
        mflags = __float_as_uint(__ldg(&M->wmin.w));
        RaySlack<SLACK> lek = wek;
        if (F_SPHERES && (mflags & RR_MF_SPHERES)) {
          lo = origin; ld = dir; linv = winv; lnoi = wnoi;
        } else {
          // WorldToLocalRay, src/Trace.cl:118-137
          const V3 rel = origin - mk(i0.w, i1.w, i2.w);
          lo = mk(dot(xyz(i0), rel), dot(xyz(i1), rel), dot(xyz(i2), rel));
          ld = mk(dot(xyz(i0), dir), dot(xyz(i1), dir), dot(xyz(i2), dir));
          if (!(mflags & RR_MF_UNIT)) {
            const float scale = __ldg(&M->r0.w);
            if (mflags & RR_MF_POW2) {  // x / 2^k == x * 2^-k exactly
              const float is = __ldg(&M->r1.w);
              lo = lo * is; ld = ld * is;
            } else if (fabsf(scale) > RR_EPSILON) {
              lo = lo / scale; ld = ld / scale;
            }
          }
          ld = normalize(ld);
          linv = mk(rcp_approx(ld.x), rcp_approx(ld.y), rcp_approx(ld.z));
          lnoi = mk(-(lo.x * linv.x), -(lo.y * linv.y), -(lo.z * linv.z));
          make_slack(lek, lo, linv);
        }
        float tn;
        if (COUNT) c_box++;
        if (box_cull(blo.x, blo.y, blo.z, bhi.x, bhi.y, bhi.z, linv, lnoi, lek, INFINITY, tn)) {
          const uint32_t first = __float_as_uint(blo.w), count = __float_as_uint(bhi.w);
          // The sphere set lives in world space and is the last entry in the tie order (highest mesh index): a sphere at or
          // beyond the closest hit so far can never win, so its walk starts with that distance as the bound.
          lt = (F_SPHERES && (mflags & RR_MF_SPHERES)) ? best_dst : INFINITY;
          lprim = NO_PRIM; lback = false;
          sp = 0;
          if (count <= RR_DIRECT_MAX) {  // no hierarchy: the primitives are tested one by one in the leaf phase
            pend_slot = (((F_SPHERES && (mflags & RR_MF_SPHERES)) ? 0u : first) << 2) | (count - 1u);  // leaf word: all 1 - 4 primitives at once
            pend_cnt = 1;
            cur = REF_END;
          } else {
            pend_cnt = 0;
            cur = (int32_t)first;  // root node
#if RR_TOP_STAGE
            if (p.top_count && cur == p.top_root) cur = RR_TOP_TAG;  // the staged copy of this root
#endif
          }
          key = trav_key();
          searching = false;
        }
      }
    } while (searching);
    return key;
  };
  // what setup and shade write back after finish_mesh / enter_next_mesh
  auto store_ray_state = [&](uint32_t key) {
    PSF(W_BDST, s, best_dst);
    PW(W_BMAT, s) = (uint32_t)best_mat | (best_back ? 0x80000000u : 0u);
    if (PRIMARY) CW(C_BPRIM, s) = (uint32_t)best_prim;
    PW(W_CAND, s) = cand;
    PW(W_M, s) = (uint32_t)(m + 1) | (lback ? WM_BACK : 0u) | ((F_SPHERES && (mflags & RR_MF_SPHERES)) ? WM_SPHERES : 0u) |
                 ((mflags & RR_MF_CULL) ? WM_CULL : 0u);
    PST3(W_LOX, s, lo);
    PST3(W_LDX, s, ld);
    PST3(W_LIX, s, linv);
    PSF(W_LT, s, lt);
    PW(W_LPRIM, s) = (uint32_t)lprim;
    PW(W_CUR, s) = (uint32_t)cur;
    PW(W_SPC, s) = (uint32_t)sp | (pend_cnt << 8);
    PW(W_PSLOT, s) = pend_slot;
    PW(W_KEY, s) = key;
  };
  auto load_ray_state = [&]() {
    best_dst = PF(W_BDST, s);
    const uint32_t bm = PW(W_BMAT, s);
    best_mat = (int32_t)(bm & 0x7fffffffu);
    best_back = (bm >> 31) != 0u;
    best_mesh = best_dst < INFINITY ? min(best_mat, p.n_meshes) : 0x7fffffff;
    if (PRIMARY) best_prim = (int32_t)CW(C_BPRIM, s);
    cand = PW(W_CAND, s);
    mflags = __float_as_uint(__ldg(&p.meshes[m].wmin.w));  // m and lback were decoded from W_M by the caller
    lo = PLD3(W_LOX, s);
    ld = PLD3(W_LDX, s);
    lt = PF(W_LT, s);
    lprim = (int32_t)PW(W_LPRIM, s);
  };
  auto load_mesh_word = [&]() {  // W_M -> m (-1: new ray), lback
    const uint32_t mw = PW(W_M, s);
    m = (int)(mw & WM_MESH) - 1;
    lback = (mw & WM_BACK) != 0u;
  };
  // shade / pixel start the next segment: reset the closest hit (src/Trace.cl:437-444), collect the candidate
  // meshes (convergent here: every lane of these phases does it) and hand the slot to the setup phase
  auto store_new_ray = [&]() {
    PST3(W_OX, s, origin);
    PST3(W_DX, s, dir);
    const V3 winv = mk(rcp_approx(dir.x), rcp_approx(dir.y), rcp_approx(dir.z));
    const V3 wnoi = mk(-(origin.x * winv.x), -(origin.y * winv.y), -(origin.z * winv.z));
    RaySlack<SLACK> wek;
    make_slack(wek, origin, winv);
    PW(W_CAND, s) = scan_meshes(0, winv, wnoi, wek, INFINITY);
    PW(W_M, s) = 1u;  // chunk 0, no backface flag
    PSF(W_BDST, s, INFINITY);
    PW(W_BMAT, s) = 0u;
    PW(W_LPRIM, s) = (uint32_t)NO_PRIM;
    if (PRIMARY) CW(C_BPRIM, s) = 0xffffffffu;
    else if (COUNT) CW(C_BPRIM, s) = CW(C_BPRIM, s) + 1u;  // instrumented kernel: path segments of the slot's pixel so far (rr_render_cost)
    n_rays++;
    PW(W_KEY, s) = K_S;
  };

  for (;;) {
    // ---- vote: one REDUX over the byte-packed keys counts the ready slots of every phase ----
    __syncwarp();  // the slots written by the previous phase are visible to every lane
    uint32_t ksum = 0, keys[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) { keys[r] = PW(W_KEY, lane + 32 * r); ksum += keys[r]; }
    const uint32_t counts = __reduce_add_sync(full, ksum);
    const uint32_t nP = more_pixels ? n_need : 0u;
    if (counts == 0 && nP == 0) break;
    const uint32_t nT = min(counts & 0xffu, 32u), nL = min((counts >> 8) & 0xffu, 32u), nS = min((counts >> 16) & 0xffu, 32u),
                   nH = min(counts >> 24, 32u);
    int phase = PH_TRAV;
    uint32_t best = nT * wT;
    if (nL * wL > best) { best = nL * wL; phase = PH_LEAF; }
    if (nS * wS > best) { best = nS * wS; phase = PH_SETUP; }
    if (nH * wH > best) { best = nH * wH; phase = PH_SHADE; }
    if (min(nP, 32u) * wP > best) { best = min(nP, 32u) * wP; phase = PH_PIXEL; }
    // ---- gather: up to 32 slots that are ready for the chosen phase go to the lanes (ONE copy of this code for all
    // phases: the phases a warp cycles through have to share a 32 KB instruction cache) ----
    int n_sel;
    {
      bool ready[ROUNDS];
      if (phase == PH_PIXEL) {
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) ready[r] = (int32_t)PW(W_PIX, lane + 32 * r) == PIX_NEED;
      } else {
        const uint32_t km = phase == PH_TRAV ? K_T : phase == PH_LEAF ? K_L : phase == PH_SETUP ? K_S : K_H;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) ready[r] = (keys[r] & km) != 0u;
      }
      n_sel = select(ready);
    }
    if (phase == PH_TRAV) {
      // ================= node steps =================
      const bool mine = s >= 0;
      if (mine) {
        cur = (int32_t)PW(W_CUR, s);
        const uint32_t spc = PW(W_SPC, s);
        sp = (int)(spc & 0xffu);
        pend_cnt = spc >> 8;
        pend_slot = PW(W_PSLOT, s);
        lt = PF(W_LT, s);
        lo = PLD3(W_LOX, s);
        linv = PLD3(W_LIX, s);
        lnoi = mk(-(lo.x * linv.x), -(lo.y * linv.y), -(lo.z * linv.z));
      } else {
        cur = REF_END; sp = 0; pend_cnt = 0;
      }
      // SLACK: the per-ray slack is folded into the offsets -- entry planes move towards the origin, exit planes away
      // (three more registers in the loop; without SLACK nnoi and fnoi ARE lnoi)
      V3 nnoi = lnoi, fnoi = lnoi;
      if (SLACK) {
        RaySlack<true> lek;
        make_slack(lek, lo, linv);
        nnoi = lnoi - lek.e;
        fnoi = lnoi + lek.e;
      }
      // which quad of a node holds the entry / exit plane of each axis (node layout: min.x min.y min.z max.x max.y max.z)
      const int qnx = linv.x < 0.0f ? 3 : 0, qny = linv.y < 0.0f ? 4 : 1, qnz = linv.z < 0.0f ? 5 : 2;
      const int qfx = 3 - qnx, qfy = 5 - qny, qfz = 7 - qnz;
      uint2* const stk = stack + (mine ? s : 0);
      // Entry sp-1 (the top) is kept in registers / shared memory, entries 0..sp-2 in the global scratch.  A pop
      // hands out the top at once and only ISSUES the load of the next entry, so that its L2 latency overlaps
      // with the node fetch that follows instead of preceding it.
      uint2 top = make_uint2(0u, 0u);
      if (mine && sp > 0) top = make_uint2(PW(W_TOPN, s), PW(W_TOPD, s));
      // ... and a register copy of entry sp-2 (`below`), fetched one pop ahead: when a popped entry turns out to lie
      // behind the closest hit, the next one is already here instead of an L2 round trip away.
      uint2 below = make_uint2(0u, 0u);
      if (mine && sp > 1) below = stk[(sp - 2) * POOL];
      // Pops the stack / postpones leaves until `cur` is an inner node, a leaf the slot must wait for, or REF_END.
      auto resolve = [&](int32_t next) {
        for (;;) {
          if (next == REF_POP) {
            bool found = false;
            while (sp > 0) {
              const uint2 e = top;
              --sp;
              top = below;
              if (sp > 1) below = stk[(sp - 2) * POOL];
              if (__uint_as_float(e.y) <= lt) { next = (int32_t)e.x; found = true; break; }
            }
            if (!found) { cur = REF_END; break; }
          }
          if (next >= 0) { cur = next; break; }
          if (pend_cnt == 0) { pend_slot = ref_leaf_word(next); pend_cnt = 1; next = REF_POP; continue; }
          cur = next;  // a second leaf while one is postponed: wait for the leaf phase
          break;
        }
      };
      // The loop is left when fewer than `keep` lanes can still step (the others idle meanwhile).  With a full pool that is
      // trav_keep of the ~25 lanes a run starts with; a run that starts with fewer (a warp whose pool is draining: the
      // tail of a frame) would leave after every single node and pay a vote + gather per step, so its
      // threshold is three quarters of the lanes it started with.
#if RR_ADAPTIVE_KEEP
      const uint32_t keep = min(trav_keep, max(((uint32_t)n_sel * (uint32_t)RR_KEEP_NUM) >> 2, 1u));
#else
      const uint32_t keep = trav_keep;
#endif
      uint32_t active;
      do {
        const bool step = mine && cur >= REF_POP && (speculate || pend_cnt == 0);
        if (COUNT) { ph_runs[PH_TRAV]++; ph_lanes[PH_TRAV] += __popc(__ballot_sync(full, step)); }
        if (step) {
          int32_t next = REF_POP;  // cur == REF_POP (the slot comes from the leaf phase): only take the next stack entry
          if (cur >= 0) {
            // one 128-byte node: the boxes of up to four children (SoA) and their references.  The quads holding the
            // planes the ray ENTERS / LEAVES through are picked by the sign of its direction (qn*/qf*, set up
            // once per run), so no per-child min/max is needed to order the two planes of a slab.
#if RR_TOP_STAGE
            float4 nx, ny, nz, fx, fy, fz, rf;
            if (cur >= RR_TOP_TAG) {  // staged node: seven 16-byte loads from shared memory
              const float4* nd = top_s + RR_NODE_QUADS * (size_t)(cur - RR_TOP_TAG);
              nx = nd[qnx]; ny = nd[qny]; nz = nd[qnz]; fx = nd[qfx]; fy = nd[qfy]; fz = nd[qfz]; rf = nd[6];
              if (COUNT) c_box += (unsigned)__float_as_int(nd[7].x);
            } else {
              const float4* nd = p.nodes + RR_NODE_QUADS * (size_t)cur;
              nx = __ldg(nd + qnx); ny = __ldg(nd + qny); nz = __ldg(nd + qnz); fx = __ldg(nd + qfx);
              fy = __ldg(nd + qfy); fz = __ldg(nd + qfz); rf = __ldg(nd + 6);
              if (COUNT) c_box += (unsigned)__float_as_int(__ldg(&nd[7].x));
            }
#else
            const float4* nd = p.nodes + RR_NODE_QUADS * (size_t)cur;
            const float4 nx = __ldg(nd + qnx), ny = __ldg(nd + qny), nz = __ldg(nd + qnz), fx = __ldg(nd + qfx),
                         fy = __ldg(nd + qfy), fz = __ldg(nd + qfz), rf = __ldg(nd + 6);
            if (COUNT) c_box += (unsigned)__float_as_int(__ldg(&nd[7].x));
#endif
            // sort key of a child: entry distance (clamped at 0, two low mantissa bits dropped) | child number;
            // a child the ray misses (or an unused NaN box) sorts last
            auto child_key = [&](float pnx, float pny, float pnz, float pfx, float pfy, float pfz, uint32_t c) -> uint32_t {
              const float tn = fmaxf(fmaxf(__fmaf_rn(pnx, linv.x, nnoi.x), __fmaf_rn(pny, linv.y, nnoi.y)), __fmaf_rn(pnz, linv.z, nnoi.z));
              const float tf = fminf(fminf(__fmaf_rn(pfx, linv.x, fnoi.x), __fmaf_rn(pfy, linv.y, fnoi.y)), __fmaf_rn(pfz, linv.z, fnoi.z));
              const float t0 = fmaxf(tn, 0.0f);
              return (tf >= t0 && tn <= lt) ? ((__float_as_uint(t0) & ~3u) | c) : 0xffffffffu;
            };
            uint32_t k0 = child_key(nx.x, ny.x, nz.x, fx.x, fy.x, fz.x, 0u), k1 = child_key(nx.y, ny.y, nz.y, fx.y, fy.y, fz.y, 1u),
                     k2 = child_key(nx.z, ny.z, nz.z, fx.z, fy.z, fz.z, 2u), k3 = child_key(nx.w, ny.w, nz.w, fx.w, fy.w, fz.w, 3u);
            {  // 5-comparator sorting network, ascending
              uint32_t t;
              t = min(k0, k1); k1 = max(k0, k1); k0 = t;
              t = min(k2, k3); k3 = max(k2, k3); k2 = t;
              t = min(k0, k2); k2 = max(k0, k2); k0 = t;
              t = min(k1, k3); k3 = max(k1, k3); k1 = t;
              t = min(k1, k2); k2 = max(k1, k2); k1 = t;
            }
            // the references in sorted order, computed once for all lanes (two-level select on the child number)
            auto ref_of = [&](uint32_t k) -> int32_t {
              const float lo2 = (k & 1u) ? rf.y : rf.x, hi2 = (k & 1u) ? rf.w : rf.z;
              return __float_as_int((k & 2u) ? hi2 : lo2);
            };
            int32_t r0 = ref_of(k0), r1 = ref_of(k1), r2 = ref_of(k2), r3 = ref_of(k3);
            int hits = (k0 != 0xffffffffu) + (k1 != 0xffffffffu) + (k2 != 0xffffffffu) + (k3 != 0xffffffffu);
            if (hits > 0 && r0 < REF_POP && pend_cnt == 0) {  // the nearest child is a leaf: postpone it, go on with the next one
              pend_slot = ref_leaf_word(r0); pend_cnt = 1;
              k0 = k1; k1 = k2; k2 = k3;
              r0 = r1; r1 = r2; r2 = r3;
              hits--;
            }
            // misses sort last, so the children to push are a suffix of the hits: farthest first, the nearest of
            // them ends up as the (register-resident) top; the previous top is spilled once
            if (hits >= 2 && sp + hits - 1 > (int)p.stack_entries) {
              tstate[5] = 1u;  // cannot happen (3 entries per wide level + 4 are allocated); flagged, not hidden: RR_ERR_BVH_DEPTH
            } else if (hits >= 2) {
              uint2* const w = stk + sp * POOL;  // one address; the stores below use constant offsets from it
              if (sp > 0) w[-POOL] = top;
              if (hits >= 3) w[0] = hits == 4 ? make_uint2((uint32_t)r3, k3 & ~3u) : make_uint2((uint32_t)r2, k2 & ~3u);
              if (hits == 4) w[POOL] = make_uint2((uint32_t)r2, k2 & ~3u);
              below = hits >= 3 ? make_uint2((uint32_t)r2, k2 & ~3u) : top;  // the entry under the new top
              top = make_uint2((uint32_t)r1, k1 & ~3u);
              sp += hits - 1;
            }
            if (hits > 0) next = r0;
          }
          resolve(next);  // the one copy of the pop / postpone logic
        }
        active = __popc(__ballot_sync(full, mine && cur >= REF_POP && (speculate || pend_cnt == 0)));
      } while (active >= keep);
      if (mine) {
        PW(W_CUR, s) = (uint32_t)cur;
        PW(W_SPC, s) = (uint32_t)sp | (pend_cnt << 8);
        PW(W_PSLOT, s) = pend_slot;
        PW(W_TOPN, s) = top.x;
        PW(W_TOPD, s) = top.y;
        PW(W_KEY, s) = trav_key();
      }
    } else if (phase == PH_LEAF) {
      // ================= leaf tests =================
      if (COUNT) { ph_runs[PH_LEAF]++; ph_lanes[PH_LEAF] += n_sel; }
      if (s >= 0) {
        cur = (int32_t)PW(W_CUR, s);
        const uint32_t spc = PW(W_SPC, s);
        sp = (int)(spc & 0xffu);
        pend_cnt = spc >> 8;
        pend_slot = PW(W_PSLOT, s);
        lt = PF(W_LT, s);
        lprim = (int32_t)PW(W_LPRIM, s);
        lo = PLD3(W_LOX, s);
        ld = PLD3(W_LDX, s);
        const uint32_t mw = PW(W_M, s);
        bool accepted = false;
        V3 n3 = mk(0, 0, 0), nbest = n3;  // normal of the primitive under test / of the closest accepted hit
        // One primitive per round for a leaf of a hierarchy; the 2 - 4 primitives of a hierarchy-less mesh (the Cornell
        // quads) are tested back to back: a round trip through the vote per triangle costs more than the idle lanes.
        uint32_t leaf_slot = pend_slot >> 2, leaf_left = (pend_slot & 3u) + 1u;  // the postponed leaf: 1 - 4 consecutive sorted slots
        // (one loop per primitive kind: a shared loop makes the compiler carry the running addresses of both kinds)
        if (F_SPHERES && (mw & WM_SPHERES)) {
#pragma unroll 1
        do {
        const uint32_t slot = leaf_slot;
        leaf_slot++;
        leaf_left--;
          // EXTENSION (the reference kernel has no sphere primitive): semantics of oracle/rr_oracle.c ray_sphere
          if (COUNT) c_sph++;
          const float4 cr = __ldg(p.sph_geom + slot);
          const int32_t prim = (int32_t)__ldg(p.sph_order + slot);
          const V3 c = xyz(cr);
          const float r = cr.w;
          const V3 oc = lo - c;
          const float b = dot(oc, ld);
          const float cc = dot(oc, oc) - r * r;
          const float disc = b * b - cc;
          if (disc >= 0.0f) {
            const float sq = sqrtf(disc);
            float t = -b - sq;
            bool back = false;
            if (t <= RR_EPSILON) { t = -b + sq; back = true; }
            if (t > RR_EPSILON && (t < lt || (t == lt && lprim != NO_PRIM && prim < lprim))) {
              const int32_t mtype = __ldg(&p.materials[p.n_meshes + prim].type);
              const bool cull = (mtype != RR_MATERIAL_GLASSY && mtype != RR_MATERIAL_INVISIBLE && mtype != RR_MATERIAL_ONESIDED);
              if (!(back && cull)) {
                const V3 hp = lo + ld * t;
                n3 = (hp - c) / r;
                if (back) n3 = -n3;
                lt = t; lprim = prim; lback = back;
                nbest = n3;
                accepted = true;
              }
            }
          }
        } while (leaf_left > 0);
        } else {
#pragma unroll 1
        do {
        const uint32_t slot = leaf_slot;
        leaf_slot++;
        leaf_left--;
          // src/Trace.cl:276-317 with the distance test hoisted before the normal (same accept set) and a
          // total order (t, prim) instead of first-found-wins
          if (COUNT) c_tri++;
          const float4* gp = p.tri_geom + 3 * (size_t)slot;
#if RR_STREAM_GEOM
          const float4 g0 = __ldcs(gp), g1 = __ldcs(gp + 1), g2 = __ldcs(gp + 2);
#else
          const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
#endif
          const V3 A = xyz(g0), edge1 = xyz(g1), edge2 = xyz(g2);
          const V3 h = cross(ld, edge2);
          const float a = dot(edge1, h);
          if (!(fabsf(a) < RR_EPSILON)) {
            const float f = 1.0f / a;
            const V3 sv = lo - A;
            const float u = f * dot(sv, h);
            if (!(u < 0.0f || u > 1.0f)) {
              const V3 q = cross(sv, edge1);
              const float v = f * dot(ld, q);
              if (!(v < 0.0f || u + v > 1.0f)) {
                const float t = f * dot(edge2, q);
                const int32_t prim = (int32_t)__float_as_uint(g0.w);
                if (t > RR_EPSILON && (t < lt || (t == lt && lprim != NO_PRIM && prim < lprim))) {
                  const float4* np = p.tri_nrm + 3 * (size_t)slot;
#if RR_STREAM_NRM
                  const float4 n0 = __ldcs(np), n1 = __ldcs(np + 1), n2 = __ldcs(np + 2);
#else
                  const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
#endif
                  n3 = normalize(xyz(n0) * (1.0f - u - v) + xyz(n1) * u + xyz(n2) * v);
                  bool back = false;
                  bool ok = true;
                  if (dot(ld, n3) > RR_EPSILON) {
                    if (mw & WM_CULL) ok = false;
                    back = true;
                    n3 = -n3;
                  }
                  if (ok) {
                    lt = t; lprim = prim; lback = back;
                    nbest = n3;
                    accepted = true;
                  }
                }
              }
            }
          }
        } while (leaf_left > 0);
        }
        pend_cnt = 0;
        if (accepted) {
          PSF(W_LT, s, lt);
          PW(W_LPRIM, s) = (uint32_t)lprim;
          PW(W_M, s) = (mw & ~WM_BACK) | (lback ? WM_BACK : 0u);
          CST3(C_LNX, s, nbest);
        }
        if (pend_cnt == 0 && ref_is_leaf(cur)) {  // the leaf this slot was waiting on becomes the postponed one
          pend_slot = ref_leaf_word(cur);
          pend_cnt = 1;
          cur = REF_POP;
          PW(W_CUR, s) = (uint32_t)cur;
        }
        PW(W_SPC, s) = (uint32_t)sp | (pend_cnt << 8);
        PW(W_PSLOT, s) = pend_slot;
        PW(W_KEY, s) = trav_key();
      }
    } else if (phase == PH_SETUP) {
      // ================= finish the current mesh, enter the next candidate (src/Trace.cl:444-482) =================
      if (COUNT) { ph_runs[PH_SETUP]++; ph_lanes[PH_SETUP] += n_sel; }
      if (s >= 0) {
        origin = PLD3(W_OX, s);
        dir = PLD3(W_DX, s);
        load_mesh_word();
        const V3 winv = mk(rcp_approx(dir.x), rcp_approx(dir.y), rcp_approx(dir.z));
        const V3 wnoi = mk(-(origin.x * winv.x), -(origin.y * winv.y), -(origin.z * winv.z));
        load_ray_state();
        finish_mesh();  // nothing to finish for a new ray (no local hit)
        RaySlack<SLACK> wek;
        make_slack(wek, origin, winv);
        const uint32_t key = enter_next_mesh(winv, wnoi, wek);
        store_ray_state(key);
      }
    } else if (phase == PH_SHADE) {
      // ================= one bounce of Trace() (src/Trace.cl:497-591) and the sample loop (:639-642) =================
      if (COUNT) { ph_runs[PH_SHADE]++; ph_lanes[PH_SHADE] += n_sel; }
      bool pixel_done = false;
      if (s >= 0) {
        origin = PLD3(W_OX, s);
        dir = PLD3(W_DX, s);
        best_dst = PF(W_BDST, s);
        const uint32_t bm = PW(W_BMAT, s);
        best_mat = (int32_t)(bm & 0x7fffffffu);
        best_back = (bm >> 31) != 0u;
        if (PRIMARY) { best_mesh = min(best_mat, p.n_meshes); best_prim = (int32_t)CW(C_BPRIM, s); }
        pix = (int32_t)PW(W_PIX, s);
        if (PRIMARY) {
          if (p.hit_mesh) p.hit_mesh[pix] = best_dst < INFINITY ? best_mesh : -1;
          if (p.hit_prim) p.hit_prim[pix] = best_dst < INFINITY ? best_prim : -1;
          if (p.hit_dst) p.hit_dst[pix] = best_dst < INFINITY ? best_dst : 0.0f;
          pixel_done = true;
        } else {
          rng = CW(C_RNG, s);
          sample = CW(C_SAMPLE, s);
          const uint32_t bw = CW(C_BOUNCE, s);
          bounce = bw & RR_MAX_BOUNCES;
          passes = bw >> 23;
          SceneHit hit;
          hit.did = best_dst < INFINITY;
          hit.dst = best_dst;
          hit.point = CLD3(C_BPX, s);
          hit.normal = CLD3(C_BNX, s);
          hit.back = best_back;
          hit.material = best_mat;
          V3 throughput = CLD3(C_THR, s), incoming = CLD3(C_INC, s);
          bool alive = shade<F_MATERIALS>(p, hit, origin, dir, throughput, incoming, bounce, passes, rng);
          alive = alive && bounce < p.max_bounces;
          if (alive) {
            CST3(C_THR, s, throughput);
            CST3(C_INC, s, incoming);
          } else {  // path finished: src/Trace.cl:639-642
            const V3 accum = CLD3(C_ACC, s) + incoming;
            sample++;
            if (sample >= p.spp) {
              const V3 c = accum / (float)p.spp;
              reinterpret_cast<uint32_t*>(p.frame)[pix] = tonemap_rgba(c);
              if (COUNT && p.cost) p.cost[pix] = CW(C_BPRIM, s);
              if (p.radiance) {
                p.radiance[3 * (size_t)pix] = c.x;
                p.radiance[3 * (size_t)pix + 1] = c.y;
                p.radiance[3 * (size_t)pix + 2] = c.z;
              }
              pixel_done = true;
            } else {
              CST3(C_ACC, s, accum);
              bounce = 0; passes = 0;  // next sample of this pixel: src/Trace.cl:488-491
              origin = cam_pos;
              dir = CLD3(C_PD, s);
              CSF(C_THR, s, 1.0f); CSF(C_THR1, s, 1.0f); CSF(C_THR2, s, 1.0f);
              CSF(C_INC, s, 0.0f); CSF(C_INC1, s, 0.0f); CSF(C_INC2, s, 0.0f);
            }
          }
        }
        if (pixel_done) {
          PW(W_KEY, s) = 0u;
          PW(W_PIX, s) = (uint32_t)PIX_NEED;
        } else {  // next segment: the setup phase collects its candidate meshes
          CW(C_RNG, s) = rng;
          CW(C_SAMPLE, s) = sample;
          CW(C_BOUNCE, s) = bounce | (passes << 23);  // bounce <= max_bounces <= RR_MAX_BOUNCES, passes <= 257
          store_new_ray();
        }
      }
      n_need += __popc(__ballot_sync(full, pixel_done));
    } else {
      // ================= hand a pixel to every slot that needs one =================
      if (COUNT) { ph_runs[PH_PIXEL]++; ph_lanes[PH_PIXEL] += n_sel; }
      bool need = s >= 0;
      bool queue_empty = false;
#if RR_PIXEL_QUEUE
      // The counter hands out PIXELS, numbered tile by tile (item i = pixel i % tile_pixels of the launch's tile
      // i / tile_pixels), and the warp takes exactly as many as it has free slots: ONE atomic per run of this phase, as
      // with whole tiles (a run hands out ~27 pixels), but no warp sits on the unstarted rest of a tile when the queue
      // runs dry -- those pixels would start up to a third of a pixel lifetime late and stretch the drain of the frame
      // (DESIGN.md section 6).  Items of a ragged border tile that fall outside the image are skipped.
      for (;;) {
        const unsigned mb = __ballot_sync(full, need);
        if (mb == 0u) break;
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd_system(p.queue, (unsigned long long)__popc(mb));
        t = __shfl_sync(full, t, 0);
        if ((uint32_t)(t >> RR_QUEUE_EPOCH_SHIFT) != p.queue_epoch) {  // see pop_tile
          if (lane == 0) atomicAdd(&p.counters->queue_errors, 1ull);
          queue_empty = true;
          break;
        }
        t &= (1ull << RR_QUEUE_EPOCH_SHIFT) - 1ull;
        if (t >= (unsigned long long)p.queue_items) { queue_empty = true; break; }
        const unsigned long long it = t + __popc(mb & lanes_below);
        bool first_of_tile = false;
        if (need && it < (unsigned long long)p.queue_items) {
          const uint32_t item = (uint32_t)it, seq0 = item / p.tile_pixels, k = item - seq0 * p.tile_pixels;
#if RR_TILE_ORDER == 1    // last tile first
          const uint32_t seq = p.queue_tiles - 1u - seq0;
#elif RR_TILE_ORDER == 2  // scattered: a multiplicative permutation of the launch's tiles (tile_mul is coprime to their number)
          const uint32_t seq = (uint32_t)(((unsigned long long)seq0 * p.tile_mul) % p.queue_tiles);
#else
          const uint32_t seq = seq0;
#endif
          // (a caller-supplied order of the launch's tiles, rr_set_tile_order: measurement of cost-ordered queues)
          const uint32_t tile = p.tile_order ? __ldg(p.tile_order + seq) : p.tile_begin + seq * p.tile_stride;
          const uint32_t x = (tile % p.tiles_x) * p.tile_w + k % p.tile_w, y = (tile / p.tiles_x) * p.tile_h + k / p.tile_w;
          first_of_tile = k == 0u;
          if (x < p.width && y < p.height) {
            pix = (int32_t)(y * p.width + x);
            if (!PRIMARY && p.max_bounces == 0) {  // no segment is ever traced: the pixel is black
              reinterpret_cast<uint32_t*>(p.frame)[pix] = tonemap_rgba(mk(0, 0, 0));
              if (p.radiance) { p.radiance[3 * (size_t)pix] = 0.0f; p.radiance[3 * (size_t)pix + 1] = 0.0f; p.radiance[3 * (size_t)pix + 2] = 0.0f; }
            } else {
              PW(W_PIX, s) = (uint32_t)pix;
              if (COUNT && !PRIMARY) CW(C_BPRIM, s) = 0u;
              CW(C_RNG, s) = make_seed((uint32_t)pix, p.frame_index, 0u);  // src/Trace.cl:631-632
              const V3 pd = primary_dir(p.cam, x, y, p.width, p.height);   // once per pixel, :634-636
              CST3(C_PD, s, pd);
              CSF(C_ACC, s, 0.0f); CSF(C_ACC1, s, 0.0f); CSF(C_ACC2, s, 0.0f);
              CW(C_SAMPLE, s) = 0u;
              CW(C_BOUNCE, s) = 0u;
              CSF(C_THR, s, 1.0f); CSF(C_THR1, s, 1.0f); CSF(C_THR2, s, 1.0f);
              CSF(C_INC, s, 0.0f); CSF(C_INC1, s, 0.0f); CSF(C_INC2, s, 0.0f);
              origin = cam_pos;
              dir = pd;
              store_new_ray();
              need = false;
            }
          }
        }
        n_tiles += __popc(__ballot_sync(full, first_of_tile));  // a tile is counted by the warp that takes its first pixel
      }
      more_pixels = !queue_empty;
      if (COUNT && queue_empty && t_empty == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_empty));
#else
      // the tile state lives in shared memory between two runs of this (rare) phase, not in six registers
      uint32_t tile_x0 = tstate[0], tile_y0 = tstate[1], tile_w = tstate[2], tile_next = tstate[3], tile_pixels = tstate[4];
      while (__any_sync(full, need)) {
        if (tile_next >= tile_pixels) {
          if (queue_empty) break;
          uint32_t tile;
          if (!pop_tile(p, tile)) { queue_empty = true; break; }
          n_tiles++;
          tile_x0 = (tile % p.tiles_x) * p.tile_w;
          tile_y0 = (tile / p.tiles_x) * p.tile_h;
          tile_w = min(p.tile_w, p.width - tile_x0);
          tile_pixels = tile_w * min(p.tile_h, p.height - tile_y0);
          tile_next = 0;
        }
        const unsigned mb = __ballot_sync(full, need);
        const unsigned rank = __popc(mb & lanes_below);
        const uint32_t k = tile_next + rank;
        if (need && k < tile_pixels) {
          const uint32_t x = tile_x0 + k % tile_w, y = tile_y0 + k / tile_w;
          pix = (int32_t)(y * p.width + x);
          if (!PRIMARY && p.max_bounces == 0) {  // no segment is ever traced: the pixel is black
            reinterpret_cast<uint32_t*>(p.frame)[pix] = tonemap_rgba(mk(0, 0, 0));
            if (p.radiance) { p.radiance[3 * (size_t)pix] = 0.0f; p.radiance[3 * (size_t)pix + 1] = 0.0f; p.radiance[3 * (size_t)pix + 2] = 0.0f; }
          } else {
            PW(W_PIX, s) = (uint32_t)pix;
            if (COUNT && !PRIMARY) CW(C_BPRIM, s) = 0u;
            CW(C_RNG, s) = make_seed((uint32_t)pix, p.frame_index, 0u);  // src/Trace.cl:631-632
            const V3 pd = primary_dir(p.cam, x, y, p.width, p.height);   // once per pixel, :634-636
            CST3(C_PD, s, pd);
            CSF(C_ACC, s, 0.0f); CSF(C_ACC1, s, 0.0f); CSF(C_ACC2, s, 0.0f);
            CW(C_SAMPLE, s) = 0u;
            CW(C_BOUNCE, s) = 0u;
            CSF(C_THR, s, 1.0f); CSF(C_THR1, s, 1.0f); CSF(C_THR2, s, 1.0f);
            CSF(C_INC, s, 0.0f); CSF(C_INC1, s, 0.0f); CSF(C_INC2, s, 0.0f);
            origin = cam_pos;
            dir = pd;
            store_new_ray();
            need = false;
          }
        }
        tile_next += __popc(mb);
      }
      more_pixels = !(queue_empty && tile_next >= tile_pixels);
      if (COUNT && queue_empty && t_empty == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_empty));
      __syncwarp();
      if (lane == 0) { tstate[0] = tile_x0; tstate[1] = tile_y0; tstate[2] = tile_w; tstate[3] = tile_next; tstate[4] = tile_pixels; }
#endif
      const bool got = s >= 0 && !need;
      n_need -= __popc(__ballot_sync(full, got));
      if (need) PW(W_PIX, s) = (uint32_t)PIX_IDLE;  // the queue is empty
    }
  }
#undef CW
#undef CF
#undef CSF
#undef CLD3
#undef CST3
#undef PW
#undef PF
#undef PSF
#undef PLD3
#undef PST3
  // counters: one atomic per warp
  unsigned long long r = n_rays;  // summed over the warp in 64 bits
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(full, r, off);
  __syncwarp();
  if (lane == 0) {
    atomicAdd(&p.counters->rays, r);
    atomicAdd(&p.counters->tiles, (unsigned long long)n_tiles);
    if (tstate[5]) atomicAdd(&p.counters->stack_overflows, 1ull);  // warps that dropped pushes
  }
  if (COUNT) {
    unsigned long long b = c_box, t = c_tri, sq = c_sph;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      b += __shfl_xor_sync(full, b, off);
      t += __shfl_xor_sync(full, t, off);
      sq += __shfl_xor_sync(full, sq, off);
    }
    if (lane == 0) {
      atomicAdd(&p.counters->box_tests, b);
      atomicAdd(&p.counters->tri_tests, t);
      atomicAdd(&p.counters->sphere_tests, sq);
      for (int k = 0; k < 5; ++k) {
        atomicAdd(&p.counters->phase_runs[k], (unsigned long long)ph_runs[k]);
        atomicAdd(&p.counters->phase_lanes[k], (unsigned long long)ph_lanes[k]);
      }
      unsigned long long t_exit;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_exit));
      const unsigned long long tail = t_empty ? t_exit - t_empty : 0ull;
      atomicAdd(&p.counters->tail_ns_sum, tail);
      atomicMax(&p.counters->tail_ns_max, tail);
      atomicAdd(&p.counters->tail_warps, 1ull);
    }
  }
}

void default_tuning(Tuning& t) {
  for (int k = 0; k < 5; ++k) t.weight[k] = 4;
  t.trav_keep = RR_TRAV_KEEP_DEFAULT;
  t.speculate = 1;
  t.ctas_per_sm = 0;
}

constexpr size_t RENDER_SMEM = (size_t)WARPS * (NW * POOL + 32 + 8) * sizeof(uint32_t) + (size_t)RR_TOP_STAGE * 128;

template <class K>
static cudaError_t launch_persistent(K kernel, const RenderParams& p, int sm_count, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RENDER_SMEM);
  if (e != cudaSuccess) return e;
  int n = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, NT, RENDER_SMEM);
  if (e != cudaSuccess) return e;
  if (n < 1) n = 1;
  if (p.tune.ctas_per_sm && (int)p.tune.ctas_per_sm < n) n = (int)p.tune.ctas_per_sm;
  int grid = sm_count * n;
  if (grid > (int)p.stack_warps / WARPS) grid = (int)p.stack_warps / WARPS;  // scratch stacks were sized for this many warps
  // persistent: as many CTAs as are resident at once, each warp loops until the tile queue is empty
  kernel<<<grid, NT, RENDER_SMEM, s>>>(p);
  return cudaGetLastError();
}

// The production kernel exists once per feature set (default knobs, no per-ray slack) and, with everything compiled in,
// with the per-ray slack and / or run-time scheduler knobs; the instrumented and the primary-hit kernels only with everything.
static bool default_knobs(const Tuning& t) {
  for (int k = 1; k < 5; ++k)
    if (t.weight[k] != t.weight[0]) return false;
  return t.trav_keep == RR_TRAV_KEEP_DEFAULT && (t.speculate & 1u);
}
template <int FEAT>
static cudaError_t launch_lean(const RenderParams& p, int sm_count, cudaStream_t s) {
  return launch_persistent(k_render<false, false, false, FEAT, false>, p, sm_count, s);
}
cudaError_t launch_render(const RenderParams& p, bool count_tests, bool slack, int feat, int sm_count, cudaStream_t s) {
  constexpr int ALL = RR_FEAT_ALL;
  if (count_tests) return slack ? launch_persistent(k_render<true, false, true, ALL, true>, p, sm_count, s) : launch_persistent(k_render<true, false, false, ALL, true>, p, sm_count, s);
  if (!default_knobs(p.tune)) return slack ? launch_persistent(k_render<false, false, true, ALL, true>, p, sm_count, s) : launch_persistent(k_render<false, false, false, ALL, true>, p, sm_count, s);
  if (slack) return launch_persistent(k_render<false, false, true, ALL, false>, p, sm_count, s);
  switch (feat & ALL) {
    case 0: return launch_lean<0>(p, sm_count, s);
    case 1: return launch_lean<1>(p, sm_count, s);
    case 2: return launch_lean<2>(p, sm_count, s);
    case 3: return launch_lean<3>(p, sm_count, s);
    case 4: return launch_lean<4>(p, sm_count, s);
    case 5: return launch_lean<5>(p, sm_count, s);
    case 6: return launch_lean<6>(p, sm_count, s);
    default: return launch_lean<7>(p, sm_count, s);
  }
}

// Primary-ray closest hit per pixel (MakeRay + CalculateRayCollisionWithTriangle): the same kernel, stopped
// at the first shade phase.
cudaError_t launch_primary(const RenderParams& p, bool slack, int sm_count, cudaStream_t s) {
  if (!p.width || !p.height) return cudaSuccess;
  return slack ? launch_persistent(k_render<false, true, true, RR_FEAT_ALL, true>, p, sm_count, s)
               : launch_persistent(k_render<false, true, false, RR_FEAT_ALL, true>, p, sm_count, s);
}

size_t render_stack_bytes_per_warp(uint32_t stack_entries) { return (size_t)stack_entries * POOL * sizeof(uint2); }
size_t render_cold_bytes_per_warp() { return (size_t)NC * POOL * sizeof(uint32_t); }
int render_max_warps_per_sm() { return (RR_MIN_CTAS + 1) * WARPS; }  // the launch clamps its grid to this

// ---- probes for the bit-level parity tests (rr_probe_math / rr_probe_rng, tests/test_gpu_parity.py) ---------
__global__ void k_math_probe(int fn, const float* x, const float* y, float* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r;
  switch (fn) {
    case 0: r = cos_c(x[i]); break;
    case 1: r = sin_c(x[i]); break;
    case 2: r = log_c(x[i]); break;
    case 3: r = exp2_c(x[i]); break;
    case 4: r = powr_c(x[i], y[i]); break;
    default: r = tan_c(x[i]); break;
  }
  out[i] = r;
}
cudaError_t launch_math_probe(int fn, const float* x, const float* y, float* out, uint64_t n, cudaStream_t s) {
  if (!n) return cudaSuccess;
  k_math_probe<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(fn, x, y, out, n);
  return cudaGetLastError();
}

// out_u32: seed, then the u32 state after each of 4 RandomValue + 2 rand01 calls; out_f32: the 6 floats + a RandomDirection
__global__ void k_rng_probe(uint32_t pixel, int32_t frame, uint32_t* out_u32, float* out_f32) {
  uint32_t s = make_seed(pixel, frame, 0u);
  out_u32[0] = s;
  for (int k = 0; k < 4; ++k) { out_f32[k] = random_value(s); out_u32[1 + k] = s; }
  for (int k = 0; k < 2; ++k) { out_f32[4 + k] = rand01(s); out_u32[5 + k] = s; }
  V3 d = random_direction(s);
  out_f32[6] = d.x; out_f32[7] = d.y; out_f32[8] = d.z;
  out_u32[7] = s;
}
cudaError_t launch_rng_probe(uint32_t pixel, int32_t frame, uint32_t* out_u32, float* out_f32, cudaStream_t s) {
  k_rng_probe<<<1, 1, 0, s>>>(pixel, frame, out_u32, out_f32);
  return cudaGetLastError();
}

}  // namespace rr
