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
//   * persistent warps pop 8x4-pixel tiles from one 64-bit atomic counter (the
//     reference's mutex-guarded std::queue, src/image.hpp:286-314); a lane that
//     finishes its pixel takes the next pixel of the warp's tile;
//   * per pixel the spp samples run serially in the lane because the RNG state
//     is carried across samples (src/Trace.cl:632,639-642);
//   * every lane is in one of five phases (pixel, shade, mesh setup, node step,
//     leaf test); each round the warp votes and runs the phase with the most
//     ready lanes, so the hot node-step loop executes with most lanes active
//     instead of each lane walking its own ray while 31 others wait;
//   * a node step is one 64-byte fetch (4 x LDG.128: both child boxes and both
//     child references) and two slab tests; a lane may postpone one leaf and
//     keep walking (speculative traversal);
//   * the per-path colour state lives in shared memory so that the traversal
//     fits 64 registers and 32 warps stay resident per SM.
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
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// fast_normalize / normalize of the numerics contract
__device__ __forceinline__ V3 normalize(V3 a) {
  float inv = 1.0f / sqrtf(dot(a, a));
  return a * inv;
}
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
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
// src/Trace.cl:179-187
__device__ __forceinline__ float random_normal(uint32_t& state) {
  float u1 = random_value(state);
  float u2 = random_value(state);
  u1 = fmaxf(u1, RR_EPSILON);
  float r = sqrtf(-2.0f * log_c(u1));
  float theta = RR_TAU * u2;
  return r * cos_c(theta);
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
__device__ __forceinline__ bool box_cull(float lox, float loy, float loz, float hix, float hiy, float hiz, const V3& inv,
                                         const V3& noi, float tbest, float& tn) {
  const float t0x = __fmaf_rn(lox, inv.x, noi.x), t1x = __fmaf_rn(hix, inv.x, noi.x);
  const float t0y = __fmaf_rn(loy, inv.y, noi.y), t1y = __fmaf_rn(hiy, inv.y, noi.y);
  const float t0z = __fmaf_rn(loz, inv.z, noi.z), t1z = __fmaf_rn(hiz, inv.z, noi.z);
  tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
  const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  return tf >= fmaxf(tn, 0.0f) && tn <= tbest;
}

constexpr int32_t NO_PRIM = 0x7fffffff;

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
__device__ __forceinline__ bool shade(const RenderParams& p, const SceneHit& hit, V3& origin, V3& dir, V3& throughput,
                                      V3& incoming, uint32_t& bounce, uint32_t& passes, uint32_t& rng) {
  if (!hit.did) return false;
  const DMaterial* M = p.materials + hit.material;
  const int32_t type = __ldg(&M->type);
  if (type == RR_MATERIAL_INVISIBLE) {
    // `continue` without counting a bounce (:502-506).  Guard: when hit.point + dir*1e-6 rounds back
    // to hit.point the reference loops forever; the path is ended after RR_MAX_INVISIBLE_PASSES.
    if (++passes > RR_MAX_INVISIBLE_PASSES) return false;
    origin = hit.point + dir * RR_EPSILON;
    return true;
  }
  V3 color = ld3(M->color);
  const V3 emissionColor = ld3(M->emissionColor);
  float emissionStrength = __ldg(&M->emissionStrength);
  const float specProb = __ldg(&M->specularProbability);
  const float reflectiveness = __ldg(&M->reflectiveness);
  if (type == RR_MATERIAL_CHECKER) {  // :509-533
    const float size = emissionStrength;
    const int xi = (int)floorf(hit.point.x / size);
    const int zi = (int)floorf(hit.point.z / size);
    const bool isEven = (((uint32_t)xi + (uint32_t)zi) & 1u) == 0u;
    color = isEven ? color : emissionColor;
    emissionStrength = 0.0f;
  }
  if (type == RR_MATERIAL_CHECKER || type == RR_MATERIAL_SOLID) {  // :525-532, :559-567
    const bool isSpec = specProb >= random_value(rng);
    const V3 diffuseDir = normalize(hit.normal + random_direction(rng));
    const V3 specularDir = reflect3(dir, hit.normal);
    dir = normalize(lerp3(diffuseDir, specularDir, reflectiveness * (isSpec ? 1.0f : 0.0f)));
  } else if (type == RR_MATERIAL_GLASSY) {  // :534-558
    const float ior = __ldg(&M->ior);
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
__device__ __forceinline__ V3 primary_dir(const DCamera& cam, uint32_t x, uint32_t y, uint32_t W, uint32_t H) {
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
__device__ __forceinline__ uint32_t tonemap_rgba(V3 c) {
  const float r = powr_c(fminf(fmaxf(c.x, 0.0f), 1.0f), 1.0f / 2.2f);
  const float g = powr_c(fminf(fmaxf(c.y, 0.0f), 1.0f), 1.0f / 2.2f);
  const float b = powr_c(fminf(fmaxf(c.z, 0.0f), 1.0f), 1.0f / 2.2f);
  const uint32_t R = (uint32_t)(unsigned char)(r * 255.0f), G = (uint32_t)(unsigned char)(g * 255.0f),
                 B = (uint32_t)(unsigned char)(b * 255.0f);
  return R | (G << 8) | (B << 16) | (255u << 24);
}


// ---- tile queue ----------------------------------------------------------------
// Tiles are numbered row-major.  With a shared counter (possibly in a peer
// GPU's memory, hence the system-scope atomic) every warp of every GPU pops the
// next tile; with a static partition rank r renders tiles r, r+world, ...
__device__ __forceinline__ bool pop_tile(const RenderParams& p, uint32_t& tile) {
  unsigned long long t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd_system(p.queue, 1ull);
  t = __shfl_sync(0xffffffffu, t, 0);
  t = (unsigned long long)p.tile_begin + t * p.tile_stride;
  tile = (uint32_t)t;
  return t < (unsigned long long)p.tiles_x * p.tiles_y;
}

constexpr int NT = 128;  // threads per CTA (warps are independent: no block-level barrier in the loop)
enum { ST_IDLE = 0, ST_PIXEL = 1, ST_SHADE = 2, ST_SETUP = 3, ST_TRAV = 4 };
enum { PH_PIXEL = 0, PH_SHADE = 1, PH_SETUP = 2, PH_TRAV = 3, PH_LEAF = 4 };
// per-thread words kept in shared memory (word k of thread t at smem[k * NT + t]: conflict-free)
enum { S_THR = 0, S_INC = 3, S_ACC = 6, S_BP = 9, S_BN = 12, S_LN = 15, S_WINV = 18, S_PD = 21, S_WORDS = 24 };
// Node references.  In the packed nodes a leaf is -(slot + 2), so that "inner node or pop" is cur >= -1.
constexpr int32_t REF_POP = -1;                   // take the next entry of the stack
constexpr int32_t REF_END = (int32_t)0x80000000;  // traversal of this mesh finished
__device__ __forceinline__ bool ref_is_leaf(int32_t r) { return r < REF_POP && r != REF_END; }
__device__ __forceinline__ uint32_t ref_slot(int32_t r) { return (uint32_t)(-r) - 2u; }

template <bool COUNT, bool PRIMARY>
__global__ void __launch_bounds__(NT, 8) k_render(const RenderParams p) {
  __shared__ float smem[S_WORDS * NT];
#define SM(k) smem[(k) * NT + threadIdx.x]
#define SM_ST3(k, v) do { SM(k) = (v).x; SM((k) + 1) = (v).y; SM((k) + 2) = (v).z; } while (0)
#define SM_LD3(k) mk(SM(k), SM((k) + 1), SM((k) + 2))
  const unsigned lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const V3 cam_pos = mk(p.cam.pos[0], p.cam.pos[1], p.cam.pos[2]);
  // traversal stack (local memory; interleaved per lane by the hardware)
  int32_t stackN[RR_STACK];
  float stackD[RR_STACK];
  // warp-uniform tile state
  bool queue_empty = false;
  uint32_t tile_x0 = 0, tile_y0 = 0, tile_w = 1, tile_next = 0, tile_pixels = 0;
  // lane state
  int state = ST_PIXEL;
  int32_t pix = -1;
  uint32_t rng = 0, sample = 0, bounce = 0, passes = 0;
  V3 origin = cam_pos, dir = mk(0, 0, 1);
  float best_dst = INFINITY;
  int32_t best_mat = 0, best_mesh = 0x7fffffff, best_prim = -1;
  bool best_back = false;
  uint32_t cand = 0;     // candidate meshes of the current 32-mesh chunk (bit k = mesh cand_base + k) not yet visited
  int32_t cand_base = 0;
  int m = 0;
  uint32_t mflags = 0;
  V3 lo = origin, ld = dir, linv = dir, lnoi = dir;
  float lt = INFINITY;
  int32_t lprim = NO_PRIM;
  bool lback = false;
  int32_t cur = REF_END;
  int sp = 0;
  uint32_t pend_slot = 0, pend_cnt = 0;
  // statistics
  unsigned long long n_rays = 0, n_tiles = 0;
  unsigned c_box = 0, c_tri = 0, c_sph = 0;
  unsigned ph_runs[5] = {0, 0, 0, 0, 0}, ph_lanes[5] = {0, 0, 0, 0, 0};

  const uint32_t wP = p.tune.weight[PH_PIXEL], wH = p.tune.weight[PH_SHADE], wS = p.tune.weight[PH_SETUP],
                 wT = p.tune.weight[PH_TRAV], wL = p.tune.weight[PH_LEAF];
  const uint32_t trav_keep = p.tune.trav_keep;
  const bool speculate = p.tune.speculate != 0;

  // World-box tests of the meshes [base, base + 32): bit k set = the ray enters mesh base + k's box before `tmax`.
  // Every lane walks the whole chunk, so the loop is convergent.
  auto scan_meshes = [&](int32_t base, const V3& winv, const V3& wnoi, float tmax) -> uint32_t {
    uint32_t mask = 0;
    const int32_t end = min(base + 32, p.last_mesh + 1);
    for (int32_t k = base; k < end; ++k) {
      const DMesh* M = p.meshes + k;
      const float4 wlo = __ldg(&M->wmin), whi = __ldg(&M->wmax);
      float tn;
      const bool hit = box_cull(wlo.x, wlo.y, wlo.z, whi.x, whi.y, whi.z, winv, wnoi, tmax, tn);
      if (hit && !(__float_as_uint(wlo.w) & RR_MF_SKIP)) mask |= 1u << (k - base);
    }
    if (COUNT) c_box += (unsigned)(end - base);
    return mask;
  };
  // A new ray starts: reset the closest hit (src/Trace.cl:437-444) and collect the candidate meshes.
  auto begin_ray = [&]() {
    best_dst = INFINITY; best_mat = 0; best_mesh = 0x7fffffff; best_prim = -1; best_back = false;
    lprim = NO_PRIM;
    const V3 winv = mk(rcp_approx(dir.x), rcp_approx(dir.y), rcp_approx(dir.z));
    SM_ST3(S_WINV, winv);
    const V3 wnoi = mk(-(origin.x * winv.x), -(origin.y * winv.y), -(origin.z * winv.z));
    cand_base = 0;
    cand = scan_meshes(0, winv, wnoi, INFINITY);
    n_rays++;
    state = ST_SETUP;
  };
  auto begin_path = [&]() {  // src/Trace.cl:488-491
    bounce = 0; passes = 0;
    origin = cam_pos;
    dir = SM_LD3(S_PD);
    SM(S_THR) = 1.0f; SM(S_THR + 1) = 1.0f; SM(S_THR + 2) = 1.0f;
    SM(S_INC) = 0.0f; SM(S_INC + 1) = 0.0f; SM(S_INC + 2) = 0.0f;
  };
  // The mesh just traversed has a closest hit (local space): LocalToWorldHit and the keep-min of
  // src/Trace.cl:465-481.  Meshes are visited in our own order, so equal distances are resolved by the
  // original mesh index, which is what the reference's in-order strict `<` does.
  auto finish_mesh = [&]() {
    if (lprim == NO_PRIM) return;
    const DMesh* M = p.meshes + m;
    const int32_t mesh_index = __float_as_int(__ldg(&M->wmax.w));
    if (mflags & RR_MF_SPHERES) {
      const int32_t mat = mesh_index + lprim;
      const int32_t type = __ldg(&p.materials[mat].type);
      if (!(type == RR_MATERIAL_ONESIDED && lback) && (lt < best_dst || (lt == best_dst && mesh_index < best_mesh))) {
        best_dst = lt; best_mat = mat; best_back = lback; best_mesh = mesh_index; best_prim = lprim;
        const V3 wp = origin + dir * lt;
        SM_ST3(S_BP, wp);
        SM(S_BN) = SM(S_LN); SM(S_BN + 1) = SM(S_LN + 1); SM(S_BN + 2) = SM(S_LN + 2);
      }
    } else {
      const int32_t type = (int32_t)(mflags >> RR_MF_TYPE_SHIFT);
      if (!(type == RR_MATERIAL_ONESIDED && lback)) {
        // LocalToWorldHit, src/Trace.cl:139-156
        const float4 r0 = __ldg(&M->r0), r1 = __ldg(&M->r1), r2 = __ldg(&M->r2);
        const V3 pos = mk(__ldg(&M->ri0.w), __ldg(&M->ri1.w), __ldg(&M->ri2.w));
        const V3 lp = (lo + ld * lt) * r0.w;
        const V3 wp = mk(dot(xyz(r0), lp), dot(xyz(r1), lp), dot(xyz(r2), lp)) + pos;
        const V3 ln = SM_LD3(S_LN);
        const V3 wn = normalize(mk(dot(xyz(r0), ln), dot(xyz(r1), ln), dot(xyz(r2), ln)));
        const float wd = length(wp - origin);
        if (wd < best_dst || (wd == best_dst && mesh_index < best_mesh)) {
          best_dst = wd; best_mat = mesh_index; best_back = lback; best_mesh = mesh_index; best_prim = lprim;
          SM_ST3(S_BP, wp);
          SM_ST3(S_BN, wn);
        }
      }
    }
    lprim = NO_PRIM;
  };
  // Enters the next candidate mesh (src/Trace.cl:444-463): world box against the closest hit so far,
  // WorldToLocalRay, root box; ST_TRAV when a traversal starts, ST_SHADE when no candidate is left.
  auto enter_next_mesh = [&](const V3& winv, const V3& wnoi) {
    for (;;) {
      if (cand == 0) {
        if (cand_base + 32 > p.last_mesh) { state = ST_SHADE; return; }
        cand_base += 32;
        cand = scan_meshes(cand_base, winv, wnoi, best_dst);
        continue;
      }
      const int k = __ffs((int)cand) - 1;
      cand &= cand - 1u;
      m = cand_base + k;
      const DMesh* M = p.meshes + m;
      float tn;
      if (best_dst < INFINITY) {  // a hit exists: the box may now lie behind it
        const float4 wlo = __ldg(&M->wmin), whi = __ldg(&M->wmax);
        if (COUNT) c_box++;
        if (!box_cull(wlo.x, wlo.y, wlo.z, whi.x, whi.y, whi.z, winv, wnoi, best_dst, tn)) continue;
      }
      mflags = __float_as_uint(__ldg(&M->wmin.w));
      if (mflags & RR_MF_SPHERES) {
        lo = origin; ld = dir; linv = winv; lnoi = wnoi;
      } else {
        // WorldToLocalRay, src/Trace.cl:118-137
        const float4 i0 = __ldg(&M->ri0), i1 = __ldg(&M->ri1), i2 = __ldg(&M->ri2);
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
      }
      const float4 blo = __ldg(&M->bmin), bhi = __ldg(&M->bmax);
      if (COUNT) c_box++;
      if (!box_cull(blo.x, blo.y, blo.z, bhi.x, bhi.y, bhi.z, linv, lnoi, INFINITY, tn)) continue;
      const uint32_t first = __float_as_uint(blo.w), count = __float_as_uint(bhi.w);
      lt = INFINITY; lprim = NO_PRIM; lback = false;
      sp = 0;
      if (count <= RR_DIRECT_MAX) {  // no hierarchy: the primitives are tested one by one in the leaf phase
        pend_slot = (mflags & RR_MF_SPHERES) ? 0u : first;
        pend_cnt = count;
        cur = REF_END;
      } else {
        pend_cnt = 0;
        cur = (int32_t)first;  // root node
      }
      state = ST_TRAV;
      return;
    }
  };
  // after the last node / leaf of a mesh: more candidates -> setup phase, none -> shade phase (which
  // finishes the mesh itself)
  auto leave_mesh = [&]() { state = (cand != 0 || cand_base + 32 <= p.last_mesh) ? ST_SETUP : ST_SHADE; };

  for (;;) {
    // ---- vote: one REDUX over 6-bit fields counts the ready lanes of every phase ----
    const bool more_pixels = !(queue_empty && tile_next >= tile_pixels);
    uint32_t key = 0;
    if (state == ST_TRAV) {
      key = ((cur >= REF_POP && (speculate || pend_cnt == 0)) ? 1u << (6 * PH_TRAV) : 0u) | (pend_cnt ? 1u << (6 * PH_LEAF) : 0u);
    } else if (state == ST_SETUP) key = 1u << (6 * PH_SETUP);
    else if (state == ST_SHADE) key = 1u << (6 * PH_SHADE);
    else if (state == ST_PIXEL && more_pixels) key = 1u << (6 * PH_PIXEL);
    const uint32_t counts = __reduce_add_sync(full, key);
    if (counts == 0) break;
    const uint32_t nP = counts & 63u, nH = (counts >> 6) & 63u, nS = (counts >> 12) & 63u, nT = (counts >> 18) & 63u,
                   nL = (counts >> 24) & 63u;
    int phase = PH_TRAV;
    uint32_t best = nT * wT;
    if (nL * wL > best) { best = nL * wL; phase = PH_LEAF; }
    if (nS * wS > best) { best = nS * wS; phase = PH_SETUP; }
    if (nH * wH > best) { best = nH * wH; phase = PH_SHADE; }
    if (nP * wP > best) { best = nP * wP; phase = PH_PIXEL; }
    if (COUNT && phase != PH_TRAV) {
      ph_runs[phase]++;
      ph_lanes[phase] += phase == PH_LEAF ? nL : phase == PH_SETUP ? nS : phase == PH_SHADE ? nH : nP;
    }

    if (phase == PH_TRAV) {
      // ================= node steps =================
      uint32_t active = nT;
      do {
        if (COUNT) { ph_runs[PH_TRAV]++; ph_lanes[PH_TRAV] += active; }
        if (state == ST_TRAV && cur >= REF_POP && (speculate || pend_cnt == 0)) {
          if (cur == REF_POP) {  // one stack entry per step, in lock-step with the other lanes
            if (sp > 0) {
              --sp;
              const float d = stackD[sp];
              const int32_t n = stackN[sp];
              cur = d <= lt ? n : REF_POP;
              if (ref_is_leaf(cur) && pend_cnt == 0) { pend_slot = ref_slot(cur); pend_cnt = 1; cur = REF_POP; }
            } else {
              cur = REF_END;
            }
          }
          if (cur >= 0) {
            const float4* nd = p.nodes + 4 * (size_t)cur;
            const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
            if (COUNT) c_box += 2;
            float tA, tB;
            const bool hA = box_cull(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, linv, lnoi, lt, tA);
            const bool hB = box_cull(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, linv, lnoi, lt, tB);
            const int32_t L = __float_as_int(q3.x), R = __float_as_int(q3.y);
            int32_t next = REF_POP;
            if (hA && hB) {
              const bool aNear = tA < tB;
              const int32_t nearRef = aNear ? L : R, farRef = aNear ? R : L;
              if (pend_cnt == 0 && nearRef < REF_POP) {  // the near child is a leaf: postpone it, go on with the far one
                if (farRef < REF_POP) {  // two leaves: adjacent slots, L first
                  pend_slot = ref_slot(L); pend_cnt = 2;
                } else {
                  pend_slot = ref_slot(nearRef); pend_cnt = 1;
                  next = farRef;
                }
              } else {
                if (sp < RR_STACK) { stackN[sp] = farRef; stackD[sp] = aNear ? tB : tA; sp++; }
                next = nearRef;
              }
            } else if (hA) next = L;
            else if (hB) next = R;
            if (next < REF_POP && pend_cnt == 0) { pend_slot = ref_slot(next); pend_cnt = 1; next = REF_POP; }
            cur = next;
          }
          if (cur == REF_END && pend_cnt == 0) leave_mesh();
        }
        active = __popc(__ballot_sync(full, state == ST_TRAV && cur >= REF_POP && (speculate || pend_cnt == 0)));
      } while (active >= trav_keep);
    } else if (phase == PH_LEAF) {
      // ================= leaf tests =================
      if (state == ST_TRAV && pend_cnt > 0) {
        const uint32_t slot = pend_slot;
        pend_slot++;
        pend_cnt--;
        if (mflags & RR_MF_SPHERES) {
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
                V3 n = (hp - c) / r;
                if (back) n = -n;
                lt = t; lprim = prim; lback = back;
                SM_ST3(S_LN, n);
              }
            }
          }
        } else {
          // src/Trace.cl:276-317 with the distance test hoisted before the normal (same accept set) and a
          // total order (t, prim) instead of first-found-wins
          if (COUNT) c_tri++;
          const float4* gp = p.tri_geom + 3 * (size_t)slot;
          const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
          const V3 A = xyz(g0), edge1 = xyz(g1), edge2 = xyz(g2);
          const V3 h = cross(ld, edge2);
          const float a = dot(edge1, h);
          if (!(fabsf(a) < RR_EPSILON)) {
            const float f = 1.0f / a;
            const V3 s = lo - A;
            const float u = f * dot(s, h);
            if (!(u < 0.0f || u > 1.0f)) {
              const V3 q = cross(s, edge1);
              const float v = f * dot(ld, q);
              if (!(v < 0.0f || u + v > 1.0f)) {
                const float t = f * dot(edge2, q);
                const int32_t prim = (int32_t)__float_as_uint(g0.w);
                if (t > RR_EPSILON && (t < lt || (t == lt && lprim != NO_PRIM && prim < lprim))) {
                  const float4* np = p.tri_nrm + 3 * (size_t)slot;
                  const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
                  V3 n = normalize(xyz(n0) * (1.0f - u - v) + xyz(n1) * u + xyz(n2) * v);
                  bool back = false;
                  bool ok = true;
                  if (dot(ld, n) > RR_EPSILON) {
                    if (mflags & RR_MF_CULL) ok = false;
                    back = true;
                    n = -n;
                  }
                  if (ok) {
                    lt = t; lprim = prim; lback = back;
                    SM_ST3(S_LN, n);
                  }
                }
              }
            }
          }
        }
        if (pend_cnt == 0) {
          if (ref_is_leaf(cur)) {  // the leaf this lane was waiting on becomes the postponed one
            pend_slot = ref_slot(cur);
            pend_cnt = 1;
            cur = REF_POP;
          } else if (cur == REF_END) {
            leave_mesh();
          }
        }
      }
    } else if (phase == PH_SETUP) {
      // ================= finish the current mesh, enter the next candidate (src/Trace.cl:444-482) =================
      if (state == ST_SETUP) {
        finish_mesh();
        const V3 winv = SM_LD3(S_WINV);
        const V3 wnoi = mk(-(origin.x * winv.x), -(origin.y * winv.y), -(origin.z * winv.z));
        enter_next_mesh(winv, wnoi);
      }
    } else if (phase == PH_SHADE) {
      // ================= one bounce of Trace() (src/Trace.cl:497-591) and the sample loop (:639-642) =================
      if (state == ST_SHADE) {
        finish_mesh();
        if (PRIMARY) {
          if (p.hit_mesh) p.hit_mesh[pix] = best_dst < INFINITY ? best_mesh : -1;
          if (p.hit_prim) p.hit_prim[pix] = best_dst < INFINITY ? best_prim : -1;
          if (p.hit_dst) p.hit_dst[pix] = best_dst < INFINITY ? best_dst : 0.0f;
          state = ST_PIXEL;
        } else {
          SceneHit hit;
          hit.did = best_dst < INFINITY;
          hit.dst = best_dst;
          hit.point = SM_LD3(S_BP);
          hit.normal = SM_LD3(S_BN);
          hit.back = best_back;
          hit.material = best_mat;
          V3 throughput = SM_LD3(S_THR), incoming = SM_LD3(S_INC);
          bool alive = shade(p, hit, origin, dir, throughput, incoming, bounce, passes, rng);
          alive = alive && bounce < p.max_bounces;
          if (alive) {
            SM_ST3(S_THR, throughput);
            SM_ST3(S_INC, incoming);
          } else {  // path finished: src/Trace.cl:639-642
            const V3 accum = SM_LD3(S_ACC) + incoming;
            sample++;
            if (sample >= p.spp) {
              const V3 c = accum / (float)p.spp;
              reinterpret_cast<uint32_t*>(p.frame)[pix] = tonemap_rgba(c);
              if (p.radiance) {
                p.radiance[3 * (size_t)pix] = c.x;
                p.radiance[3 * (size_t)pix + 1] = c.y;
                p.radiance[3 * (size_t)pix + 2] = c.z;
              }
              state = ST_PIXEL;
            } else {
              SM_ST3(S_ACC, accum);
              begin_path();
            }
          }
          if (state != ST_PIXEL) {  // next segment: candidates, then straight into its first mesh
            begin_ray();
            const V3 winv = SM_LD3(S_WINV);
            const V3 wnoi = mk(-(origin.x * winv.x), -(origin.y * winv.y), -(origin.z * winv.z));
            enter_next_mesh(winv, wnoi);
          }
        }
      }
    } else {
      // ================= hand a pixel to every lane that needs one =================
      bool need = state == ST_PIXEL;
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
        const unsigned rank = __popc(mb & ((1u << lane) - 1u));
        const uint32_t k = tile_next + rank;
        if (need && k < tile_pixels) {
          const uint32_t x = tile_x0 + k % tile_w, y = tile_y0 + k / tile_w;
          pix = (int32_t)(y * p.width + x);
          if (!PRIMARY && p.max_bounces == 0) {  // no segment is ever traced: the pixel is black
            reinterpret_cast<uint32_t*>(p.frame)[pix] = tonemap_rgba(mk(0, 0, 0));
            if (p.radiance) { p.radiance[3 * (size_t)pix] = 0.0f; p.radiance[3 * (size_t)pix + 1] = 0.0f; p.radiance[3 * (size_t)pix + 2] = 0.0f; }
          } else {
            rng = make_seed((uint32_t)pix, p.frame_index, 0u);          // src/Trace.cl:631-632
            const V3 pd = primary_dir(p.cam, x, y, p.width, p.height);  // once per pixel, :634-636
            SM_ST3(S_PD, pd);
            SM(S_ACC) = 0.0f; SM(S_ACC + 1) = 0.0f; SM(S_ACC + 2) = 0.0f;
            sample = 0;
            begin_path();
            begin_ray();  // -> ST_SETUP
            need = false;
          }
        }
        tile_next += __popc(mb);
      }
      if (need) state = ST_IDLE;  // the queue is empty
    }
  }
#undef SM
#undef SM_ST3
#undef SM_LD3
  // counters: one atomic per warp
  unsigned long long r = n_rays;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(full, r, off);
  if (lane == 0) {
    atomicAdd(&p.counters->rays, r);
    atomicAdd(&p.counters->tiles, n_tiles);
  }
  if (COUNT) {
    unsigned long long b = c_box, t = c_tri, s = c_sph;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      b += __shfl_xor_sync(full, b, off);
      t += __shfl_xor_sync(full, t, off);
      s += __shfl_xor_sync(full, s, off);
    }
    if (lane == 0) {
      atomicAdd(&p.counters->box_tests, b);
      atomicAdd(&p.counters->tri_tests, t);
      atomicAdd(&p.counters->sphere_tests, s);
      for (int k = 0; k < 5; ++k) {
        atomicAdd(&p.counters->phase_runs[k], (unsigned long long)ph_runs[k]);
        atomicAdd(&p.counters->phase_lanes[k], (unsigned long long)ph_lanes[k]);
      }
    }
  }
}

void default_tuning(Tuning& t) {
  for (int k = 0; k < 5; ++k) t.weight[k] = 4;
  t.trav_keep = 12;
  t.speculate = 1;
  t.ctas_per_sm = 0;
}

static int resident_ctas(const void* fn, const RenderParams& p) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, NT, 0) != cudaSuccess || n < 1) n = 1;
  if (p.tune.ctas_per_sm && (int)p.tune.ctas_per_sm < n) n = (int)p.tune.ctas_per_sm;
  return n;
}

cudaError_t launch_render(const RenderParams& p, bool count_tests, int sm_count, cudaStream_t s) {
  // persistent: as many CTAs as are resident at once, each warp loops until the tile queue is empty
  if (count_tests) {
    const int grid = sm_count * resident_ctas((const void*)k_render<true, false>, p);
    k_render<true, false><<<grid, NT, 0, s>>>(p);
  } else {
    const int grid = sm_count * resident_ctas((const void*)k_render<false, false>, p);
    k_render<false, false><<<grid, NT, 0, s>>>(p);
  }
  return cudaGetLastError();
}

// Primary-ray closest hit per pixel (MakeRay + CalculateRayCollisionWithTriangle): the same kernel, stopped
// at the first shade phase.
cudaError_t launch_primary(const RenderParams& p, int sm_count, cudaStream_t s) {
  if (!p.width || !p.height) return cudaSuccess;
  const int grid = sm_count * resident_ctas((const void*)k_render<false, true>, p);
  k_render<false, true><<<grid, NT, 0, s>>>(p);
  return cudaGetLastError();
}

// ---- probes for the bit-level parity tests (tests/test_math_parity.py) ---------
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
