// rr_render.cu -- the per-pixel path-tracing kernels (sm_100a).
//
// B200-native replacement of the reference's `raytrace` OpenCL kernel
// (src/Trace.cl:623-653) and everything it calls.  Each device function names
// the reference lines whose arithmetic it reproduces; the arithmetic is kept in
// the reference's operation order and this file is compiled with -fmad=false
// so the result is bit-identical to the CPU oracle.
//
// Structure:
//   * persistent warps pop 8x4-pixel tiles from one 64-bit atomic counter (the
//     reference's mutex-guarded std::queue, src/image.hpp:286-314); a lane that
//     finishes its pixel takes the next pixel of the warp's tile at once;
//   * per pixel the spp samples run serially in the lane because the RNG state
//     is carried across samples (src/Trace.cl:632,639-642);
//   * closest hit = loop over meshes in mesh-local space (src/Trace.cl:444-482)
//     with our LBVH instead of the reference's SAH tree: one 64-byte node fetch
//     (4 x LDG.128) brings both child boxes and both child references.
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
struct Ray {
  V3 o, d, inv;
};

// src/Trace.cl:259-274
__device__ __forceinline__ bool ray_box(const Ray& r, float minx, float miny, float minz, float maxx, float maxy,
                                        float maxz, float& dist) {
  float t0x = (minx - r.o.x) * r.inv.x, t0y = (miny - r.o.y) * r.inv.y, t0z = (minz - r.o.z) * r.inv.z;
  float t1x = (maxx - r.o.x) * r.inv.x, t1y = (maxy - r.o.y) * r.inv.y, t1z = (maxz - r.o.z) * r.inv.z;
  float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
  float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  dist = tmin;
  return tmax >= fmaxf(tmin, 0.0f);
}

struct Hit {  // closest hit inside one mesh (local space) or the sphere set
  float t;
  int32_t prim;  // uploaded index; INT_MAX while empty
  V3 n;
  bool back;
  bool did;
};

// src/Trace.cl:276-317 with the distance test hoisted before the normal (same
// accept set) and a total order (t, prim) instead of first-found-wins.
__device__ __forceinline__ void ray_triangle(const Ray& ray, const float4* __restrict__ geom,
                                             const float4* __restrict__ nrm, uint32_t slot, bool cull, Hit& best) {
  const float4 g0 = __ldg(geom + 3 * (size_t)slot), g1 = __ldg(geom + 3 * (size_t)slot + 1),
               g2 = __ldg(geom + 3 * (size_t)slot + 2);
  const V3 A = xyz(g0), edge1 = xyz(g1), edge2 = xyz(g2);
  const V3 h = cross(ray.d, edge2);
  const float a = dot(edge1, h);
  if (fabsf(a) < RR_EPSILON) return;
  const float f = 1.0f / a;
  const V3 s = ray.o - A;
  const float u = f * dot(s, h);
  if (u < 0.0f || u > 1.0f) return;
  const V3 q = cross(s, edge1);
  const float v = f * dot(ray.d, q);
  if (v < 0.0f || u + v > 1.0f) return;
  const float t = f * dot(edge2, q);
  if (t <= RR_EPSILON) return;
  const int32_t prim = (int32_t)__float_as_uint(g0.w);
  if (!(t < best.t || (t == best.t && best.did && prim < best.prim))) return;
  const float4 n0 = __ldg(nrm + 3 * (size_t)slot), n1 = __ldg(nrm + 3 * (size_t)slot + 1),
               n2 = __ldg(nrm + 3 * (size_t)slot + 2);
  V3 n = normalize(xyz(n0) * (1.0f - u - v) + xyz(n1) * u + xyz(n2) * v);
  bool back = false;
  if (dot(ray.d, n) > RR_EPSILON) {
    if (cull) return;
    back = true;
    n = -n;
  }
  best.did = true;
  best.t = t;
  best.prim = prim;
  best.n = n;
  best.back = back;
}

// EXTENSION (the reference kernel has no sphere primitive): semantics defined by oracle/rr_oracle.c ray_sphere.
__device__ __forceinline__ void ray_sphere(const Ray& ray, float4 cr, int32_t prim, int32_t mtype, Hit& best) {
  const V3 c = xyz(cr);
  const float r = cr.w;
  const V3 oc = ray.o - c;
  const float b = dot(oc, ray.d);
  const float cc = dot(oc, oc) - r * r;
  const float disc = b * b - cc;
  if (!(disc >= 0.0f)) return;
  const float sq = sqrtf(disc);
  float t = -b - sq;
  bool back = false;
  if (t <= RR_EPSILON) { t = -b + sq; back = true; }
  if (t <= RR_EPSILON) return;
  if (!(t < best.t || (t == best.t && best.did && prim < best.prim))) return;
  const bool cull = (mtype != RR_MATERIAL_GLASSY && mtype != RR_MATERIAL_INVISIBLE && mtype != RR_MATERIAL_ONESIDED);
  if (back && cull) return;
  const V3 hp = ray.o + ray.d * t;
  V3 n = (hp - c) / r;
  if (back) n = -n;
  best.did = true; best.t = t; best.prim = prim; best.n = n; best.back = back;
}

struct TraversalCounters {
  unsigned box, tri, sph;
};

// LBVH traversal of one segment; replaces src/Trace.cl:319-397.  PRIM = 0 triangles, 1 spheres.
template <int PRIM, bool COUNT>
__device__ __forceinline__ void traverse(const RenderParams& p, const Ray& ray, const float4* __restrict__ nodes,
                                         uint32_t sfirst, uint32_t count, bool cull, Hit& best, TraversalCounters& tc) {
  auto leaf = [&](uint32_t slot) {
    if (PRIM == 0) {
      if (COUNT) tc.tri++;
      ray_triangle(ray, p.tri_geom, p.tri_nrm, slot, cull, best);
    } else {
      if (COUNT) tc.sph++;
      const uint32_t prim = __ldg(p.sph_order + slot);
      const int32_t mtype = __ldg(&p.materials[p.n_meshes + prim].type);
      ray_sphere(ray, __ldg(p.sph_geom + slot), (int32_t)prim, mtype, best);
    }
  };
  if (count <= RR_DIRECT_MAX) {
    for (uint32_t k = 0; k < count; ++k) leaf(sfirst + k);
    return;
  }
  int32_t stackN[RR_STACK];
  float stackD[RR_STACK];
  int sp = 0;
  int32_t cur = (int32_t)sfirst;
  for (;;) {
    const float4* nd = nodes + 4 * (size_t)cur;
    const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
    const int32_t L = __float_as_int(q3.x), R = __float_as_int(q3.y);
    float dA, dB;
    if (COUNT) tc.box += 2;
    const bool hA = ray_box(ray, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, dA) && dA < best.t;
    const bool hB = ray_box(ray, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, dB) && dB < best.t;
    int32_t next = 0;
    bool have = false;
    if (hA && hB) {
      int32_t far;
      float dfar;
      if (dA < dB) { next = L; far = R; dfar = dB; } else { next = R; far = L; dfar = dA; }
      if (sp < RR_STACK) { stackN[sp] = far; stackD[sp] = dfar; sp++; }
      have = true;
    } else if (hA) { next = L; have = true; }
    else if (hB) { next = R; have = true; }
    for (;;) {
      if (have) {
        if (next >= 0) { cur = next; break; }
        leaf((uint32_t)~next);
        have = false;
      }
      bool found = false;
      while (sp > 0) {
        --sp;
        if (stackD[sp] < best.t) { next = stackN[sp]; found = true; break; }
      }
      if (!found) return;
      have = true;
    }
  }
}

struct SceneHit {
  bool did;
  float dst;
  V3 point, normal;
  bool back;
  int32_t mesh;      // mesh index, n_meshes for a sphere
  int32_t prim;      // uploaded primitive index
  int32_t material;  // index into the material table
};

// src/Trace.cl:434-485 (+ the sphere extension after the mesh loop).
template <bool COUNT>
__device__ __forceinline__ void scene_closest(const RenderParams& p, V3 origin, V3 dir, SceneHit& out, TraversalCounters& tc) {
  out.did = false;
  out.dst = INFINITY;
  out.mesh = -1;
  out.prim = -1;
  out.material = 0;
  for (int m = 0; m < p.n_meshes; ++m) {
    const DMesh* M = p.meshes + m;
    if (__ldg(&M->skip)) continue;
    const float scale = __ldg(&M->scale);
    const V3 pos = ld3(M->pos);
    const V3 i0 = ld3(M->Rinv), i1 = ld3(M->Rinv + 3), i2 = ld3(M->Rinv + 6);
    // WorldToLocalRay, src/Trace.cl:118-137
    const V3 rel = origin - pos;
    V3 lo = mk(dot(i0, rel), dot(i1, rel), dot(i2, rel));
    V3 ld = mk(dot(i0, dir), dot(i1, dir), dot(i2, dir));
    if (fabsf(scale) > RR_EPSILON) {
      lo = lo / scale;
      ld = ld / scale;
    }
    ld = normalize(ld);
    Ray lr;
    lr.o = lo;
    lr.d = ld;
    lr.inv = mk(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
    Hit lh;
    lh.did = false; lh.t = INFINITY; lh.prim = 0x7fffffff; lh.back = false; lh.n = mk(0, 0, 0);
    float dRoot;
    if (COUNT) tc.box++;
    if (!ray_box(lr, __ldg(M->bmin), __ldg(M->bmin + 1), __ldg(M->bmin + 2), __ldg(M->bmax), __ldg(M->bmax + 1),
                 __ldg(M->bmax + 2), dRoot))
      continue;
    const bool cull = __ldg(&M->cull) != 0;
    traverse<0, COUNT>(p, lr, p.tri_nodes, __ldg(&M->sfirst), __ldg(&M->count), cull, lh, tc);
    if (!lh.did) continue;
    const int32_t type = __ldg(&M->type);
    if (type == RR_MATERIAL_ONESIDED && lh.back) continue;
    // LocalToWorldHit, src/Trace.cl:139-156
    const V3 r0 = ld3(M->R), r1 = ld3(M->R + 3), r2 = ld3(M->R + 6);
    const V3 lp = (lr.o + lr.d * lh.t) * scale;
    const V3 wp = mk(dot(r0, lp), dot(r1, lp), dot(r2, lp)) + pos;
    const V3 wn = normalize(mk(dot(r0, lh.n), dot(r1, lh.n), dot(r2, lh.n)));
    const float wd = length(wp - origin);
    if (wd < out.dst) {
      out.did = true; out.dst = wd; out.point = wp; out.normal = wn; out.back = lh.back;
      out.mesh = m; out.prim = lh.prim; out.material = __ldg(&M->material);
    }
  }
  if (p.n_spheres > 0) {
    Ray wr;
    wr.o = origin;
    wr.d = dir;
    wr.inv = mk(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    float dRoot;
    if (COUNT) tc.box++;
    if (ray_box(wr, p.sph_bmin[0], p.sph_bmin[1], p.sph_bmin[2], p.sph_bmax[0], p.sph_bmax[1], p.sph_bmax[2], dRoot)) {
      Hit sh;
      sh.did = false; sh.t = INFINITY; sh.prim = 0x7fffffff; sh.back = false; sh.n = mk(0, 0, 0);
      traverse<1, COUNT>(p, wr, p.sph_nodes, 0u, (uint32_t)p.n_spheres, false, sh, tc);
      if (sh.did) {
        const int32_t mat = p.n_meshes + sh.prim;
        const int32_t type = __ldg(&p.materials[mat].type);
        if (!(type == RR_MATERIAL_ONESIDED && sh.back) && sh.t < out.dst) {
          out.did = true; out.dst = sh.t; out.point = origin + dir * sh.t; out.normal = sh.n; out.back = sh.back;
          out.mesh = p.n_meshes; out.prim = sh.prim; out.material = mat;
        }
      }
    }
  }
}

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
// next tile; with a static partition (queue == nullptr is not used: the local
// counter is scaled by tile_stride) rank r renders tiles r, r+world, ...
__device__ __forceinline__ bool pop_tile(const RenderParams& p, uint32_t& tile) {
  unsigned long long t = 0;
  if ((threadIdx.x & 31) == 0) t = atomicAdd_system(p.queue, 1ull);
  t = __shfl_sync(0xffffffffu, t, 0);
  t = (unsigned long long)p.tile_begin + t * p.tile_stride;
  tile = (uint32_t)t;
  return t < (unsigned long long)p.tiles_x * p.tiles_y;
}

constexpr int RENDER_THREADS = 256;

template <bool COUNT>
__global__ void __launch_bounds__(RENDER_THREADS, 2) k_render(const RenderParams p) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  TraversalCounters tc = {0u, 0u, 0u};
  unsigned long long n_rays = 0, n_tiles = 0;
  const V3 cam_pos = mk(p.cam.pos[0], p.cam.pos[1], p.cam.pos[2]);
  // warp-uniform tile state
  bool queue_empty = false;
  uint32_t tile_x0 = 0, tile_y0 = 0, tile_w = 0, tile_h = 0, tile_next = 0, tile_pixels = 0;
  // lane state
  int32_t pix = -1;
  uint32_t rng = 0, sample = 0, bounce = 0, passes = 0;
  V3 pd = mk(0, 0, 1), origin = cam_pos, dir = pd, throughput = mk(1, 1, 1), incoming = mk(0, 0, 0), accum = mk(0, 0, 0);

  for (;;) {
    // ---- hand a pixel to every idle lane ----
    bool need = pix < 0;
    while (__any_sync(full, need)) {
      if (tile_next >= tile_pixels) {
        if (queue_empty) break;
        uint32_t tile;
        if (!pop_tile(p, tile)) { queue_empty = true; break; }
        n_tiles++;
        tile_x0 = (tile % p.tiles_x) * p.tile_w;
        tile_y0 = (tile / p.tiles_x) * p.tile_h;
        tile_w = min(p.tile_w, p.width - tile_x0);
        tile_h = min(p.tile_h, p.height - tile_y0);
        tile_pixels = tile_w * tile_h;
        tile_next = 0;
      }
      const unsigned m = __ballot_sync(full, need);
      const unsigned rank = __popc(m & ((1u << lane) - 1u));
      const uint32_t k = tile_next + rank;
      if (need && k < tile_pixels) {
        const uint32_t x = tile_x0 + k % tile_w, y = tile_y0 + k / tile_w;
        pix = (int32_t)(y * p.width + x);
        rng = make_seed((uint32_t)pix, p.frame_index, 0u);  // src/Trace.cl:631-632
        pd = primary_dir(p.cam, x, y, p.width, p.height);   // once per pixel, :634-636
        accum = mk(0, 0, 0);
        sample = 0; bounce = 0; passes = 0;
        origin = cam_pos; dir = pd; throughput = mk(1, 1, 1); incoming = mk(0, 0, 0);
        need = false;
      }
      tile_next += __popc(m);
    }
    if (__all_sync(full, pix < 0)) break;

    if (pix >= 0) {
      bool alive = bounce < p.max_bounces && sample < p.spp;
      if (alive) {
        SceneHit hit;
        n_rays++;
        scene_closest<COUNT>(p, origin, dir, hit, tc);
        alive = shade(p, hit, origin, dir, throughput, incoming, bounce, passes, rng);
        alive = alive && bounce < p.max_bounces;
      }
      if (!alive) {  // path finished: src/Trace.cl:639-642
        if (sample < p.spp) {
          accum = accum + incoming;
          sample++;
        }
        if (sample >= p.spp) {
          const V3 c = accum / (float)p.spp;
          reinterpret_cast<uint32_t*>(p.frame)[pix] = tonemap_rgba(c);
          if (p.radiance) {
            p.radiance[3 * (size_t)pix] = c.x;
            p.radiance[3 * (size_t)pix + 1] = c.y;
            p.radiance[3 * (size_t)pix + 2] = c.z;
          }
          pix = -1;
        } else {
          bounce = 0; passes = 0;
          origin = cam_pos; dir = pd; throughput = mk(1, 1, 1); incoming = mk(0, 0, 0);
        }
      }
    }
  }
  // counters: one atomic per warp
  unsigned long long r = n_rays;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) r += __shfl_xor_sync(full, r, off);
  if (lane == 0) {
    atomicAdd(&p.counters->rays, r);
    atomicAdd(&p.counters->tiles, n_tiles);
  }
  if (COUNT) {
    unsigned long long b = tc.box, t = tc.tri, s = tc.sph;
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
    }
  }
}

// Primary-ray closest hit per pixel (MakeRay + CalculateRayCollisionWithTriangle).
__global__ void __launch_bounds__(256) k_primary(const RenderParams p) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.width * p.height) return;
  const uint32_t x = i % p.width, y = i / p.width;
  const V3 d = primary_dir(p.cam, x, y, p.width, p.height);
  SceneHit hit;
  TraversalCounters tc = {0u, 0u, 0u};
  scene_closest<false>(p, mk(p.cam.pos[0], p.cam.pos[1], p.cam.pos[2]), d, hit, tc);
  if (p.hit_mesh) p.hit_mesh[i] = hit.did ? hit.mesh : -1;
  if (p.hit_prim) p.hit_prim[i] = hit.did ? hit.prim : -1;
  if (p.hit_dst) p.hit_dst[i] = hit.did ? hit.dst : 0.0f;
}

cudaError_t launch_render(const RenderParams& p, bool count_tests, int sm_count, cudaStream_t s) {
  const int grid = sm_count * 2;  // persistent: launch_bounds(256, 2) -> two resident CTAs per SM
  if (count_tests) k_render<true><<<grid, RENDER_THREADS, 0, s>>>(p);
  else k_render<false><<<grid, RENDER_THREADS, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_primary(const RenderParams& p, cudaStream_t s) {
  const uint64_t n = (uint64_t)p.width * p.height;
  if (!n) return cudaSuccess;
  k_primary<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
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
