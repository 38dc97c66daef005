/*
 * oracle/rr_oracle.c -- TEST INFRASTRUCTURE.  Not product code.
 *
 * CPU restatement ("Oracle A'") of the render hot path of ripoff-raytracer
 * (/root/reference/src/Trace.cl `raytrace`) plus the CPU statement of OUR
 * LBVH ("Oracle B": Morton codes -> stable sort -> Karras hierarchy -> refit),
 * which the reference does not have (it ships a host SAH builder,
 * src/readobj.hpp:96-267).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (ripoff_raytracer_b200/csrc) never links or calls it.
 *
 * Parity pinning: the reference has no tests or golden vectors.
 * This restatement is pinned against the reference's OWN kernel text compiled
 * for the host (oracle/_ref/libref_strict.so, built by oracle/Makefile `ref`)
 * in tests/test_oracle_vs_reference.py and against fixtures generated from it
 * (tests/golden/, script tests/golden/make_golden.py).  Both sides evaluate
 * the OpenCL builtins whose precision is implementation-defined through the
 * numerics contract of oracle/rr_math_ref.h (IEEE binary32, round-to-nearest,
 * no FMA contraction, left-to-right dot products).
 * SPHERES: the reference kernel has no sphere primitive; the
 * sphere semantics below are an extension DEFINED here -- "parity unpinned"
 * for that primitive.
 *
 * Every function cites the reference lines it restates.
 * Build: gcc -std=c11 -O2 -ffp-contract=off (oracle/Makefile).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rr_api.h"
#include "rr_math_ref.h"

#define RRO_EPSILON 1e-6f            /* src/Trace.cl:6 */
#define RRO_TAU 6.28318530717958647692f /* src/Trace.cl:5 */
#define RRO_IOR_AIR 1.0f             /* src/Trace.cl:7 */
#define RRO_STACK 64                 /* src/Trace.cl:2 */
#define RRO_MAX_INVISIBLE_PASSES 256u /* see trace_path */
#define RRO_DIRECT_MAX 4             /* segments this small are tested without a hierarchy */

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vscale_l(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
static inline v3 vdivs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* fast_normalize / normalize of the numerics contract */
static inline v3 vnormalize(v3 a) {
  float inv = 1.0f / sqrtf(vdot(a, a));
  return vscale(a, inv);
}
static inline float vlength(v3 a) { return sqrtf(vdot(a, a)); }
static inline v3 from_f3(const rr_float3* f) { return V(f->s[0], f->s[1], f->s[2]); }

/* ---------------------------------------------------------------- RNG ---- */
/* src/Trace.cl:158-161 */
static inline float map_u32(uint32_t s) { return (float)(s + 1u) * (1.0f / 4294967296.0f); }
/* src/Trace.cl:163-168 */
static inline float random_value(uint32_t* state) {
  *state = *state * 747796405u + 2891336453u;
  uint32_t result = ((*state >> ((*state >> 28) + 4u)) ^ *state) * 277803737u;
  result = (result >> 22) ^ result;
  return map_u32(result);
}
/* src/Trace.cl:170-177 */
static inline uint32_t make_seed(uint32_t pixelIndex, int32_t frameIndex, uint32_t rayIdx) {
  uint32_t s = pixelIndex * 1664525u + (uint32_t)frameIndex * 1013904223u;
  s ^= (rayIdx + 0x9E3779B9u);
  s = s * 22695477u + 1u;
  return s;
}
/* src/Trace.cl:209-217 */
static inline float rand01(uint32_t* state) {
  *state = *state * 747796405u + 2891336453u;
  uint32_t z = *state;
  z = (z ^ (z >> 16)) * 0x7feb352du;
  z = (z ^ (z >> 15)) * 0x846ca68bu;
  z = z ^ (z >> 16);
  return map_u32(z);
}
/* src/Trace.cl:179-187 */
static inline float random_normal(uint32_t* state) {
  float u1 = random_value(state);
  float u2 = random_value(state);
  u1 = fmaxf(u1, RRO_EPSILON);
  float r = sqrtf(-2.0f * rr_logf_ref(u1));
  float theta = RRO_TAU * u2;
  return r * rr_cosf_ref(theta);
}
static inline int rro_finite(float x) { return (rrm_bits(x) & 0x7f800000u) != 0x7f800000u; }
/* src/Trace.cl:189-200 */
static inline v3 random_direction(uint32_t* state) {
  float x = random_normal(state);
  float y = random_normal(state);
  float z = random_normal(state);
  v3 v = vnormalize(V(x, y, z));
  if (!rro_finite(v.x) || !rro_finite(v.y) || !rro_finite(v.z)) v = V(0.0f, 1.0f, 0.0f);
  return v;
}

/* ------------------------------------------------------------ scene ------ */
typedef struct {
  v3 r0, r1, r2; /* rows of R (makeRotation) */
} m33;

/* src/Trace.cl:90-100 */
static m33 make_rotation(float pitch, float yaw, float roll) {
  float cx = rr_cosf_ref(pitch), sx = rr_sinf_ref(pitch);
  float cy = rr_cosf_ref(yaw), sy = rr_sinf_ref(yaw);
  float cz = rr_cosf_ref(roll), sz = rr_sinf_ref(roll);
  m33 m;
  m.r0 = V(cy * cz, cy * sz, -sy);
  m.r1 = V(cz * sy * sx - cx * sz, cx * cz + sx * sy * sz, cy * sx);
  m.r2 = V(sx * sz + cx * cz * sy, cx * sy * sz - cz * sx, cx * cy);
  return m;
}
/* src/Trace.cl:109-116 */
static m33 transpose(m33 m) {
  m33 t;
  t.r0 = V(m.r0.x, m.r1.x, m.r2.x);
  t.r1 = V(m.r0.y, m.r1.y, m.r2.y);
  t.r2 = V(m.r0.z, m.r1.z, m.r2.z);
  return t;
}
/* src/Trace.cl:105-107 */
static inline v3 mul_mat_vec(const m33* m, v3 v) { return V(vdot(m->r0, v), vdot(m->r1, v), vdot(m->r2, v)); }

/* Our LBVH over a set of primitive boxes, segmented (one hierarchy per
 * segment).  All indices are GLOBAL (into the arrays below). */
typedef struct {
  uint64_t n;        /* primitives */
  uint64_t* codes;   /* [n] sorted Morton keys                          */
  uint32_t* order;   /* [n] primitive index at sorted slot              */
  int32_t* left;     /* [n] children of inner node i; <0 = ~sorted slot */
  int32_t* right;
  int32_t* parent;   /* [n] parent inner node of inner node i, -1 none  */
  float* bounds;     /* [n*6] inner node boxes                          */
  float* prim_box;   /* [n*6] primitive boxes in ORIGINAL order          */
  uint32_t max_depth;
} rro_lbvh;

typedef struct {
  uint64_t first, count;
  float bmin[3], bmax[3]; /* local-space root box, delta-inflated */
  float delta;            /* conservative slack of every box of this segment (box_delta) */
  m33 R, Rinv;
  int cull;
} rro_meshx;

typedef struct rro_scene {
  uint64_t n_tris, n_meshes, n_spheres;
  rr_triangle* tris;
  rr_mesh* meshes;
  rr_mesh_range* ranges;
  rr_sphere* spheres;
  rro_meshx* mx;
  rro_lbvh tb;  /* triangles, one segment per mesh */
  rro_lbvh sb;  /* spheres, one segment */
  float sph_bmin[3], sph_bmax[3];
  float sph_delta;
  /* optional: the reference's own node list (GPUNode layout) for validation */
  const void* ref_nodes;
  int brute_force; /* 1: closest hit by testing every primitive (no hierarchy): defines the result */
} rro_scene;

/* ---- Oracle B: LBVH --------------------------------------------------- */
static inline uint64_t expand21(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
static inline uint32_t quant21(float c, float lo, float ext) {
  if (!(ext > 0.0f)) return 0u;
  float q = (c - lo) / ext;
  float g = q * 2097152.0f;
  if (!(g > 0.0f)) return 0u; /* also NaN */
  if (g >= 2097151.0f) return 2097151u;
  return (uint32_t)g;
}
/* 63-bit Morton key of the centre of a primitive box inside the segment box */
static uint64_t morton63(const float* pb, const float* smin, const float* smax) {
  uint32_t g[3];
  for (int a = 0; a < 3; ++a) {
    float c = (pb[a] + pb[3 + a]) * 0.5f;
    g[a] = quant21(c, smin[a], smax[a] - smin[a]);
  }
  return (expand21(g[0]) << 2) | (expand21(g[1]) << 1) | expand21(g[2]);
}
static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int lbvh_delta(const uint64_t* codes, int64_t n, int64_t i, int64_t j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = codes[i], b = codes[j];
  if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
  return clz64(a ^ b);
}

typedef struct { uint64_t code; uint32_t idx; } keyidx;
static int cmp_keyidx(const void* pa, const void* pb) {
  const keyidx* a = (const keyidx*)pa;
  const keyidx* b = (const keyidx*)pb;
  if (a->code != b->code) return a->code < b->code ? -1 : 1;
  if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1;
  return 0;
}

static void lbvh_alloc(rro_lbvh* b, uint64_t n) {
  memset(b, 0, sizeof(*b));
  b->n = n;
  size_t m = n ? n : 1;
  b->codes = (uint64_t*)calloc(m, sizeof(uint64_t));
  b->order = (uint32_t*)calloc(m, sizeof(uint32_t));
  b->left = (int32_t*)calloc(m, sizeof(int32_t));
  b->right = (int32_t*)calloc(m, sizeof(int32_t));
  b->parent = (int32_t*)malloc(m * sizeof(int32_t));
  for (size_t i = 0; i < m; ++i) b->parent[i] = -1;
  b->bounds = (float*)calloc(m * 6, sizeof(float));
  b->prim_box = (float*)calloc(m * 6, sizeof(float));
}
static void lbvh_free(rro_lbvh* b) {
  free(b->codes); free(b->order); free(b->left); free(b->right); free(b->parent); free(b->bounds); free(b->prim_box);
  memset(b, 0, sizeof(*b));
}

static void box_of_ref(const rro_lbvh* b, int32_t ref, float* out) {
  if (ref < 0) memcpy(out, b->prim_box + 6 * (size_t)b->order[~ref], 24);
  else memcpy(out, b->bounds + 6 * (size_t)ref, 24);
}

/* Conservative slack added to every box a ray is tested against (same statement as
 * ripoff_raytracer_b200/csrc/rr_internal.h box_delta): 2^-18 of the largest |coordinate| of the
 * segment box.  With it the hierarchy only culls what cannot hold the closest hit, so the result
 * equals the brute-force minimum over all primitives whatever the traversal order. */
static float box_delta(const float* smin, const float* smax) {
  float m = 0.0f;
  for (int k = 0; k < 3; ++k) {
    float a = fabsf(smin[k]), c = fabsf(smax[k]);
    if (a > m && a < 3.0e38f) m = a;
    if (c > m && c < 3.0e38f) m = c;
  }
  return m * 3.814697265625e-06f;
}
static void inflate_box(float* box6, float d) {
  for (int k = 0; k < 3; ++k) { box6[k] -= d; box6[3 + k] += d; }
}

/* Builds the hierarchy of one segment [first, first+n) whose prim_box entries
 * are filled.  Writes the segment box to smin/smax. */
static void lbvh_build_segment(rro_lbvh* b, uint64_t first, uint64_t n, float* smin, float* smax) {
  for (int a = 0; a < 3; ++a) { smin[a] = INFINITY; smax[a] = -INFINITY; }
  for (uint64_t i = 0; i < n; ++i) {
    const float* pb = b->prim_box + 6 * (first + i);
    for (int a = 0; a < 3; ++a) {
      smin[a] = fminf(smin[a], pb[a]);
      smax[a] = fmaxf(smax[a], pb[3 + a]);
    }
  }
  if (n == 0) return;
  keyidx* ki = (keyidx*)malloc(n * sizeof(keyidx));
  for (uint64_t i = 0; i < n; ++i) {
    ki[i].code = morton63(b->prim_box + 6 * (first + i), smin, smax);
    ki[i].idx = (uint32_t)(first + i);
  }
  qsort(ki, n, sizeof(keyidx), cmp_keyidx); /* == stable sort by code */
  for (uint64_t i = 0; i < n; ++i) {
    b->codes[first + i] = ki[i].code;
    b->order[first + i] = ki[i].idx;
  }
  free(ki);
  if (n < 2) return;
  const uint64_t* codes = b->codes + first;
  const int64_t N = (int64_t)n;
  for (int64_t i = 0; i < N - 1; ++i) { /* Karras 2012, fig. 4 */
    int d = (lbvh_delta(codes, N, i, i + 1) - lbvh_delta(codes, N, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(codes, N, i, i - d);
    int64_t lmax = 2;
    while (lbvh_delta(codes, N, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
      if (lbvh_delta(codes, N, i, i + (l + t) * d) > dmin) l += t;
    int64_t j = i + l * d;
    int dnode = lbvh_delta(codes, N, i, j);
    int64_t s = 0;
    for (int64_t t = (l + 1) >> 1;; t = (t + 1) >> 1) {
      if (lbvh_delta(codes, N, i, i + (s + t) * d) > dnode) s += t;
      if (t <= 1) break;
    }
    int64_t gamma = i + s * d + (d < 0 ? -1 : 0);
    int64_t lo = i < j ? i : j, hi = i < j ? j : i;
    int32_t L = (lo == gamma) ? ~(int32_t)(first + gamma) : (int32_t)(first + gamma);
    int32_t R = (hi == gamma + 1) ? ~(int32_t)(first + gamma + 1) : (int32_t)(first + gamma + 1);
    b->left[first + i] = L;
    b->right[first + i] = R;
    if (L >= 0) b->parent[L] = (int32_t)(first + i);
    if (R >= 0) b->parent[R] = (int32_t)(first + i);
  }
  /* refit, post-order with an explicit stack; also measures the depth */
  int32_t* st = (int32_t*)malloc((size_t)n * 2 * sizeof(int32_t) + 16);
  uint32_t* dep = (uint32_t*)malloc((size_t)n * 2 * sizeof(uint32_t) + 16);
  size_t sp = 0;
  st[sp] = (int32_t)first; dep[sp] = 1; sp++;
  /* first pass: preorder list */
  int32_t* pre = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  size_t np = 0;
  uint32_t maxd = 1;
  while (sp) {
    --sp;
    int32_t nd = st[sp];
    uint32_t dd = dep[sp];
    pre[np++] = nd;
    if (dd + 1 > maxd) maxd = dd + 1; /* children (leaf or inner) sit one level deeper */
    if (b->left[nd] >= 0) { st[sp] = b->left[nd]; dep[sp] = dd + 1; sp++; }
    if (b->right[nd] >= 0) { st[sp] = b->right[nd]; dep[sp] = dd + 1; sp++; }
  }
  for (size_t k = np; k-- > 0;) { /* reverse preorder: children before parents */
    int32_t nd = pre[k];
    float a[6], c[6];
    box_of_ref(b, b->left[nd], a);
    box_of_ref(b, b->right[nd], c);
    float* o = b->bounds + 6 * (size_t)nd;
    for (int x = 0; x < 3; ++x) {
      o[x] = fminf(a[x], c[x]);
      o[3 + x] = fmaxf(a[3 + x], c[3 + x]);
    }
  }
  if (maxd > b->max_depth) b->max_depth = maxd;
  free(st); free(dep); free(pre);
}

/* ---- scene ------------------------------------------------------------ */
rro_scene* rro_scene_create(const rr_triangle* tris, uint64_t n_tris, const rr_mesh* meshes,
                            const rr_mesh_range* ranges, uint64_t n_meshes, const rr_sphere* spheres,
                            uint64_t n_spheres) {
  rro_scene* s = (rro_scene*)calloc(1, sizeof(rro_scene));
  s->n_tris = n_tris; s->n_meshes = n_meshes; s->n_spheres = n_spheres;
  s->tris = (rr_triangle*)aligned_alloc(16, (n_tris ? n_tris : 1) * sizeof(rr_triangle));
  s->meshes = (rr_mesh*)aligned_alloc(16, (n_meshes ? n_meshes : 1) * sizeof(rr_mesh));
  s->ranges = (rr_mesh_range*)calloc(n_meshes ? n_meshes : 1, sizeof(rr_mesh_range));
  s->spheres = (rr_sphere*)aligned_alloc(16, (n_spheres ? n_spheres : 1) * sizeof(rr_sphere));
  if (n_tris) memcpy(s->tris, tris, n_tris * sizeof(rr_triangle));
  if (n_meshes) memcpy(s->meshes, meshes, n_meshes * sizeof(rr_mesh));
  if (n_meshes) memcpy(s->ranges, ranges, n_meshes * sizeof(rr_mesh_range));
  if (n_spheres) memcpy(s->spheres, spheres, n_spheres * sizeof(rr_sphere));
  s->mx = (rro_meshx*)calloc(n_meshes ? n_meshes : 1, sizeof(rro_meshx));
  lbvh_alloc(&s->tb, n_tris);
  for (uint64_t i = 0; i < n_tris; ++i) {
    const rr_triangle* t = &s->tris[i];
    float* pb = s->tb.prim_box + 6 * i;
    for (int a = 0; a < 3; ++a) {
      pb[a] = fminf(fminf(t->posA.s[a], t->posB.s[a]), t->posC.s[a]);
      pb[3 + a] = fmaxf(fmaxf(t->posA.s[a], t->posB.s[a]), t->posC.s[a]);
    }
  }
  for (uint64_t m = 0; m < n_meshes; ++m) {
    rro_meshx* x = &s->mx[m];
    x->first = ranges[m].firstTriangle;
    x->count = ranges[m].numTriangles;
    if (x->first + x->count > n_tris) { x->count = 0; }
    lbvh_build_segment(&s->tb, x->first, x->count, x->bmin, x->bmax);
    x->delta = box_delta(x->bmin, x->bmax);
    for (int a = 0; a < 3; ++a) { x->bmin[a] -= x->delta; x->bmax[a] += x->delta; }
    x->R = make_rotation(meshes[m].pitch, meshes[m].yaw, meshes[m].roll);
    x->Rinv = transpose(x->R);
    int ty = meshes[m].material.type;
    x->cull = (ty != RR_MATERIAL_GLASSY && ty != RR_MATERIAL_INVISIBLE && ty != RR_MATERIAL_ONESIDED);
  }
  lbvh_alloc(&s->sb, n_spheres);
  for (uint64_t i = 0; i < n_spheres; ++i) {
    const rr_sphere* sp = &s->spheres[i];
    float* pb = s->sb.prim_box + 6 * i;
    for (int a = 0; a < 3; ++a) {
      pb[a] = sp->center.s[a] - sp->radius;
      pb[3 + a] = sp->center.s[a] + sp->radius;
    }
  }
  lbvh_build_segment(&s->sb, 0, n_spheres, s->sph_bmin, s->sph_bmax);
  s->sph_delta = box_delta(s->sph_bmin, s->sph_bmax);
  for (int a = 0; a < 3; ++a) { s->sph_bmin[a] -= s->sph_delta; s->sph_bmax[a] += s->sph_delta; }
  return s;
}

void rro_scene_destroy(rro_scene* s) {
  if (!s) return;
  free(s->tris); free(s->meshes); free(s->ranges); free(s->spheres); free(s->mx);
  lbvh_free(&s->tb); lbvh_free(&s->sb);
  free(s);
}

/* GPUNode list of the reference (src/image.hpp:116-125), 48 B per node; the
 * caller keeps it alive.  When set, triangle meshes are traversed with the
 * reference's own hierarchy and tie rule (validation against libref). */
void rro_scene_set_ref_nodes(rro_scene* s, const void* gpunodes) { s->ref_nodes = gpunodes; }
void rro_scene_set_brute_force(rro_scene* s, int on) { s->brute_force = on; }

uint64_t rro_lbvh_size(const rro_scene* s, int which) { return which ? s->sb.n : s->tb.n; }
uint32_t rro_lbvh_depth(const rro_scene* s, int which) { return which ? s->sb.max_depth : s->tb.max_depth; }
void rro_lbvh_read(const rro_scene* s, int which, uint64_t* codes, uint32_t* order, int32_t* left, int32_t* right,
                   int32_t* parent, float* bounds) {
  const rro_lbvh* b = which ? &s->sb : &s->tb;
  if (codes) memcpy(codes, b->codes, b->n * 8);
  if (order) memcpy(order, b->order, b->n * 4);
  if (left) memcpy(left, b->left, b->n * 4);
  if (right) memcpy(right, b->right, b->n * 4);
  if (parent) memcpy(parent, b->parent, b->n * 4);
  if (bounds) memcpy(bounds, b->bounds, b->n * 24);
}

/* ------------------------------------------------------- intersection ---- */
typedef struct {
  v3 origin, direction, invDir;
} ray_t;

typedef struct {
  uint64_t rays, box_tests, tri_tests, sphere_tests;
} rro_counters;

/* src/Trace.cl:259-274 */
static inline int ray_box(const ray_t* r, const float* bmin, const float* bmax, float* outDist) {
  float t0x = (bmin[0] - r->origin.x) * r->invDir.x, t0y = (bmin[1] - r->origin.y) * r->invDir.y,
        t0z = (bmin[2] - r->origin.z) * r->invDir.z;
  float t1x = (bmax[0] - r->origin.x) * r->invDir.x, t1y = (bmax[1] - r->origin.y) * r->invDir.y,
        t1z = (bmax[2] - r->origin.z) * r->invDir.z;
  float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
  float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  *outDist = tmin;
  return tmax >= fmaxf(tmin, 0.0f);
}

/* Per-ray part of the conservative slack (same statement as rr_internal.h ray_slack): the rounding error of the
 * slab test AND of the triangle test's `origin - A` grows with the ray ORIGIN, not with the box, so every box the
 * ray is tested against is additionally widened by 2^-18 of the largest |origin coordinate| (mesh-local origin for
 * the walk inside a mesh).  Without it a camera 10^4 scene extents away loses hits to culling
 * (tests/test_oracle.py::test_far_origin_hierarchy_only_culls). */
static inline float ray_slack(const ray_t* r) {
  float m = fmaxf(fmaxf(fabsf(r->origin.x), fabsf(r->origin.y)), fabsf(r->origin.z));
  return (m < 3.0e38f ? m : 0.0f) * 3.814697265625e-06f;
}
static inline int ray_box_slack(const ray_t* r, const float* bmin, const float* bmax, float slack, float* outDist) {
  float lo[3] = {bmin[0] - slack, bmin[1] - slack, bmin[2] - slack};
  float hi[3] = {bmax[0] + slack, bmax[1] + slack, bmax[2] + slack};
  return ray_box(r, lo, hi, outDist);
}

typedef struct {
  int didHit;
  float dst;
  v3 hitPoint, normal;
  int isBackface;
  int32_t prim; /* index in the uploaded triangle / sphere array */
} hit_t;

/* src/Trace.cl:276-317.  Returns 0 = no hit.  `best`/`bestPrim`: the candidate
 * is accepted only if (t, prim) orders before the current closest -- the
 * reference accepts on `hit.dst < closestHit.dst` in ITS traversal order
 * (src/Trace.cl:355); our total order (t, then uploaded index) makes the result
 * independent of the hierarchy.  strictRef != 0 reproduces the reference rule. */
static inline int ray_triangle(const ray_t* ray, const rr_triangle* tri, int cull, int32_t prim, hit_t* best,
                               int strictRef) {
  v3 A = from_f3(&tri->posA);
  v3 edge1 = vsub(from_f3(&tri->posB), A);
  v3 edge2 = vsub(from_f3(&tri->posC), A);
  v3 h = vcross(ray->direction, edge2);
  float a = vdot(edge1, h);
  if (fabsf(a) < RRO_EPSILON) return 0;
  float f = 1.0f / a;
  v3 s = vsub(ray->origin, A);
  float u = f * vdot(s, h);
  if (u < 0.0f || u > 1.0f) return 0;
  v3 q = vcross(s, edge1);
  float v = f * vdot(ray->direction, q);
  if (v < 0.0f || u + v > 1.0f) return 0;
  float t = f * vdot(edge2, q);
  if (t <= RRO_EPSILON) return 0;
  /* early distance test (moved before the normal: same accept set) */
  if (strictRef) {
    if (!(t < best->dst)) return 0;
  } else {
    if (!(t < best->dst || (t == best->dst && best->didHit && prim < best->prim))) return 0;
  }
  v3 n = vnormalize(vadd(vadd(vscale(from_f3(&tri->normalA), (1.0f - u - v)), vscale(from_f3(&tri->normalB), u)),
                         vscale(from_f3(&tri->normalC), v)));
  int back = 0;
  if (vdot(ray->direction, n) > RRO_EPSILON) {
    if (cull) return 0;
    back = 1;
    n = vneg(n);
  }
  best->didHit = 1;
  best->dst = t;
  best->hitPoint = vadd(ray->origin, vscale(ray->direction, t));
  best->normal = n;
  best->isBackface = back;
  best->prim = prim;
  return 1;
}

/* Closest hit inside one mesh with OUR hierarchy (statement of the traversal
 * the CUDA kernel performs; replaces src/Trace.cl:319-397). */
static void mesh_closest_lbvh(const rro_scene* sc, const rro_meshx* mx, const ray_t* ray, int cull, float tmax,
                              hit_t* best, rro_counters* c) {
  best->didHit = 0;
  best->dst = tmax;
  best->prim = 0x7fffffff;
  float distRoot;
  const float slack = ray_slack(ray);
  if (!sc->brute_force) {
    c->box_tests++;
    if (!ray_box_slack(ray, mx->bmin, mx->bmax, slack, &distRoot)) return;
  }
  const rro_lbvh* b = &sc->tb;
  if (mx->count <= RRO_DIRECT_MAX || sc->brute_force) {
    for (uint64_t k = 0; k < mx->count; ++k) {
      uint32_t prim = b->order[mx->first + k];
      c->tri_tests++;
      ray_triangle(ray, &sc->tris[prim], cull, (int32_t)prim, best, 0);
    }
    return;
  }
  int32_t stackN[RRO_STACK];
  float stackD[RRO_STACK];
  int sp = 0;
  int32_t cur = (int32_t)mx->first;
  for (;;) {
    int32_t L = b->left[cur], R = b->right[cur];
    float ba[6], bb[6], dA, dB;
    box_of_ref(b, L, ba);
    box_of_ref(b, R, bb);
    inflate_box(ba, mx->delta + slack);
    inflate_box(bb, mx->delta + slack);
    c->box_tests += 2;
    int hA = ray_box(ray, ba, ba + 3, &dA) && dA <= best->dst;
    int hB = ray_box(ray, bb, bb + 3, &dB) && dB <= best->dst;
    int32_t next = 0;
    int have = 0;
    if (hA && hB) {
      int32_t far;
      float dfar;
      if (dA < dB) { next = L; far = R; dfar = dB; } else { next = R; far = L; dfar = dA; }
      if (sp < RRO_STACK) { stackN[sp] = far; stackD[sp] = dfar; sp++; }
      have = 1;
    } else if (hA) { next = L; have = 1; }
    else if (hB) { next = R; have = 1; }
    for (;;) {
      if (have) {
        if (next >= 0) { cur = next; break; }
        uint32_t prim = b->order[~next];
        c->tri_tests++;
        ray_triangle(ray, &sc->tris[prim], cull, (int32_t)prim, best, 0);
        have = 0;
      }
      /* pop */
      int found = 0;
      while (sp > 0) {
        --sp;
        if (stackD[sp] <= best->dst) { next = stackN[sp]; found = 1; break; }
      }
      if (!found) return;
      have = 1;
    }
  }
}

/* Closest hit inside one mesh with the REFERENCE's node list: literal
 * restatement of src/Trace.cl:319-397 (used only to pin this file against
 * libref with identical tie behaviour). */
typedef struct { float bmin[4], bmax[4]; uint64_t index, numTriangles; } ref_gpunode;
static void mesh_closest_refbvh(const rro_scene* sc, uint64_t nodeIdx, const ray_t* ray, int cull, hit_t* best,
                                rro_counters* c) {
  const ref_gpunode* nodes = (const ref_gpunode*)sc->ref_nodes;
  best->didHit = 0;
  best->dst = INFINITY;
  best->prim = 0x7fffffff;
  int32_t stN[RRO_STACK];
  float stD[RRO_STACK];
  size_t sp = 0;
  float distRoot;
  c->box_tests++;
  if (!ray_box(ray, nodes[nodeIdx].bmin, nodes[nodeIdx].bmax, &distRoot)) return;
  stN[sp] = (int32_t)nodeIdx; stD[sp] = distRoot; sp++;
  while (sp > 0) {
    --sp;
    int32_t ni = stN[sp];
    float nd = stD[sp];
    const ref_gpunode* node = &nodes[ni];
    if (node->numTriangles == 0 && node->index == 0) continue;
    if (nd >= best->dst) continue;
    if (node->numTriangles > 0) {
      for (uint64_t i = 0; i < node->numTriangles; ++i) {
        uint64_t ti = node->index + i;
        c->tri_tests++;
        ray_triangle(ray, &sc->tris[ti], cull, (int32_t)ti, best, 1);
      }
    } else {
      const ref_gpunode* A = &nodes[node->index];
      const ref_gpunode* B = &nodes[node->index + 1];
      float dA, dB;
      c->box_tests += 2;
      int hA = ray_box(ray, A->bmin, A->bmax, &dA);
      int hB = ray_box(ray, B->bmin, B->bmax, &dB);
      if (!hB && !hA) continue;
      if (!hB && hA) { if (dA < best->dst) { stN[sp] = (int32_t)node->index; stD[sp] = dA; sp++; } continue; }
      if (hB && !hA) { if (dB < best->dst) { stN[sp] = (int32_t)node->index + 1; stD[sp] = dB; sp++; } continue; }
      if (dA < dB) {
        stN[sp] = (int32_t)node->index + 1; stD[sp] = dB; sp++;
        stN[sp] = (int32_t)node->index; stD[sp] = dA; sp++;
      } else {
        stN[sp] = (int32_t)node->index; stD[sp] = dA; sp++;
        stN[sp] = (int32_t)node->index + 1; stD[sp] = dB; sp++;
      }
    }
  }
}

/* EXTENSION (no reference counterpart): ray/sphere.  Nearest root with
 * t > EPSILON; the far root means the origin is inside (backface). */
static inline int ray_sphere(const ray_t* ray, const rr_sphere* sp, int32_t prim, hit_t* best) {
  v3 c = from_f3(&sp->center);
  float r = sp->radius;
  v3 oc = vsub(ray->origin, c);
  float b = vdot(oc, ray->direction);
  float cc = vdot(oc, oc) - r * r;
  float disc = b * b - cc;
  if (!(disc >= 0.0f)) return 0;
  float sq = sqrtf(disc);
  float t = -b - sq;
  int back = 0;
  if (t <= RRO_EPSILON) { t = -b + sq; back = 1; }
  if (t <= RRO_EPSILON) return 0;
  if (!(t < best->dst || (t == best->dst && best->didHit && prim < best->prim))) return 0;
  int ty = sp->material.type;
  int cull = (ty != RR_MATERIAL_GLASSY && ty != RR_MATERIAL_INVISIBLE && ty != RR_MATERIAL_ONESIDED);
  if (back && cull) return 0;
  v3 hp = vadd(ray->origin, vscale(ray->direction, t));
  v3 n = vdivs(vsub(hp, c), r);
  if (back) n = vneg(n);
  best->didHit = 1; best->dst = t; best->hitPoint = hp; best->normal = n; best->isBackface = back; best->prim = prim;
  return 1;
}

static void spheres_closest(const rro_scene* sc, const ray_t* ray, float tmax, hit_t* best, rro_counters* c) {
  best->didHit = 0;
  best->dst = tmax;
  best->prim = 0x7fffffff;
  if (sc->n_spheres == 0) return;
  float distRoot;
  const float slack = ray_slack(ray);
  if (!sc->brute_force) {
    c->box_tests++;
    if (!ray_box_slack(ray, sc->sph_bmin, sc->sph_bmax, slack, &distRoot)) return;
  }
  const rro_lbvh* b = &sc->sb;
  if (sc->n_spheres <= RRO_DIRECT_MAX || sc->brute_force) {
    for (uint64_t k = 0; k < sc->n_spheres; ++k) {
      uint32_t prim = b->order[k];
      c->sphere_tests++;
      ray_sphere(ray, &sc->spheres[prim], (int32_t)prim, best);
    }
    return;
  }
  int32_t stackN[RRO_STACK];
  float stackD[RRO_STACK];
  int sp = 0;
  int32_t cur = 0;
  for (;;) {
    int32_t L = b->left[cur], R = b->right[cur];
    float ba[6], bb[6], dA, dB;
    box_of_ref(b, L, ba);
    box_of_ref(b, R, bb);
    inflate_box(ba, sc->sph_delta + slack);
    inflate_box(bb, sc->sph_delta + slack);
    c->box_tests += 2;
    int hA = ray_box(ray, ba, ba + 3, &dA) && dA <= best->dst;
    int hB = ray_box(ray, bb, bb + 3, &dB) && dB <= best->dst;
    int32_t next = 0;
    int have = 0;
    if (hA && hB) {
      int32_t far;
      float dfar;
      if (dA < dB) { next = L; far = R; dfar = dB; } else { next = R; far = L; dfar = dA; }
      if (sp < RRO_STACK) { stackN[sp] = far; stackD[sp] = dfar; sp++; }
      have = 1;
    } else if (hA) { next = L; have = 1; }
    else if (hB) { next = R; have = 1; }
    for (;;) {
      if (have) {
        if (next >= 0) { cur = next; break; }
        uint32_t prim = b->order[~next];
        c->sphere_tests++;
        ray_sphere(ray, &sc->spheres[prim], (int32_t)prim, best);
        have = 0;
      }
      int found = 0;
      while (sp > 0) {
        --sp;
        if (stackD[sp] <= best->dst) { next = stackN[sp]; found = 1; break; }
      }
      if (!found) return;
      have = 1;
    }
  }
}

typedef struct {
  hit_t h;
  int32_t mesh; /* mesh index; n_meshes for a sphere; -1 miss */
  const rr_material* material;
} scene_hit;

/* src/Trace.cl:434-485 (+ the sphere extension after the mesh loop). */
static void scene_closest(const rro_scene* sc, const ray_t* worldRay, scene_hit* out, rro_counters* c) {
  out->h.didHit = 0;
  out->h.dst = INFINITY;
  out->mesh = -1;
  out->material = NULL;
  c->rays++;
  for (uint64_t m = 0; m < sc->n_meshes; ++m) {
    const rr_mesh* info = &sc->meshes[m];
    const rro_meshx* mx = &sc->mx[m];
    if (info->scale <= RRO_EPSILON) continue;
    /* WorldToLocalRay, src/Trace.cl:118-137 */
    v3 pos = from_f3(&info->pos);
    v3 lo = mul_mat_vec(&mx->Rinv, vsub(worldRay->origin, pos));
    v3 ld = mul_mat_vec(&mx->Rinv, worldRay->direction);
    if (fabsf(info->scale) > RRO_EPSILON) {
      lo = vdivs(lo, info->scale);
      ld = vdivs(ld, info->scale);
    }
    ld = vnormalize(ld);
    ray_t lr;
    lr.invDir = V(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
    lr.origin = lo;
    lr.direction = ld;
    hit_t lh;
    if (sc->ref_nodes) mesh_closest_refbvh(sc, info->nodeIdx, &lr, mx->cull, &lh, c);
    else mesh_closest_lbvh(sc, mx, &lr, mx->cull, INFINITY, &lh, c);
    if (!lh.didHit) continue;
    if (info->material.type == RR_MATERIAL_ONESIDED && lh.isBackface) continue;
    /* LocalToWorldHit, src/Trace.cl:139-156 */
    hit_t wh = lh;
    wh.hitPoint = vadd(mul_mat_vec(&mx->R, vscale(lh.hitPoint, info->scale)), pos);
    wh.normal = vnormalize(mul_mat_vec(&mx->R, lh.normal));
    wh.dst = vlength(vsub(wh.hitPoint, worldRay->origin));
    if (wh.dst < out->h.dst) {
      out->h = wh;
      out->mesh = (int32_t)m;
      out->material = &info->material;
    }
  }
  if (sc->n_spheres) {
    ray_t wr = *worldRay;
    wr.invDir = V(1.0f / wr.direction.x, 1.0f / wr.direction.y, 1.0f / wr.direction.z);
    hit_t sh;
    spheres_closest(sc, &wr, INFINITY, &sh, c);
    if (sh.didHit) {
      const rr_sphere* sp = &sc->spheres[sh.prim];
      if (!(sp->material.type == RR_MATERIAL_ONESIDED && sh.isBackface)) {
        if (sh.dst < out->h.dst) {
          out->h = sh;
          out->mesh = (int32_t)sc->n_meshes;
          out->material = &sp->material;
        }
      }
    }
  }
}

/* ------------------------------------------------------------ shading ---- */
/* src/Trace.cl:84 */
static inline v3 lerp3(v3 a, v3 b, float t) { return vadd(vscale(a, (1.0f - t)), vscale(b, t)); }
/* src/Trace.cl:234-236 */
static inline v3 reflect3(v3 inDir, v3 normal) {
  float k = 2 * vdot(inDir, normal);
  return vsub(inDir, vscale_l(k, normal));
}
/* src/Trace.cl:219-232 */
static inline v3 refract3(v3 inDir, v3 normal, float iorA, float iorB) {
  float refractRatio = iorA / iorB;
  float cosAngleIn = -vdot(inDir, normal);
  float sinSqr = refractRatio * refractRatio * (1 - cosAngleIn * cosAngleIn);
  if (sinSqr > 1) return V(0.0f, 0.0f, 0.0f);
  return vadd(vscale_l(refractRatio, inDir), vscale_l(refractRatio * cosAngleIn - sqrtf(1 - sinSqr), normal));
}
/* src/Trace.cl:401-432 (both denominators use the same expression, as in the reference) */
static inline float reflectance(v3 inDir, v3 normal, float iorA, float iorB) {
  float refractRatio = iorA / iorB;
  float cosAngleIn = -vdot(inDir, normal);
  if (cosAngleIn <= 0) return 1;
  float sinSqr = refractRatio * refractRatio * (1 - cosAngleIn * cosAngleIn);
  if (sinSqr >= 1) return 1;
  float cosR = sqrtf(1 - sinSqr);
  float denPerp = iorA * cosAngleIn + iorB * cosR;
  float denPar = iorA * cosAngleIn + iorB * cosR;
  if (fminf(denPerp, denPar) < RRO_EPSILON) return 1;
  float rPerp = (iorA * cosAngleIn - iorB * cosR) / denPerp;
  rPerp *= rPerp;
  float rPar = (iorB * cosAngleIn - iorA * cosR) / denPar;
  rPar *= rPar;
  return (rPerp + rPar) / 2;
}
static inline float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
/* (int)floor(x): contract = saturating conversion, NaN -> 0 (what cvt.rzi.s32.f32 does) */
static inline int32_t f2i_sat(float x) {
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (int32_t)0x80000000;
  return (int32_t)x;
}

/* src/Trace.cl:487-594 */
static v3 trace_path(const rro_scene* sc, ray_t ray, uint32_t* rng, uint32_t maxBounce, rro_counters* c) {
  v3 incoming = V(0.0f, 0.0f, 0.0f);
  v3 throughput = V(1.0f, 1.0f, 1.0f);
  uint32_t bounce = 0;
  uint32_t passes = 0;
  while (bounce < maxBounce) {
    scene_hit sh;
    scene_closest(sc, &ray, &sh, c);
    if (!sh.h.didHit) break;
    rr_material mat = *sh.material;
    hit_t* hit = &sh.h;
    if (mat.type == RR_MATERIAL_INVISIBLE) {
      /* GUARD (deviation): the reference `continue`s without counting a bounce (src/Trace.cl:502-506);
       * when hitPoint + dir*1e-6 rounds back to hitPoint it loops forever (a GPU hang).  Both sides of
       * this repo end the path after RR_MAX_INVISIBLE_PASSES pass-throughs. */
      if (++passes > RRO_MAX_INVISIBLE_PASSES) break;
      ray.origin = vadd(hit->hitPoint, vscale(ray.direction, RRO_EPSILON));
      continue;
    }
    v3 color = from_f3(&mat.color);
    v3 emissionColor = from_f3(&mat.emissionColor);
    float emissionStrength = mat.emissionStrength;
    if (mat.type == RR_MATERIAL_CHECKER) {
      float checkerSize = mat.emissionStrength;
      int32_t xi = f2i_sat(floorf(hit->hitPoint.x / checkerSize));
      int32_t zi = f2i_sat(floorf(hit->hitPoint.z / checkerSize));
      int isEven = (((uint32_t)xi + (uint32_t)zi) & 1u) == 0;
      color = isEven ? color : emissionColor;
      emissionStrength = 0.0f;
      int isSpec = mat.specularProbability >= random_value(rng);
      v3 diffuseDir = vnormalize(vadd(hit->normal, random_direction(rng)));
      v3 specularDir = reflect3(ray.direction, hit->normal);
      ray.direction = vnormalize(lerp3(diffuseDir, specularDir, mat.reflectiveness * (float)isSpec));
    }
    if (mat.type == RR_MATERIAL_GLASSY) {
      float iorCur = hit->isBackface ? mat.ior : RRO_IOR_AIR;
      float iorNext = hit->isBackface ? RRO_IOR_AIR : mat.ior;
      v3 reflectDir = reflect3(ray.direction, hit->normal);
      v3 refractDir = refract3(ray.direction, hit->normal, iorCur, iorNext);
      float reflectWeight = reflectance(ray.direction, hit->normal, iorCur, iorNext);
      float refractWeight = 1.0f - reflectWeight;
      int willReflect = rand01(rng) < reflectWeight;
      ray.direction = willReflect ? reflectDir : refractDir;
      /* origin of :553-554 is overwritten at :579 -- no observable effect */
      float w = willReflect ? reflectWeight : refractWeight;
      throughput = vscale(throughput, w);
    }
    if (mat.type == RR_MATERIAL_SOLID) {
      int isSpec = mat.specularProbability >= random_value(rng);
      v3 diffuseDir = vnormalize(vadd(hit->normal, random_direction(rng)));
      v3 specularDir = reflect3(ray.direction, hit->normal);
      ray.direction = vnormalize(lerp3(diffuseDir, specularDir, mat.reflectiveness * (float)isSpec));
    }
    incoming = vadd(incoming, vmul(throughput, vscale(emissionColor, emissionStrength)));
    ray.origin = vadd(hit->hitPoint, vscale(ray.direction, RRO_EPSILON));
    throughput = vmul(throughput, color);
    float p = fmaxf(throughput.x, fmaxf(throughput.y, throughput.z));
    if (bounce > 3) {
      float q = fmaxf(0.05f, 1.0f - p);
      if (rand01(rng) < q) break;
      throughput = vdivs(throughput, (1.0f - q));
    }
    bounce++;
  }
  return incoming;
}

/* src/Trace.cl:596-621 */
static ray_t make_ray(const rr_camera* cam, float u, float v) {
  float ndc0 = u * 2.0f - 1.0f;
  float ndc1 = v * 2.0f - 1.0f;
  ndc0 *= cam->aspectRatio;
  float scale = rr_tanf_ref((cam->fov * 0.5f) * 0.017453292519943295f);
  v3 dc = vnormalize(V(ndc0 * scale, ndc1 * scale, 1.0f));
  float cx = rr_cosf_ref(cam->pitch), sx = rr_sinf_ref(cam->pitch);
  float cy = rr_cosf_ref(cam->yaw), sy = rr_sinf_ref(cam->yaw);
  float cz = rr_cosf_ref(cam->roll), sz = rr_sinf_ref(cam->roll);
  v3 r0 = V(cy * cz, cz * sy * sx - cx * sz, sx * sz + cx * cz * sy);
  v3 r1 = V(cy * sz, cx * cz + sx * sy * sz, cx * sy * sz - cz * sx);
  v3 r2 = V(-sy, cy * sx, cx * cy);
  v3 dw = vnormalize(V(vdot(r0, dc), vdot(r1, dc), vdot(r2, dc)));
  ray_t r;
  r.origin = from_f3(&cam->position);
  r.direction = dw;
  r.invDir = V(0.0f, 0.0f, 0.0f);
  return r;
}

/* src/Trace.cl:634-635 */
static inline void pixel_uv(uint32_t x, uint32_t y, uint32_t W, uint32_t H, float* u, float* v) {
  *u = (float)x / (float)W;
  *v = (float)(1.0f - (float)y / (float)H);
}

/* ------------------------------------------------------------ drivers ---- */
typedef struct {
  const rro_scene* sc;
  const rr_camera* cam;
  uint32_t W, H, spp, bounces;
  int32_t frameIndex;
  uint8_t* rgba;
  float* radiance;
  int32_t *mesh_out, *prim_out;
  float* dst_out;
  int mode; /* 0 render, 1 primary */
  volatile int64_t* next_row;
  rro_counters counters;
} job_t;

/* src/Trace.cl:623-653 */
static void render_pixel(job_t* j, uint32_t x, uint32_t y, rro_counters* c) {
  uint32_t pixelIndex = y * j->W + x;
  uint32_t rng = make_seed(pixelIndex, j->frameIndex, 0);
  float u, v;
  pixel_uv(x, y, j->W, j->H, &u, &v);
  ray_t ray = make_ray(j->cam, u, v);
  v3 accum = V(0.0f, 0.0f, 0.0f);
  for (uint32_t s = 0; s < j->spp; ++s) accum = vadd(accum, trace_path(j->sc, ray, &rng, j->bounces, c));
  v3 col = vdivs(accum, (float)j->spp);
  if (j->radiance) {
    float* o = j->radiance + (size_t)pixelIndex * 3;
    o[0] = col.x; o[1] = col.y; o[2] = col.z;
  }
  float ch[3] = {col.x, col.y, col.z};
  uint8_t* px = j->rgba + (size_t)pixelIndex * 4;
  for (int k = 0; k < 3; ++k) {
    float cc = fminf(fmaxf(ch[k], 0.0f), 1.0f);
    cc = rr_powrf_ref(cc, 1.0f / 2.2f);
    px[k] = (uint8_t)(cc * 255.0f);
  }
  px[3] = 255; /* host side forces alpha, src/image.hpp:271 */
}

static void primary_pixel(job_t* j, uint32_t x, uint32_t y, rro_counters* c) {
  float u, v;
  pixel_uv(x, y, j->W, j->H, &u, &v);
  ray_t ray = make_ray(j->cam, u, v);
  scene_hit sh;
  scene_closest(j->sc, &ray, &sh, c);
  size_t p = (size_t)y * j->W + x;
  if (j->mesh_out) j->mesh_out[p] = sh.h.didHit ? sh.mesh : -1;
  if (j->prim_out) j->prim_out[p] = sh.h.didHit ? sh.h.prim : -1;
  if (j->dst_out) j->dst_out[p] = sh.h.didHit ? sh.h.dst : 0.0f;
}

static void* worker(void* arg) {
  job_t* shared = (job_t*)arg;
  rro_counters c;
  memset(&c, 0, sizeof(c));
  for (;;) {
    int64_t y = __sync_fetch_and_add(shared->next_row, 1);
    if (y >= (int64_t)shared->H) break;
    for (uint32_t x = 0; x < shared->W; ++x) {
      if (shared->mode == 0) render_pixel(shared, x, (uint32_t)y, &c);
      else primary_pixel(shared, x, (uint32_t)y, &c);
    }
  }
  __sync_fetch_and_add(&shared->counters.rays, c.rays);
  __sync_fetch_and_add(&shared->counters.box_tests, c.box_tests);
  __sync_fetch_and_add(&shared->counters.tri_tests, c.tri_tests);
  __sync_fetch_and_add(&shared->counters.sphere_tests, c.sphere_tests);
  return NULL;
}

static void run_job(job_t* j, int nthreads) {
  volatile int64_t next = 0;
  j->next_row = &next;
  memset(&j->counters, 0, sizeof(j->counters));
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int i = 1; i < nthreads; ++i) pthread_create(&th[i], NULL, worker, j);
  worker(j);
  for (int i = 1; i < nthreads; ++i) pthread_join(th[i], NULL);
}

/* stats4: rays, box tests, triangle tests, sphere tests (may be NULL) */
int rro_render(const rro_scene* sc, const rr_camera* cam, uint32_t W, uint32_t H, uint32_t spp, uint32_t bounces,
               int32_t frameIndex, uint8_t* rgba, float* radiance, uint64_t* stats4, int nthreads) {
  job_t j;
  memset(&j, 0, sizeof(j));
  j.sc = sc; j.cam = cam; j.W = W; j.H = H; j.spp = spp; j.bounces = bounces; j.frameIndex = frameIndex;
  j.rgba = rgba; j.radiance = radiance; j.mode = 0;
  run_job(&j, nthreads);
  if (stats4) { stats4[0] = j.counters.rays; stats4[1] = j.counters.box_tests; stats4[2] = j.counters.tri_tests; stats4[3] = j.counters.sphere_tests; }
  return 0;
}

int rro_primary(const rro_scene* sc, const rr_camera* cam, uint32_t W, uint32_t H, int32_t* mesh_out,
                int32_t* prim_out, float* dst_out, int nthreads) {
  job_t j;
  memset(&j, 0, sizeof(j));
  j.sc = sc; j.cam = cam; j.W = W; j.H = H; j.mode = 1;
  j.mesh_out = mesh_out; j.prim_out = prim_out; j.dst_out = dst_out;
  run_job(&j, nthreads);
  return 0;
}

/* ---- known-answer hooks ------------------------------------------------- */
uint32_t rro_make_seed(uint32_t pixelIndex, int32_t frameIndex, uint32_t rayIdx) { return make_seed(pixelIndex, frameIndex, rayIdx); }
float rro_random_value(uint32_t* state) { return random_value(state); }
float rro_rand01(uint32_t* state) { return rand01(state); }
void rro_random_direction(uint32_t* state, float* out3) {
  v3 d = random_direction(state);
  out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
}
/* numerics contract, vectorised for the tests (test_numerics_contract_accuracy, test_numerics_contract_is_bit_identical_on_device): fn 0 cos, 1 sin, 2 log, 3 exp2, 4 powr(x, y[i]), 5 tan */
void rro_math(int fn, const float* x, const float* y, float* out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    switch (fn) {
      case 0: out[i] = rr_cosf_ref(x[i]); break;
      case 1: out[i] = rr_sinf_ref(x[i]); break;
      case 2: out[i] = rr_logf_ref(x[i]); break;
      case 3: out[i] = rr_exp2f_ref(x[i]); break;
      case 4: out[i] = rr_powrf_ref(x[i], y[i]); break;
      default: out[i] = rr_tanf_ref(x[i]); break;
    }
  }
}
void rro_make_ray(const rr_camera* cam, uint32_t x, uint32_t y, uint32_t W, uint32_t H, float* dir3) {
  float u, v;
  pixel_uv(x, y, W, H, &u, &v);
  ray_t r = make_ray(cam, u, v);
  dir3[0] = r.direction.x; dir3[1] = r.direction.y; dir3[2] = r.direction.z;
}
