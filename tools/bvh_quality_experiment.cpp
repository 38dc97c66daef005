// tools/bvh_quality_experiment.cpp -- CPU experiment behind DESIGN.md section 5.1 (round 2): how much would a better
// hierarchy than the plain LBVH save on the bench scene?  NOT product code and not part of any test.
//
//   python tools/bvh_quality_experiment.py            # dumps the C4 mesh, compiles this file, runs it
//
// Builds, over the 1 012 000 triangles of the C4 mesh: the Morton / Karras LBVH (as csrc/rr_lbvh.cu), the same with
// subtrees of <= 2 / 4 / 8 triangles collapsed into leaves, with 1-3 passes of bottom-up tree rotations (Kensler 2008),
// a 16-bin top-down SAH build as the quality ceiling, and the hybrid that was then built on the GPU (csrc/rr_lbvh.cu
// k_find_clusters / lbvh_sah_top): Karras subtrees of <= T primitives kept as clusters, binned SAH over the clusters above them.  Every tree is collapsed 4-wide by surface area (as
// k_pack_wide) and walked front to back by ~186 k path-like rays (camera rays of the C4 camera, then diffuse bounces
// inside the Cornell walls); prints wide-node visits, box tests, leaf visits and triangle tests per ray.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <vector>
using namespace std;

struct V3 { float x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline V3 norm(V3 a) { float l = sqrtf(dot(a, a)); return a * (1.0f / l); }

struct Tri { V3 a, b, c; };
struct Box {
  float lo[3], hi[3];
  void init() { for (int k = 0; k < 3; ++k) { lo[k] = 1e30f; hi[k] = -1e30f; } }
  void grow(const Box& o) { for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], o.lo[k]); hi[k] = max(hi[k], o.hi[k]); } }
  float area() const { float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2]; return dx * dy + dy * dz + dz * dx; }
};
static Box uni(const Box& a, const Box& b) { Box r = a; r.grow(b); return r; }

vector<Tri> tris, others;
vector<Box> pbox;

// binary tree: child ref >= 0 inner node index; < 0: leaf ~leafIndex
struct Node { Box b; int l, r; int count; };
struct Leaf { vector<int> prims; Box b; };
struct Tree { vector<Node> nodes; vector<Leaf> leaves; int root; };

static Box refbox(const Tree& t, int ref) { return ref >= 0 ? t.nodes[ref].b : t.leaves[~ref].b; }
static int refcount(const Tree& t, int ref) { return ref >= 0 ? t.nodes[ref].count : (int)t.leaves[~ref].prims.size(); }

// ---------- LBVH
static inline uint64_t expand21(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
static int delta(const vector<uint64_t>& k, int i, int j) {
  int n = (int)k.size();
  if (j < 0 || j >= n) return -1;
  if (k[i] == k[j]) return 64 + __builtin_clz((uint32_t)i ^ (uint32_t)j);
  return __builtin_clzll(k[i] ^ k[j]);
}
Tree build_lbvh() {
  int n = (int)tris.size();
  Box sb; sb.init();
  for (auto& b : pbox) sb.grow(b);
  vector<pair<uint64_t, int>> ki(n);
  for (int i = 0; i < n; ++i) {
    uint32_t g[3];
    for (int a = 0; a < 3; ++a) {
      float c = (pbox[i].lo[a] + pbox[i].hi[a]) * 0.5f;
      float ext = sb.hi[a] - sb.lo[a];
      float q = ext > 0 ? (c - sb.lo[a]) / ext * 2097152.0f : 0.0f;
      g[a] = q <= 0 ? 0u : q >= 2097151.0f ? 2097151u : (uint32_t)q;
    }
    ki[i] = {(expand21(g[0]) << 2) | (expand21(g[1]) << 1) | expand21(g[2]), i};
  }
  sort(ki.begin(), ki.end());
  vector<uint64_t> keys(n);
  for (int i = 0; i < n; ++i) keys[i] = ki[i].first;
  Tree t;
  t.nodes.resize(n - 1);
  t.leaves.resize(n);
  for (int i = 0; i < n; ++i) { t.leaves[i].prims = {ki[i].second}; t.leaves[i].b = pbox[ki[i].second]; }
  for (int i = 0; i < n - 1; ++i) {
    int d = (delta(keys, i, i + 1) - delta(keys, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, i, i - d);
    int lmax = 2;
    while (delta(keys, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int tt = lmax / 2; tt >= 1; tt /= 2) if (delta(keys, i, i + (l + tt) * d) > dmin) l += tt;
    int j = i + l * d;
    int dn = delta(keys, i, j);
    int s = 0;
    for (int tt = (l + 1) >> 1;; tt = (tt + 1) >> 1) { if (delta(keys, i, i + (s + tt) * d) > dn) s += tt; if (tt <= 1) break; }
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    int lo = min(i, j), hi = max(i, j);
    t.nodes[i].l = lo == gamma ? ~gamma : gamma;
    t.nodes[i].r = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
  }
  t.root = 0;
  // refit (post-order)
  function<void(int)> fit = [&](int i) {
    Node& nd = t.nodes[i];
    if (nd.l >= 0) fit(nd.l);
    if (nd.r >= 0) fit(nd.r);
    nd.b = uni(refbox(t, nd.l), refbox(t, nd.r));
    nd.count = refcount(t, nd.l) + refcount(t, nd.r);
  };
  fit(0);
  return t;
}

// ---------- binned SAH (quality reference)
Tree build_sah(int maxLeaf) {
  int n = (int)tris.size();
  Tree t;
  vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  function<int(int, int)> rec = [&](int lo, int hi) -> int {
    Box b; b.init();
    Box cb; cb.init();
    for (int i = lo; i < hi; ++i) {
      b.grow(pbox[idx[i]]);
      Box c; for (int k = 0; k < 3; ++k) c.lo[k] = c.hi[k] = 0.5f * (pbox[idx[i]].lo[k] + pbox[idx[i]].hi[k]);
      cb.grow(c);
    }
    int cnt = hi - lo;
    auto mkleaf = [&]() { Leaf L; L.b = b; for (int i = lo; i < hi; ++i) L.prims.push_back(idx[i]); t.leaves.push_back(L); return ~(int)(t.leaves.size() - 1); };
    if (cnt <= 1) return mkleaf();
    const int NB = 16;
    float bestCost = 1e30f; int bestAxis = -1, bestBin = -1;
    for (int a = 0; a < 3; ++a) {
      float ext = cb.hi[a] - cb.lo[a];
      if (!(ext > 0)) continue;
      Box bb[NB]; int bc[NB];
      for (int k = 0; k < NB; ++k) { bb[k].init(); bc[k] = 0; }
      for (int i = lo; i < hi; ++i) {
        float c = 0.5f * (pbox[idx[i]].lo[a] + pbox[idx[i]].hi[a]);
        int k = min(NB - 1, (int)((c - cb.lo[a]) / ext * NB));
        bb[k].grow(pbox[idx[i]]); bc[k]++;
      }
      float la[NB], ra[NB]; int lc[NB], rc[NB];
      Box acc; acc.init(); int c = 0;
      for (int k = 0; k < NB; ++k) { acc.grow(bb[k]); c += bc[k]; la[k] = c ? acc.area() : 0; lc[k] = c; }
      acc.init(); c = 0;
      for (int k = NB - 1; k >= 0; --k) { acc.grow(bb[k]); c += bc[k]; ra[k] = c ? acc.area() : 0; rc[k] = c; }
      for (int k = 0; k < NB - 1; ++k) {
        if (!lc[k] || !rc[k + 1]) continue;
        float cost = la[k] * lc[k] + ra[k + 1] * rc[k + 1];
        if (cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = k; }
      }
    }
    int mid;
    if (bestAxis < 0) {
      if (cnt <= maxLeaf) return mkleaf();
      mid = (lo + hi) / 2;
    } else {
      float leafCost = b.area() * cnt;
      if (cnt <= maxLeaf && leafCost <= bestCost + 1.0f * b.area()) return mkleaf();
      float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
      mid = (int)(partition(idx.begin() + lo, idx.begin() + hi, [&](int p) {
        float c = 0.5f * (pbox[p].lo[bestAxis] + pbox[p].hi[bestAxis]);
        int k = min(NB - 1, (int)((c - cb.lo[bestAxis]) / ext * NB));
        return k <= bestBin; }) - idx.begin());
      if (mid == lo || mid == hi) mid = (lo + hi) / 2;
    }
    int me = (int)t.nodes.size();
    t.nodes.push_back(Node());
    int l = rec(lo, mid), r = rec(mid, hi);
    t.nodes[me].l = l; t.nodes[me].r = r; t.nodes[me].b = b; t.nodes[me].count = cnt;
    return me;
  };
  t.root = rec(0, n);
  return t;
}


// ---------- hybrid: LBVH subtrees of <= T prims as clusters, binned SAH (over cluster boxes) above them
Tree build_hybrid(const Tree& lb, int T) {
  Tree t = lb;  // copy nodes/leaves; we add new top nodes
  vector<int> clusters;  // refs (inner node or leaf) that are cluster roots
  function<void(int)> cut = [&](int ref) {
    if (ref < 0 || t.nodes[ref].count <= T) { clusters.push_back(ref); return; }
    cut(t.nodes[ref].l); cut(t.nodes[ref].r);
  };
  cut(t.root);
  int K = (int)clusters.size();
  vector<int> idx(K);
  for (int i = 0; i < K; ++i) idx[i] = i;
  auto cbox = [&](int c) { return refbox(t, clusters[c]); };
  auto ccnt = [&](int c) { return refcount(t, clusters[c]); };
  function<int(int, int)> rec = [&](int lo, int hi) -> int {
    if (hi - lo == 1) return clusters[idx[lo]];
    Box b; b.init(); Box cb; cb.init();
    for (int i = lo; i < hi; ++i) { Box x = cbox(idx[i]); b.grow(x); Box c; for (int k = 0; k < 3; ++k) c.lo[k] = c.hi[k] = 0.5f * (x.lo[k] + x.hi[k]); cb.grow(c); }
    const int NB = 16;
    float bestCost = 1e30f; int bestAxis = -1, bestBin = -1;
    for (int a = 0; a < 3; ++a) {
      float ext = cb.hi[a] - cb.lo[a];
      if (!(ext > 0)) continue;
      Box bb[NB]; int bc[NB];
      for (int k = 0; k < NB; ++k) { bb[k].init(); bc[k] = 0; }
      for (int i = lo; i < hi; ++i) { Box x = cbox(idx[i]); float c = 0.5f * (x.lo[a] + x.hi[a]); int k = min(NB - 1, (int)((c - cb.lo[a]) / ext * NB)); bb[k].grow(x); bc[k] += ccnt(idx[i]); }
      float la[NB], ra[NB]; int lc[NB], rc[NB];
      Box acc; acc.init(); int c = 0;
      for (int k = 0; k < NB; ++k) { acc.grow(bb[k]); c += bc[k]; la[k] = c ? acc.area() : 0; lc[k] = c; }
      acc.init(); c = 0;
      for (int k = NB - 1; k >= 0; --k) { acc.grow(bb[k]); c += bc[k]; ra[k] = c ? acc.area() : 0; rc[k] = c; }
      for (int k = 0; k < NB - 1; ++k) { if (!lc[k] || !rc[k + 1]) continue; float cost = la[k] * lc[k] + ra[k + 1] * rc[k + 1]; if (cost < bestCost) { bestCost = cost; bestAxis = a; bestBin = k; } }
    }
    int mid;
    if (bestAxis < 0) mid = (lo + hi) / 2;
    else {
      float ext = cb.hi[bestAxis] - cb.lo[bestAxis];
      mid = (int)(partition(idx.begin() + lo, idx.begin() + hi, [&](int c) { Box x = cbox(c); float cc = 0.5f * (x.lo[bestAxis] + x.hi[bestAxis]); int k = min(NB - 1, (int)((cc - cb.lo[bestAxis]) / ext * NB)); return k <= bestBin; }) - idx.begin());
      if (mid == lo || mid == hi) mid = (lo + hi) / 2;
    }
    int l = rec(lo, mid), r = rec(mid, hi);
    Node nd; nd.l = l; nd.r = r; nd.b = b; nd.count = refcount(t, l) + refcount(t, r);
    t.nodes.push_back(nd);
    return (int)t.nodes.size() - 1;
  };
  t.root = rec(0, K);
  printf("   hybrid T=%d: %d clusters\n", T, K);
  return t;
}

// ---------- rotations (post-order, deterministic)
int rotate_pass(Tree& t) {
  int applied = 0;
  function<void(int)> rec = [&](int i) {
    Node& P = t.nodes[i];
    if (P.l >= 0) rec(P.l);
    if (P.r >= 0) rec(P.r);
    // options: swap L with R.l / R.r ; swap R with L.l / L.r
    float best = 0.0f; int opt = -1;
    Box bl = refbox(t, P.l), br = refbox(t, P.r);
    if (P.r >= 0) {
      Node& R = t.nodes[P.r];
      float cur = br.area();
      float a0 = uni(bl, refbox(t, R.r)).area();  // L <-> R.l : R' = (L, R.r)
      float a1 = uni(refbox(t, R.l), bl).area();  // L <-> R.r : R' = (R.l, L)
      if (cur - a0 > best) { best = cur - a0; opt = 0; }
      if (cur - a1 > best) { best = cur - a1; opt = 1; }
    }
    if (P.l >= 0) {
      Node& L = t.nodes[P.l];
      float cur = bl.area();
      float a2 = uni(br, refbox(t, L.r)).area();  // R <-> L.l : L' = (R, L.r)
      float a3 = uni(refbox(t, L.l), br).area();  // R <-> L.r : L' = (L.l, R)
      if (cur - a2 > best) { best = cur - a2; opt = 2; }
      if (cur - a3 > best) { best = cur - a3; opt = 3; }
    }
    if (opt < 0) return;
    applied++;
    if (opt == 0) { Node& R = t.nodes[P.r]; swap(P.l, R.l); R.b = uni(refbox(t, R.l), refbox(t, R.r)); R.count = refcount(t, R.l) + refcount(t, R.r); }
    if (opt == 1) { Node& R = t.nodes[P.r]; swap(P.l, R.r); R.b = uni(refbox(t, R.l), refbox(t, R.r)); R.count = refcount(t, R.l) + refcount(t, R.r); }
    if (opt == 2) { Node& L = t.nodes[P.l]; swap(P.r, L.l); L.b = uni(refbox(t, L.l), refbox(t, L.r)); L.count = refcount(t, L.l) + refcount(t, L.r); }
    if (opt == 3) { Node& L = t.nodes[P.l]; swap(P.r, L.r); L.b = uni(refbox(t, L.l), refbox(t, L.r)); L.count = refcount(t, L.l) + refcount(t, L.r); }
  };
  rec(t.root);
  return applied;
}

// ---------- leaf collapse: subtree with count <= K becomes one leaf (mode 0: always, mode 1: SAH test)
void collapse_leaves(Tree& t, int K, int mode, float Ct) {
  // cost of subtree: SAH with Cb per node box-pair... computed bottom-up
  function<float(int, vector<int>&)> rec = [&](int ref, vector<int>& prims) -> float {
    if (ref < 0) { for (int p : t.leaves[~ref].prims) prims.push_back(p); return t.leaves[~ref].b.area() * Ct * t.leaves[~ref].prims.size(); }
    Node& N = t.nodes[ref];
    vector<int> pl, pr;
    float cl = rec(N.l, pl), cr = rec(N.r, pr);
    // N.l / N.r may have been replaced below
    prims = pl; prims.insert(prims.end(), pr.begin(), pr.end());
    float sub = N.b.area() * 1.0f + cl + cr;  // traversal step cost 1 per node
    return sub;
  };
  // simpler: do it top-down: at node with count<=K collapse (mode 0)
  function<int(int)> td = [&](int ref) -> int {
    if (ref < 0) return ref;
    Node& N = t.nodes[ref];
    if (N.count <= K) {
      vector<int> prims;
      function<void(int)> gather = [&](int r) { if (r < 0) { for (int p : t.leaves[~r].prims) prims.push_back(p); } else { gather(t.nodes[r].l); gather(t.nodes[r].r); } };
      gather(ref);
      Leaf L; L.prims = prims; L.b = N.b;
      t.leaves.push_back(L);
      return ~(int)(t.leaves.size() - 1);
    }
    int l = td(N.l), r = td(N.r);
    t.nodes[ref].l = l; t.nodes[ref].r = r;
    return ref;
  };
  t.root = td(t.root);
}

float sah_cost(const Tree& t, float Ct) {
  double c = 0; float ra = refbox(t, t.root).area();
  function<void(int)> rec = [&](int ref) {
    if (ref < 0) { c += t.leaves[~ref].b.area() * Ct * t.leaves[~ref].prims.size(); return; }
    c += t.nodes[ref].b.area(); rec(t.nodes[ref].l); rec(t.nodes[ref].r);
  };
  rec(t.root);
  return (float)(c / ra);
}

// ---------- 4-wide collapse
struct WNode { Box cb[4]; int ref[4]; int n; };  // ref >=0 wide node index, <0 leaf
struct WTree { vector<WNode> nodes; int root; int levels; };
WTree widen(const Tree& t, int policy /*0: largest area first*/) {
  WTree w;
  w.levels = 0;
  function<int(int, int)> rec = [&](int bref, int level) -> int {
    w.levels = max(w.levels, level + 1);
    int kids[4] = {t.nodes[bref].l, t.nodes[bref].r, 0, 0};
    int cnt = 2;
    while (cnt < 4) {
      int best = -1; float ba = -1;
      for (int k = 0; k < cnt; ++k) if (kids[k] >= 0) { float a = t.nodes[kids[k]].b.area(); if (a > ba) { ba = a; best = k; } }
      if (best < 0) break;
      int c = kids[best];
      kids[best] = t.nodes[c].l; kids[cnt++] = t.nodes[c].r;
    }
    int me = (int)w.nodes.size();
    w.nodes.push_back(WNode());
    WNode nd; nd.n = cnt;
    for (int k = 0; k < cnt; ++k) { nd.cb[k] = refbox(t, kids[k]); nd.ref[k] = kids[k] >= 0 ? rec(kids[k], level + 1) : kids[k]; }
    w.nodes[me] = nd;
    return me;
  };
  if (t.root < 0) { w.root = t.root; return w; }
  w.root = rec(t.root, 0);
  return w;
}

// ---------- tracing
struct Stats { double rays = 0, visits = 0, boxtests = 0, leafvisits = 0, tritests = 0, maxstack = 0; };
static bool tri_hit(const Tri& T, V3 o, V3 d, float& t) {
  V3 e1 = T.b - T.a, e2 = T.c - T.a;
  V3 h = cross(d, e2);
  float a = dot(e1, h);
  if (fabsf(a) < 1e-9f) return false;
  float f = 1.0f / a;
  V3 s = o - T.a;
  float u = f * dot(s, h);
  if (u < 0 || u > 1) return false;
  V3 q = cross(s, e1);
  float v = f * dot(d, q);
  if (v < 0 || u + v > 1) return false;
  float tt = f * dot(e2, q);
  if (tt > 1e-4f && tt < t) { t = tt; return true; }
  return false;
}
static bool box_hit(const Box& b, V3 o, V3 inv, float tbest, float& tn) {
  float t0x = (b.lo[0] - o.x) * inv.x, t1x = (b.hi[0] - o.x) * inv.x;
  float t0y = (b.lo[1] - o.y) * inv.y, t1y = (b.hi[1] - o.y) * inv.y;
  float t0z = (b.lo[2] - o.z) * inv.z, t1z = (b.hi[2] - o.z) * inv.z;
  tn = max(max(min(t0x, t1x), min(t0y, t1y)), min(t0z, t1z));
  float tf = min(min(max(t0x, t1x), max(t0y, t1y)), max(t0z, t1z));
  return tf >= max(tn, 0.0f) && tn <= tbest;
}
int trace(const Tree& t, const WTree& w, V3 o, V3 d, float& tbest, Stats& st) {
  V3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
  int hit = -1;
  struct E { int ref; float d; };
  E stack[256]; int sp = 0;
  int cur = w.root;
  int maxsp = 0;
  auto leaf = [&](int ref) {
    st.leafvisits++;
    for (int p : t.leaves[~ref].prims) { st.tritests++; if (tri_hit(tris[p], o, d, tbest)) hit = p; }
  };
  if (cur < 0) { leaf(cur); return hit; }
  for (;;) {
    const WNode& n = w.nodes[cur];
    st.visits++; st.boxtests += n.n;
    E h[4]; int nh = 0;
    for (int k = 0; k < n.n; ++k) { float tn; if (box_hit(n.cb[k], o, inv, tbest, tn)) { h[nh++] = {n.ref[k], max(tn, 0.0f)}; } }
    sort(h, h + nh, [](const E& a, const E& b) { return a.d < b.d; });
    for (int k = nh - 1; k >= 0; --k) stack[sp++] = h[k];
    maxsp = max(maxsp, sp);
    cur = -1;
    bool found = false;
    while (sp > 0) {
      E e = stack[--sp];
      if (e.d > tbest) continue;
      if (e.ref < 0) { leaf(e.ref); continue; }
      cur = e.ref; found = true; break;
    }
    if (!found) break;
  }
  st.maxstack = max(st.maxstack, (double)maxsp);
  return hit;
}

int main(int argc, char** argv) {
  const char* name = argc > 1 ? argv[1] : "c4";
  char path[256];
  snprintf(path, 256, "%s_big.bin", name);
  FILE* f = fopen(path, "rb"); fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  tris.resize(sz / sizeof(Tri)); fread(tris.data(), sizeof(Tri), tris.size(), f); fclose(f);
  snprintf(path, 256, "%s_other.bin", name);
  f = fopen(path, "rb"); fseek(f, 0, SEEK_END); sz = ftell(f); fseek(f, 0, SEEK_SET);
  others.resize(sz / sizeof(Tri)); fread(others.data(), sizeof(Tri), others.size(), f); fclose(f);
  printf("%zu tris, %zu others\n", tris.size(), others.size());
  pbox.resize(tris.size());
  for (size_t i = 0; i < tris.size(); ++i) {
    Box b; b.init();
    const V3* v = &tris[i].a;
    for (int k = 0; k < 3; ++k) { b.lo[0] = min(b.lo[0], v[k].x); b.lo[1] = min(b.lo[1], v[k].y); b.lo[2] = min(b.lo[2], v[k].z);
      b.hi[0] = max(b.hi[0], v[k].x); b.hi[1] = max(b.hi[1], v[k].y); b.hi[2] = max(b.hi[2], v[k].z); }
    pbox[i] = b;
  }
  // rays: generated once with the SAH tree (paths), stored, then replayed on every variant
  struct Ray { V3 o, d; };
  vector<Ray> rays;
  {
    Tree t = build_sah(4);
    WTree w = widen(t, 0);
    mt19937 rng(1);
    uniform_real_distribution<float> U(0, 1);
    normal_distribution<float> N(0, 1);
    V3 cam = {0, 260, 470};
    Stats st;
    int W = 240, H = 135;
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
      // pitch 0.35 down, looking -z
      float u = ((x + 0.5f) / W * 2 - 1) * 1.7777f, v = 1 - (y + 0.5f) / H * 2;
      V3 d = norm({u, v, -1.0f});
      float cp = cosf(0.35f), sp = sinf(0.35f);
      d = {d.x, d.y * cp + d.z * sp * 1.0f, -d.y * sp * 1.0f + d.z * cp};
      d = norm({d.x, d.y - 0.0f, d.z});
      // tilt downwards
      d = norm({d.x, d.y - 0.35f, d.z});
      V3 o = cam;
      for (int b = 0; b < 7; ++b) {
        rays.push_back({o, d});
        float tb = 1e30f;
        int hp = trace(t, w, o, d, tb, st);
        V3 nrm; bool any = hp >= 0;
        if (any) nrm = norm(cross(tris[hp].b - tris[hp].a, tris[hp].c - tris[hp].a));
        for (auto& T : others) if (tri_hit(T, o, d, tb)) { any = true; nrm = norm(cross(T.b - T.a, T.c - T.a)); }
        if (!any) break;
        if (dot(nrm, d) > 0) nrm = nrm * -1.0f;
        V3 p = o + d * tb;
        V3 r = norm({N(rng), N(rng), N(rng)});
        d = norm(nrm + r * 0.999f);
        o = p + nrm * 1e-2f;
        if (b >= 3 && U(rng) < 0.25f) break;
      }
    }
    printf("%zu rays generated\n", rays.size());
  }
  auto eval = [&](const char* label, const Tree& t) {
    WTree w = widen(t, 0);
    Stats st;
    for (auto& r : rays) { float tb = 1e30f; trace(t, w, r.o, r.d, tb, st); st.rays++; }
    printf("%-34s sah %8.2f | wide nodes %8zu levels %2d | per ray: visits %6.2f boxtests %6.2f leafvisits %5.2f tritests %5.2f maxstack %.0f\n", label,
           sah_cost(t, 1.0f), w.nodes.size(), w.levels, st.visits / st.rays, st.boxtests / st.rays, st.leafvisits / st.rays, st.tritests / st.rays, st.maxstack);
    fflush(stdout);
  };
  {
    Tree t = build_lbvh();
    eval("lbvh", t);
    { Tree c = t; collapse_leaves(c, 2, 0, 1); eval("lbvh leaf<=2", c); }
    { Tree c = t; collapse_leaves(c, 4, 0, 1); eval("lbvh leaf<=4", c); }
    { Tree c = t; collapse_leaves(c, 8, 0, 1); eval("lbvh leaf<=8", c); }
    Tree r = t;
    for (int pass = 0; pass < 3; ++pass) {
      int a = rotate_pass(r);
      // refit counts/boxes fully
      function<void(int)> fit = [&](int i) { Node& nd = r.nodes[i]; if (nd.l >= 0) fit(nd.l); if (nd.r >= 0) fit(nd.r); nd.b = uni(refbox(r, nd.l), refbox(r, nd.r)); nd.count = refcount(r, nd.l) + refcount(r, nd.r); };
      fit(r.root);
      char lab[64]; snprintf(lab, 64, "lbvh + rot pass %d (%d applied)", pass + 1, a);
      eval(lab, r);
      if (pass == 0 || pass == 2) { Tree c = r; collapse_leaves(c, 4, 0, 1); snprintf(lab, 64, "lbvh + rot x%d + leaf<=4", pass + 1); eval(lab, c); }
    }
  }
  {
    Tree lb = build_lbvh();
    for (int T : {64, 512, 4096, 32768}) {
      Tree h = build_hybrid(lb, T);
      char lab[64]; snprintf(lab, 64, "hybrid SAH-over-clusters T=%d", T); eval(lab, h);
      Tree c = h; collapse_leaves(c, 2, 0, 1); snprintf(lab, 64, "hybrid T=%d + leaf<=2", T); eval(lab, c);
    }
  }
  { Tree t = build_sah(1); eval("binned sah leaf=1", t); }
  { Tree t = build_sah(4); eval("binned sah leaf<=4", t); }
  return 0;
}
