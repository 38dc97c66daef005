// rr_host.cpp -- host-side formats either side of the render path: the OBJ
// triangle input (reference src/readobj.hpp:270-376), the scene assembly
// helpers (addQuad src/readobj.hpp:378-408, addCornellBoxToScene
// src/image.hpp:401-448), the default camera (src/main.cpp:299-304,
// src/settings.hpp:23-28) and the output.bmp writer (src/math.hpp:117-164).
// Pure C++; no CUDA in this file.
#include <algorithm>
#include <cfloat>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/rr_api.h"

// Output array whose pages are first touched by the threads that fill it (std::vector::resize would zero -- and
// page-fault -- every byte on the calling thread, which costs more than the parse itself at 16 threads).
template <class T>
struct RawArray {
  T* p = nullptr;
  size_t n = 0;
  RawArray() = default;
  RawArray(const RawArray&) = delete;
  RawArray& operator=(const RawArray&) = delete;
  ~RawArray() { free(p); }
  bool resize_keep(size_t count) {  // contents up to min(n, count) are kept
    if (count == 0) { free(p); p = nullptr; n = 0; return true; }  // realloc(p, 0) is not portable
    T* q = static_cast<T*>(realloc(p, count * sizeof(T)));
    if (!q) return false;
    p = q; n = count;
    return true;
  }
  size_t size() const { return n; }
  T* data() { return p; }
  const T* data() const { return p; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

struct rr_obj {
  RawArray<float> positions, normals;  // x,y,z triples
  RawArray<uint32_t> corners;          // v0 v1 v2 n0 n1 n2 per triangle, 0-based
};

struct rr_scene {
  std::vector<rr_triangle> tris;
  std::vector<rr_mesh> meshes;
  std::vector<rr_mesh_range> ranges;
  std::vector<rr_sphere> spheres;
};

namespace {

rr_float3 f3(float x, float y, float z) {
  rr_float3 r;
  r.s[0] = x; r.s[1] = y; r.s[2] = z; r.s[3] = 0.0f;
  return r;
}

rr_material solid(rr_float3 color) {
  rr_material m;
  memset(&m, 0, sizeof(m));
  m.type = RR_MATERIAL_SOLID;
  m.ior = 1.0f;  // struct default, src/readobj.hpp:50
  m.color = color;
  return m;
}

rr_mesh default_mesh() {
  rr_mesh m;
  memset(&m, 0, sizeof(m));
  m.scale = 1.0f;  // src/readobj.hpp:79
  m.material = solid(f3(1.0f, 1.0f, 1.0f));
  return m;
}

// ---- OBJ text -> indexed arrays, in parallel ---------------------------------------------------------------
// The file is read once and cut into one chunk per thread at line boundaries.  Pass 1 (parallel) parses the `v` /
// `vn` lines of a chunk into chunk-local arrays and notes where its `f` lines are, together with how many `v` / `vn`
// lines precede each of them inside the chunk; a prefix sum over the chunks then gives every face line the number of
// vertices / normals defined BEFORE it in the file -- what 1-based, relative (negative) and out-of-range indices are
// resolved against, exactly as a sequential reader does.  Pass 2 (parallel) parses the faces against the merged
// vertex array.  Two dialects:
//   strict     the reference loader's (src/readobj.hpp:289-344): `v `, `vn `, `f a/b/c` or `f a//c`, first three
//              corners only, normals mandatory, no negative indices;
//   permissive polygons (fan), `f a` / `f a/b` (face normals generated), negative indices, tabs.
struct ObjFace { size_t offset; uint32_t lv, ln; };  // line start in the file; `v` / `vn` lines before it in its chunk
struct ObjChunk {
  size_t begin = 0, end = 0;
  std::vector<float> pos, nrm, gen;  // gen: face normals generated for faces without `vn`
  std::vector<ObjFace> faces;
  std::vector<uint32_t> corners;     // v0 v1 v2 n0 n1 n2; generated normals flagged with bit 30 (chunk-local index)
  size_t base_v = 0, base_n = 0, base_gen = 0, base_tri = 0;
};

inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// "%f" of sscanf on [p, end): optional blanks, then a float; correctly rounded like strtof.  Anything from_chars does
// not take (a leading '+', hex floats, out-of-range values) goes through strtof on a bounded copy.
bool parse_float(const char*& p, const char* end, float& out) {
  while (p < end && is_blank(*p)) ++p;
  if (p >= end) return false;
  const char* d = (*p == '-') ? p + 1 : p;
  const bool hex = end - d >= 2 && d[0] == '0' && (d[1] == 'x' || d[1] == 'X');  // "%f" reads hex floats, from_chars would stop at the x
  if (!hex) {
    const auto r = std::from_chars(p, end, out);
    if (r.ec == std::errc()) { p = r.ptr; return true; }
  }
  char buf[96];
  const size_t n = std::min<size_t>((size_t)(end - p), sizeof(buf) - 1);
  memcpy(buf, p, n);
  buf[n] = 0;
  char* e;
  out = strtof(buf, &e);
  if (e == buf) return false;
  p += e - buf;
  return true;
}
bool parse_float3(const char* p, const char* end, float* xyz) {
  return parse_float(p, end, xyz[0]) && parse_float(p, end, xyz[1]) && parse_float(p, end, xyz[2]);
}
// strtol(p, &end, 10) on [p, end): blanks, optional sign, digits.  false when no digit was read (p unchanged).
bool parse_long(const char*& p, const char* end, long& out) {
  const char* q = p;
  while (q < end && is_blank(*q)) ++q;
  bool neg = false;
  if (q < end && (*q == '-' || *q == '+')) { neg = *q == '-'; ++q; }
  if (q >= end || *q < '0' || *q > '9') return false;
  long v = 0;
  while (q < end && *q >= '0' && *q <= '9') { v = v * 10 + (*q - '0'); if (v > (1L << 40)) v = 1L << 40; ++q; }
  out = neg ? -v : v;
  p = q;
  return true;
}

void obj_pass1(const char* data, ObjChunk& c, bool strict) {
  const char* p = data + c.begin;
  const char* const stop = data + c.end;
  uint32_t lv = 0, ln = 0;
  while (p < stop) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(stop - p));
    const char* e = nl ? nl : stop;
    if (e - p >= 2) {
      const bool sep1 = p[1] == ' ' || (!strict && p[1] == '\t');
      if (p[0] == 'v' && sep1) {
        float f[3];
        if (parse_float3(p + 1, e, f)) { c.pos.insert(c.pos.end(), f, f + 3); ++lv; }
      } else if (p[0] == 'v' && p[1] == 'n' && e - p >= 3 && (p[2] == ' ' || (!strict && p[2] == '\t'))) {
        float f[3];
        if (parse_float3(p + 2, e, f)) { c.nrm.insert(c.nrm.end(), f, f + 3); ++ln; }
      } else if (p[0] == 'f' && sep1) {
        c.faces.push_back({(size_t)(p - data), lv, ln});
      }
    }
    p = nl ? nl + 1 : stop;
  }
}

void obj_pass2(const char* data, size_t size, ObjChunk& c, bool strict, const float* positions) {
  std::vector<long> fv, fn;
  for (const ObjFace& face : c.faces) {
    const char* p = data + face.offset;
    const char* nl = (const char*)memchr(p, '\n', size - face.offset);
    const char* e = nl ? nl : data + size;
    const long nv = (long)(c.base_v + face.lv), nn = (long)(c.base_n + face.ln);  // defined before this line
    if (strict) {
      // three corners, ALL `v/vt/vn` or ALL `v//vn`: the reference tries one sscanf pattern per form
      // (src/readobj.hpp:307-312), so a line that mixes the two is an "Unsupported face format"; a 4th corner is ignored
      p += 2;
      long v[3], n[3];
      bool ok = true;
      int form = -1;  // 0: v//vn, 1: v/vt/vn
      for (int k = 0; k < 3 && ok; ++k) {
        long vt;
        ok = parse_long(p, e, v[k]) && p < e && *p == '/';
        if (!ok) break;
        ++p;
        int f = 0;
        if (p < e && *p == '/') ++p;
        else { f = 1; ok = parse_long(p, e, vt) && p < e && *p == '/'; if (!ok) break; ++p; }
        if (form < 0) form = f;
        ok = f == form && parse_long(p, e, n[k]);
      }
      if (!ok) continue;
      for (int k = 0; k < 3; ++k) {
        v[k] -= 1; n[k] -= 1;
        if (v[k] < 0 || v[k] >= nv || n[k] < 0 || n[k] >= nn) ok = false;
      }
      if (!ok) continue;
      for (int k = 0; k < 3; ++k) c.corners.push_back((uint32_t)v[k]);
      for (int k = 0; k < 3; ++k) c.corners.push_back((uint32_t)n[k]);
      continue;
    }
    fv.clear(); fn.clear();
    ++p;
    bool ok = true;
    for (;;) {  // corners: v, v/vt, v//vn, v/vt/vn
      while (p < e && is_blank(*p)) ++p;
      if (p >= e) break;
      long v, n = 0, vt;
      if (!parse_long(p, e, v)) { ok = false; break; }
      if (p < e && *p == '/') {
        ++p;
        if (!(p < e && *p == '/')) parse_long(p, e, vt);  // texture index, unused
        if (p < e && *p == '/') {
          ++p;
          if (!parse_long(p, e, n)) { ok = false; break; }
        }
      }
      v = v < 0 ? nv + v : v - 1;  // negative indices count from the end
      n = n < 0 ? nn + n : n - 1;  // n == 0 (absent) becomes -1
      if (v < 0 || v >= nv || n >= nn) { ok = false; break; }
      fv.push_back(v);
      fn.push_back(n);
    }
    if (!ok || fv.size() < 3) continue;
    for (size_t k = 1; k + 1 < fv.size(); ++k) {  // fan triangulation; a triangle is a fan of one
      const long v3[3] = {fv[0], fv[k], fv[k + 1]};
      long n3[3] = {fn[0], fn[k], fn[k + 1]};
      if (n3[0] < 0 || n3[1] < 0 || n3[2] < 0) {  // no normals: one face normal for the triangle
        const float* A = &positions[3 * v3[0]];
        const float* B = &positions[3 * v3[1]];
        const float* C = &positions[3 * v3[2]];
        const float e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        float nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
        const float len = std::sqrt(nx * nx + ny * ny + nz * nz);
        if (len > 0.0f) { nx /= len; ny /= len; nz /= len; } else { nx = 0.0f; ny = 1.0f; nz = 0.0f; }
        const long idx = (long)(c.gen.size() / 3);
        c.gen.push_back(nx); c.gen.push_back(ny); c.gen.push_back(nz);
        n3[0] = n3[1] = n3[2] = idx | 0x40000000L;
      }
      for (int q = 0; q < 3; ++q) c.corners.push_back((uint32_t)v3[q]);
      for (int q = 0; q < 3; ++q) c.corners.push_back((uint32_t)n3[q]);
    }
  }
}

template <class F>
void parallel_chunks(size_t n, F&& body) {
  std::vector<std::thread> pool;
  for (size_t k = 1; k < n; ++k) pool.emplace_back([&body, k] { body(k); });
  body(0);
  for (auto& t : pool) t.join();
}

// The file's bytes, mapped (no copy of the page cache); an empty or unmappable file falls back to read().
struct FileText {
  const char* data = nullptr;
  size_t size = 0;
  void* map = nullptr;
  std::string fallback;
  ~FileText() { if (map) munmap(map, size); }
  bool open(const char* path) {
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd); return read_all(path); }
    size = (size_t)st.st_size;
    if (size == 0) { ::close(fd); data = ""; return true; }
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    ::close(fd);
    if (m == MAP_FAILED) return read_all(path);
    map = m;
    data = static_cast<const char*>(m);
    return true;
  }
  bool read_all(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    fallback.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    data = fallback.data();
    size = fallback.size();
    return true;
  }
};

int obj_parse(const char* path, bool strict, rr_obj* o) {
  FileText text;
  if (!text.open(path)) return RR_ERR_IO;
  const char* data = text.data;
  const size_t size = text.size;
  const bool dbg = getenv("RR_OBJ_DEBUG") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double tl = now();
  auto lap = [&](const char* what) { if (dbg) { const double t = now(); fprintf(stderr, "[obj] %s %.1f ms\n", what, (t - tl) * 1e3); tl = t; } };
  size_t threads = std::max(1u, std::thread::hardware_concurrency());
  if (const char* e = getenv("RR_OBJ_THREADS")) threads = (size_t)std::max(1, atoi(e));
  size_t min_chunk = 1u << 20;  // at least 1 MiB of text per thread
  if (const char* e = getenv("RR_OBJ_MIN_CHUNK")) min_chunk = (size_t)std::max(1, atoi(e));
  threads = std::min(threads, std::max<size_t>(1, size / min_chunk));
  std::vector<ObjChunk> chunks(threads);
  size_t at = 0;
  for (size_t k = 0; k < threads; ++k) {  // cut at line boundaries
    chunks[k].begin = at;
    size_t want = k + 1 == threads ? size : std::max(at, size / threads * (k + 1));
    if (want < size) {
      const char* nl = (const char*)memchr(data + want, '\n', size - want);
      want = nl ? (size_t)(nl - data) + 1 : size;
    }
    chunks[k].end = at = want;
  }
  lap("read+cut");
  parallel_chunks(threads, [&](size_t k) { obj_pass1(data, chunks[k], strict); });
  lap("pass1");
  size_t nv = 0, nn = 0;
  for (ObjChunk& c : chunks) { c.base_v = nv; c.base_n = nn; nv += c.pos.size() / 3; nn += c.nrm.size() / 3; }
  if (nv >= 0x3fffffffull || nn >= 0x3fffffffull) return RR_ERR_UNSUPPORTED;
  if (!o->positions.resize_keep(3 * nv) || !o->normals.resize_keep(3 * nn)) return RR_ERR_OUT_OF_MEMORY;
  parallel_chunks(threads, [&](size_t k) {
    ObjChunk& c = chunks[k];
    if (!c.pos.empty()) memcpy(&o->positions[3 * c.base_v], c.pos.data(), c.pos.size() * sizeof(float));
    if (!c.nrm.empty()) memcpy(&o->normals[3 * c.base_n], c.nrm.data(), c.nrm.size() * sizeof(float));
    std::vector<float>().swap(c.pos);
    std::vector<float>().swap(c.nrm);
  });
  lap("merge v/vn");
  parallel_chunks(threads, [&](size_t k) { obj_pass2(data, size, chunks[k], strict, o->positions.data()); });
  lap("pass2");
  size_t ngen = 0, ntri = 0;
  for (ObjChunk& c : chunks) { c.base_gen = ngen; c.base_tri = ntri; ngen += c.gen.size() / 3; ntri += c.corners.size() / 6; }
  if (ntri >= 0x7fffffffull || nn + ngen >= 0x3fffffffull) return RR_ERR_UNSUPPORTED;
  // generated face normals go behind the file's normals, so that they never shift the file's own normal indices
  if (!o->normals.resize_keep(3 * (nn + ngen)) || !o->corners.resize_keep(6 * ntri)) return RR_ERR_OUT_OF_MEMORY;
  parallel_chunks(threads, [&](size_t k) {
    ObjChunk& c = chunks[k];
    if (!c.gen.empty()) memcpy(&o->normals[3 * (nn + c.base_gen)], c.gen.data(), c.gen.size() * sizeof(float));
    uint32_t* dst = o->corners.data() + 6 * c.base_tri;
    const uint32_t gen0 = (uint32_t)(nn + c.base_gen);
    for (size_t i = 0; i < c.corners.size(); ++i) {
      const uint32_t v = c.corners[i];
      dst[i] = (i % 6 >= 3 && (v & 0x40000000u)) ? gen0 + (v & 0x3fffffffu) : v;
    }
  });
  lap("merge faces");
  return RR_OK;
}

}  // namespace

extern "C" {

// OBJ -> indexed arrays with the reference loader's limits lifted (see include/rr_api.h).
int rr_obj_load(const char* path, rr_obj** out) {
  if (!path || !out) return RR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  rr_obj* o = new rr_obj();
  const int rc = obj_parse(path, false, o);
  if (rc) { delete o; return rc; }
  *out = o;
  return RR_OK;
}
void rr_obj_destroy(rr_obj* o) { delete o; }
size_t rr_obj_position_count(const rr_obj* o) { return o ? o->positions.size() / 3 : 0; }
size_t rr_obj_normal_count(const rr_obj* o) { return o ? o->normals.size() / 3 : 0; }
size_t rr_obj_triangle_count(const rr_obj* o) { return o ? o->corners.size() / 6 : 0; }
const float* rr_obj_positions(const rr_obj* o) { return o ? o->positions.data() : nullptr; }
const float* rr_obj_normals(const rr_obj* o) { return o ? o->normals.data() : nullptr; }
const uint32_t* rr_obj_corners(const rr_obj* o) { return o ? o->corners.data() : nullptr; }

int rr_scene_create(rr_scene** out) {
  if (!out) return RR_ERR_INVALID_ARGUMENT;
  *out = new rr_scene();
  return RR_OK;
}
void rr_scene_destroy(rr_scene* s) { delete s; }

int rr_scene_add_triangles(rr_scene* s, const rr_triangle* tris, size_t n, rr_mesh_range* range_out) {
  if (!s || (n && !tris)) return RR_ERR_INVALID_ARGUMENT;
  const size_t first = s->tris.size();
  s->tris.insert(s->tris.end(), tris, tris + n);
  if (range_out) { range_out->firstTriangle = first; range_out->numTriangles = n; }
  return RR_OK;
}

// OBJ dialect of the reference loader: `v x y z`, `vn x y z`, `f a/b/c ...` or
// `f a//c ...` with exactly the first three corners used (a 4th is ignored,
// src/readobj.hpp:307-312); 1-based indices; normals mandatory.  Unlike the
// reference, a face that fails to parse or indexes out of bounds is skipped
// WITHOUT corrupting the mesh's triangle range (the reference counts it in
// triCount before validating, src/readobj.hpp:305,346).
int rr_scene_load_obj(rr_scene* s, const char* path, rr_mesh* mesh_out, rr_mesh_range* range_out) {
  if (!s || !path) return RR_ERR_INVALID_ARGUMENT;
  rr_obj o;
  const int rc = obj_parse(path, true, &o);
  if (rc) return rc;
  const size_t first = s->tris.size(), n = o.corners.size() / 6;
  s->tris.resize(first + n);
  // the 96-byte Triangle records of src/readobj.hpp:313-343, expanded in parallel
  const size_t threads = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, n >> 16));
  parallel_chunks(threads, [&](size_t k) {
    for (size_t i = n * k / threads; i < n * (k + 1) / threads; ++i) {
      const uint32_t* c = &o.corners[6 * i];
      rr_triangle& t = s->tris[first + i];
      rr_float3* dst[6] = {&t.posA, &t.posB, &t.posC, &t.normalA, &t.normalB, &t.normalC};
      for (int q = 0; q < 6; ++q) {
        const float* src = q < 3 ? &o.positions[3 * (size_t)c[q]] : &o.normals[3 * (size_t)c[q]];
        *dst[q] = f3(src[0], src[1], src[2]);
      }
    }
  });
  if (mesh_out) *mesh_out = default_mesh();  // src/readobj.hpp:369-375
  if (range_out) { range_out->firstTriangle = first; range_out->numTriangles = n; }
  return RR_OK;
}

// Root-node bounds of the OBJ loader (src/readobj.hpp:353-362), including its
// initial value max = +FLT_MIN (src/readobj.hpp:16-17): geometry entirely below
// zero on an axis reports max = 1.18e-38 there, which the Cornell box inherits.
int rr_scene_range_bounds(const rr_scene* s, const rr_mesh_range* range, float* min3, float* max3) {
  if (!s || !range || !min3 || !max3) return RR_ERR_INVALID_ARGUMENT;
  if (range->firstTriangle + range->numTriangles > s->tris.size()) return RR_ERR_BAD_MESH_RANGE;
  for (int a = 0; a < 3; ++a) { min3[a] = FLT_MAX; max3[a] = FLT_MIN; }
  for (size_t i = 0; i < range->numTriangles; ++i) {
    const rr_triangle& t = s->tris[range->firstTriangle + i];
    for (int a = 0; a < 3; ++a) {
      min3[a] = std::min(min3[a], std::min(t.posA.s[a], std::min(t.posB.s[a], t.posC.s[a])));
      max3[a] = std::max(max3[a], std::max(t.posA.s[a], std::max(t.posB.s[a], t.posC.s[a])));
    }
  }
  return RR_OK;
}

int rr_scene_add_mesh(rr_scene* s, const rr_mesh* mesh, const rr_mesh_range* range) {
  if (!s || !mesh || !range) return RR_ERR_INVALID_ARGUMENT;
  if (range->firstTriangle + range->numTriangles > s->tris.size()) return RR_ERR_BAD_MESH_RANGE;
  s->meshes.push_back(*mesh);
  s->ranges.push_back(*range);
  return RR_OK;
}

int rr_scene_add_quad(rr_scene* s, const float* a, const float* b, const float* c, const float* d, const float* normal,
                      const float* color) {
  if (!s || !a || !b || !c || !d || !normal || !color) return RR_ERR_INVALID_ARGUMENT;
  rr_mesh m = default_mesh();
  m.material = solid(f3(color[0], color[1], color[2]));  // reflectiveness 0, specularProbability 0 (:394-402)
  rr_mesh_range r;
  r.firstTriangle = s->tris.size();
  r.numTriangles = 2;
  const rr_float3 A = f3(a[0], a[1], a[2]), B = f3(b[0], b[1], b[2]), C = f3(c[0], c[1], c[2]), D = f3(d[0], d[1], d[2]),
                  N = f3(normal[0], normal[1], normal[2]);
  rr_triangle t1 = {A, B, C, N, N, N}, t2 = {A, C, D, N, N, N};  // src/readobj.hpp:405-406
  s->tris.push_back(t1);
  s->tris.push_back(t2);
  s->meshes.push_back(m);
  s->ranges.push_back(r);
  return RR_OK;
}

#define RR_CORNELL_BREATHING_ROOM 100.0f /* src/settings.hpp:52 */

int rr_scene_add_cornell(rr_scene* s, const rr_mesh* mesh, const rr_mesh_range* range) {
  if (!s || !mesh || !range) return RR_ERR_INVALID_ARGUMENT;
  float bmin[3], bmax[3];
  int rc = rr_scene_range_bounds(s, range, bmin, bmax);
  if (rc) return rc;
  // src/image.hpp:403-408
  const float minX = (bmin[0] * mesh->scale) - RR_CORNELL_BREATHING_ROOM, maxX = (bmax[0] * mesh->scale) + RR_CORNELL_BREATHING_ROOM;
  const float minY = (bmin[1] * mesh->scale), maxY = (bmax[1] * mesh->scale) + RR_CORNELL_BREATHING_ROOM;
  const float minZ = (bmin[2] * mesh->scale) - RR_CORNELL_BREATHING_ROOM, maxZ = (bmax[2] * mesh->scale) + RR_CORNELL_BREATHING_ROOM;
  auto quad = [&](float ax, float ay, float az, float bx, float by, float bz, float cx, float cy, float cz, float dx, float dy,
                  float dz, float nx, float ny, float nz, float r, float g, float b) {
    const float A[3] = {ax, ay, az}, B[3] = {bx, by, bz}, C[3] = {cx, cy, cz}, D[3] = {dx, dy, dz}, N[3] = {nx, ny, nz},
                col[3] = {r, g, b};
    rr_scene_add_quad(s, A, B, C, D, N, col);
  };
  // Floor (:411-421)
  quad(minX, minY, minZ, maxX, minY, minZ, maxX, minY, maxZ, minX, minY, maxZ, 0, 1, 0, 0, 0, 0);
  {
    rr_material& m = s->meshes.back().material;
    memset(&m, 0, sizeof(m));
    m.type = RR_MATERIAL_SOLID; m.ior = 1.0f;
    m.color = f3(0.1f, 0.1f, 0.1f);
    m.specularProbability = 1.0f;
  }
  // Ceiling (:424)
  quad(minX, maxY, minZ, maxX, maxY, minZ, maxX, maxY, maxZ, minX, maxY, maxZ, 0, -1, 0, 1, 1, 1);
  // Front wall, one-sided (:427-428)
  quad(minX, minY, maxZ, maxX, minY, maxZ, maxX, maxY, maxZ, minX, maxY, maxZ, 0, 0, -1, 1.0f, 1.0f, 1.0f);
  s->meshes.back().material.type = RR_MATERIAL_ONESIDED;
  // Back wall, green (:432)
  quad(minX, minY, minZ, maxX, minY, minZ, maxX, maxY, minZ, minX, maxY, minZ, 0, 0, 1, 0.1f, 0.8f, 0.1f);
  // Left wall, blue (:435)
  quad(minX, minY, minZ, minX, minY, maxZ, minX, maxY, maxZ, minX, maxY, minZ, 1, 0, 0, 0.1f, 0.1f, 1.f);
  // Right wall, red (:438)
  quad(maxX, minY, minZ, maxX, minY, maxZ, maxX, maxY, maxZ, maxX, maxY, minZ, -1, 0, 0, 1.f, 0.2f, 0.2f);
  // Light quad just below the ceiling (:441-447)
  const float lx = 50, lz = 50, ly = maxY - 1;
  quad(-lx, ly, -lz, lx, ly, -lz, lx, ly, lz, -lx, ly, lz, 0, -1, 0, 0.0f, 0.0f, 0.0f);
  {
    rr_material& m = s->meshes.back().material;
    memset(&m, 0, sizeof(m));
    m.type = RR_MATERIAL_SOLID; m.ior = 1.0f;
    m.color = f3(1, 1, 1);
    m.emissionColor = f3(1.0f, 1.0f, 1.0f);
    m.emissionStrength = 8.0f;
    m.specularProbability = 1.0f;
  }
  return RR_OK;
}

int rr_scene_add_sphere(rr_scene* s, const rr_sphere* sphere) {
  if (!s || !sphere) return RR_ERR_INVALID_ARGUMENT;
  s->spheres.push_back(*sphere);
  return RR_OK;
}

rr_mesh* rr_scene_mesh(rr_scene* s, size_t index) { return (s && index < s->meshes.size()) ? &s->meshes[index] : nullptr; }
size_t rr_scene_mesh_count(const rr_scene* s) { return s ? s->meshes.size() : 0; }
size_t rr_scene_triangle_count(const rr_scene* s) { return s ? s->tris.size() : 0; }
size_t rr_scene_sphere_count(const rr_scene* s) { return s ? s->spheres.size() : 0; }
const rr_triangle* rr_scene_triangles(const rr_scene* s) { return s ? s->tris.data() : nullptr; }
const rr_mesh* rr_scene_meshes(const rr_scene* s) { return s ? s->meshes.data() : nullptr; }
const rr_mesh_range* rr_scene_ranges(const rr_scene* s) { return s ? s->ranges.data() : nullptr; }
const rr_sphere* rr_scene_spheres(const rr_scene* s) { return s ? s->spheres.data() : nullptr; }

int rr_scene_upload(rr_ctx* ctx, const rr_scene* s) {
  if (!s) return RR_ERR_INVALID_ARGUMENT;
  return rr_upload_scene(ctx, s->tris.data(), s->tris.size(), s->meshes.data(), s->ranges.data(), s->meshes.size(),
                         s->spheres.data(), s->spheres.size());
}

void rr_default_camera(rr_camera* cam, uint32_t width, uint32_t height) {
  if (!cam) return;
  memset(cam, 0, sizeof(*cam));
  cam->position = f3(0.0f, 150.0f, 250.0f);  // src/settings.hpp:23-25
  cam->pitch = 0.0f;
  cam->yaw = 3.14f;                           // src/settings.hpp:27
  cam->roll = 0.0f;
  cam->fov = 90.0f;                           // src/main.cpp:303
  cam->aspectRatio = (float)width / (float)height;
}

int rr_video_frame_setup(rr_mesh* meshes, size_t n_meshes, int32_t frame_index, int32_t frame_count) {
  if (!meshes || n_meshes == 0 || frame_count <= 0) return RR_ERR_INVALID_ARGUMENT;
  // src/image.hpp:387-390, same float operations in the same order
  const float anglePerFrame = (3.14159265359f * 2.0f) / (float)frame_count;
  const float currentRotation = anglePerFrame * (float)frame_index;
  meshes[n_meshes - 1].yaw = currentRotation + 5.5f;
  return RR_OK;
}

int rr_video_frame_path(const char* dir, int32_t frame_number, char* out, size_t out_len) {
  if (!dir || !out || out_len == 0) return RR_ERR_INVALID_ARGUMENT;
  const int n = snprintf(out, out_len, "%s/output_%d.bmp", dir, (int)frame_number);  // src/main.cpp:701
  return (n < 0 || (size_t)n >= out_len) ? RR_ERR_INVALID_ARGUMENT : RR_OK;
}

int rr_write_bmp(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height) {
  if (!path || !rgba) return RR_ERR_INVALID_ARGUMENT;
  FILE* f = fopen(path, "wb");
  if (!f) return RR_ERR_IO;
  const int padSize = (4 - (int)((width * 3u) % 4u)) % 4;
  const int rowSize = 3 * (int)width + padSize;
  const int dataSize = rowSize * (int)height;
  const int fileSize = 54 + dataSize;
  unsigned char header[54];
  memset(header, 0, sizeof(header));
  header[0] = 'B'; header[1] = 'M';
  for (int k = 0; k < 4; ++k) header[2 + k] = (unsigned char)((fileSize >> (8 * k)) & 0xFF);
  header[10] = 54;
  header[14] = 40;
  for (int k = 0; k < 4; ++k) header[18 + k] = (unsigned char)((width >> (8 * k)) & 0xFF);
  for (int k = 0; k < 4; ++k) header[22 + k] = (unsigned char)((height >> (8 * k)) & 0xFF);
  header[26] = 1;
  header[28] = 24;
  bool ok = fwrite(header, 1, 54, f) == 54;
  std::vector<unsigned char> row((size_t)rowSize, 0);
  for (int y = (int)height - 1; y >= 0 && ok; --y) {  // bottom-up, BGR
    const uint8_t* src = rgba + (size_t)y * width * 4;
    for (uint32_t x = 0; x < width; ++x) {
      row[3 * x + 0] = src[4 * x + 2];
      row[3 * x + 1] = src[4 * x + 1];
      row[3 * x + 2] = src[4 * x + 0];
    }
    ok = fwrite(row.data(), 1, (size_t)rowSize, f) == (size_t)rowSize;
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? RR_OK : RR_ERR_IO;
}

}  // extern "C"
