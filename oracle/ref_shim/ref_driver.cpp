// oracle/ref_shim/ref_driver.cpp -- TEST INFRASTRUCTURE (Oracle A, "the
// reference's own code on the CPU").  Not product code; nothing in the
// product may link or call this.
//
// Builds into oracle/_ref/libref_*.so (git-ignored).  At BUILD time the
// Makefile reads three things straight from /root/reference (nothing is copied
// into the repo or kept after the build):
//   * src/Trace.cl          -> trace_cl_adapted.inc   (4 sed rules, see Makefile)
//   * src/image.hpp:383-449 -> image_scene_part.inc   (setupNextVideoFrame,
//                                                     addCornellBoxToScene: CL-free)
//   * src/readobj.hpp, src/math.hpp, src/settings.hpp  included unmodified.
// What this file adds is only what the reference keeps inside OpenCL calls or
// main(): the scene assembly of src/main.cpp:246-272,298-304,706, the
// Node->GPUNode repack of src/image.hpp:116-125, the host-side alpha=255 of
// src/image.hpp:267-271, and a thread pool over image rows in place of the
// NDRange (one work-item = one pixel, src/Trace.cl:629-631).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <typeinfo>
#include <vector>

#include "../rr_math_ref.h"
#include "cl_host_types.hpp"

// ---- reference host code, unmodified ------------------------------------
#include "settings.hpp"
#include "readobj.hpp"  // pulls math.hpp
#include "image_scene_part.inc"
// The same reference text once more with the compile-time VIDEO_FRAME_COUNT (src/settings.hpp:55) bound to a
// variable, so that setupNextVideoFrame (src/image.hpp:385-390) can be checked for any frame count.
static int g_video_frame_count = 1;
namespace vid {
#undef VIDEO_FRAME_COUNT
#define VIDEO_FRAME_COUNT g_video_frame_count
#include "image_scene_part.inc"
#undef VIDEO_FRAME_COUNT
#define VIDEO_FRAME_COUNT 1
}  // namespace vid

// ---- reference kernel text, compiled as C++ -----------------------------
namespace clk {
#include "cl_kernel_compat.hpp"
#include "trace_cl_adapted.inc"
}  // namespace clk

static_assert(sizeof(::Triangle) == 96 && sizeof(clk::Triangle) == 96, "Triangle layout");
static_assert(sizeof(::MeshInfo) == 112 && sizeof(clk::MeshInfo) == 112, "MeshInfo layout");
static_assert(sizeof(::GPUNode) == 48 && sizeof(clk::Node) == 48, "Node layout");
static_assert(sizeof(::CameraInformation) == 48 && sizeof(clk::CameraInformation) == 48, "Camera layout");
static_assert(sizeof(::RayTracingMaterial) == 64 && sizeof(clk::RayTracingMaterial) == 64, "Material layout");
static_assert(offsetof(::MeshInfo, material) == 48 && offsetof(::MeshInfo, scale) == 44, "MeshInfo offsets");

namespace {
std::vector<clk::Triangle> g_tris;
std::vector<clk::MeshInfo> g_meshes;
std::vector<clk::Node> g_nodes;

template <class F> void parallel_rows(int H, int nthreads, F&& body) {
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  auto work = [&]() {
    for (;;) {
      int y = next.fetch_add(1);
      if (y >= H) break;
      body(y);
    }
  };
  for (int i = 1; i < nthreads; ++i) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
}
}  // namespace

extern "C" {

void ref_scene_reset() {
  meshCaches.clear();
  meshList.clear();
  triangleList.clear();
  nodeList.clear();
  g_tris.clear();
  g_meshes.clear();
  g_nodes.clear();
}

// Scene assembly exactly as the reference's main() does it.
int ref_scene_default(const char* obj_path) {
  ref_scene_reset();
  MeshInfo mesh = loadMeshFromOBJFile(obj_path);  // src/main.cpp:246
  mesh.material = {
      .type = MaterialType_Solid,
      .ior = 1.0f,
      .color = {1.0f, 1.0f, 1.0f},
      .emissionColor = {0.0f, 0.0f, 0.0f},
      .emissionStrength = 0.0f,
      .reflectiveness = 0.0f,
      .specularProbability = 1.0f,
  };                                    // src/main.cpp:256-264
  mesh.scale = 0.5f;                    // src/main.cpp:266
  addCornellBoxToScene(mesh);           // src/main.cpp:272
  meshList.emplace_back(mesh);          // src/main.cpp:298
  CameraInformation cam{};
  setupNextVideoFrame(cam, 0);          // src/main.cpp:706

  // src/image.hpp:116-125 (Node -> GPUNode)
  std::vector<GPUNode> gpuNodes(nodeList.size());
  for (size_t i = 0; i < nodeList.size(); ++i) {
    GPUNode n;
    n.bounds = nodeList[i].bounds;
    n.index = nodeList[i].childIndex == 0 ? nodeList[i].firstTriangleIdx : nodeList[i].childIndex;
    n.numTriangles = nodeList[i].childIndex == 0 ? nodeList[i].numTriangles : 0;
    gpuNodes[i] = n;
  }
  g_tris.resize(triangleList.size());
  g_meshes.resize(meshList.size());
  g_nodes.resize(gpuNodes.size());
  if (!g_tris.empty()) memcpy((void*)g_tris.data(), triangleList.data(), g_tris.size() * 96);
  if (!g_meshes.empty()) memcpy((void*)g_meshes.data(), meshList.data(), g_meshes.size() * 112);
  if (!g_nodes.empty()) memcpy((void*)g_nodes.data(), gpuNodes.data(), g_nodes.size() * 48);
  return 0;
}

// Arbitrary scene in the reference's upload layouts (what generateBuffers copies).
int ref_scene_set(const void* tris, size_t ntris, const void* meshes, size_t nmeshes, const void* gpunodes, size_t nnodes) {
  g_tris.resize(ntris);
  g_meshes.resize(nmeshes);
  g_nodes.resize(nnodes);
  if (ntris) memcpy((void*)g_tris.data(), tris, ntris * 96);
  if (nmeshes) memcpy((void*)g_meshes.data(), meshes, nmeshes * 112);
  if (nnodes) memcpy((void*)g_nodes.data(), gpunodes, nnodes * 48);
  return 0;
}

// Scene from caller arrays, hierarchy built by the REFERENCE's own SplitBVH
// (src/readobj.hpp:206-267).  For each mesh the root node is formed exactly as
// loadMeshFromOBJFile does after parsing (src/readobj.hpp:346-367: childIndex =
// size+1, bounds grown from the struct defaults, SplitBVH(root, 64)); meshes of
// <= 2 triangles get the single leaf node addQuad makes (src/readobj.hpp:379-392).
// ranges: (firstTriangle, numTriangles) pairs, uint64 each.
int ref_scene_from_arrays(const void* tris, size_t ntris, const void* meshes, const unsigned long long* ranges, size_t nmeshes) {
  ref_scene_reset();
  triangleList.resize(ntris);
  if (ntris) memcpy((void*)triangleList.data(), tris, ntris * 96);
  meshList.resize(nmeshes);
  if (nmeshes) memcpy((void*)meshList.data(), meshes, nmeshes * 112);
  for (size_t m = 0; m < nmeshes; ++m) {
    size_t first = ranges[2 * m], count = ranges[2 * m + 1];
    Node c;
    c.firstTriangleIdx = first;
    c.numTriangles = count;
    c.childIndex = count <= 2 ? 0 : nodeList.size() + 1;
    for (size_t t = 0; t < count; ++t) GrowToInclude(c.bounds, triangleList[first + t]);
    size_t rootIdx = nodeList.size();
    nodeList.emplace_back(c);
    if (count > 2) SplitBVH(rootIdx, 64);
    meshList[m].nodeIdx = rootIdx;
  }
  std::vector<GPUNode> gpuNodes(nodeList.size());
  for (size_t i = 0; i < nodeList.size(); ++i) {  // src/image.hpp:116-125
    GPUNode n;
    n.bounds = nodeList[i].bounds;
    n.index = nodeList[i].childIndex == 0 ? nodeList[i].firstTriangleIdx : nodeList[i].childIndex;
    n.numTriangles = nodeList[i].childIndex == 0 ? nodeList[i].numTriangles : 0;
    gpuNodes[i] = n;
  }
  g_tris.resize(triangleList.size());
  g_meshes.resize(meshList.size());
  g_nodes.resize(gpuNodes.size());
  if (!g_tris.empty()) memcpy((void*)g_tris.data(), triangleList.data(), g_tris.size() * 96);
  if (!g_meshes.empty()) memcpy((void*)g_meshes.data(), meshList.data(), g_meshes.size() * 112);
  if (!g_nodes.empty()) memcpy((void*)g_nodes.data(), gpuNodes.data(), g_nodes.size() * 48);
  return 0;
}

// setupNextVideoFrame of the reference on the current mesh list (src/main.cpp:691, 706).
int ref_video_frame_setup(int frameIndex, int frameCount) {
  if (g_meshes.empty() || frameCount <= 0) return 1;
  meshList.resize(g_meshes.size());
  memcpy((void*)meshList.data(), g_meshes.data(), g_meshes.size() * 112);
  g_video_frame_count = frameCount;
  CameraInformation cam{};
  vid::setupNextVideoFrame(cam, frameIndex);
  memcpy((void*)g_meshes.data(), meshList.data(), g_meshes.size() * 112);
  return 0;
}

size_t ref_count(int what) { return what == 0 ? g_tris.size() : what == 1 ? g_meshes.size() : g_nodes.size(); }

void ref_copy(int what, void* out) {
  if (what == 0 && !g_tris.empty()) memcpy(out, g_tris.data(), g_tris.size() * 96);
  if (what == 1 && !g_meshes.empty()) memcpy(out, g_meshes.data(), g_meshes.size() * 112);
  if (what == 2 && !g_nodes.empty()) memcpy(out, g_nodes.data(), g_nodes.size() * 48);
}

// src/main.cpp:299-304 with the settings.hpp start pose.
void ref_default_camera(void* cam48, int W, int H) {
  CameraInformation cam = {.position = {CAMERA_START_X, CAMERA_START_Y, CAMERA_START_Z},
                           .pitch = CAMERA_START_PITCH,
                           .yaw = CAMERA_START_YAW,
                           .roll = CAMERA_START_ROLL,
                           .fov = 90.0f,
                           .aspectRatio = (float)W / (float)H};
  memcpy(cam48, &cam, 48);
}

void ref_default_settings(unsigned* out5) {
  out5[0] = WIDTH;
  out5[1] = HEIGHT;
  out5[2] = RAYS_PER_PIXEL;
  out5[3] = MAX_BOUNCE_COUNT;
  out5[4] = TILE_SIZE;
}

// Runs the reference kernel `raytrace` over the whole frame.  rgba gets what
// the host-side `pixels` vector holds after renderTile (alpha forced to 255,
// src/image.hpp:267-271).  If `radiance` is non-null it also receives the
// float accumulator mean (accum / spp, i.e. the value at src/Trace.cl:643),
// obtained by re-running the kernel body's own helpers (MakeSeed, MakeRay,
// Trace) -- the kernel itself only stores 8-bit output.
int ref_render(const void* cam48, int W, int H, unsigned spp, unsigned bounces, int frameIndex, unsigned char* rgba,
               float* radiance, int nthreads) {
  clk::CameraInformation cam;
  memcpy((void*)&cam, cam48, 48);
  const clk::MeshInfo* meshes = g_meshes.data();
  const clk::Triangle* tris = g_tris.data();
  const clk::Node* nodes = g_nodes.data();
  const int meshCount = (int)g_meshes.size();
  parallel_rows(H, nthreads, [&](int y) {
    for (int x = 0; x < W; ++x) {
      clk::g_global_id[0] = (size_t)x;
      clk::g_global_id[1] = (size_t)y;
      clk::raytrace(meshes, tris, meshCount, (clk::uchar4*)rgba, W, H, cam, frameIndex, nodes, bounces, spp);
      rgba[((size_t)y * W + x) * 4 + 3] = 255;
      if (radiance) {
        clk::uint pixelIndex = (clk::uint)y * W + (clk::uint)x;
        clk::uint rng = clk::MakeSeed(pixelIndex, frameIndex, 0);
        clk::float2 uv = clk::mk2((float)(clk::uint)x / (float)W, (float)(1.0f - (clk::uint)y / (float)H));
        clk::Ray ray = clk::MakeRay(cam, uv);
        clk::float3 accum = clk::mk3(0.0f, 0.0f, 0.0f);
        for (unsigned s = 0; s < spp; ++s) accum += clk::Trace(ray, &rng, meshes, meshCount, tris, nodes, bounces);
        clk::float3 c = accum / (float)spp;
        float* o = radiance + ((size_t)y * W + x) * 3;
        o[0] = c.x;
        o[1] = c.y;
        o[2] = c.z;
      }
    }
  });
  return 0;
}

// Primary-ray closest hit as the reference computes it
// (MakeRay + CalculateRayCollisionWithTriangle, src/Trace.cl:596-621,434-485).
// out: 8 floats per pixel = didHit(0/1), dst, hitPoint.xyz, normal.xyz ;
// flags: 1 int per pixel = isBackface | material.type << 8 (or -1 on miss).
int ref_primary(const void* cam48, int W, int H, float* out, int* flags, int nthreads) {
  clk::CameraInformation cam;
  memcpy((void*)&cam, cam48, 48);
  parallel_rows(H, nthreads, [&](int y) {
    for (int x = 0; x < W; ++x) {
      clk::float2 uv = clk::mk2((float)(clk::uint)x / (float)W, (float)(1.0f - (clk::uint)y / (float)H));
      clk::Ray ray = clk::MakeRay(cam, uv);
      clk::HitInfo h =
          clk::CalculateRayCollisionWithTriangle(ray, g_meshes.data(), (int)g_meshes.size(), g_tris.data(), g_nodes.data());
      float* o = out + ((size_t)y * W + x) * 8;
      if (!h.didHit) {
        for (int i = 0; i < 8; ++i) o[i] = 0.0f;
        flags[(size_t)y * W + x] = -1;
      } else {
        o[0] = 1.0f;
        o[1] = h.dst;
        o[2] = h.hitPoint.x;
        o[3] = h.hitPoint.y;
        o[4] = h.hitPoint.z;
        o[5] = h.normal.x;
        o[6] = h.normal.y;
        o[7] = h.normal.z;
        flags[(size_t)y * W + x] = (h.isBackface ? 1 : 0) | ((int)h.material.type << 8);
      }
    }
  });
  return 0;
}

// RNG known-answer hooks (src/Trace.cl:158-177, 209-217): returns the state
// after the call and the float produced.
unsigned ref_make_seed(unsigned pixelIndex, int frameIndex, unsigned rayIdx) { return clk::MakeSeed(pixelIndex, frameIndex, rayIdx); }
float ref_random_value(unsigned* state) { return clk::RandomValue(state); }
float ref_rand01(unsigned* state) { return clk::rand01(state); }
void ref_random_direction(unsigned* state, float* out3) {
  clk::float3 d = clk::RandomDirection(state);
  out3[0] = d.x;
  out3[1] = d.y;
  out3[2] = d.z;
}

// BMP writer of the reference (src/math.hpp:117-164), called on a caller's buffer.
void ref_write_bmp(const unsigned char* rgba, int W, int H, const char* path) {
  std::vector<unsigned char> px(rgba, rgba + (size_t)W * H * 4);
  placeImageDataIntoBMP(px, W, H, path);
}

}  // extern "C"
