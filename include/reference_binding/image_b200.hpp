// image_b200.hpp -- the binding a maintainer of ripoff-raytracer adds to use librr_b200.so.
// It replaces the OpenCL half of the reference's src/image.hpp (lines 11-381); the CL-free half
// (setupNextVideoFrame, addCornellBoxToScene, lines 383-449) stays as it is.  Include it from src/main.cpp AFTER
// settings.hpp and readobj.hpp (it uses their WIDTH / HEIGHT / RAYS_PER_PIXEL / MAX_BOUNCE_COUNT and their structs).
// tests/test_host.py::test_reference_binding_compiles_against_the_reference_headers compiles this file against the
// reference's own headers, so the static_asserts below are checked, not prose.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "rr_api.h"

static_assert(sizeof(Triangle) == sizeof(rr_triangle) && alignof(Triangle) == alignof(rr_triangle), "Triangle wire format");
static_assert(sizeof(MeshInfo) == sizeof(rr_mesh) && offsetof(MeshInfo, pos) == offsetof(rr_mesh, pos) &&
                  offsetof(MeshInfo, pitch) == offsetof(rr_mesh, pitch) && offsetof(MeshInfo, scale) == offsetof(rr_mesh, scale) &&
                  offsetof(MeshInfo, material) == offsetof(rr_mesh, material),
              "MeshInfo wire format");
static_assert(sizeof(RayTracingMaterial) == sizeof(rr_material) && offsetof(RayTracingMaterial, color) == offsetof(rr_material, color) &&
                  offsetof(RayTracingMaterial, emissionColor) == offsetof(rr_material, emissionColor) &&
                  offsetof(RayTracingMaterial, emissionStrength) == offsetof(rr_material, emissionStrength) &&
                  offsetof(RayTracingMaterial, reflectiveness) == offsetof(rr_material, reflectiveness) &&
                  offsetof(RayTracingMaterial, specularProbability) == offsetof(rr_material, specularProbability),
              "RayTracingMaterial wire format");
static_assert(sizeof(Node) == sizeof(rr_ref_node) && offsetof(Node, childIndex) == offsetof(rr_ref_node, childIndex) &&
                  offsetof(Node, firstTriangleIdx) == offsetof(rr_ref_node, firstTriangleIdx) &&
                  offsetof(Node, numTriangles) == offsetof(rr_ref_node, numTriangles),
              "Node wire format");
static_assert(sizeof(CameraInformation) == sizeof(rr_camera) && offsetof(CameraInformation, pitch) == offsetof(rr_camera, pitch) &&
                  offsetof(CameraInformation, fov) == offsetof(rr_camera, fov) &&
                  offsetof(CameraInformation, aspectRatio) == offsetof(rr_camera, aspectRatio),
              "CameraInformation wire format");

static inline void rrCheck(int status, const char* where) {  // the reference's convention: print and exit(1)
  if (status != RR_OK) {                                     // (src/image.hpp:33-36, 236-239)
    std::cerr << where << ": " << rr_error_string(status) << " (" << rr_last_error() << ")" << std::endl;
    exit(1);
  }
}

struct KernelContext { rr_ctx* ctx = nullptr; };  // was: cl_context, queue, program, kernel

// generateKernelForDevice(cl_device_id)            src/image.hpp:30-71, called src/main.cpp:239-244
inline KernelContext generateKernelForDevices(const std::vector<int>& cudaOrdinals) {
  KernelContext k;
  rrCheck(rr_create(cudaOrdinals.data(), (int)cudaOrdinals.size(), &k.ctx), "rr_create");
  return k;
}

// generateBuffers(triangleList, meshList, nodeList, ctx, kernel)   src/image.hpp:97-175, src/main.cpp:711
inline void generateBuffers(KernelContext& k, std::vector<Triangle>& tris, std::vector<MeshInfo>& meshes, std::vector<Node>& nodes) {
  // same three vectors; the reference's SAH nodes are read only to recover each mesh's triangle range --
  // the hierarchy itself is rebuilt on the GPU (LBVH)
  rrCheck(rr_upload_scene_ref(k.ctx, (const rr_triangle*)tris.data(), tris.size(), (const rr_mesh*)meshes.data(), meshes.size(),
                              (const rr_ref_node*)nodes.data(), nodes.size()),
          "rr_upload_scene_ref");
}

// singleThreadedCompute / multiThreadedCompute     src/image.hpp:280-381, dispatch src/main.cpp:719-723
// (tiles, the work queue over devices and the per-tile read-back all live inside rr_render)
inline void compute(KernelContext& k, const CameraInformation& cam, unsigned char* pixels) {
  rrCheck(rr_render(k.ctx, (const rr_camera*)&cam, WIDTH, HEIGHT, RAYS_PER_PIXEL, MAX_BOUNCE_COUNT,
                    /*frameIndex, always 0: src/image.hpp:228*/ 0, /*tile: library default*/ 0, pixels),
          "rr_render");
}

// releaseBuffers + releaseKernelContext            src/image.hpp:73-95, 186-209, src/main.cpp:728-730
inline void release(KernelContext& k) {
  rr_destroy(k.ctx);
  k.ctx = nullptr;
}
