/*
 * rr_api.h -- C ABI of the B200-native render path (librr_b200.so).
 *
 * This is the drop-in boundary for the render hot path of ripoff-raytracer:
 * the free functions of the reference's host driver (src/image.hpp) that
 * main() calls, plus the POD layouts they upload.  Each entry point names the
 * reference interface it replaces (paths relative to the reference checkout).
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Every function returns an rr_status (0 = RR_OK).  Where the reference
 * prints getCLErrorString(err) and calls exit(1) (e.g. src/image.hpp:33-36,
 * 236-239), this library returns a code; rr_error_string() gives the text and
 * rr_last_error() the detail (CUDA error string) of the calling thread.
 */
#ifndef RR_API_H
#define RR_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Wire format: byte-identical to the reference's host structs
 * (src/readobj.hpp:15-89) and their kernel mirrors (src/Trace.cl:9-74).
 * cl_float3 == cl_float4: 16 bytes, 16-byte aligned (src/math.hpp:4).
 * ---------------------------------------------------------------------- */
#if defined(__GNUC__) || defined(__clang__) || defined(__CUDACC__)
#define RR_ALIGN16 __attribute__((aligned(16)))
#else
#define RR_ALIGN16
#endif

typedef struct RR_ALIGN16 rr_float3 {
  float s[4]; /* x, y, z, (pad) */
} rr_float3;

/* MaterialType, src/readobj.hpp:40-46 / src/Trace.cl:28-34 */
enum {
  RR_MATERIAL_SOLID = 0,
  RR_MATERIAL_CHECKER = 1,
  RR_MATERIAL_INVISIBLE = 2,
  RR_MATERIAL_GLASSY = 3,
  RR_MATERIAL_ONESIDED = 4
};

/* RayTracingMaterial, 64 bytes (src/readobj.hpp:48-56) */
typedef struct RR_ALIGN16 rr_material {
  int32_t type;            /* @0  */
  float ior;               /* @4  */
  float _pad0[2];          /* @8  */
  rr_float3 color;         /* @16 */
  rr_float3 emissionColor; /* @32 */
  float emissionStrength;  /* @48 */
  float reflectiveness;    /* @52 */
  float specularProbability; /* @56 */
  float _pad1;             /* @60 */
} rr_material;

/* Triangle, 96 bytes (src/readobj.hpp:69-73) */
typedef struct RR_ALIGN16 rr_triangle {
  rr_float3 posA, posB, posC;
  rr_float3 normalA, normalB, normalC;
} rr_triangle;

/* MeshInfo, 112 bytes (src/readobj.hpp:75-81).  nodeIdx is the index of the
 * mesh's root in the reference's nodeList; this library reads it only in
 * rr_upload_scene_ref(), to recover the mesh's triangle range. */
typedef struct RR_ALIGN16 rr_mesh {
  uint64_t nodeIdx;     /* @0  */
  uint64_t _pad0;       /* @8  */
  rr_float3 pos;        /* @16 */
  float pitch, yaw, roll; /* @32 */
  float scale;          /* @44 */
  rr_material material; /* @48 */
} rr_mesh;

/* Host-side BVH node of the reference, 64 bytes (src/readobj.hpp:20-25). */
typedef struct RR_ALIGN16 rr_ref_node {
  rr_float3 bmin, bmax;
  uint64_t childIndex;
  uint64_t firstTriangleIdx;
  uint64_t numTriangles;
  uint64_t _pad0;
} rr_ref_node;

/* CameraInformation, 48 bytes (src/readobj.hpp:33-38) */
typedef struct RR_ALIGN16 rr_camera {
  rr_float3 position;
  float pitch, yaw, roll;
  float fov;
  float aspectRatio;
  float _pad0[3];
} rr_camera;

/* Sphere, 96 bytes (src/readobj.hpp:58-62).  The reference declares this
 * struct but its kernel has no sphere routine; sphere rendering is an
 * extension whose semantics are defined by oracle/rr_oracle.c. */
typedef struct RR_ALIGN16 rr_sphere {
  rr_float3 center;     /* @0  */
  float radius;         /* @16 */
  float _pad0[3];
  rr_material material; /* @32 */
} rr_sphere;

/* Triangle range of one mesh inside the uploaded triangle array. */
typedef struct rr_mesh_range {
  uint64_t firstTriangle;
  uint64_t numTriangles;
} rr_mesh_range;

/* ------------------------------------------------------------------------ */
typedef enum rr_status {
  RR_OK = 0,
  RR_ERR_INVALID_ARGUMENT = 1,
  RR_ERR_NO_DEVICE = 2,      /* reference: "Failed to select a usable device" (src/main.cpp:186-189) */
  RR_ERR_CUDA = 3,           /* any CUDA runtime failure; see rr_last_error() */
  RR_ERR_OUT_OF_MEMORY = 4,
  RR_ERR_NO_SCENE = 5,       /* rr_render before rr_upload_scene */
  RR_ERR_BAD_MESH_RANGE = 6, /* mesh range outside the triangle array */
  RR_ERR_BVH_DEPTH = 7,      /* hierarchy deeper than the traversal stack */
  RR_ERR_IO = 8,             /* file could not be opened / parsed */
  RR_ERR_UNSUPPORTED = 9,
  RR_ERR_QUEUE = 10          /* shared tile queue misuse: reset while a frame was in flight, frame larger than the exported one */
} rr_status;

const char* rr_error_string(int status);
/* Detail text of the last failure on the calling thread ("" if none). */
const char* rr_last_error(void);
/* Library/ABI version: major*10000 + minor*100 + patch. */
int rr_version(void);

typedef struct rr_ctx rr_ctx;

/* Replaces generateKernelForDevice (src/image.hpp:30-71; called
 * src/main.cpp:239-244): one context over `n` CUDA devices (ordinals).  The
 * reference JIT-builds its OpenCL program here; this library ships sm_100a
 * SASS and only creates streams.  Fails with RR_ERR_NO_DEVICE when no CUDA
 * device is usable -- there is no CPU fallback. */
int rr_create(const int* cuda_ordinals, int n, rr_ctx** out);

/* Replaces releaseBuffers + releaseKernelContext (src/image.hpp:73-95,
 * 186-209; src/main.cpp:728-730). */
void rr_destroy(rr_ctx* ctx);

int rr_device_count(int* out);
/* Device table line of src/main.cpp:137-140: name, SM count, memory. */
int rr_device_info(int ordinal, char* name, size_t name_len, int* sm_count, uint64_t* mem_bytes);

/* Replaces generateBuffers (src/image.hpp:97-175; called src/main.cpp:711).
 * COPIES the host arrays (CL_MEM_COPY_HOST_PTR semantics) to every device of
 * the context and builds the per-mesh LBVH on the device.  `ranges[i]` is the
 * triangle range of meshes[i].  spheres may be NULL (n_spheres == 0). */
int rr_upload_scene(rr_ctx* ctx, const rr_triangle* tris, size_t n_tris, const rr_mesh* meshes,
                    const rr_mesh_range* ranges, size_t n_meshes, const rr_sphere* spheres, size_t n_spheres);

/* Same, with the reference's exact argument list (triangleList, meshList,
 * nodeList): the triangle range of each mesh is recovered from the subtree
 * under nodes[meshes[i].nodeIdx]; the reference's SAH hierarchy itself is
 * ignored (the device builds its own LBVH). */
int rr_upload_scene_ref(rr_ctx* ctx, const rr_triangle* tris, size_t n_tris, const rr_mesh* meshes, size_t n_meshes,
                        const rr_ref_node* nodes, size_t n_nodes);

/* Indexed upload (the step before the path).  The raw OBJ arrays go to the device and the
 * Triangle array of src/readobj.hpp:69-73 is assembled THERE (one gather kernel), instead of being expanded to
 * 96 bytes per triangle on the host (src/readobj.hpp:313-343) and copied.  positions / normals: x,y,z triples;
 * corners: six 0-based indices per triangle, v0 v1 v2 n0 n1 n2.  Everything after the gather (LBVH build, render)
 * is the same as for rr_upload_scene, and so is the image. */
int rr_upload_scene_indexed(rr_ctx* ctx, const float* positions, size_t n_positions, const float* normals,
                            size_t n_normals, const uint32_t* corners, size_t n_tris, const rr_mesh* meshes,
                            const rr_mesh_range* ranges, size_t n_meshes, const rr_sphere* spheres, size_t n_spheres);

/* OBJ text -> indexed arrays, with the limits of the reference loader lifted (src/readobj.hpp:289-344 accepts only
 * `f a/b/c` or `f a//c` triangles with normals and silently drops a 4th corner): polygons are fan-triangulated,
 * `f a`, `f a/b` (no normals) get one face normal per triangle, negative indices count from the end.  On the
 * dialect the reference accepts, the triangles are the ones rr_scene_load_obj / the reference produce. */
typedef struct rr_obj rr_obj;
int rr_obj_load(const char* path, rr_obj** out);
void rr_obj_destroy(rr_obj* o);
size_t rr_obj_position_count(const rr_obj* o);
size_t rr_obj_normal_count(const rr_obj* o);
size_t rr_obj_triangle_count(const rr_obj* o);
const float* rr_obj_positions(const rr_obj* o);
const float* rr_obj_normals(const rr_obj* o);
const uint32_t* rr_obj_corners(const rr_obj* o);

/* Re-poses the meshes of the uploaded scene.  The reference's video loop calls
 * setupNextVideoFrame and then generateBuffers again for every frame (src/main.cpp:691-693, src/image.hpp:385-390),
 * re-uploading every triangle and node although only MeshInfo.yaw changed.  The hierarchies here live in mesh-local
 * space, so a new position / rotation / scale / material only rewrites the per-mesh records (one small kernel);
 * triangles and LBVH stay in HBM.  n_meshes must equal the uploaded count; nodeIdx is ignored.  The images are
 * the ones a fresh rr_upload_scene with the same arrays would give. */
int rr_update_meshes(rr_ctx* ctx, const rr_mesh* meshes, size_t n_meshes);

/* Counters of one render (exact, from device atomics). */
typedef struct rr_stats {
  uint64_t samples;      /* Trace() calls = W*H*spp                         */
  uint64_t rays;         /* path segments traced (closest-hit queries)      */
  uint64_t stack_overflows; /* always 0: a non-zero count makes the render fail with RR_ERR_BVH_DEPTH */
  uint64_t box_tests;    /* ray/AABB tests                                  */
  uint64_t tri_tests;    /* ray/triangle tests                              */
  uint64_t sphere_tests; /* ray/sphere tests                                */
  uint64_t tiles;        /* tiles whose first pixel this context took from the queue (sums to the tiles of the frame over all ranks) */
  float render_ms;       /* device time of the render kernel(s), CUDA events */
  float build_ms;        /* device time of the last LBVH build              */
  /* warp-scheduler statistics of the instrumented kernel (count_tests != 0): how often each phase
   * (0 pixel, 1 shade, 2 mesh setup, 3 node step, 4 leaf test) ran in a warp, and the lanes active in it */
  uint64_t phase_runs[5];
  uint64_t phase_lanes[5];
  /* drain of the persistent slot pool (instrumented kernel): time between a warp's first failed queue pop (no new
   * pixels) and its exit -- mean and maximum over the warps, milliseconds.  A pixel's samples are serial
   * (src/Trace.cl:632, 639-642), so this tail does not shrink with more GPUs: DESIGN.md section 6. */
  float tail_avg_ms;
  float tail_max_ms;
} rr_stats;

/* Replaces singleThreadedCompute / multiThreadedCompute + renderTile
 * (src/image.hpp:218-381; dispatch src/main.cpp:719-723) and kernel args
 * 4-10 (src/image.hpp:161-172, src/main.cpp:657-676).
 * rgba_out: width*height*4 bytes, row 0 = top, RGBA, alpha = 255
 * (src/image.hpp:267-271).  Blocking.  frameIndex is kernel arg 7, which the
 * reference always evaluates to 0 (src/image.hpp:228).  tile_size == 0 picks
 * the library default (8 x 4 pixels); larger values are honoured up to 32 x 32.
 * The queue hands out the pixels of the frame tile by tile (row-major tiles,
 * row-major pixels inside a tile) and a warp takes as many as it has free path
 * slots at a time, so a tile only sets the ORDER of the pixels (and the unit of
 * rr_render_strided and of the progress line).  The image does not depend on it.
 * Supported geometry range: the closest hit is the brute-force minimum over all primitives (the hierarchy only
 * culls) as long as ray origins -- the camera and every hit point, taken into each mesh's local space, i.e.
 * (origin - pos) / scale -- stay within 10^4 times the mesh's extent; the library switches to a kernel with a
 * per-ray culling slack by itself when a frame leaves the range its build-time box slack covers (16x).  Beyond
 * 10^4 float32 no longer resolves the mesh's triangles from the origin and isolated pixels may differ from the
 * brute-force result (tests/test_gpu_parity.py::test_far_origin_is_bit_exact).
 * max_bounces <= 8388607. */
int rr_render(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_bounces,
              int32_t frame_index, uint32_t tile_size, uint8_t* rgba_out);

/* rr_render plus optional extras: radiance_out (width*height*3 floats, the
 * mean accumulator of src/Trace.cl:643 before clamp/gamma) and stats_out.
 * count_tests != 0 selects the instrumented kernel (box/tri counters). */
int rr_render_ex(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                 uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, uint8_t* rgba_out,
                 float* radiance_out, rr_stats* stats_out, int count_tests);

/* Scheduler knobs of the render kernel (performance only; the image does not depend on them).
 * values[0..4]: vote weight of the phases pixel/shade/setup/node-step/leaf, [5]: the node-step loop keeps
 * running while at least this many lanes can step, [6]: speculative traversal on/off, [7]: persistent
 * CTAs per SM (0 = as many as fit).  n < 8 leaves the rest unchanged; values == NULL restores defaults. */
int rr_set_tuning(rr_ctx* ctx, const uint32_t* values, size_t n);

/* Progress of the render that is running on this context (callable from a second host thread while rr_render blocks;
 * the analogue of the "Rendering tile i of n" line of src/image.hpp:316-323, 363-377): tiles taken from the queue so
 * far and tiles of the frame.  Reads the device's tile counter with an 8-byte copy on a stream of its own. */
int rr_render_progress(rr_ctx* ctx, uint64_t* tiles_popped, uint64_t* tiles_total);

/* Device-resident variant used for kernel-only timing: renders into the
 * context's own frame buffer on the device and does not copy it back.
 * rr_read_frame() fetches it afterwards. */
int rr_render_device(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                     uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, rr_stats* stats_out);
int rr_read_frame(rr_ctx* ctx, uint8_t* rgba_out, size_t bytes);

/* ------------------------------------------------------------------------
 * Progressive mode: the frame-averaging loop of the
 * reference (FRAME_TOTAL, src/settings.hpp:29-31; the live copy of the loop is
 * the viewer's, src/main.cpp:481, 535, 575-582): every frame is rendered with
 * its own seed term (kernel arg 7 = the 1-based frame number), its 8-bit RGB is
 * ADDED to per-pixel integer sums and the displayed image is sum / frames
 * (integer division).  Here the sums stay in HBM (three u32 planes) and one
 * byte-streaming kernel does add + divide, so a frame costs one render plus
 * 32 B/pixel of HBM traffic instead of a read-back and a host loop.
 * ---------------------------------------------------------------------- */
/* Zeroes the sums (numFrames = 0, src/main.cpp:356). */
int rr_accum_reset(rr_ctx* ctx, uint32_t width, uint32_t height);
/* Renders one frame with `frame_index` as the seed term, adds it to the sums and, if rgba_avg_out != NULL,
 * returns the running average (alpha 255).  width / height must match rr_accum_reset. */
int rr_accum_add_frame(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                       uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, uint8_t* rgba_avg_out,
                       rr_stats* stats_out);
/* Frames added since the last reset. */
int rr_accum_frame_count(rr_ctx* ctx, uint32_t* frames_out);
/* Device time (CUDA events) of the accumulation kernel of the last rr_accum_add_frame: 32 B of HBM traffic per pixel. */
int rr_accum_last_ms(rr_ctx* ctx, float* ms_out);
/* reset + n_frames x add_frame with frame_index = first_frame_index + k (the reference counts from 1);
 * rgba_out is the final average.  stats_out sums the frames (render_ms: total kernel time). */
int rr_render_progressive(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                          uint32_t max_bounces, int32_t first_frame_index, uint32_t n_frames, uint32_t tile_size,
                          uint8_t* rgba_out, rr_stats* stats_out);

/* Primary-ray closest hit per pixel (MakeRay + CalculateRayCollisionWithTriangle,
 * src/Trace.cl:596-621, 434-485).  mesh_out: mesh index (spheres: n_meshes),
 * -1 on miss.  prim_out: index of the triangle in the UPLOADED array (or
 * sphere index).  dst_out: world distance.  Any output may be NULL. */
int rr_primary_hits(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, int32_t* mesh_out,
                    int32_t* prim_out, float* dst_out);

/* LBVH read-back for build-order parity (oracle/rr_oracle.c rro_lbvh_build).
 * which = 0: triangles, 1: spheres.  Arrays sized by rr_bvh_size(). Any
 * output may be NULL.
 *   codes[n], order[n]     sorted Morton keys and the primitive index at each sorted slot
 *   left/right[n]          children of inner node i: >= 0 inner index, < 0 ~sorted-slot (leaf)
 *   parent[n]              parent inner index of inner node i (-1 root / unused slot)
 *   bounds[n*6]            min.xyz,max.xyz of inner node i                                  */
int rr_bvh_size(rr_ctx* ctx, int which, uint64_t* n_prims);
int rr_bvh_read(rr_ctx* ctx, int which, uint64_t* codes, uint32_t* order, int32_t* left, int32_t* right,
                int32_t* parent, float* bounds);

/* ------------------------------------------------------------------------
 * Multi-GPU tile queue (replaces the mutex-guarded std::queue of
 * src/image.hpp:280-350).  One process per GPU: rank 0 owns a 64-bit queue
 * counter (frame epoch << 48 | pixels handed out, in tile-major order) and the
 * frame buffer in its HBM and exports CUDA IPC handles; the other ranks import
 * them and their persistent warps take pixels with system-scope atomics (one
 * atomicAdd per warp and refill: exactly as many pixels as it has free path
 * slots, so no rank sits on unstarted pixels when the queue runs dry) and store
 * finished pixels straight into rank 0's frame over NVLink.  handle buffers are
 * RR_IPC_HANDLE_BYTES each.
 * ---------------------------------------------------------------------- */
#define RR_IPC_HANDLE_BYTES 64
/* Rank 0: allocates the shared frame (width*height*4 bytes, an allocation of its own that no later render of this
 * context frees or moves) and exports it with the tile counter.  Exporting again replaces the shared frame: handles
 * given out before are dead and every peer has to import the new ones. */
int rr_queue_export(rr_ctx* ctx, uint32_t width, uint32_t height, uint8_t* queue_handle, uint8_t* frame_handle);
/* Other ranks: width / height must be the exported ones (they bound what rr_render_shared may write). */
int rr_queue_import(rr_ctx* ctx, uint32_t width, uint32_t height, const uint8_t* queue_handle,
                    const uint8_t* frame_handle);
/* Protocol of one shared frame (every rank, in this order):
 *     barrier A   -- every rank has RETURNED from rr_render_shared of the previous frame
 *     rank 0: rr_queue_reset
 *     barrier B   -- the reset is done before anybody pops
 *     every rank: rr_render_shared
 * The counter carries a 16-bit frame epoch (reset k sets epoch k; the k-th rr_render_shared of a rank after its
 * export / import expects epoch k), so a reset that overtakes a kernel still popping, or a rank that skipped a frame,
 * is reported as RR_ERR_QUEUE by the rr_render_shared that saw it instead of painting tiles of the wrong frame. */
int rr_queue_reset(rr_ctx* ctx);
/* Render the tiles this rank manages to pop from the shared queue; pixels go to the shared frame (peer stores over
 * NVLink on the importing ranks).  width*height*4 must not exceed the exported / imported frame (RR_ERR_QUEUE).
 * Returns when this rank's kernel has drained the queue. */
int rr_render_shared(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                     uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, rr_stats* stats_out);
/* Static partition fallback (no peer access): render only tiles t with
 * t % world == rank into the local frame buffer. */
int rr_render_strided(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp,
                      uint32_t max_bounces, int32_t frame_index, uint32_t tile_size, uint32_t rank, uint32_t world,
                      rr_stats* stats_out);
/* Measurement hooks of the cost-ordered queue experiment.
 * rr_render_cost: renders the frame with the instrumented kernel and returns, per pixel, the path segments its spp samples
 * traced (the pixel's cost; the image goes to the context's frame buffer as usual).
 * rr_set_tile_order: the following renders of this context hand out exactly the n_tiles tiles of the table (row-major tile
 * numbers of the frame they will render, 8 x 4 pixels unless a tile_size is passed), in that order -- every rank of a
 * shared-queue frame has to set the same table; n_tiles = 0 restores the row-major order. */
int rr_render_cost(rr_ctx* ctx, const rr_camera* cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_bounces,
                   uint32_t* segments_out);
int rr_set_tile_order(rr_ctx* ctx, const uint32_t* tiles, uint32_t n_tiles);
/* Device pointer of the local frame buffer (for an NCCL gather by the caller). */
int rr_frame_device_ptr(rr_ctx* ctx, uint64_t* ptr_out, uint64_t* bytes_out);

/* ------------------------------------------------------------------------
 * Test hooks (tests/test_gpu_parity.py): the device side of the numerics contract and of the RNG, evaluated on
 * the GPU for arrays of inputs.  fn: 0 cos, 1 sin, 2 log, 3 exp2, 4 powr(x, y), 5 tan.  rr_probe_rng: the seed of
 * `pixel` (src/Trace.cl:158-166), then state / value after 4 x RandomValue and 2 x rand01, then a RandomDirection.
 * ---------------------------------------------------------------------- */
int rr_probe_math(int fn, const float* x, const float* y, float* out, uint64_t n);
int rr_probe_rng(uint32_t pixel, int32_t frame, uint32_t* out_u32_8, float* out_f32_9);
/* Measured roofline denominators on the current device (bench.py): what = 0: FP32 FMA issue rate in TFLOP/s (8
 * independent FFMA chains per thread, 8 CTAs of 256 threads per SM); what = 1: L2 read bandwidth in GB/s (a 32 MB
 * buffer read 64 times with 16-byte loads that bypass L1). */
int rr_probe_peak(int what, float* value_out);

/* ------------------------------------------------------------------------
 * Host helpers that sit either side of the path (same formats as the reference).
 * ---------------------------------------------------------------------- */

/* placeImageDataIntoBMP (src/math.hpp:117-164): 24-bit BGR, bottom-up, rows
 * padded to 4 bytes, 54-byte header.  Unlike the reference (which silently
 * returns) an unopenable file yields RR_ERR_IO. */
int rr_write_bmp(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height);

/* Scene builder mirroring the reference's global triangleList / meshList
 * (src/readobj.hpp:91-94) without the globals. */
typedef struct rr_scene rr_scene;
int rr_scene_create(rr_scene** out);
void rr_scene_destroy(rr_scene* s);
/* loadMeshFromOBJFile (src/readobj.hpp:270-376): `v`, `vn`, `f a/b/c` or
 * `f a//c` triangles.  The mesh is NOT appended to the mesh list (the
 * reference appends it last, src/main.cpp:298): mesh_out receives the
 * MeshInfo defaults, range_out its triangles; add it with rr_scene_add_mesh. */
int rr_scene_load_obj(rr_scene* s, const char* path, rr_mesh* mesh_out, rr_mesh_range* range_out);
/* Local-space bounds of a triangle range (root node bounds, src/readobj.hpp:353-362). */
int rr_scene_range_bounds(const rr_scene* s, const rr_mesh_range* range, float* min3, float* max3);
/* addQuad (src/readobj.hpp:378-408): two triangles + a Solid mesh, appended. */
int rr_scene_add_quad(rr_scene* s, const float* a3, const float* b3, const float* c3, const float* d3,
                      const float* normal3, const float* color3);
/* addCornellBoxToScene (src/image.hpp:401-448) around the given mesh bounds. */
int rr_scene_add_cornell(rr_scene* s, const rr_mesh* mesh, const rr_mesh_range* range);
int rr_scene_add_mesh(rr_scene* s, const rr_mesh* mesh, const rr_mesh_range* range);
int rr_scene_add_triangles(rr_scene* s, const rr_triangle* tris, size_t n, rr_mesh_range* range_out);
int rr_scene_add_sphere(rr_scene* s, const rr_sphere* sphere);
/* Last mesh's material / transform (meshList.back() edits in src/image.hpp:389, 413-421). */
rr_mesh* rr_scene_mesh(rr_scene* s, size_t index);
size_t rr_scene_mesh_count(const rr_scene* s);
size_t rr_scene_triangle_count(const rr_scene* s);
size_t rr_scene_sphere_count(const rr_scene* s);
const rr_triangle* rr_scene_triangles(const rr_scene* s);
const rr_mesh* rr_scene_meshes(const rr_scene* s);
const rr_mesh_range* rr_scene_ranges(const rr_scene* s);
const rr_sphere* rr_scene_spheres(const rr_scene* s);
/* rr_upload_scene with the builder's arrays. */
int rr_scene_upload(rr_ctx* ctx, const rr_scene* s);
/* Default camera of src/main.cpp:299-304 + src/settings.hpp:23-28. */
void rr_default_camera(rr_camera* cam, uint32_t width, uint32_t height);
/* setupNextVideoFrame (src/image.hpp:385-390): the scene change before video frame `frame_index` of
 * `frame_count` (VIDEO_FRAME_COUNT): the LAST mesh's yaw = 2*pi/frame_count * frame_index + 5.5, in float. */
int rr_video_frame_setup(rr_mesh* meshes, size_t n_meshes, int32_t frame_index, int32_t frame_count);
/* Path of video frame number `frame_number` (1-based, src/main.cpp:701): "<dir>/output_<n>.bmp". */
int rr_video_frame_path(const char* dir, int32_t frame_number, char* out, size_t out_len);

#ifdef __cplusplus
}
#endif
#endif /* RR_API_H */
