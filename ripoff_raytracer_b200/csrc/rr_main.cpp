// rr_main.cpp -- gputest_b200: the reference's command-line program (src/main.cpp, the live
// RENDER_AND_GET_OUT path) on top of the C ABI of include/rr_api.h.  SURVEY.md 8f rank 2: same six stdin
// prompts and defaults (src/main.cpp:159-229, src/settings.hpp), device table (src/main.cpp:137-140), scene
// assembly (src/main.cpp:246-272, 298-304; src/image.hpp:385-390), timing line (src/image.hpp:340-344) and
// output.bmp (src/main.cpp:725).  Nothing of the render path lives here: it is ~150 lines of host glue.
//
//   printf '\n\n\n\n\nknight.obj\n' | ./gputest_b200        # empty line = default, as in the reference
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/rr_api.h"

namespace {

// defaults of src/settings.hpp:34-35, 42-43, 48, 50
unsigned int RAYS_PER_PIXEL = 50, MAX_BOUNCE_COUNT = 50, WIDTH = 512, HEIGHT = 512, TILE_SIZE = 512;
std::string OBJECT_PATH = "knight.obj";

void die(int status, const char* where) {  // the reference prints the error string and exits (src/image.hpp:33-36)
  std::cerr << where << ": " << rr_error_string(status) << " (" << rr_last_error() << ")" << std::endl;
  std::exit(1);
}

// parseDefaultInput (src/math.hpp:182-218): an empty line keeps the default
bool ask_uint(unsigned int* out) {
  std::string line;
  std::getline(std::cin, line);
  if (line.empty()) return true;
  try {
    *out = (unsigned int)std::stoul(line);
    return true;
  } catch (...) {
    return false;
  }
}

}  // namespace

int main() {
  int n_dev = 0;
  if (rr_device_count(&n_dev) != RR_OK || n_dev == 0) {
    std::cerr << "Failed to select a usable device on any platform." << std::endl;  // src/main.cpp:186-189
    return 1;
  }
  for (int i = 0; i < n_dev; ++i) {  // device table, src/main.cpp:137-140
    char name[256];
    int sms = 0;
    uint64_t mem = 0;
    rr_device_info(i, name, sizeof(name), &sms, &mem);
    std::cout << "[" << i << "] " << name << " (" << sms << " SMs, " << (mem >> 20) << " MiB)" << (i == 0 ? " [chosen]" : "") << "\n";
  }
  std::cout << "Enter the device numbers to use, separated by commas. (0-" << n_dev - 1 << ")\n(0) > " << std::flush;
  std::vector<int> devices;
  {
    std::string line, token;
    std::getline(std::cin, line);
    std::stringstream ss(line);
    while (std::getline(ss, token, ',')) {
      try {
        const int idx = std::stoi(token);
        if (idx >= 0 && idx < n_dev) devices.push_back(idx);
        else std::cerr << "Invalid GPU index: " << idx << ". Skipping." << std::endl;
      } catch (...) {
        std::cerr << "Invalid input: " << token << ". Skipping." << std::endl;
      }
    }
    if (devices.empty()) devices.push_back(0);
  }
  std::cout << "Please enter a width, in pixels. For example, 1920, 3840, ...\n(" << WIDTH << ") > " << std::flush;
  if (!ask_uint(&WIDTH)) { std::cerr << "Invalid input for width. Please enter a numeric value.\nExiting..." << std::endl; return 1; }
  std::cout << "Please enter a height, in pixels. For example, 1080, 2160, ...\n(" << HEIGHT << ") > " << std::flush;
  if (!ask_uint(&HEIGHT)) { std::cerr << "Invalid input for height. Please enter a numeric value.\nExiting..." << std::endl; return 1; }
  std::cout << "Please enter how many rays per pixel to shoot. Higher = better quality,but slower. 100-500 is a bare minimum.\n("
            << RAYS_PER_PIXEL << ") > " << std::flush;
  if (!ask_uint(&RAYS_PER_PIXEL)) { std::cerr << "Invalid input for rays per pixel. Please enter a numeric value.\nExiting..." << std::endl; return 1; }
  std::cout << "Please enter the maximum number of bounces per ray. Higher = better quality,but slower, with diminishing returns. "
               "50+ is a good trade-off.\n(" << MAX_BOUNCE_COUNT << ") > " << std::flush;
  if (!ask_uint(&MAX_BOUNCE_COUNT)) { std::cerr << "Invalid input for max bounce count. Please enter a numeric value.\nExiting..." << std::endl; return 1; }
  std::cout << "Please enter path (rel. or abs.) to the .obj file to load.\n(" << OBJECT_PATH << ")> " << std::flush;
  {
    std::string line;
    std::getline(std::cin, line);
    if (!line.empty()) OBJECT_PATH = line;
  }

  rr_ctx* ctx = nullptr;  // generateKernelForDevice per chosen device, src/main.cpp:239-244
  int rc = rr_create(devices.data(), (int)devices.size(), &ctx);
  if (rc) die(rc, "rr_create");

  // scene assembly, src/main.cpp:246-272, 298: OBJ mesh (Solid white, scale 0.5), Cornell box around it, mesh last
  rr_scene* scene = nullptr;
  rc = rr_scene_create(&scene);
  if (rc) die(rc, "rr_scene_create");
  rr_mesh mesh;
  rr_mesh_range range;
  rc = rr_scene_load_obj(scene, OBJECT_PATH.c_str(), &mesh, &range);
  if (rc) die(rc, "rr_scene_load_obj");
  mesh.material.type = RR_MATERIAL_SOLID;
  mesh.material.ior = 1.0f;
  mesh.material.color.s[0] = mesh.material.color.s[1] = mesh.material.color.s[2] = 1.0f;
  mesh.material.emissionColor.s[0] = mesh.material.emissionColor.s[1] = mesh.material.emissionColor.s[2] = 0.0f;
  mesh.material.emissionStrength = 0.0f;
  mesh.material.reflectiveness = 0.0f;
  mesh.material.specularProbability = 1.0f;
  mesh.scale = 0.5f;
  rc = rr_scene_add_cornell(scene, &mesh, &range);
  if (rc) die(rc, "rr_scene_add_cornell");
  rc = rr_scene_add_mesh(scene, &mesh, &range);
  if (rc) die(rc, "rr_scene_add_mesh");
  rr_scene_mesh(scene, rr_scene_mesh_count(scene) - 1)->yaw = 5.5f;  // setupNextVideoFrame, src/image.hpp:385-390

  rr_camera cam;  // src/main.cpp:299-304
  rr_default_camera(&cam, WIDTH, HEIGHT);

  std::cout << rr_scene_triangle_count(scene) << " triangles, " << rr_scene_mesh_count(scene) << " meshes" << std::endl;
  rc = rr_scene_upload(ctx, scene);  // generateBuffers, src/main.cpp:709-717
  if (rc) die(rc, "rr_scene_upload");

  std::vector<uint8_t> pixels((size_t)WIDTH * HEIGHT * 4);
  const auto t0 = std::chrono::high_resolution_clock::now();  // src/image.hpp:283
  rr_stats st;
  // tile_size 0 = library default (8x4 warp tiles): the reference's TILE_SIZE = 512 only bounds the length of
  // one OpenCL launch (src/settings.hpp:44-48) and does not change the image (src/image.hpp:228: seed term 0)
  (void)TILE_SIZE;
  rc = rr_render_ex(ctx, &cam, WIDTH, HEIGHT, RAYS_PER_PIXEL, MAX_BOUNCE_COUNT, 0, 0, pixels.data(), nullptr, &st, 0);
  if (rc) die(rc, "rr_render");
  const auto t1 = std::chrono::high_resolution_clock::now();
  const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  std::cout << "Rendered " << st.tiles << " tiles, " << st.samples << " samples, " << st.rays << " path segments in " << ms
            << " ms (kernel " << st.render_ms << " ms, LBVH build " << st.build_ms << " ms): "
            << st.rays / (st.render_ms * 1e3) << " Mrays/s" << std::endl;

  rc = rr_write_bmp("output.bmp", pixels.data(), WIDTH, HEIGHT);  // placeImageDataIntoBMP, src/main.cpp:725
  if (rc) die(rc, "rr_write_bmp");
  std::cout << "Wrote output.bmp" << std::endl;
  rr_scene_destroy(scene);
  rr_destroy(ctx);  // src/main.cpp:728-730
  return 0;
}
