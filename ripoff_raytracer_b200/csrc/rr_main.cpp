// rr_main.cpp -- gputest_b200: the reference's command-line program (src/main.cpp, the live
// RENDER_AND_GET_OUT path) on top of the C ABI of include/rr_api.h.  SURVEY.md 8f rank 2: same six stdin
// prompts and defaults (src/main.cpp:159-229, src/settings.hpp), device table (src/main.cpp:137-140), scene
// assembly (src/main.cpp:246-272, 298-304; src/image.hpp:385-390), timing line (src/image.hpp:340-344) and
// output.bmp (src/main.cpp:725).  Nothing of the render path lives here: it is ~150 lines of host glue.
//
//   printf '\n\n\n\n\nknight.obj\n' | ./gputest_b200        # empty line = default, as in the reference
//
// The reference's compile-time switches FRAME_TOTAL, VIDEO_FRAME_COUNT and VIDEO_FRAME_OUTPUT_DIR
// (src/settings.hpp:29-31, 52-62) are read from the environment here (RR_FRAME_TOTAL, RR_VIDEO_FRAME_COUNT,
// RR_VIDEO_FRAME_OUTPUT_DIR; defaults 1, 1, "img"), so that the stdin dialogue stays the reference's.  With
// RR_VIDEO_FRAME_COUNT > 1 the video loop the reference has commented out (src/main.cpp:686-704) runs: per frame
// setupNextVideoFrame, a re-pose of the meshes (rr_update_meshes instead of a full generateBuffers), the render
// and <dir>/output_<n>.bmp (the numbering render.sh feeds to ffmpeg).  RR_FRAME_TOTAL > 1 averages that many
// differently seeded frames per image (src/main.cpp:575-582).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
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

// The progress line of the reference's tile loops (src/image.hpp:316-323, 363-377), same text and the same estimate
// (remaining = elapsed * (100 / percent - 1)).  The reference prints it from the loop that enqueues one OpenCL launch
// per tile; here ONE persistent kernel renders the frame while rr_render blocks, so a second host thread polls the
// device's tile counter (rr_render_progress) and prints the line a few times per second.
class ProgressLine {
 public:
  explicit ProgressLine(rr_ctx* ctx) : ctx_(ctx), start_(std::chrono::high_resolution_clock::now()), thread_([this] { run(); }) {}
  // joins the poller and prints the closing line of src/image.hpp:343-344 / 376-377
  void finish(uint64_t tiles_total) {
    stop_.store(true);
    thread_.join();
    std::cout << "\rRendering tile " << tiles_total << " of " << tiles_total << " (100%) " << elapsed_ms() << " ms elapsed; 0 ms remaining"
              << std::endl;
  }

 private:
  unsigned long elapsed_ms() const {
    return (unsigned long)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::high_resolution_clock::now() - start_).count();
  }
  void run() {
    while (!stop_.load()) {
      uint64_t done = 0, total = 0;
      if (rr_render_progress(ctx_, &done, &total) == RR_OK && total > 0 && done > 0) {
        const double percentCompleted = (double)done / (double)total * 100.0;
        const unsigned long millisecondsPassed = elapsed_ms();
        std::cout << "\033[2K\rRendering tile " << done << " of " << total << " (" << percentCompleted << "%) " << millisecondsPassed
                  << "ms elapsed; " << (unsigned long)(millisecondsPassed * ((100.0 / percentCompleted) - 1.0)) << "ms remaining" << std::flush;
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(200));
    }
  }
  rr_ctx* ctx_;
  std::chrono::high_resolution_clock::time_point start_;
  std::atomic<bool> stop_{false};
  std::thread thread_;
};

}  // namespace

int main() {
  int FRAME_TOTAL = 1, VIDEO_FRAME_COUNT = 1;
  std::string VIDEO_FRAME_OUTPUT_DIR = "img";
  if (const char* e = std::getenv("RR_FRAME_TOTAL")) FRAME_TOTAL = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("RR_VIDEO_FRAME_COUNT")) VIDEO_FRAME_COUNT = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("RR_VIDEO_FRAME_OUTPUT_DIR")) VIDEO_FRAME_OUTPUT_DIR = e;
  if (VIDEO_FRAME_COUNT > 1) {  // src/main.cpp:31-50: checked first, before anything is allocated
    namespace fs = std::filesystem;
    if (!fs::exists(VIDEO_FRAME_OUTPUT_DIR)) {
      std::cout << "Output directory for video frames does not exist: " << VIDEO_FRAME_OUTPUT_DIR << std::endl;
      std::cout << "Should one be made automatically in the current working directory? (y/N)\n> " << std::flush;
      char response = 'n';
      std::cin >> response;  // as the reference: the rest of this line then answers the device prompt (empty = default)
      if (response == 'y' || response == 'Y') {
        fs::create_directory(VIDEO_FRAME_OUTPUT_DIR);
      } else {
        std::cout << "Exiting..." << std::endl;
        return 1;
      }
    } else if (!fs::is_empty(VIDEO_FRAME_OUTPUT_DIR)) {
      std::cout << "Output directory for video frames is not empty: " << VIDEO_FRAME_OUTPUT_DIR << std::endl;
      std::cout << "Files will not be overwritten, just in case you have something important in there.\n";
      std::cout << "Please empty it and try again.\nExiting..." << std::endl;
      return 1;
    }
  }
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
  const size_t n_meshes = rr_scene_mesh_count(scene);
  rr_video_frame_setup(rr_scene_mesh(scene, 0), n_meshes, 0, VIDEO_FRAME_COUNT);  // setupNextVideoFrame(camInfo, 0), src/main.cpp:706

  rr_camera cam;  // src/main.cpp:299-304
  rr_default_camera(&cam, WIDTH, HEIGHT);

  std::cout << rr_scene_triangle_count(scene) << " triangles, " << n_meshes << " meshes" << std::endl;
  rc = rr_scene_upload(ctx, scene);  // generateBuffers, src/main.cpp:709-717
  if (rc) die(rc, "rr_scene_upload");

  std::vector<uint8_t> pixels((size_t)WIDTH * HEIGHT * 4);
  // tile_size 0 = library default (8x4 warp tiles): the reference's TILE_SIZE = 512 only bounds the length of
  // one OpenCL launch (src/settings.hpp:44-48) and does not change the image (src/image.hpp:228: seed term 0)
  (void)TILE_SIZE;
  // one image: a single frame with seed term 0 (src/image.hpp:228), or FRAME_TOTAL frames seeded 1.. and averaged
  auto render_image = [&](rr_stats* st) {
    ProgressLine progress(ctx);
    int rc2;
    if (FRAME_TOTAL > 1)
      rc2 = rr_render_progressive(ctx, &cam, WIDTH, HEIGHT, RAYS_PER_PIXEL, MAX_BOUNCE_COUNT, 1, (uint32_t)FRAME_TOTAL, 0, pixels.data(), st);
    else
      rc2 = rr_render_ex(ctx, &cam, WIDTH, HEIGHT, RAYS_PER_PIXEL, MAX_BOUNCE_COUNT, 0, 0, pixels.data(), nullptr, st, 0);
    progress.finish(rc2 == RR_OK ? st->tiles / (uint64_t)FRAME_TOTAL : 0);
    return rc2;
  };
  auto report = [&](const rr_stats& st, double ms) {
    std::cout << "Rendered " << st.tiles << " tiles, " << st.samples << " samples, " << st.rays << " path segments in " << ms
              << " ms (kernel " << st.render_ms << " ms, LBVH build " << st.build_ms << " ms): "
              << st.rays / (st.render_ms * 1e3) << " Mrays/s" << std::endl;
  };
  if (VIDEO_FRAME_COUNT > 1) {  // the loop of src/main.cpp:686-704
    for (int videoFrameIdx = 0; videoFrameIdx < VIDEO_FRAME_COUNT;) {
      rr_video_frame_setup(rr_scene_mesh(scene, 0), n_meshes, videoFrameIdx++, VIDEO_FRAME_COUNT);
      rc = rr_update_meshes(ctx, rr_scene_meshes(scene), n_meshes);
      if (rc) die(rc, "rr_update_meshes");
      std::cout << "Rendering video frame " << videoFrameIdx << " of " << VIDEO_FRAME_COUNT << std::endl;
      const auto t0 = std::chrono::high_resolution_clock::now();
      rr_stats st;
      rc = render_image(&st);
      if (rc) die(rc, "rr_render");
      report(st, std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
      char path[4096];
      rc = rr_video_frame_path(VIDEO_FRAME_OUTPUT_DIR.c_str(), videoFrameIdx, path, sizeof(path));
      if (!rc) rc = rr_write_bmp(path, pixels.data(), WIDTH, HEIGHT);
      if (rc) die(rc, "rr_write_bmp");
    }
    std::cout << "Wrote " << VIDEO_FRAME_COUNT << " frames to " << VIDEO_FRAME_OUTPUT_DIR << std::endl;
  } else {
    const auto t0 = std::chrono::high_resolution_clock::now();  // src/image.hpp:283
    rr_stats st;
    rc = render_image(&st);
    if (rc) die(rc, "rr_render");
    report(st, std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
    rc = rr_write_bmp("output.bmp", pixels.data(), WIDTH, HEIGHT);  // placeImageDataIntoBMP, src/main.cpp:725
    if (rc) die(rc, "rr_write_bmp");
    std::cout << "Wrote output.bmp" << std::endl;
  }
  rr_scene_destroy(scene);
  rr_destroy(ctx);  // src/main.cpp:728-730
  return 0;
}
