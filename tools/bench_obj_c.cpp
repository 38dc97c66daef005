#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../include/rr_api.h"
int main(int argc, char** argv) {
  for (int rep = 0; rep < 3; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    rr_obj* o = nullptr;
    int rc = rr_obj_load(argv[1], &o);
    auto t1 = std::chrono::steady_clock::now();
    printf("rc=%d tris=%zu %.3f s\n", rc, rr_obj_triangle_count(o), std::chrono::duration<double>(t1 - t0).count());
    rr_obj_destroy(o);
  }
}
