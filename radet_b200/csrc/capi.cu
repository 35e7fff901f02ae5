// Library-level entry points of libradet_b200.so (version, launch accounting, grid helpers).
#include "common.cuh"

namespace radet {
std::atomic<uint64_t> g_launch_count{0};
void* g_debug_buf = nullptr;
}

extern "C" const char* radet_version(void) { return "radet_b200 0.1.0 (sm_100a)"; }

extern "C" uint64_t radet_launch_count(void) { return radet::g_launch_count.load(); }

extern "C" int64_t radet_num_points(const radet_grid_t* grid) {
  radet::GridDev g;
  if (radet::make_grid_dev(grid, &g) != RADET_OK) return -1;
  return g.off[g.num_levels];
}

// Development aid (not part of the documented ABI): when set, some kernels write clock64() phase stamps there.
extern "C" void radet_debug_set_buffer(void* device_buffer) { radet::g_debug_buf = device_buffer; }
