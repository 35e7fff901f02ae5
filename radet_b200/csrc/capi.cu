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

// Measurement aid: a one-thread kernel that holds `stream` until *flag becomes non-zero (flag lives in pinned, mapped
// host memory; the host opens the gate with a plain store) or until max_wait_ns have passed.  bench.py enqueues a whole
// timed block behind it so that host launch stalls fall outside the CUDA-event window.
namespace radet {
__global__ void stream_gate_kernel(const volatile int* flag, unsigned long long max_wait_ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*flag == 0) {
    __nanosleep(256);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > max_wait_ns) break;
  }
}
}  // namespace radet

extern "C" int radet_stream_gate(const int32_t* flag, int64_t max_wait_ns, void* stream) {
  if (!flag || max_wait_ns <= 0) return RADET_E_BADARG;
  radet::stream_gate_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, (unsigned long long)max_wait_ns);
  cudaError_t e = cudaGetLastError();     // not counted in radet_launch_count(): not part of the hot path
  return e == cudaSuccess ? RADET_OK : (int)e;
}
