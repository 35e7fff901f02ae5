// Shared device/host helpers for the radet_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/radet_b200.h"

namespace radet {

extern std::atomic<uint64_t> g_launch_count;
extern void* g_debug_buf;  // optional device buffer for phase timestamps (development aid, see radet_debug_set_buffer)

#define RADET_LAUNCH_CHECK()                      \
  do {                                            \
    ::radet::g_launch_count.fetch_add(1);         \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Programmatic dependent launch (sm_90+).  A kernel launched with launch_after_trigger() may start once every CTA of the
// preceding kernel in the stream has executed griddep_launch_dependents() (or exited); it must execute griddep_wait()
// before it reads or overwrites anything the preceding kernel touches (the wait returns when that grid has completed and
// its writes are visible).  Both instructions are no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_after_trigger(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool programmatic,
                                        Args&&... args) {
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid;
  lc.blockDim = block;
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = programmatic ? 1u : 0u;
  return cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);
}

// Device-side copy of the grid with prefix offsets (passed by value as a kernel parameter).
struct GridDev {
  int num_levels;
  int h[RADET_MAX_LEVELS], w[RADET_MAX_LEVELS], stride[RADET_MAX_LEVELS];
  int off[RADET_MAX_LEVELS + 1];  // point offset of each level within one image; off[num_levels] = P
  float lo[RADET_MAX_LEVELS], hi[RADET_MAX_LEVELS];
  float anchor_scale, nrm;
};

inline int make_grid_dev(const radet_grid_t* g, GridDev* d) {
  if (!g || g->num_levels <= 0 || g->num_levels > RADET_MAX_LEVELS) return RADET_E_BADARG;
  d->num_levels = g->num_levels;
  int64_t off = 0;
  for (int l = 0; l < g->num_levels; ++l) {
    if (g->level_h[l] <= 0 || g->level_w[l] <= 0 || g->stride[l] <= 0) return RADET_E_BADARG;
    d->h[l] = g->level_h[l];
    d->w[l] = g->level_w[l];
    d->stride[l] = g->stride[l];
    d->lo[l] = g->range_lo[l];
    d->hi[l] = g->range_hi[l];
    d->off[l] = (int)off;
    off += (int64_t)g->level_h[l] * g->level_w[l];
    if (off > (1ll << 30)) return RADET_E_BADARG;
  }
  for (int l = g->num_levels; l <= RADET_MAX_LEVELS; ++l) d->off[l] = (int)off;
  for (int l = g->num_levels; l < RADET_MAX_LEVELS; ++l) {
    d->h[l] = d->w[l] = 0;
    d->stride[l] = 1;
    d->lo[l] = d->hi[l] = 0.f;
  }
  d->anchor_scale = g->anchor_scale;
  d->nrm = g->tblr_normalizer;
  if (!(g->anchor_scale > 0.f) || !(g->tblr_normalizer > 0.f)) return RADET_E_BADARG;
  return RADET_OK;
}

__device__ __forceinline__ int level_of(const GridDev& g, int p) {
  int l = 0;
#pragma unroll
  for (int k = 1; k < RADET_MAX_LEVELS; ++k) l += (k < g.num_levels && p >= g.off[k]) ? 1 : 0;
  return l;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Exclusive prefix sum of one int per thread over the whole CTA (blockDim.x multiple of 32, <= 1024).
// `scratch` holds 33 ints.  Returns the exclusive prefix; *total receives the CTA-wide sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* scratch, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // scratch reuse across calls
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < nw ? scratch[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(kFull, si, o);
      if (lane >= o) si += t;
    }
    scratch[lane] = si - s;  // exclusive warp offsets
    if (lane == 31) scratch[32] = si;
  }
  __syncthreads();
  *total = scratch[32];
  return scratch[wid] + inc - v;
}

// TMA (bulk async copy engine) 1-D global -> shared staging with an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// float -> uint32 whose unsigned order equals torch's ascending float order (NaN largest).
__device__ __forceinline__ uint32_t float_order_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (f != f) return 0xffffffffu;  // NaN sorts as the largest value (torch.sort semantics)
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace radet
