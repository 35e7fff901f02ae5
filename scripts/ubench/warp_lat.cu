// Development aid: single-warp dependent-chain latencies of the warp primitives the sampling loop is made of (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/warp_lat scripts/ubench/warp_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
constexpr int N = 512;
__global__ void k(long long* out, int* sink, int distinct) {
  __shared__ int sm[64];
  const int lane = threadIdx.x & 31;
  sm[lane] = (lane + 1) & 31;
  sm[32 + lane] = lane;
  __syncwarp();
  int v = lane, acc = 0;
  long long t0, t1;
  int idx = 0;
#define RUN(slot, ...)                      \
  __syncwarp();                             \
  t0 = clock64();                           \
  _Pragma("unroll 1") for (int i = 0; i < N; ++i) { __VA_ARGS__; } \
  t1 = clock64();                           \
  if (threadIdx.x == 0) out[slot] = t1 - t0; \
  acc += v;
  RUN(0, v = v + i)                                             // loop + IADD
  RUN(1, v = __shfl_sync(FULL, v, (lane + 1) & 31) + i)         // SHFL.IDX (register index) + IADD
  RUN(2, v = __shfl_sync(FULL, v, 3) + i)                       // SHFL.IDX (immediate)
  RUN(3, { int a = __shfl_sync(FULL, v, 1), b = __shfl_sync(FULL, v, 2), c = __shfl_sync(FULL, v, 3), d = __shfl_sync(FULL, v, 4); v = a + b + c + d + i; })
  RUN(4, v = (int)__ballot_sync(FULL, (v + lane) & 1) + i)      // VOTE
  RUN(5, v = __popc((unsigned)v) + i)                           // POPC
  RUN(6, v = (int)__match_any_sync(FULL, (v + lane) % distinct) + i)   // MATCH.ANY
  RUN(7, v = sm[v & 31] + i)                                    // LDS chain
  RUN(8, v = (int)__reduce_add_sync(FULL, (unsigned)v) + i)     // REDUX
  RUN(9, v = (int)__umul64hi((unsigned long long)v << 11, (unsigned long long)(v | 1)) + i)
  RUN(10, { sm[32 + lane] = v; __syncwarp(); v = sm[32 + ((lane + 1) & 31)] + i; __syncwarp(); })   // STS, syncwarp, LDS
  RUN(11, v = __ffs(v | 1024) + i)
  RUN(12, { if (lane == (v & 31)) idx = atomicAdd(&sm[40], 1); v = __shfl_sync(FULL, idx, v & 31) + i; })   // ATOMS w/ return + SHFL
  RUN(13, v = (int)__match_any_sync(FULL, lane < 10 ? (v + lane * 7) & 0xffff : -1) + i)   // MATCH.ANY, 11 groups
  RUN(14, { unsigned long long x = ((unsigned long long)v << 21) | 12345ull; unsigned long long hi = __umul64hi(x << 11, 37ull + (v & 7)); v = (int)hi + i; })
  sink[threadIdx.x] = acc;
}
int main() {
  long long* out; int* sink;
  cudaMalloc(&out, 16 * 8); cudaMalloc(&sink, 32 * 4);
  const char* names[] = {"loop+IADD", "SHFL.IDX reg idx", "SHFL.IDX imm", "4 indep SHFL", "VOTE(ballot)", "POPC", "MATCH.ANY", "LDS chain", "REDUX.SUM", "umul64hi chain", "STS+syncwarp+LDS+syncwarp", "FFS", "ATOMS ret + SHFL", "MATCH.ANY 11 groups", "umul64hi(search)"};
  for (int distinct : {1, 32}) {
    k<<<1, 32>>>(out, sink, distinct); cudaDeviceSynchronize();
    k<<<1, 32>>>(out, sink, distinct); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("distinct=%d\n", distinct);
    for (int i = 0; i < 15; ++i) printf("  %-28s %7.1f cycles/iter (minus loop: %6.1f)\n", names[i], (double)h[i] / N, (double)(h[i] - h[0]) / N);
  }
  return 0;
}
