// Visibility-guided sample assignment on sm_100a.
//
// Reference semantics: radet/datasets/pipelines/label_assignment.py:57-201 (numpy, one image at a time in a
// DataLoader worker).  Here a whole batch is assigned in two launches:
//
//   assign_pairs    grid (ceil(P/256), B): every point x GT pair test (inside-box, level range, visible bit).
//                   GT boxes (area-sorted) and the image's bit-packed visible masks are staged in shared memory
//                   with one TMA bulk copy each; the result is two bit sets per point (candidate / visible
//                   candidate, bit r = GT of area rank r).
//   assign_resolve  one CTA per image: resolves the sequential min-area claiming as a fixed point over the set of
//                   "fallback" GTs (no visible unclaimed candidate), ranks members, draws the numpy-legacy
//                   weighted choice from an MT19937 stream, writes points_to_gt_index / points_weight.
//
// Why the claiming is a fixed point: GT r (area order) claims N_r = R_r&vis if any(R_r&vis) else R_r, where
// R_r = candidates not claimed by smaller GTs.  Let F = {r : no visible unclaimed candidate}.  Then the owner of
// point p is the lowest set bit of  vis[p] | (cand[p] & F), and r is in F iff no point is owned by r through a
// visible bit.  A(F) = {owners through visible bits} is monotone decreasing in F, so iterating
// F <- ~A(F) from F = {} converges (usually in 2 passes) to the unique solution, which is the sequential one.
#include "common.cuh"

namespace radet {

// ------------------------------------------------------------------------------------------------ mask packing
__global__ void pack_masks_kernel(const uint8_t* __restrict__ src, int64_t num_gt, int src_h, int src_w, int step,
                                  int grid_h, int grid_w, int pitch, uint32_t* __restrict__ bits,
                                  int* __restrict__ status) {
  // one warp per output word: lane i tests sample gx = word*32 + i (general form: any step, any source size)
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t total = num_gt * grid_h * pitch;
  if (warp >= total) return;
  int wx, gy;
  int64_t g;
  if (total < (1ll << 31)) {                                         // 32-bit index arithmetic (the usual case)
    const unsigned w32 = (unsigned)warp, row = w32 / (unsigned)pitch;
    wx = (int)(w32 - row * (unsigned)pitch);
    const unsigned gg = row / (unsigned)grid_h;
    gy = (int)(row - gg * (unsigned)grid_h);
    g = gg;
  } else {
    wx = (int)(warp % pitch);
    gy = (int)((warp / pitch) % grid_h);
    g = warp / ((int64_t)pitch * grid_h);
  }
  const int gx = wx * 32 + lane;
  int v = 0;
  if (gx < grid_w) {
    const int sy = gy * step, sx = gx * step;
    if (sy < src_h && sx < src_w) v = src[(g * src_h + sy) * (int64_t)src_w + sx];
  }
  const unsigned word = __ballot_sync(kFull, v != 0);
  // binary-mask check (cheap, per word): two different non-zero values -> not a visible mask
  const int vmax = __reduce_max_sync(kFull, v);
  const int vmin = __reduce_min_sync(kFull, v ? v : 256);
  if (lane == 0) {
    bits[warp] = word;
    if (status && word && vmin != vmax) atomicOr(status, 1);
  }
}

// Pre-sampled grids (step 1, source = grid, width a multiple of 4): one THREAD per output word, eight 32-bit loads,
// byte-wise SIMD compares.  ~45 instructions per 32 samples instead of a warp's worth.
__global__ void pack_grid_kernel(const uint32_t* __restrict__ src4, unsigned total, int grid_h, int grid_w, int pitch,
                                 uint32_t* __restrict__ bits, int* __restrict__ status) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const unsigned row = t / (unsigned)pitch, wx = t - row * (unsigned)pitch;   // row = g * grid_h + gy
  const int n4 = min(8, (grid_w - (int)wx * 32) >> 2);                       // 32-bit groups of this word
  const uint32_t* p = src4 + ((size_t)row * grid_w >> 2) + wx * 8;
  uint32_t v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = k < n4 ? p[k] : 0u;
  uint32_t word = 0u, mx = 0u, mn = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t nz = __vcmpne4(v[k], 0u);                                  // 0xff per non-zero byte
    const uint32_t x = nz & 0x01010101u;
    word |= ((x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xfu) << (4 * k);
    mx = __vmaxu4(mx, v[k]);
    mn = __vminu4(mn, v[k] | ~nz);                                            // zero bytes count as 255
  }
  bits[t] = word;
  if (status && word) {                                                       // same per-word binary-mask check
    const uint32_t hi = max(max(mx & 0xffu, (mx >> 8) & 0xffu), max((mx >> 16) & 0xffu, mx >> 24));
    const uint32_t lo = min(min(mn & 0xffu, (mn >> 8) & 0xffu), min((mn >> 16) & 0xffu, mn >> 24));
    if (lo != hi) atomicOr(status, 1);
  }
}

// ------------------------------------------------------------------------------------------------ MT19937 (numpy legacy)
struct MtWarp {
  // 624-word key in shared memory, driven by ONE warp.  pos = next unread word (624 = exhausted).
  uint32_t* key;
  int pos;

  __device__ static uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  __device__ static uint32_t tw(uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  // numpy mt19937_seed(): init_genrand recurrence; inherently sequential (lane 0).
  __device__ void seed(uint32_t s, int lane) {
    if (lane == 0) {
#pragma unroll 8
      for (int i = 0; i < 624; ++i) {
        key[i] = s;
        s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1);
      }
    }
    pos = 624;
    __syncwarp();
  }
  // mt19937_gen(): regenerate the block, warp-parallel in three dependent phases.
  __device__ void twist(int lane) {
    for (int base = 0; base < 227; base += 32) {  // key[k] = key[k+397] ^ tw(key[k], key[k+1])
      const int k = base + lane;
      uint32_t v = 0;
      if (k < 227) v = key[k + 397] ^ tw(key[k], key[k + 1]);
      __syncwarp();
      if (k < 227) key[k] = v;
      __syncwarp();
    }
    for (int base = 227; base < 623; base += 32) {  // key[k] = key[k-227] ^ tw(key[k], key[k+1])
      const int k = base + lane;
      uint32_t v = 0;
      if (k < 623) v = key[k - 227] ^ tw(key[k], key[k + 1]);
      __syncwarp();
      if (k < 623) key[k] = v;
      __syncwarp();
    }
    if (lane == 0) key[623] = key[396] ^ tw(key[623], key[0]);
    __syncwarp();
  }
  // Lane i < count receives the i-th of the next `count` doubles of random_sample() (count <= 32).
  __device__ double draw(int count, int lane) { return (double)draw53(count, lane) / 9007199254740992.0; }
  // 53-bit integer X of the next doubles: random_sample() = X / 2^53 with X = (a >> 5) << 26 | (b >> 6)
  __device__ unsigned long long draw53(int count, int lane) {
    const int qa = pos + 2 * lane, qb = qa + 1;
    uint32_t a = 0, b = 0;
    const bool act = lane < count;
    if (act && qa < 624) a = key[qa];
    if (act && qb < 624) b = key[qb];
    __syncwarp();
    if (pos + 2 * count > 624) {
      twist(lane);
      if (act && qa >= 624) a = key[qa - 624];
      if (act && qb >= 624) b = key[qb - 624];
      pos = pos + 2 * count - 624;
    } else {
      pos += 2 * count;
    }
    __syncwarp();
    const uint32_t ha = temper(a) >> 5, hb = temper(b) >> 6;
    return ((unsigned long long)ha << 26) | (unsigned long long)hb;
  }
};

// Pre-tempered view of the stream for the sampling loop: the 53-bit integers of all doubles left in the current
// 624-word block are produced in one parallel sweep (all 32 lanes), so a choice() round fetches its uniforms with a
// single shared-memory load instead of load + temper + combine on its critical path.  Requires an even word
// position (always true for np.random.seed / random_sample streams); odd positions use MtWarp::draw53 directly.
struct MtStream {
  MtWarp mt;
  unsigned long long* xbuf;   // [312]
  int nx, xi;                 // doubles buffered / consumed
  bool fast;

  __device__ void fill(int lane) {
    fast = (mt.pos & 1) == 0;
    nx = fast ? (624 - mt.pos) >> 1 : 0;
    xi = 0;
    for (int i = lane; i < nx; i += 32) {
      const uint32_t a = MtWarp::temper(mt.key[mt.pos + 2 * i]) >> 5, b = MtWarp::temper(mt.key[mt.pos + 2 * i + 1]) >> 6;
      xbuf[i] = ((unsigned long long)a << 26) | (unsigned long long)b;
    }
    __syncwarp();
  }
  // word position of the underlying generator (for handing the state back)
  __device__ int word_pos() const { return fast ? mt.pos + 2 * xi : mt.pos; }
  __device__ unsigned long long draw53(int count, int lane) {
    if (!fast) return mt.draw53(count, lane);
    const int rem = nx - xi;
    unsigned long long X = 0ull;
    if (lane < count && lane < rem) X = xbuf[xi + lane];
    if (count > rem) {           // block exhausted: regenerate, take the rest from the new block
      __syncwarp();
      mt.twist(lane);
      mt.pos = 0;
      fill(lane);
      if (lane >= rem && lane < count) X = xbuf[lane - rem];
      xi = count - rem;
    } else {
      xi += count;
    }
    return X;
  }
};

__global__ void mt19937_uniforms_kernel(const uint32_t* __restrict__ seeds, int n, double* __restrict__ out) {
  __shared__ uint32_t key[624];
  const int lane = threadIdx.x;
  MtWarp mt{key, 624};
  mt.seed(seeds[blockIdx.x], lane);
  double* o = out + (int64_t)blockIdx.x * n;
  for (int base = 0; base < n; base += 32) {
    const int cnt = min(32, n - base);
    const double u = mt.draw(cnt, lane);
    if (lane < cnt) o[base + lane] = u;
  }
}

// np.random.seed(seeds[b]) + first regeneration -> legacy state words [B][625] (pos = 0).  One warp per image; meant
// to be enqueued early / on a side stream: the 624-step recurrence is sequential (~5 us) but depends on nothing.
__global__ void mt19937_seed_kernel(const uint32_t* __restrict__ seeds, uint32_t* __restrict__ states) {
  __shared__ uint32_t key[624];
  const int lane = threadIdx.x;
  MtWarp mt{key, 624};
  mt.seed(seeds[blockIdx.x], lane);
  mt.twist(lane);
  uint32_t* st = states + (int64_t)blockIdx.x * RADET_MT_STATE_WORDS;
  for (int i = lane; i < 624; i += 32) st[i] = key[i];
  if (lane == 0) st[624] = 0u;
}

// ------------------------------------------------------------------------------------------------ pair tests
constexpr int kPairThreads = 256;

// stable ascending-area rank (label_assignment.py:156,170): rank = #{k : a_k < a_g or (a_k == a_g and k < g)}
__device__ __forceinline__ void area_ranks(const float* __restrict__ boxes, int G, float* s_area, int* s_rank2gt) {
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float4 b = *reinterpret_cast<const float4*>(boxes + 4 * g);
    s_area[g] = __fmul_rn(b.z - b.x, b.w - b.y);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float a = s_area[g];
    int r = 0;
    for (int k = 0; k < G; ++k) r += (s_area[k] < a || (s_area[k] == a && k < g)) ? 1 : 0;
    s_rank2gt[r] = g;
  }
  __syncthreads();
}

template <int W32>
__global__ void __launch_bounds__(kPairThreads)
assign_pairs_kernel(GridDev grid, const int* __restrict__ gt_offsets, const float* __restrict__ gt_bboxes,
                    const uint32_t* __restrict__ mask_bits, int mask_h, int mask_pitch, int mask_step, int stage_masks,
                    uint32_t* __restrict__ pair_bits /* [B][P][2*W32] */, const uint32_t* __restrict__ seeds,
                    uint32_t* __restrict__ seeded_states /* [B][625] */, int pair_blocks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.y;
  if ((int)blockIdx.x == pair_blocks) {
    // Extra CTA per image (seeded mode only): np.random.seed(seeds[b]) is a sequential 624-step recurrence; running
    // it here, in the shadow of the pair tests on the other SMs, takes it off assign_resolve's critical path.
    __shared__ uint32_t s_mtkey[624];
    if (threadIdx.x < 32) {
      MtWarp mt{s_mtkey, 624};
      mt.seed(seeds[b], threadIdx.x);
      mt.twist(threadIdx.x);   // first block of the stream
      uint32_t* st = seeded_states + (int64_t)b * RADET_MT_STATE_WORDS;
      for (int i = threadIdx.x; i < 624; i += 32) st[i] = s_mtkey[i];
      if (threadIdx.x == 0) st[624] = 0u;
    }
    return;
  }
  const int P = grid.off[grid.num_levels];
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  const int p = blockIdx.x * kPairThreads + threadIdx.x;
  uint32_t* out = pair_bits + ((int64_t)b * P + p) * (2 * W32);
  if (G <= 0) {
    if (p < P) {
#pragma unroll
      for (int w = 0; w < 2 * W32; ++w) out[w] = 0u;
    }
    return;
  }
  // shared: raw boxes [G][4] (TMA destination) | sorted boxes [G] float4 | area [G] | rank2gt [G] | mbarrier | masks
  const int Gp = (G + 3) & ~3;
  float* s_raw = reinterpret_cast<float*>(smem_raw);
  float4* s_box = reinterpret_cast<float4*>(s_raw + 4 * Gp);
  float* s_area = reinterpret_cast<float*>(s_box + Gp);
  int* s_rank2gt = reinterpret_cast<int*>(s_area + Gp);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_rank2gt + Gp);
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_bar + 2);
  const int words_per_gt = mask_h * mask_pitch;
  const uint32_t box_bytes = (uint32_t)G * 16u;
  const uint32_t mask_bytes = (uint32_t)G * (uint32_t)words_per_gt * 4u;
  // TMA bulk staging needs 16-byte aligned source + size; the per-image slices are (16 B boxes; masks checked on host)
  const uint32_t* img_masks = mask_bits + (int64_t)g0 * words_per_gt;
  const bool tma_masks = stage_masks && (mask_bytes % 16u == 0) && ((reinterpret_cast<uintptr_t>(img_masks) & 15) == 0);
  if (threadIdx.x == 0) {
    mbar_init(s_bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(s_bar, box_bytes + (tma_masks ? mask_bytes : 0u));
    tma_bulk_g2s(s_raw, gt_bboxes + 4 * (int64_t)g0, box_bytes, s_bar);
    if (tma_masks) tma_bulk_g2s(s_mask, img_masks, mask_bytes, s_bar);
  }
  if (stage_masks && !tma_masks) {
    for (int i = threadIdx.x; i < G * words_per_gt; i += kPairThreads) s_mask[i] = img_masks[i];
  }
  const uint32_t* mask_src = stage_masks ? s_mask : img_masks;  // huge G x resolution: read the bits through L2 instead
  mbar_wait(s_bar, 0);
  area_ranks(s_raw, G, s_area, s_rank2gt);
  for (int r = threadIdx.x; r < G; r += kPairThreads) s_box[r] = *reinterpret_cast<const float4*>(s_raw + 4 * s_rank2gt[r]);
  __syncthreads();
  if (p >= P) return;

  const int l = level_of(grid, p);
  const int q = p - grid.off[l];
  const int y = q / grid.w[l], x = q - y * grid.w[l];
  const int s = grid.stride[l];
  const float cx = (float)(x * s), cy = (float)(y * s);
  const float lo = grid.lo[l], hi = grid.hi[l];
  const int ratio = s / mask_step;                               // stride is a multiple of the sample step (checked by the entry)
  const int my = y * ratio, mx = x * ratio;                      // int(cy), int(cx) on the sample grid
  const int mword = my * mask_pitch + (mx >> 5);
  const uint32_t mbit = 1u << (mx & 31);

  uint32_t c[W32], v[W32];
#pragma unroll
  for (int w = 0; w < W32; ++w) c[w] = v[w] = 0u;
#pragma unroll
  for (int w = 0; w < W32; ++w) {
    const int rend = min(G, (w + 1) * 32);
    for (int r = w * 32; r < rend; ++r) {
      const float4 bx = s_box[r];
      const float le = cx - bx.x, ri = bx.z - cx, to = cy - bx.y, bo = bx.w - cy;  // label_assignment.py:65-68
      const float mn = fminf(fminf(le, to), fminf(ri, bo));
      const float mxs = fmaxf(fmaxf(le, to), fmaxf(ri, bo));
      if (mn > 0.01f && mxs >= lo && mxs <= hi) {                                   // :71-75
        c[w] |= 1u << (r & 31);
        if (mask_src[s_rank2gt[r] * words_per_gt + mword] & mbit) v[w] |= 1u << (r & 31);
      }
    }
  }
  if (W32 == 1) {
    *reinterpret_cast<uint2*>(out) = make_uint2(c[0], v[0]);
  } else {
#pragma unroll
    for (int w = 0; w < W32; ++w) {
      out[w] = c[w];
      out[W32 + w] = v[w];
    }
  }
}

// ------------------------------------------------------------------------------------------------ resolve + sample
constexpr int kResolveThreads = 512;
constexpr int kListCap = 16384;  // shared-memory capacity of the candidate-point list

// t = max{c in [0,m] : fl64(c/m) <= x}: position in numpy's normalised cumulative sum of m equal weights
// (cdf[j] = fl64((j+1)*p / (m*p)) = fl64((j+1)/m) exactly, because both products are exact in binary64).
// fl64(c/m) <= x is decided from the sign of the exactly-rounded residual x*m - c (one DFMA); only inside the
// half-ulp band above x is the IEEE division itself evaluated, so the result is identical to numpy's.
__device__ __forceinline__ bool cdf_le(double c, double m, double x) {
  const double s = fma(x, m, -c);
  if (s >= 0.0) return true;                  // c/m <= x  =>  RN(c/m) <= x (x is representable)
  if (s < -2.3e-16 * (m * x)) return false;   // c/m > x (1 + 2^-52): rounds above x
  return __ddiv_rn(c, m) <= x;
}
// Same position from the 53-bit integer X (x = X / 2^53): floor(x*m) exactly by a 64x64 high multiply; the only other
// possible answer is floor+1, and only when (t0+1)/m lies within half an ulp above x, i.e. when the residual
// R = 2^53 - frac(X*m / 2^53)*2^53 is <= m (probability ~m*2^-53): then the exact fp64 test decides.
__device__ __forceinline__ int cdf_search(double x, int m);
__device__ __forceinline__ int cdf_search53(unsigned long long X, int m) {
  const unsigned long long hi = __umul64hi(X << 11, (unsigned long long)m);   // floor(X*m / 2^53)
  const unsigned long long lo = (X << 11) * (unsigned long long)m;            // frac * 2^64
  const unsigned long long R = (0ull - lo) >> 11;                             // (1 - frac) * 2^53 (2^53 when frac == 0)
  if (lo != 0ull && R <= (unsigned long long)m) return cdf_search((double)X / 9007199254740992.0, m);
  return (int)hi;
}
__device__ __forceinline__ int cdf_search(double x, int m) {
  const double dm = (double)m;
  int t = (int)(x * dm);
  t = max(0, min(t, m));
  while (t < m && cdf_le((double)(t + 1), dm, x)) ++t;
  while (t > 0 && !cdf_le((double)t, dm, x)) --t;
  return t;
}

// Out-of-line copies for the sampling fast path: warp 0 runs it alone, one dependent instruction after the other, so
// what it costs is the number of instructions fetched -- the rare branches must not sit in the loop body.
__device__ __noinline__ int cdf_search_slow(unsigned long long X, int m) { return cdf_search((double)X / 9007199254740992.0, m); }
struct StreamRefill {
  unsigned long long X;
  int xi, nx;
};
// the draw that crosses the end of the current 624-word block: regenerate, re-buffer, take the rest from the new block
__device__ __noinline__ StreamRefill stream_draw_slow(uint32_t* key, unsigned long long* xbuf, int xi, int nx, int count, int lane) {
  MtStream ms{MtWarp{key, 624 - 2 * nx}, xbuf, nx, xi, true};
  StreamRefill r;
  r.X = ms.draw53(count, lane);
  r.xi = ms.xi;
  r.nx = ms.nx;
  return r;
}

struct ResolveSmem {
  // fixed part; dynamic arrays follow (see resolve_smem_bytes)
  uint32_t F[8], A[8];
  int scan[34];
  int M;
  int changed;
  int found[RADET_MAX_POSITIVE_NUM];
};

// warps 1..31 of the resolve CTA synchronise among themselves on named barrier 1 (warp 0 runs the RNG bring-up)
constexpr int kWorkers = kResolveThreads - 32;
__device__ __forceinline__ void worker_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory"); }

__host__ __device__ inline size_t resolve_smem_bytes(int maxG, int K, bool list_in_smem) {
  size_t s = sizeof(ResolveSmem);
  s = (s + 15) & ~size_t(15);
  s += 624 * 4 + 312 * 8;                   // MT key + pre-tempered 53-bit stream
  s += (size_t)maxG * 4 * 3;                // area, rank2gt, n_r
  s += (size_t)maxG * K * 4;                // sel_pos
  s += (size_t)maxG * K;                    // sel_cnt
  s += (size_t)maxG * 4;                    // n_sel
  s += (size_t)maxG * RADET_MAX_LEVELS * 4; // remaining candidates per (GT, level)
  s += (size_t)maxG * 4;                    // running member count per GT (member positions, phase 3b)
  s = (s + 15) & ~size_t(15);
  if (list_in_smem) s += (size_t)kListCap * 4 + (size_t)kListCap * 2;
  return s;
}

template <int W32>
__global__ void __maxnreg__(96)
assign_resolve_kernel(GridDev grid, const int* __restrict__ gt_offsets, const float* __restrict__ gt_bboxes,
                      const uint32_t* __restrict__ pair_bits, int maxG, int Kcap /* smem stride, <= 32 */, int Kbase /* positive_num */,
                      int flags /* RADET_ASSIGN_* */,
                      const double* __restrict__ uniforms, int n_uniform, const uint32_t* __restrict__ seeds,
                      uint32_t* __restrict__ mt_states, uint32_t* __restrict__ g_list, uint16_t* __restrict__ g_own,
                      int list_in_smem, int64_t* __restrict__ out_idx, float* __restrict__ out_w,
                      int* __restrict__ consumed, double* __restrict__ weight_sums, int state_writeback,
                      long long* __restrict__ dbg) {
#define RESOLVE_DBG(k) do { if (dbg && lane == 0) dbg[(int64_t)blockIdx.x * 16 + (k)] = clock64(); } while (0)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int P = grid.off[grid.num_levels];
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  int64_t* idx = out_idx + (int64_t)b * P;
  float* wt = out_w + (int64_t)b * P;

  // The worker warps' first batch of candidate words does not depend on anything this CTA computes: issue the loads now,
  // they land while the GT offsets / boxes are fetched and ranked (three dependent round trips to L2 otherwise).
  constexpr int kWW0 = (kResolveThreads - 32) / 32;
  uint32_t pre_any[8];
  {
    const uint32_t* bits0 = pair_bits + (int64_t)b * P * (2 * W32);
    const int span0 = min(P, kWW0 * 1024);
    const int chunk0 = (((span0 + kWW0 - 1) / kWW0) + 31) & ~31;
    const int c00 = (wid - 1) * chunk0;
    const int nit0 = max(0, min(chunk0, span0 - c00) + 31) >> 5;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = c00 + u * 32 + lane;
      pre_any[u] = 0u;
      if (wid > 0 && u < nit0 && p < span0) {
#pragma unroll
        for (int w = 0; w < W32; ++w) pre_any[u] |= bits0[(int64_t)p * (2 * W32) + w];
      }
    }
  }
  ResolveSmem* S = reinterpret_cast<ResolveSmem*>(smem_raw);
  unsigned char* cur = smem_raw + ((sizeof(ResolveSmem) + 15) & ~size_t(15));
  unsigned long long* s_xbuf = reinterpret_cast<unsigned long long*>(cur); cur += 312 * 8;
  uint32_t* s_key = reinterpret_cast<uint32_t*>(cur); cur += 624 * 4;
  float* s_area = reinterpret_cast<float*>(cur); cur += (size_t)maxG * 4;
  int* s_rank2gt = reinterpret_cast<int*>(cur); cur += (size_t)maxG * 4;
  int* s_nr = reinterpret_cast<int*>(cur); cur += (size_t)maxG * 4;
  int* s_selpos = reinterpret_cast<int*>(cur); cur += (size_t)maxG * Kcap * 4;
  int* s_nsel = reinterpret_cast<int*>(cur); cur += (size_t)maxG * 4;
  int* s_lcnt = reinterpret_cast<int*>(cur); cur += (size_t)maxG * RADET_MAX_LEVELS * 4;   // remaining candidates per (GT, level)
  int* s_run = reinterpret_cast<int*>(cur); cur += (size_t)maxG * 4;
  unsigned char* s_selcnt = cur; cur += (size_t)maxG * Kcap;
  cur = smem_raw + (((size_t)(cur - smem_raw) + 15) & ~size_t(15));
  uint32_t* list;
  uint16_t* own;
  if (list_in_smem) {
    list = reinterpret_cast<uint32_t*>(cur);
    own = reinterpret_cast<uint16_t*>(cur + (size_t)kListCap * 4);
  } else {
    list = g_list + (int64_t)b * P;
    own = g_own + (int64_t)b * P;
  }

  if (G <= 0) {  // label_assignment.py:166-167 defaults
    for (int p = tid; p < P; p += kResolveThreads) {
      idx[p] = -1;
      wt[p] = 1.0f;
    }
    if (tid == 0) {
      consumed[b] = 0;
      if (weight_sums) weight_sums[b] = 0.0;
    }
    return;
  }
  area_ranks(gt_bboxes + 4 * (int64_t)g0, G, s_area, s_rank2gt);
  const int balance = flags & RADET_ASSIGN_BALANCE;
  const bool adapt = (flags & RADET_ASSIGN_ADAPT_K) != 0, mult = (flags & RADET_ASSIGN_WEIGHT_BY_PRO) != 0;
  for (int r = tid; r < G; r += kResolveThreads) {
    s_nr[r] = 0;
    s_nsel[r] = 0;
  }
  if (adapt)
    for (int i = tid; i < G * RADET_MAX_LEVELS; i += kResolveThreads) s_lcnt[i] = 0;
  if (tid < 8) S->F[tid] = 0u;
  __syncthreads();

  const uint32_t* bits = pair_bits + (int64_t)b * P * (2 * W32);
  const double* ub = uniforms ? uniforms + (int64_t)b * n_uniform : nullptr;
  MtStream ms{MtWarp{s_key, 624}, s_xbuf, 0, 0, false};
  MtWarp& mt = ms.mt;
  int M = 0;
  const bool pack_pos = P < 65536;          // member positions ride in the upper half of the list entries (phase 3b)
  if (wid == 0) RESOLVE_DBG(0);
  if (wid == 0) {
    // ---- warp 0 (specialised): bring up the MT19937 state while the worker warps resolve the claiming.
    // np.random.seed() is an inherently sequential 624-step recurrence (~6 us); it is fully hidden here.
    if (!ub) {
      if (mt_states) {
        const uint32_t* st = mt_states + (int64_t)b * RADET_MT_STATE_WORDS;
        uint32_t tmp[20];
#pragma unroll
        for (int k = 0; k < 20; ++k) tmp[k] = (lane + 32 * k < 624) ? st[lane + 32 * k] : 0u;   // all loads in flight
        mt.pos = (int)st[624];
#pragma unroll
        for (int k = 0; k < 20; ++k)
          if (lane + 32 * k < 624) s_key[lane + 32 * k] = tmp[k];
        __syncwarp();
      } else {
        mt.seed(seeds[b], lane);
      }
      if (mt.pos >= 624) {  // first block of the stream
        mt.twist(lane);
        mt.pos = 0;
      }
      ms.fill(lane);
    }
    RESOLVE_DBG(1);
  } else {
    // ---- worker warps 1.. (kWorkers threads, named barrier 1)
    const int wt_ = tid - 32;
    // 1. ordered compaction of points that are a candidate of at least one GT; everything else keeps the defaults.
    //    Worker warp ww owns a contiguous chunk of the points and walks it 32 points at a time (coalesced loads and
    //    default stores, independent iterations); lane `it` keeps the ballot word of iteration `it`, so one tile
    //    covers up to kWW * 1024 points with a single cross-warp scan.
    const int ww = wid - 1;
    constexpr int kWW = kWorkers / 32;
    for (int base = 0; base < P; base += kWW * 1024) {
      const int span = min(P - base, kWW * 1024);
      const int chunk = (((span + kWW - 1) / kWW) + 31) & ~31;     // points per warp, multiple of 32, <= 1024
      const int c0 = base + ww * chunk;
      const int nit = max(0, min(chunk, base + span - c0) + 31) >> 5;
      uint32_t myword = 0u;
      for (int it0 = 0; it0 < nit; it0 += 8) {                      // 8 iterations' loads in flight
        uint32_t any[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = c0 + (it0 + u) * 32 + lane;
          any[u] = 0u;
          if (base == 0 && it0 == 0) {
            any[u] = pre_any[u];                                    // loaded at the top of the kernel
          } else if (it0 + u < nit && p < base + span) {
#pragma unroll
            for (int w = 0; w < W32; ++w) any[u] |= bits[(int64_t)p * (2 * W32) + w];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = c0 + (it0 + u) * 32 + lane;
          if (it0 + u < nit && p < base + span && !any[u]) {
            idx[p] = -1;
            wt[p] = 1.0f;
          }
          const uint32_t bal = __ballot_sync(kFull, any[u] != 0u);
          if (lane == it0 + u) myword = bal;
        }
      }
      // exclusive prefix of the per-iteration counts inside the warp, then across the worker warps
      const int cnt = __popc(myword);
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
      }
      worker_barrier();                                             // S->scan reuse across tiles
      if (lane == 31) S->scan[ww] = inc;
      worker_barrier();
      const int wtot = lane < kWW ? S->scan[lane] : 0;
      const int before = __reduce_add_sync(kFull, lane < ww ? wtot : 0);
      const int total = __reduce_add_sync(kFull, wtot);
      const int ex = M + before + inc - cnt;
      const int wsum = __shfl_sync(kFull, inc, 31);                 // candidates of this warp's chunk
      if (wsum <= 4 * nit) {
        // sparse (the usual case: a few percent of the points are candidates): lane `it` walks the set bits of its own word
        uint32_t wbits = myword;
        int o = ex;
        while (wbits) {
          const int bpos = __ffs((int)wbits) - 1;
          wbits &= wbits - 1u;
          list[o++] = (uint32_t)(c0 + lane * 32 + bpos);
        }
      } else {
        for (int it = 0; it < nit; ++it) {
          const uint32_t bal = __shfl_sync(kFull, myword, it);
          const int off = __shfl_sync(kFull, ex, it);
          if ((bal >> lane) & 1u) list[off + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)(c0 + it * 32 + lane);
        }
      }
      M += total;
    }
    worker_barrier();
    if (wid == 1) RESOLVE_DBG(11);

    // 2. fixed point over the fallback set F
    for (int round = 0; round <= G; ++round) {
      if (wt_ < 8) S->A[wt_] = 0u;
      if (wt_ == 0) S->changed = 0;
      worker_barrier();
      uint32_t acc[W32];
#pragma unroll
      for (int w = 0; w < W32; ++w) acc[w] = 0u;
      for (int e = wt_; e < M; e += kWorkers) {
        const uint32_t* pb = bits + (int64_t)list[e] * (2 * W32);
#pragma unroll
        for (int w = 0; w < W32; ++w) {
          const uint32_t c = pb[w], v = pb[W32 + w];
          const uint32_t cl = v | (c & S->F[w]);
          if (cl) {
            const uint32_t low = cl & (0u - cl);
            if (v & low) acc[w] |= low;
            break;
          }
        }
      }
#pragma unroll
      for (int w = 0; w < W32; ++w) {
        const uint32_t r = __reduce_or_sync(kFull, acc[w]);
        if (lane == 0 && r) atomicOr(&S->A[w], r);
      }
      worker_barrier();
      if (wt_ < W32) {
        const int rem = G - wt_ * 32;
        const uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
        const uint32_t nf = ~S->A[wt_] & valid;
        if (nf != S->F[wt_]) {
          S->F[wt_] = nf;
          S->changed = 1;
        }
      }
      worker_barrier();
      const int changed = S->changed;
      worker_barrier();
      if (!changed) break;
    }

    if (wid == 1) RESOLVE_DBG(12);
    // 3. owners, member counts; unclaimed candidate points keep the defaults
    for (int e = wt_; e < M; e += kWorkers) {
      const int p = (int)list[e];
      const uint32_t* pb = bits + (int64_t)p * (2 * W32);
      int r = -1;
#pragma unroll
      for (int w = 0; w < W32; ++w) {
        const uint32_t cl = pb[W32 + w] | (pb[w] & S->F[w]);
        if (cl) {
          r = w * 32 + __ffs((int)cl) - 1;
          break;
        }
      }
      own[e] = (uint16_t)(r < 0 ? 0xffff : r);
      if (adapt) {
        // adapt_cal_k (label_assignment.py:88-95) looks at ALL remaining candidates of a GT at its turn: the point
        // counts for every GT it is a candidate of up to and including its owner
        const int lvl = level_of(grid, p);
#pragma unroll
        for (int w = 0; w < W32; ++w) {
          uint32_t cb = pb[w];
          while (cb) {
            const int g_ = w * 32 + __ffs((int)cb) - 1;
            cb &= cb - 1u;
            if (r < 0 || g_ <= r) atomicAdd(&s_lcnt[g_ * RADET_MAX_LEVELS + lvl], 1);
          }
        }
      }
      if (r >= 0) {
        atomicAdd(&s_nr[r], 1);
      } else {
        idx[p] = -1;
        wt[p] = 1.0f;
      }
    }
    if (wt_ == 0) S->M = M;
    if (wid == 1) RESOLVE_DBG(2);
  }
  __syncthreads();
  M = S->M;
  if (wid == 0) RESOLVE_DBG(3);

  // 4. numpy-legacy weighted choice per GT, in area order, from one sequential MT19937 stream (warp 0).
  //    Lanes = draws of one choice() round; found[] lives in shared memory.
  if (wid == 0) {
    int* s_found = S->found;
    int used = 0;
    bool overflow = false, adapt_too_large = false;
    // next `count` uniforms as 53-bit integers (bit 63 set: not of the form X/2^53 -> use the fp64 search on ux)
    double ux = 0.0;
    auto draw = [&](int count) -> unsigned long long {
      unsigned long long X = 0ull;
      if (ub) {
        if (used + count > n_uniform) overflow = true;
        else if (lane < count) {
          ux = ub[used + lane];
          const double sc = ux * 9007199254740992.0;
          X = (unsigned long long)sc;
          if ((double)X != sc || !(ux >= 0.0 && ux < 1.0)) X = 1ull << 63;
        }
      } else {
        X = ms.draw53(count, lane);
      }
      used += count;
      return X;
    };
    auto search = [&](unsigned long long X, int m) -> int { return (X >> 63) ? cdf_search(ux, m) : cdf_search53(X, m); };
    const bool tight = !ub && !adapt && ms.fast;
    if (tight) {
      // The usual configuration (MT19937 stream buffered in s_xbuf, fixed positive_num).  Warp 0 runs this alone, one
      // dependent instruction after the other, and on this machine every warp-wide primitive is 30-40 cycles of latency
      // (SHFL 38, VOTE 33, POPC 29, FLO/BREV 29 each, LDS 40; MATCH.ANY ~95 + 6 per distinct value: scripts/ubench/warp_lat.cu),
      // so the loop is written to keep them off the critical path: the uniforms of the NEXT draw are fetched as soon as the
      // stream position is known (the next draw starts there whether it is this GT's next round or the next GT's first),
      // the next GT's first search runs in the shadow of this GT's MATCH, a round without collisions ends on a mask
      // compare instead of a population count, the lowest-peer test needs no FLO, and the (f, g) bookkeeping of the
      // later rounds reads its operands as shared-memory broadcasts instead of a chain of shuffles.
      const int K = Kbase;
      const unsigned lt = (1u << lane) - 1u, fullK = K >= 32 ? 0xffffffffu : (1u << K) - 1u;
      int* s_g = S->found;
      int xi = ms.xi, nx = ms.nx;
      auto search_t = [&](unsigned long long X, int m) -> int {
        const unsigned long long Xs = X << 11, um = (unsigned long long)(unsigned)m;   // m < 2^31: two 32 x 32 products
        const unsigned long long p0 = (Xs & 0xffffffffull) * um, p1 = (Xs >> 32) * um + (p0 >> 32);
        const unsigned long long lo = (p1 << 32) | (p0 & 0xffffffffull);            // low 64 bits of Xs * m
        const unsigned long long R = (0ull - lo) >> 11;
        int t = (int)(p1 >> 32);                                                    // floor(X*m / 2^53), see cdf_search53
        if (lo != 0ull && R <= um) t = cdf_search_slow(X, m);
        return t;
      };
      unsigned long long Xc = s_xbuf[min(xi + lane, 311)];     // uniforms at the current stream position
      // consumes `count` uniforms: returns them (lane i < count holds the i-th) and prefetches the following ones
      auto draw_t = [&](int count, bool* refilled) -> unsigned long long {
        unsigned long long X = Xc;
        *refilled = false;
        if (xi + count > nx) {                                  // crosses the end of the 624-word block
          const StreamRefill rf = stream_draw_slow(s_key, s_xbuf, xi, nx, count, lane);
          X = rf.X;
          xi = rf.xi;
          nx = rf.nx;
          *refilled = true;
        } else {
          xi += count;
        }
        used += count;
        Xc = s_xbuf[min(xi + lane, 311)];
        return X;
      };
      int n_next = s_nr[0];
      int t_spec = search_t(Xc, n_next);                        // first search of the next GT that draws
#pragma unroll 1
      for (int r = 0; r < G; ++r) {
        const int n = n_next;
        n_next = r + 1 < G ? s_nr[r + 1] : 0;
        if (n == 0 || (n < K && !balance)) {                    // label_assignment.py:182-183 no RNG use; :116 chosen = arange(n)
          if (n != 0 && lane == 0) s_nsel[r] = -1;
          t_spec = search_t(Xc, n_next);
          continue;
        }
        int* selpos = s_selpos + r * Kcap;
        unsigned char* selcnt = s_selcnt + r * Kcap;
        bool refilled;
        const unsigned long long x = draw_t(K, &refilled);
        const int t = refilled ? search_t(x, n) : t_spec;
        const int pos = lane < K ? t : -1;
        const unsigned peers = __match_any_sync(kFull, pos);    // idle lanes all hold -1
        t_spec = search_t(Xc, n_next);                          // in the shadow of the MATCH
        const bool first = lane < K && (peers & lt) == 0u;
        const unsigned fm = __ballot_sync(kFull, first);
        if (n < K) {                                            // :112 choice(n, K, p, replace=True); np.unique(return_counts), :125
          if (first) {
            const int slot = __popc(fm & lt);
            selpos[slot] = pos;
            selcnt[slot] = (unsigned char)__popc(peers);
          }
          if (lane == 0) s_nsel[r] = __popc(fm);
          continue;
        }
        // :119 choice(n, K, p, replace=False); see the general loop below for the (f, g) bookkeeping
        if (fm == fullK) {                                      // no collision: done after one round
          if (lane < K) {
            selpos[lane] = pos;
            selcnt[lane] = 1;
          }
          if (lane == 0) s_nsel[r] = K;
          continue;
        }
        if (first) selpos[__popc(fm & lt)] = pos;
        int n_uniq = __popc(fm);
#pragma unroll 1
        while (true) {
          __syncwarp();
          const int f = lane < n_uniq ? selpos[lane] : 0x7fffffff;
          int rk = 0;
#pragma unroll 4
          for (int j = 0; j < n_uniq; ++j) rk += selpos[j] < f ? 1 : 0;
          s_g[lane] = lane < n_uniq ? f - rk : 0x7fffffff;
          __syncwarp();
          const int d = K - n_uniq, m = n - n_uniq;
          const unsigned long long x2 = draw_t(d, &refilled);
          const int t2 = search_t(x2, m);                       // idle lanes: result unused
          int c = 0;
#pragma unroll 4
          for (int j = 0; j < n_uniq; ++j) c += s_g[j] <= t2 ? 1 : 0;
          const int pos2 = lane < d ? t2 + c : -1;
          const unsigned peers2 = __match_any_sync(kFull, pos2);
          const bool first2 = lane < d && (peers2 & lt) == 0u;
          const unsigned fm2 = __ballot_sync(kFull, first2);
          if (first2) selpos[n_uniq + __popc(fm2 & lt)] = pos2;
          n_uniq += __popc(fm2);
          if (n_uniq >= K) break;
        }
        if (lane < K) selcnt[lane] = 1;
        if (lane == 0) s_nsel[r] = K;
        t_spec = search_t(Xc, n_next);
      }
      ms.xi = xi;
      ms.nx = nx;
      ms.mt.pos = 624 - 2 * nx;
      __syncwarp();
    }
    for (int r = 0; r < G && !overflow && !tight; ++r) {
      const int n = s_nr[r];
      if (n == 0) continue;                                         // label_assignment.py:182-183: no RNG use
      int K = Kbase;
      if (adapt) {                                                  // :104-105 positive_num = adapt_cal_k(...)
        int kk = 0;
        if (lane == 0) {
          const float4 gb = *reinterpret_cast<const float4*>(gt_bboxes + 4 * (int64_t)(g0 + s_rank2gt[r]));
          const float obj = fmaxf(gb.z - gb.x, gb.w - gb.y);
          int ntot = 0;
          for (int l = 0; l < grid.num_levels; ++l) ntot += s_lcnt[r * RADET_MAX_LEVELS + l];
          double dk = 0.0;
          for (int l = 0; l < grid.num_levels; ++l) {               // np.unique: ascending anchor size = ascending level
            const int c_ = s_lcnt[r * RADET_MAX_LEVELS + l];
            if (c_ == 0) continue;
            const float size = __fmul_rn(grid.anchor_scale, (float)grid.stride[l]);
            const float ex = (float)exp((double)__fdiv_rn(__fsub_rn(obj, size), __fmul_rn(2.f, size)));   // float32 np.exp
            dk += ((double)c_ / (double)ntot) * (double)ex;
          }
          kk = (int)((double)Kbase * dk + 0.5);
        }
        K = __shfl_sync(kFull, kk, 0);
        if (K > Kcap) {                                             // more draws than lanes / slots: not representable here
          overflow = true;
          adapt_too_large = true;
          break;
        }
        if (K == 0) {                                               // choice(size=0): no draws, everything non-neg is ignored
          if (lane == 0) s_nsel[r] = 0;
          continue;
        }
      }
      int* selpos = s_selpos + r * Kcap;
      unsigned char* selcnt = s_selcnt + r * Kcap;
      if (n < K) {
        if (!balance) {                                             // :116 chosen = arange(n)
          if (lane == 0) s_nsel[r] = -1;
          continue;
        }
        const unsigned long long x = draw(K);                       // :112 choice(n, K, p, replace=True)
        if (overflow) break;
        const int c = lane < K ? search(x, n) : -1;
        const unsigned peers = __match_any_sync(kFull, c);          // np.unique(chosen, return_counts=True), :125 (idle lanes share -1)
        const bool first = lane < K && lane == __ffs((int)peers) - 1;
        const unsigned fm = __ballot_sync(kFull, first);
        if (first) {
          const int slot = __popc(fm & ((1u << lane) - 1u));
          selpos[slot] = c;
          selcnt[slot] = (unsigned char)__popc(peers);
        }
        if (lane == 0) s_nsel[r] = __popc(fm);
      } else {                                                      // :119 choice(n, K, p, replace=False)
        // Lane j < n_uniq keeps, in registers, the j-th kept position f (draw order) and g = f - #{kept < f}, the
        // number of not-yet-kept entries below it.  A draw of rank t among the not-yet-kept entries then lands on
        // position t + #{j : g_j <= t} — one pass of shuffles, no fixed-point iteration over found[].
        int n_uniq = 0, f = 0x7fffffff, g = 0x7fffffff;
        while (true) {
          const int d = K - n_uniq, m = n - n_uniq;
          const unsigned long long x = draw(d);
          if (overflow) break;
          const int t = search(x, m);                               // idle lanes hold x = 0: harmless, result unused
          int c = 0;
          for (int j = 0; j < n_uniq; ++j) c += (__shfl_sync(kFull, g, j) <= t) ? 1 : 0;
          const int pos = lane < d ? t + c : -1;
          // keep the first occurrence of each value, in draw order (np.unique(return_index) + sort in choice())
          const unsigned peers = __match_any_sync(kFull, pos);      // idle lanes all hold -1: at most d+1 distinct values
          const bool first = lane < d && lane == __ffs((int)peers) - 1;
          const unsigned fm = __ballot_sync(kFull, first);
          if (first) s_found[n_uniq + __popc(fm & ((1u << lane) - 1u))] = pos;
          __syncwarp();
          n_uniq += __popc(fm);
          if (lane < n_uniq) f = s_found[lane];
          if (n_uniq >= K) break;
          int rk = 0;
          for (int j = 0; j < n_uniq; ++j) rk += (__shfl_sync(kFull, f, j) < f) ? 1 : 0;
          g = lane < n_uniq ? f - rk : 0x7fffffff;
        }
        if (overflow) break;
        if (lane < K) {
          selpos[lane] = f;
          selcnt[lane] = 1;
        }
        if (lane == 0) s_nsel[r] = K;
      }
      __syncwarp();
    }
    if (lane == 0) consumed[b] = adapt_too_large ? -2 : (overflow ? -1 : used);
    RESOLVE_DBG(4);
    if (dbg && lane == 0) {
      dbg[(int64_t)blockIdx.x * 16 + 8] = G;
      dbg[(int64_t)blockIdx.x * 16 + 9] = used;
      dbg[(int64_t)blockIdx.x * 16 + 10] = M;
    }
    if (!ub && mt_states && state_writeback && used > 0) {  // hand the advanced generator back (untouched if nothing was drawn)
      uint32_t* st = mt_states + (int64_t)b * RADET_MT_STATE_WORDS;
      for (int i = lane; i < 624; i += 32) st[i] = s_key[i];
      if (lane == 0) st[624] = (uint32_t)ms.word_pos();
    }
  } else if (wid == 1 && pack_pos) {
    // 3b. (in the shadow of the sampling) the position of every member inside its GT's member list, in ascending point
    //     order -- what choice() returns indices into (label_assignment.py:117-126).  One warp walks the list 32 entries at a
    //     time with a running count per GT; the position rides in the upper half of the list entry (P < 65536).
    for (int r = lane; r < G; r += 32) s_run[r] = 0;
    __syncwarp();
    for (int base = 0; base < M; base += 32) {
      const int e = base + lane;
      const int r = e < M ? (int)own[e] : 0xffff;
      const unsigned peers = __match_any_sync(kFull, r);
      if (r != 0xffff) {
        const int before = s_run[r];
        const int mpos = before + __popc(peers & ((1u << lane) - 1u));
        list[e] = (list[e] & 0xffffu) | ((uint32_t)mpos << 16);
        __syncwarp(peers);
        if (lane == __ffs((int)peers) - 1) s_run[r] = before + __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // 5. selected -> positive, other members -> ignore (label_assignment.py:193-196)
  if (pack_pos) {
    // one thread per member: its position inside the GT's list was prepared in 3b
    for (int e = tid; e < M; e += kResolveThreads) {
      const int r = (int)own[e];
      if (r == 0xffff) continue;
      const uint32_t v = list[e];
      const int p = (int)(v & 0xffffu), mpos = (int)(v >> 16);
      const int nsel = s_nsel[r];
      const int* selpos = s_selpos + r * Kcap;
      const unsigned char* selcnt = s_selcnt + r * Kcap;
      int cnt = nsel < 0 ? 1 : 0;
      for (int j = 0; j < nsel; ++j)
        if (selpos[j] == mpos) cnt = selcnt[j];
      // multiply_sample_pro_for_weight (:127-128): binary masks give pro = 1 for a visible point, clip(min=1e-8) otherwise
      // (a GT sampling from its fallback set has only such points)
      const float pro = (mult && ((S->F[r >> 5] >> (r & 31)) & 1u)) ? 1e-8f : 1.0f;
      idx[p] = cnt > 0 ? s_rank2gt[r] + 1 : 0;
      wt[p] = __fmul_rn((float)cnt, pro);
    }
    if (weight_sums) {
      // sum of the weights written for each GT, from the selection counts (the same terms as the member walk below adds;
      // sums of <= 32 equal-scale float values are exact in fp64, so the order does not matter)
      for (int r = tid; r < G; r += kResolveThreads) {
        double wsum = 0.0;
        if (s_nr[r] != 0) {
          const float pro = (mult && ((S->F[r >> 5] >> (r & 31)) & 1u)) ? 1e-8f : 1.0f;
          const int nsel = s_nsel[r];
          if (nsel < 0) wsum = (double)s_nr[r] * (double)__fmul_rn(1.0f, pro);
          for (int j = 0; j < nsel; ++j) wsum += (double)__fmul_rn((float)s_selcnt[r * Kcap + j], pro);
        }
        reinterpret_cast<double*>(s_xbuf)[r] = wsum;       // the sampling is over: its pre-tempered stream buffer is free
      }
    }
  } else {
  // (P >= 65536) one warp per GT: walk its members in ascending point order
  for (int r = wid; r < G; r += kResolveThreads / 32) {
    if (s_nr[r] == 0) continue;
    const int gt1 = s_rank2gt[r] + 1;
    const int nsel = s_nsel[r];
    const int* selpos = s_selpos + r * Kcap;
    const unsigned char* selcnt = s_selcnt + r * Kcap;
    const float pro = (mult && ((S->F[r >> 5] >> (r & 31)) & 1u)) ? 1e-8f : 1.0f;
    int running = 0;
    double wsum = 0.0;                                    // sum of the weights written for this GT (weight_sums)
    for (int base = 0; base < M; base += 32) {
      const int e = base + lane;
      const bool mine = e < M && own[e] == (uint16_t)r;
      const unsigned ball = __ballot_sync(kFull, mine);
      if (mine) {
        const int mpos = running + __popc(ball & ((1u << lane) - 1u));
        int cnt = 0;
        if (nsel < 0) cnt = 1;
        for (int j = 0; j < nsel; ++j)
          if (selpos[j] == mpos) cnt = selcnt[j];
        const int p = (int)list[e];
        idx[p] = cnt > 0 ? gt1 : 0;          // label_assignment.py:193-194
        const float wv = __fmul_rn((float)cnt, pro);
        wt[p] = wv;                          // :195-196, :127-128
        wsum += (double)wv;
      }
      running += __popc(ball);
    }
    if (weight_sums) {                                    // the sampling is over: its pre-tempered stream buffer is free
      wsum = warp_sum(wsum);
      if (lane == 0) reinterpret_cast<double*>(s_xbuf)[r] = wsum;
    }
  }
  }
  if (weight_sums) {
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int r = 0; r < G; ++r)
        if (s_nr[r] != 0) tot += reinterpret_cast<double*>(s_xbuf)[r];
      weight_sums[b] = tot;
    }
  }
  if (wid == 1) RESOLVE_DBG(5);
}

}  // namespace radet

// ================================================================================================ C ABI
using namespace radet;

extern "C" int radet_pack_masks(const uint8_t* src, int64_t num_gt, int32_t src_h, int32_t src_w, int32_t step,
                                int32_t grid_h, int32_t grid_w, uint32_t* bits, int32_t* status, void* stream) {
  if (num_gt == 0) return RADET_OK;
  if (!src || !bits || num_gt < 0 || src_h <= 0 || src_w <= 0 || step <= 0 || grid_h <= 0 || grid_w <= 0) return RADET_E_BADARG;
  const int pitch = (grid_w + 31) / 32;
  const int64_t warps = num_gt * grid_h * pitch;
  const int threads = 256;
  if (step == 1 && src_h == grid_h && src_w == grid_w && (grid_w & 3) == 0 && warps < (1ll << 31) &&
      (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    pack_grid_kernel<<<(unsigned)((warps + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint32_t*>(src), (unsigned)warps, grid_h, grid_w, pitch, bits, status);
    RADET_LAUNCH_CHECK();
    return RADET_OK;
  }
  const int64_t blocks = (warps * 32 + threads - 1) / threads;
  pack_masks_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(src, num_gt, src_h, src_w, step, grid_h, grid_w,
                                                                           pitch, bits, status);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_mt19937_uniforms(const uint32_t* seeds, int32_t batch, int32_t n, double* out, void* stream) {
  if (batch == 0 || n == 0) return RADET_OK;
  if (!seeds || !out || batch < 0 || n < 0) return RADET_E_BADARG;
  mt19937_uniforms_kernel<<<batch, 32, 0, (cudaStream_t)stream>>>(seeds, n, out);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_mt19937_seed(const uint32_t* seeds, int32_t batch, uint32_t* mt_states, void* stream) {
  if (batch == 0) return RADET_OK;
  if (!seeds || !mt_states || batch < 0) return RADET_E_BADARG;
  mt19937_seed_kernel<<<batch, 32, 0, (cudaStream_t)stream>>>(seeds, mt_states);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

static int words_for(int maxG) { return maxG <= 32 ? 1 : maxG <= 64 ? 2 : maxG <= 128 ? 4 : 8; }

extern "C" size_t radet_assign_workspace_bytes(const radet_grid_t* grid, int32_t batch) {
  GridDev g;
  if (make_grid_dev(grid, &g) != RADET_OK || batch <= 0) return 0;
  const size_t P = (size_t)g.off[g.num_levels];
  // pair bits (worst case 8 words x 2) + global fallback list (u32) + owners (u16)
  return align_up((size_t)batch * P * 16 * 4, 256) + align_up((size_t)batch * P * 4, 256) + align_up((size_t)batch * P * 2, 256) +
         align_up((size_t)batch * RADET_MT_STATE_WORDS * 4, 256);
}

extern "C" int radet_assign(const radet_grid_t* grid, int32_t batch, const int32_t* gt_offsets, const int32_t* gt_offsets_host,
                            const float* gt_bboxes, const uint32_t* mask_bits, int32_t mask_h, int32_t mask_w,
                            int32_t mask_step, const double* uniforms, int32_t n_uniform, const uint32_t* seeds,
                            uint32_t* mt_states, int32_t positive_num, int32_t balance_sample, int64_t* points_to_gt_index,
                            float* points_weight, int32_t* consumed, double* weight_sums, void* workspace, size_t workspace_bytes,
                            void* stream) {
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc != RADET_OK) return rc;
  if (batch == 0) return RADET_OK;
  if (batch < 0 || !gt_offsets || !gt_offsets_host || !points_to_gt_index || !points_weight || !consumed || !workspace)
    return RADET_E_BADARG;
  if (positive_num <= 0 || positive_num > RADET_MAX_POSITIVE_NUM) return RADET_E_UNSUPPORTED;
  const int nrng = (uniforms ? 1 : 0) + (seeds ? 1 : 0) + (mt_states ? 1 : 0);
  if (nrng != 1 || (uniforms && n_uniform < 0)) return RADET_E_BADARG;
  int maxG = 0;
  for (int b = 0; b < batch; ++b) {
    const int G = gt_offsets_host[b + 1] - gt_offsets_host[b];
    if (G < 0) return RADET_E_BADARG;
    maxG = G > maxG ? G : maxG;
  }
  if (maxG > RADET_MAX_GT_PER_IMAGE) return RADET_E_TOO_MANY_GT;
  if (maxG > 0 && (!gt_bboxes || !mask_bits || mask_h <= 0 || mask_w <= 0 || mask_step <= 0)) return RADET_E_BADARG;
  for (int l = 0; l < g.num_levels && maxG > 0; ++l) {
    if (g.stride[l] % mask_step != 0) return RADET_E_BADARG;  // sample grid must contain every (y*s, x*s)
    if ((g.h[l] - 1) * g.stride[l] / mask_step >= mask_h || (g.w[l] - 1) * g.stride[l] / mask_step >= mask_w) return RADET_E_BADARG;
  }
  if (workspace_bytes < radet_assign_workspace_bytes(grid, batch) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return RADET_E_WORKSPACE;
  const int P = g.off[g.num_levels];
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  uint32_t* pair_bits = reinterpret_cast<uint32_t*>(ws);
  ws += align_up((size_t)batch * P * 16 * 4, 256);
  uint32_t* g_list = reinterpret_cast<uint32_t*>(ws);
  ws += align_up((size_t)batch * P * 4, 256);
  uint16_t* g_own = reinterpret_cast<uint16_t*>(ws);
  ws += align_up((size_t)batch * P * 2, 256);
  uint32_t* seeded_states = reinterpret_cast<uint32_t*>(ws);
  cudaStream_t st = (cudaStream_t)stream;
  const int W = words_for(maxG > 0 ? maxG : 1);
  const int mask_pitch = (mask_w + 31) / 32;
  const int Gp = (maxG + 3) & ~3;
  size_t pair_smem = (size_t)Gp * (16 + 16 + 4 + 4) + 16 + (size_t)maxG * mask_h * mask_pitch * 4 + 16;
  int stage_masks = 1;
  if (pair_smem > 96 * 1024) {  // keep >= 2 CTAs/SM; otherwise read mask bits through L2
    stage_masks = 0;
    pair_smem = (size_t)Gp * (16 + 16 + 4 + 4) + 32;
  }
  const int pair_blocks = (P + kPairThreads - 1) / kPairThreads;
  dim3 pgrid(pair_blocks + (seeds ? 1 : 0), batch);
  uint32_t* states_arg = seeds ? seeded_states : mt_states;
  const bool list_in_smem = P <= kListCap;
  const int mg = maxG > 0 ? maxG : 1;
  const int kcap = (balance_sample & RADET_ASSIGN_ADAPT_K) ? RADET_MAX_POSITIVE_NUM : positive_num;
  const size_t rs = resolve_smem_bytes(mg, kcap, list_in_smem);
#define RADET_ASSIGN_LAUNCH(WW)                                                                                          \
  do {                                                                                                                   \
    cudaFuncSetAttribute(assign_pairs_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pair_smem);          \
    assign_pairs_kernel<WW><<<pgrid, kPairThreads, pair_smem, st>>>(g, gt_offsets, gt_bboxes, mask_bits, mask_h,         \
                                                                     mask_pitch, mask_step, stage_masks, pair_bits,      \
                                                                     seeds, seeded_states, pair_blocks);                 \
    RADET_LAUNCH_CHECK();                                                                                                \
    cudaFuncSetAttribute(assign_resolve_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs);               \
    assign_resolve_kernel<WW><<<batch, kResolveThreads, rs, st>>>(g, gt_offsets, gt_bboxes, pair_bits, mg, kcap,          \
                                                                   positive_num, balance_sample, uniforms, n_uniform,     \
                                                                   nullptr,                                               \
                                                                   states_arg,                                            \
                                                                   g_list, g_own, list_in_smem ? 1 : 0,                  \
                                                                   points_to_gt_index, points_weight, consumed,          \
                                                                   weight_sums, seeds ? 0 : 1,                          \
                                                                   static_cast<long long*>(g_debug_buf));               \
    RADET_LAUNCH_CHECK();                                                                                                \
  } while (0)
  switch (W) {
    case 1: RADET_ASSIGN_LAUNCH(1); break;
    case 2: RADET_ASSIGN_LAUNCH(2); break;
    case 4: RADET_ASSIGN_LAUNCH(4); break;
    default: RADET_ASSIGN_LAUNCH(8); break;
  }
#undef RADET_ASSIGN_LAUNCH
  return RADET_OK;
}
