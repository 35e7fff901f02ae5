// Inference: score threshold -> per-level top-k -> TBLR decode -> class-aware (vote-)NMS, on sm_100a.
//
// Reference semantics: ATSSHead.get_bboxes (atss_head.py:326-387) + RADetHead._get_bboxes_single
// (radet_head.py:55-169), which loops over images and levels in Python, syncs the host per level, copies four
// tensors per image to the CPU and runs the single-threaded O(n^2) vote_ext.cpp (:70-353).  Class-aware NMS never
// couples boxes of different labels, so the batch is decomposed into (image, class) problems.  Four launches,
// nothing leaves the device:
//
//   detect_select_kernel   one CTA per (image, level, class chunk): HBM-bound scan of the class logits IN PLACE
//                          (NCHW planes, 128-bit streaming loads), sigmoid, strict `> score_thr`; survivors are
//                          staged in shared memory and appended with ONE global atomic per CTA.
//   detect_bin_kernel      one CTA per (image, level): exact top-k (8-bit radix select on the unique
//                          score|index keys) when a level has more than nms_pre candidates, centerness gather,
//                          cluster-score key, append to the (image, class) bin.
//   class_nms_kernel       one CTA per (image, class): bitonic sort by cluster score, decode+clamp+rescale,
//                          warp-ballot IoU bit matrix in shared memory, sequential seed scan on bit rows by one
//                          warp (member sets fall out as `row & alive`), sigma-filtered weighted box vote of every
//                          cluster in member (= score) order.
//   detect_rank_kernel     one CTA per image: top-max_per_img seeds over all classes (radix select + small sort),
//                          writes dets / labels / count.
//
// All arithmetic that feeds a discrete decision or an output box uses explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn): no FMA contraction, same operation order as vote_ext.cpp, so keep
// sets and voted boxes are bit-exact with the reference.
#include <math.h>

#include "common.cuh"

namespace radet {

struct MapsDev {
  const float* cls[RADET_MAX_LEVELS];
  const float* bbox[RADET_MAX_LEVELS];
  const float* iou[RADET_MAX_LEVELS];
};

// torch's CUDA sigmoid: 1 / (1 + exp(-x)) in fp32 with IEEE division (radet_head.py:106-109 run on CUDA tensors)
__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

constexpr int kOrdLevelShift = 27;  // ord = level << 27 | flat (point*C + class)
typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------ 1. select
constexpr int kSelThreads = 256;
constexpr int kSelChunk = 4;  // classes staged per flush

struct SelPlan {
  int bpl[RADET_MAX_LEVELS];       // CTAs per (image, level) plane = ceil(ceil(hw/4) / 256)
  int boff[RADET_MAX_LEVELS + 1];  // CTA offset of each level inside one (image, class-chunk)
  int64_t coff[RADET_MAX_LEVELS + 1];  // candidate-buffer offset of each level inside one image (= C * off[l])
  int cc, nj;
};

__global__ void __launch_bounds__(kSelThreads)
detect_select_kernel(GridDev grid, SelPlan plan, int C, MapsDev maps, float thr, float x_lo, u64* __restrict__ cand,
                     int* __restrict__ counts) {
  __shared__ u64 s_buf[kSelThreads * 4 * kSelChunk];   // staged (logit, flat) entries, overwritten in place by the keys
  __shared__ int s_cnt, s_keep, s_base;
  const int b = blockIdx.y, j = blockIdx.z, tid = threadIdx.x, lane = tid & 31;
  int l = 0;
#pragma unroll
  for (int k = 1; k < RADET_MAX_LEVELS; ++k) l += (k < grid.num_levels && (int)blockIdx.x >= plan.boff[k]) ? 1 : 0;
  const int hw = grid.h[l] * grid.w[l];
  const int q0 = 4 * (((int)blockIdx.x - plan.boff[l]) * kSelThreads + tid);
  const int nv = min(4, hw - q0);  // <= 0: idle thread
  const bool vec = (hw & 3) == 0;
  const float* cp = maps.cls[l] + ((int64_t)b * C) * hw + q0;
  u64* out = cand + (int64_t)b * plan.coff[grid.num_levels] + plan.coff[l];
  int* cnt = counts + b * RADET_MAX_LEVELS + l;
  const int c0 = j * plan.cc, c1 = min(C, c0 + plan.cc);
  for (int cb = c0; cb < c1; cb += kSelChunk) {
    if (tid == 0) {
      s_cnt = 0;
      s_keep = 0;
    }
    __syncthreads();
    const int ce = min(c1, cb + kSelChunk);
    float xv[kSelChunk][4];
#pragma unroll
    const float* pk = cp + (int64_t)cb * hw;
    for (int k = 0; k < kSelChunk; ++k, pk += hw) {   // all loads of the chunk in flight first
      xv[k][0] = xv[k][1] = xv[k][2] = xv[k][3] = -INFINITY;
      if (cb + k < ce && nv > 0) {
        if (vec) {
          const float4 v4 = ldg_stream4(pk);
          xv[k][0] = v4.x; xv[k][1] = v4.y; xv[k][2] = v4.z; xv[k][3] = v4.w;
        } else {
          for (int i = 0; i < nv; ++i) xv[k][i] = cp[(int64_t)(cb + k) * hw + i];
        }
      }
    }
    // A. conservative prefilter on the raw logit: the few survivors are staged so that the sigmoid below runs on
    //    dense warps instead of inside a branch that one lane in 32 takes.  Each lane first collects a bit mask of
    //    its survivors; one warp scan and one shared-memory atomic per warp then give every lane its slots.
    unsigned pm = 0u;
#pragma unroll
    for (int k = 0; k < kSelChunk; ++k) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nv && xv[k][i] > x_lo) pm |= 1u << (k * 4 + i);
    }
    {
      const int cnt_ = __popc(pm);
      int inc = cnt_;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
      }
      int wbase = 0;
      if (lane == 31 && inc) wbase = atomicAdd(&s_cnt, inc);
      wbase = __shfl_sync(kFull, wbase, 31);
      int pos = wbase + inc - cnt_;
      // lanes with survivors peel them off one per trip (typically one or two trips per warp); the value is picked
      // by a 4-level select tree so that xv[][] stays in registers
      static_assert(kSelChunk == 4, "the select tree below picks one of 16 values");
      while (__any_sync(kFull, pm != 0u)) {
        if (pm) {
          const int e = __ffs((int)pm) - 1;
          pm &= pm - 1u;
          const bool b0 = e & 1, b1 = e & 2, b2 = e & 4, b3 = e & 8;
          float t8[8], t4[4], t2[2];
#pragma unroll
          for (int q = 0; q < 8; ++q) t8[q] = b0 ? xv[(2 * q + 1) >> 2][(2 * q + 1) & 3] : xv[(2 * q) >> 2][(2 * q) & 3];
#pragma unroll
          for (int q = 0; q < 4; ++q) t4[q] = b1 ? t8[2 * q + 1] : t8[2 * q];
#pragma unroll
          for (int q = 0; q < 2; ++q) t2[q] = b2 ? t4[2 * q + 1] : t4[2 * q];
          const float v = b3 ? t2[1] : t2[0];
          const unsigned flat = (unsigned)(q0 + (e & 3)) * (unsigned)C + (unsigned)(cb + (e >> 2));
          s_buf[pos++] = ((u64)__float_as_uint(v) << 32) | (u64)flat;
        }
      }
    }
    __syncthreads();
    const int n = s_cnt;
    // B. exact test on the staged entries; kept keys are compacted in place (a tile's writes land below its reads)
    for (int base = 0; base < n; base += kSelThreads) {
      const int i = base + tid;
      bool keep = false;
      u64 key = 0ull;
      if (i < n) {
        const u64 e = s_buf[i];
        const float sc = sigmoid_rn(__uint_as_float((unsigned)(e >> 32)));
        keep = sc > thr;                              // radet_head.py:111 (strict)
        key = ((u64)__float_as_uint(sc) << 32) | (u64)(0xffffffffu - (unsigned)(e & 0xffffffffull));
      }
      const unsigned bal = __ballot_sync(kFull, keep);
      int wbase = 0;
      if (lane == 0 && bal) wbase = atomicAdd(&s_keep, __popc(bal));
      wbase = __shfl_sync(kFull, wbase, 0);
      __syncthreads();
      if (keep) s_buf[wbase + __popc(bal & ((1u << lane) - 1u))] = key;
    }
    __syncthreads();
    const int nk = s_keep;
    if (tid == 0 && nk) s_base = atomicAdd(cnt, nk);
    __syncthreads();
    for (int i = tid; i < nk; i += kSelThreads) out[s_base + i] = s_buf[i];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ 2. top-k + binning
constexpr int kBinThreads = 512;
constexpr int kBinNK = 12;   // keys a thread keeps in registers: levels with <= 6144 candidates read them from memory once

// A CTA's view of `n` 64-bit keys: thread t keeps keys t, t + T, ..., t + (NK-1) T in registers (one batch of
// independent loads); keys beyond NK * T are re-read from global memory, eight loads in flight per thread.  Every sweep
// below used to be a chain of dependent load -> shared-memory-atomic iterations (~10 round trips to L2 per sweep on the
// finest level); from registers a sweep costs a few hundred cycles.
template <int NK>
struct KeyCache {
  u64 r[NK];
  const u64* src;
  int n;
  __device__ __forceinline__ void load(const u64* s, int n_) {
    src = s;
    n = n_;
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      const int i = (int)threadIdx.x + j * (int)blockDim.x;
      r[j] = i < n ? src[i] : 0ull;
    }
  }
  // f(key, valid) for the key slots of this thread; all threads of a WARP make the same number of calls, so f may use
  // full-warp collectives (not block barriers).
  template <typename F>
  __device__ __forceinline__ void each(F&& f) const {
    const int T = (int)blockDim.x;
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      if (((int)threadIdx.x & ~31) + j * T >= n) break;   // warp-uniform: none of this warp's keys in slot j or beyond
      f(r[j], (int)threadIdx.x + j * T < n);
    }
    for (int base = NK * T; base < n; base += 8 * T) {
      u64 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * T + (int)threadIdx.x;
        t[u] = i < n ? src[i] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) f(t[u], base + u * T + (int)threadIdx.x < n);
    }
  }
};

// Common leading bits of the view's keys: (AND of all keys, OR of all keys), block-wide.  s_red: 128 unsigned.  One barrier.
template <int NK>
__device__ __forceinline__ void key_and_or(const KeyCache<NK>& kc, unsigned* s_red, u64* and_out, u64* or_out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (int)blockDim.x >> 5;
  u64 a = ~0ull, o = 0ull;
  kc.each([&](u64 key, bool v) {
    if (v) {
      a &= key;
      o |= key;
    }
  });
  const unsigned alo = __reduce_and_sync(kFull, (unsigned)a), ahi = __reduce_and_sync(kFull, (unsigned)(a >> 32));
  const unsigned olo = __reduce_or_sync(kFull, (unsigned)o), ohi = __reduce_or_sync(kFull, (unsigned)(o >> 32));
  if (lane == 0) {
    s_red[wid] = alo;
    s_red[32 + wid] = ahi;
    s_red[64 + wid] = olo;
    s_red[96 + wid] = ohi;
  }
  __syncthreads();
  const unsigned xa = lane < nw ? s_red[lane] : 0xffffffffu, xb = lane < nw ? s_red[32 + lane] : 0xffffffffu;
  const unsigned xc = lane < nw ? s_red[64 + lane] : 0u, xd = lane < nw ? s_red[96 + lane] : 0u;
  *and_out = ((u64)__reduce_and_sync(kFull, xb) << 32) | (u64)__reduce_and_sync(kFull, xa);
  *or_out = ((u64)__reduce_or_sync(kFull, xd) << 32) | (u64)__reduce_or_sync(kFull, xc);
}

// k-th largest of the view's unique 64-bit keys (1 <= k <= n): returns the smallest key to keep.  8-bit radix passes over
// a shared-memory histogram that start right below the bits all keys have in common (scores above a threshold share their
// top byte) and stop as soon as the selected bucket is needed entirely.  One block barrier per pass: the histograms rotate
// through three buffers and every warp does the 256-bin suffix scan itself.  The increments compile to ATOMS.POPC.INC
// (lanes of a warp that hit the same bin are added in one operation).  s_hist: 3 * 256 ints, s_red: 128 unsigned.
// (Measured and dropped: 4-bit passes counted in packed register fields + REDUX instead of shared-memory atomics --
// 20 k cycles instead of 7.4 k on the finest level of the bench workload; MATCH.ANY-aggregated increments -- slower too.)
template <int NK>
__device__ u64 radix_select_kth(const KeyCache<NK>& kc, int k, int* s_hist, unsigned* s_red) {
  const int tid = threadIdx.x, lane = tid & 31;
  u64 a, o;
  for (int i = tid; i < 256; i += (int)blockDim.x) s_hist[i] = 0;
  key_and_or(kc, s_red, &a, &o);
  const u64 diff = a ^ o;
  if (diff == 0ull) return a;
  const int top = 63 - __clzll((long long)diff);
  u64 pmask = top == 63 ? 0ull : ~((2ull << top) - 1ull);
  u64 prefix = a & pmask;
  int shift = max(top - 7, 0);
  for (int pass = 0;; ++pass) {
    int* h = s_hist + (pass % 3) * 256;
    int* hz = s_hist + ((pass + 1) % 3) * 256;
    for (int i = tid; i < 256; i += (int)blockDim.x) hz[i] = 0;
    kc.each([&](u64 key, bool v) {
      if (v && (key & pmask) == prefix) atomicAdd(&h[(int)((key >> shift) & 0xffull)], 1);
    });
    __syncthreads();
    int hq[8], s = 0;                                     // lane holds bins [8 lane, 8 lane + 8)
    {
      const int4 h0 = *reinterpret_cast<const int4*>(h + 8 * lane), h1 = *reinterpret_cast<const int4*>(h + 8 * lane + 4);
      hq[0] = h0.x; hq[1] = h0.y; hq[2] = h0.z; hq[3] = h0.w;
      hq[4] = h1.x; hq[5] = h1.y; hq[6] = h1.z; hq[7] = h1.w;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) s += hq[q];
    int suf = s;                                          // inclusive suffix over lanes >= lane
#pragma unroll
    for (int of = 1; of < 32; of <<= 1) {
      const int t = __shfl_down_sync(kFull, suf, of);
      if (lane + of < 32) suf += t;
    }
    const int above = suf - s;                            // keys in the bins of higher lanes
    const bool hit = above < k && k <= suf;               // exactly one lane
    int bin = 7, acc = above;
#pragma unroll
    for (int q = 7; q > 0; --q) {
      if (bin == q && acc + hq[q] < k) {
        acc += hq[q];
        bin = q - 1;
      }
    }
    int pop = hq[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) pop = bin == q ? hq[q] : pop;
    const int src_lane = __ffs((int)__ballot_sync(kFull, hit)) - 1;
    const int digit = __shfl_sync(kFull, 8 * lane + bin, src_lane);
    const int krem = __shfl_sync(kFull, k - acc, src_lane);   // rank inside the bucket
    pop = __shfl_sync(kFull, pop, src_lane);                  // bucket population
    prefix |= (u64)digit << shift;
    pmask |= 0xffull << shift;
    k = krem;
    if (pop == k || shift == 0) break;
    shift = max(shift - 8, 0);
  }
  return prefix;
}

struct SelectSmem {
  int hist[3 * 256];
  unsigned red[128];
};
template <int NK>
__device__ __forceinline__ u64 select_kth(const KeyCache<NK>& kc, int k, SelectSmem* S) {
  return radix_select_kth(kc, k, S->hist, S->red);
}

struct BinParams {
  GridDev grid;
  MapsDev maps;
  int64_t coff[RADET_MAX_LEVELS + 1];
  int C, nms_pre, cs_mode, class_cap;
  const u64* cand;
  int* counts;        // [B][8] candidates per (image, level); re-armed here
  u64* bins;          // [B][C][class_cap]
  int* class_counts;  // [B][C]; re-armed by detect_rank_kernel
  long long* dbg;
};

__global__ void __launch_bounds__(kBinThreads, 2)
detect_bin_kernel(BinParams p) {
  __shared__ __align__(16) SelectSmem s_sel;
  const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const GridDev& g = p.grid;
#define BIN_DBG(k) do { if (p.dbg && tid == 0) p.dbg[(int64_t)(b * g.num_levels + l) * 16 + 10 + (k)] = clock64(); } while (0)
  BIN_DBG(0);
  const u64* src = p.cand + (int64_t)b * p.coff[g.num_levels] + p.coff[l];
  const int nl = p.counts[b * RADET_MAX_LEVELS + l];
  __syncthreads();
  if (tid == 0) p.counts[b * RADET_MAX_LEVELS + l] = 0;  // re-arm for the next call
  if (nl == 0) return;
  KeyCache<kBinNK> kc;
  kc.load(src, nl);
  u64 kth = 0ull;
  BIN_DBG(1);
  if (p.nms_pre > 0 && nl > p.nms_pre) kth = select_kth(kc, p.nms_pre, &s_sel);  // radet_head.py:112-122
  BIN_DBG(2);
  const int hw = g.h[l] * g.w[l];
  const float* ioumap = p.maps.iou[l] + (int64_t)b * hw;
  const unsigned C = (unsigned)p.C;
  const bool need_ctr = p.cs_mode != 1;
  // flat = point * C + class: division by the (runtime) class count as a multiplication, exact for flat * C < 2^40
  // (flat < 2^27, C <= 2^13; plain division beyond): q = flat * ceil(2^40 / C) >> 40
  const u64 cmagic = ((1ull << 40) + C - 1ull) / C;
  auto split = [&](unsigned flat, unsigned& q, int& cl) {
    q = C <= 8192u ? (unsigned)(((u64)flat * cmagic) >> 40) : flat / C;
    cl = (int)(flat - q * C);
  };
  // Class binning.  The kept candidates (a fifth of the keys on the finest level) are first compacted into shared memory
  // -- warp ballots, no atomics -- so that the two sweeps below run over ~2 keys per thread with full lanes instead of 12
  // mostly idle ones.  The CTA counts them per class in shared memory (ATOMS.POPC.INC: lanes of a warp that hit the same
  // class are added in one operation), reserves one range per class with ONE global atomic each, and hands out the slots
  // from shared memory.  (The order inside a bin is irrelevant: class_nms_kernel sorts by key.)
  constexpr int kBinClasses = 1024, kKeepCap = 2048;
  __shared__ int s_ccnt[kBinClasses], s_cbase[kBinClasses];
  __shared__ u64 s_keep[kKeepCap];
  __shared__ int s_wcnt[32], s_cursor;
  const bool local = p.C <= kBinClasses;
  const int lane = tid & 31, wid = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int nkeep = kth != 0ull ? p.nms_pre : nl;            // unique keys: exactly nms_pre survive the selection
  auto emit = [&](u64 key, float cx, int cl, int sl) {
    const float S = __uint_as_float((unsigned)(key >> 32));
    const unsigned flat = 0xffffffffu - (unsigned)(key & 0xffffffffull);
    float cs = S;
    if (need_ctr) {
      const float ctr = sigmoid_rn(cx);                                       // radet_head.py:109
      cs = p.cs_mode == 0 ? __fmul_rn(S, ctr) : ctr;                          // vote_wrapper.py:14-21
    }
    const unsigned ord = ((unsigned)l << kOrdLevelShift) | flat;
    if (sl < p.class_cap)
      p.bins[((int64_t)b * p.C + cl) * p.class_cap + sl] = ((u64)float_order_key(cs) << 32) | (u64)(0xffffffffu - ord);
  };
  if (local && nkeep <= kKeepCap) {
    for (int c = tid; c < p.C; c += kBinThreads) s_ccnt[c] = 0;
    int wc = 0;
#pragma unroll
    for (int j = 0; j < kBinNK; ++j) {
      if (wid * 32 + j * kBinThreads >= nl) break;          // warp-uniform: no key of this warp in slot j or beyond
      wc += __popc(__ballot_sync(kFull, tid + j * kBinThreads < nl && kc.r[j] >= kth));
    }
    if (lane == 0) s_wcnt[wid] = wc;
    __syncthreads();
    int wbase = 0, total = 0;
    {
      const int v = lane < kBinThreads / 32 ? s_wcnt[lane] : 0;
      wbase = __reduce_add_sync(kFull, lane < wid ? v : 0);
      total = __reduce_add_sync(kFull, v);
    }
    if (tid == 0) s_cursor = total;
#pragma unroll
    for (int j = 0; j < kBinNK; ++j) {
      if (wid * 32 + j * kBinThreads >= nl) break;
      const bool keep = tid + j * kBinThreads < nl && kc.r[j] >= kth;
      const unsigned bal = __ballot_sync(kFull, keep);
      if (keep) s_keep[wbase + __popc(bal & lt)] = kc.r[j];
      wbase += __popc(bal);
    }
    __syncthreads();
    for (int base = kBinNK * kBinThreads; base < nl; base += 8 * kBinThreads) {   // keys beyond the register-resident ones
      u64 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * kBinThreads + tid;
        t[u] = i < nl ? src[i] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (base + u * kBinThreads + tid < nl && t[u] >= kth) s_keep[atomicAdd(&s_cursor, 1)] = t[u];
    }
    if (nl > kBinNK * kBinThreads) __syncthreads();
    // per-class counts, one reservation per class, slots
    for (int i = tid; i < nkeep; i += kBinThreads) {
      unsigned q;
      int cl;
      split(0xffffffffu - (unsigned)(s_keep[i] & 0xffffffffull), q, cl);
      atomicAdd(&s_ccnt[cl], 1);
    }
    __syncthreads();
    for (int c = tid; c < p.C; c += kBinThreads) {
      const int n = s_ccnt[c];
      s_cbase[c] = n ? atomicAdd(&p.class_counts[b * p.C + c], n) : 0;
      s_ccnt[c] = 0;
    }
    BIN_DBG(3);
    // the centerness gathers of a thread's (<= 4) keys are in flight while the reservations return
    static_assert(kKeepCap == 4 * kBinThreads, "one trip of four keys per thread covers the staging area");
    u64 key[4];
    float cx[4];
    int cl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = tid + u * kBinThreads;
      key[u] = i < nkeep ? s_keep[i] : 0ull;
      unsigned q;
      split(0xffffffffu - (unsigned)(key[u] & 0xffffffffull), q, cl[u]);
      cx[u] = (i < nkeep && need_ctr) ? ioumap[q] : 0.f;
    }
    __syncthreads();                                         // s_cbase
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (tid + u * kBinThreads < nkeep) emit(key[u], cx[u], cl[u], s_cbase[cl[u]] + atomicAdd(&s_ccnt[cl[u]], 1));
  } else {
    // more kept candidates than the staging area holds (nms_pre > 2048 or no per-level limit), or more than 1024 classes:
    // every key takes its slot with an atomic of its own
    if (local) {
      for (int c = tid; c < p.C; c += kBinThreads) s_ccnt[c] = 0;
      __syncthreads();
      kc.each([&](u64 key, bool v) {
        if (v && key >= kth) {
          unsigned q;
          int cl;
          split(0xffffffffu - (unsigned)(key & 0xffffffffull), q, cl);
          atomicAdd(&s_ccnt[cl], 1);
        }
      });
      __syncthreads();
      for (int c = tid; c < p.C; c += kBinThreads) {
        const int n = s_ccnt[c];
        s_cbase[c] = n ? atomicAdd(&p.class_counts[b * p.C + c], n) : 0;
        s_ccnt[c] = 0;
      }
      __syncthreads();
    }
    BIN_DBG(3);
    kc.each([&](u64 key, bool v) {
      if (v && key >= kth) {
        unsigned q;
        int cl;
        split(0xffffffffu - (unsigned)(key & 0xffffffffull), q, cl);
        const float cx = need_ctr ? ioumap[q] : 0.f;
        const int sl = local ? s_cbase[cl] + atomicAdd(&s_ccnt[cl], 1) : atomicAdd(&p.class_counts[b * p.C + cl], 1);
        emit(key, cx, cl, sl);
      }
    });
  }
  BIN_DBG(4);
}

// ------------------------------------------------------------------------------------------------ 3. per-class NMS + vote
constexpr int kClsThreads = 256;
constexpr int kMaskItems = 512;                  // bit-matrix path: up to 512 boxes of one class in one image
constexpr int kMaskWords = kMaskItems / 32;

struct ClsParams {
  GridDev grid;
  MapsDev maps;
  int C, class_cap, rescale, cs_mode, vs_mode, mode, iou_enable;
  float thr, sigma;
  const int* img_shapes;
  const float* scale_factors;
  const int* class_counts;
  int* class_counts_rw; // same array (detect_emit_kernel re-arms it)
  u64* bins;           // sorted in place
  int img_cap;         // seeds per image <= candidates per image
  float* seed_out;     // [B][img_cap][5]  voted box + score of every cluster of the image (any order)
  u64* seed_keys;      // [B][img_cap]     sort key of the seed
  int* img_seed_count; // [B]              append cursor; re-armed by detect_rank_kernel
  long long* dbg;      // optional phase timestamps (radet_debug_set_buffer)
  // large-class fallback (m > kMaskItems): records in global memory
  float4* g_box;       // [B][C][class_cap]
  float* g_vs;         // [B][C][class_cap]
  int* g_owner;        // [B][C][class_cap]
};

__device__ __forceinline__ void bitonic_sort_desc(u64* keys, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const bool desc = (i & k) == 0;
        const u64 a = keys[i], b = keys[ixj];
        if ((a < b) == desc) {
          keys[i] = b;
          keys[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float iou_rn(const float4& bi, float area_i, const float4& bj) {
  // vote_ext.cpp:152-162 (no +1, no eps; 0/0 = NaN compares false)
  const float xl = fmaxf(bj.x, bi.x), yt = fmaxf(bj.y, bi.y), xr = fminf(bj.z, bi.z), yb = fminf(bj.w, bi.w);
  const float iw = fmaxf(0.f, __fsub_rn(xr, xl)), ih = fmaxf(0.f, __fsub_rn(yb, yt));
  const float inter = __fmul_rn(iw, ih);
  const float area_j = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_j, area_i), inter));
}
__device__ __forceinline__ float box_area_rn(const float4& b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }
// `iou > thr` with the division skipped for disjoint boxes: inter == 0 gives iou = +0 (or NaN for two empty boxes),
// which is never > thr when thr >= 0 -- same decision, no rounding involved.
__device__ __forceinline__ bool iou_gt(const float4& bi, float area_i, const float4& bj, float thr) {
  const float xl = fmaxf(bj.x, bi.x), yt = fmaxf(bj.y, bi.y), xr = fminf(bj.z, bi.z), yb = fminf(bj.w, bi.w);
  const float iw = fmaxf(0.f, __fsub_rn(xr, xl)), ih = fmaxf(0.f, __fsub_rn(yb, yt));
  const float inter = __fmul_rn(iw, ih);
  if (inter == 0.f && thr >= 0.f) return false;
  const float area_j = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_j, area_i), inter)) > thr;
}

// The same decision, branch-free, with the division replaced by rcp.approx + one multiply: |q - iou| <= ~2 ulp, so
// outside the band [lo, hi] = thr * (1 -+ 1e-6) (8 ulp on each side) the quotient is clearly on one side of thr.  Inside
// the band, for NaN and for a (near-)denormal union the result is flagged ambiguous and the caller falls back to the
// exact division (iou_gt).  Requires thr > 0; area_j is the precomputed box_area_rn of bj.
__device__ __forceinline__ bool iou_gt_fast(const float4& bi, float area_i, const float4& bj, float area_j, float lo, float hi,
                                            bool& ambiguous) {
  const float xl = fmaxf(bj.x, bi.x), yt = fmaxf(bj.y, bi.y), xr = fminf(bj.z, bi.z), yb = fminf(bj.w, bi.w);
  const float iw = fmaxf(0.f, __fsub_rn(xr, xl)), ih = fmaxf(0.f, __fsub_rn(yb, yt));
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_j, area_i), inter);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(uni));
  const float q = __fmul_rn(inter, r);
  const bool pos = inter != 0.f;                                     // inter == 0: iou is 0 or NaN, never > thr
  const bool ok = uni > 1e-30f;
  const bool yes = q > hi && ok, no = q < lo && ok;
  ambiguous = pos && !yes && !no;
  return pos && yes;
}

// decode one key into box / cluster score / vote score (radet_head.py:123-143, tblr_bbox_coder.py:154-171); split
// into the global loads and the arithmetic so a caller can put independent work between the two
struct DecodeRaw {
  float t, b, l, r, cls, iou, H, W, cx, cy, side;
  float4 sf;
};
__device__ __forceinline__ DecodeRaw decode_load(const ClsParams& p, int b, int cls, u64 key) {
  const GridDev& g = p.grid;
  const unsigned ord = 0xffffffffu - (unsigned)(key & 0xffffffffull);
  const int l = (int)(ord >> kOrdLevelShift);
  const unsigned flat = ord & ((1u << kOrdLevelShift) - 1u);
  const int q = (int)(flat / (unsigned)p.C);
  const int hw = g.h[l] * g.w[l];
  const int y = q / g.w[l], x = q - y * g.w[l];
  const float st = (float)g.stride[l];
  DecodeRaw d;
  d.cx = (float)x * st;
  d.cy = (float)y * st;
  d.side = __fmul_rn(g.anchor_scale, st);
  const float* bp = p.maps.bbox[l] + (int64_t)b * 4 * hw + q;
  d.t = bp[0];
  d.b = bp[hw];
  d.l = bp[2 * hw];
  d.r = bp[3 * hw];
  d.cls = p.maps.cls[l][((int64_t)b * p.C + cls) * hw + q];
  d.iou = p.maps.iou[l][(int64_t)b * hw + q];
  d.H = (float)p.img_shapes[2 * b];
  d.W = (float)p.img_shapes[2 * b + 1];
  d.sf = p.rescale ? *reinterpret_cast<const float4*>(p.scale_factors + 4 * b) : make_float4(1.f, 1.f, 1.f, 1.f);
  return d;
}
__device__ __forceinline__ void decode_finish(const ClsParams& p, const DecodeRaw& d, float4& bx, float& cs, float& vs) {
  const float nrm = p.grid.nrm;
  const float T = __fmul_rn(__fmul_rn(d.t, nrm), d.side), Bt = __fmul_rn(__fmul_rn(d.b, nrm), d.side);
  const float L = __fmul_rn(__fmul_rn(d.l, nrm), d.side), R = __fmul_rn(__fmul_rn(d.r, nrm), d.side);
  bx.x = fminf(fmaxf(__fsub_rn(d.cx, L), 0.f), d.W);
  bx.y = fminf(fmaxf(__fsub_rn(d.cy, T), 0.f), d.H);
  bx.z = fminf(fmaxf(__fadd_rn(d.cx, R), 0.f), d.W);
  bx.w = fminf(fmaxf(__fadd_rn(d.cy, Bt), 0.f), d.H);
  if (p.rescale) {                                                    // radet_head.py:141-143
    bx.x = __fdiv_rn(bx.x, d.sf.x); bx.y = __fdiv_rn(bx.y, d.sf.y);
    bx.z = __fdiv_rn(bx.z, d.sf.z); bx.w = __fdiv_rn(bx.w, d.sf.w);
  }
  const float S = sigmoid_rn(d.cls);
  const float ctr = sigmoid_rn(d.iou);
  cs = p.cs_mode == 0 ? __fmul_rn(S, ctr) : (p.cs_mode == 1 ? S : ctr);
  vs = p.vs_mode == 0 ? __fmul_rn(S, ctr) : (p.vs_mode == 1 ? S : ctr);
}
__device__ __forceinline__ void decode_item(const ClsParams& p, int b, int cls, u64 key, float4& bx, float& cs, float& vs) {
  decode_finish(p, decode_load(p, b, cls, key), bx, cs, vs);
}

// vote_single_dim (vote_ext.cpp:8-35) over a member list given as "seed + set bits of `mem` words" (ascending index =
// descending cluster score = the reference's member order).  fp32, one rounding per operation.
struct MemberIter {
  const unsigned* words;
  int nw, seed;
  int w;
  unsigned cur;
  bool first;
  __device__ MemberIter(const unsigned* words_, int nw_, int seed_) : words(words_), nw(nw_), seed(seed_), w(-1), cur(0u), first(true) {}
  __device__ int next() {  // -1 when exhausted
    if (first) {
      first = false;
      return seed;
    }
    while (cur == 0u) {
      if (++w >= nw) return -1;
      cur = words[w];
    }
    const int bit = __ffs((int)cur) - 1;
    cur &= cur - 1u;
    return w * 32 + bit;
  }
};

template <typename GetS, typename GetX>
__device__ float vote_axis_members(const unsigned* words, int nw, int seed, GetS gs, GetX gx) {
  float ss = 0.f, acc = 0.f;
  {
    MemberIter it(words, nw, seed);
    for (int j = it.next(); j >= 0; j = it.next()) {
      const float s = gs(j), x = gx(j);
      ss = __fadd_rn(ss, s);
      acc = __fadd_rn(acc, __fmul_rn(s, x));
    }
  }
  const float mean = __fdiv_rn(acc, ss);
  float var = 0.f;
  {
    MemberIter it(words, nw, seed);
    for (int j = it.next(); j >= 0; j = it.next()) {
      const float d = __fsub_rn(gx(j), mean);
      var = __fadd_rn(var, __fmul_rn(__fmul_rn(gs(j), d), d));
    }
  }
  const float sd = __fsqrt_rn(__fdiv_rn(var, ss));
  const float lo = __fsub_rn(mean, sd), hi = __fadd_rn(mean, sd);
  float fs = 0.f, fx = 0.f;
  {
    MemberIter it(words, nw, seed);
    for (int j = it.next(); j >= 0; j = it.next()) {
      const float x = gx(j);
      if (lo <= x && x <= hi) {
        const float s = gs(j);
        fx = __fadd_rn(fx, __fmul_rn(s, x));
        fs = __fadd_rn(fs, s);
      }
    }
  }
  return __fdiv_rn(fx, fs);
}

__global__ void __launch_bounds__(kClsThreads, 4)
class_nms_kernel(ClsParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_nseed, s_base;
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int bc = b * p.C + c;
  const int m = min(p.class_counts[bc], p.class_cap);
  if (m == 0) return;
  u64* gkeys = p.bins + (int64_t)bc * p.class_cap;
#define RADET_DBG(k) do { if (p.dbg && tid == 0) p.dbg[(int64_t)bc * 16 + (k)] = clock64(); } while (0)
  RADET_DBG(0);

  if (m <= kMaskItems) {
    // ---------------- shared-memory path
    u64* keys = reinterpret_cast<u64*>(smem_raw);                                   // [512]
    float4* box = reinterpret_cast<float4*>(keys + kMaskItems);                     // [512]
    float* vs = reinterpret_cast<float*>(box + kMaskItems);                         // [512]
    float* cs = vs + kMaskItems;                                                    // [512]
    unsigned* mask = reinterpret_cast<unsigned*>(cs + kMaskItems);                  // [512][W]
    short* seeds = reinterpret_cast<short*>(mask + kMaskItems * kMaskWords);        // [512]
    float* area = reinterpret_cast<float*>(seeds + kMaskItems) + 2 * kMaskWords;    // [512] (after nzrow, seedbits)
    const int W = (m + 31) >> 5;
    // kItems items per thread (m <= 512).  Sort by counting: rank_i = #{j : key_j > key_i} (keys are unique), m
    // broadcast reads of shared memory per item, no barriers inside.  The items' global loads (4 regression planes,
    // class and IoU logits) are issued first and land while the rank loops run.
    constexpr int kItems = kMaskItems / kClsThreads;
    u64* ukeys = reinterpret_cast<u64*>(mask);                     // unsorted keys; the mask area is free until later
    u64 mykey[kItems];
    DecodeRaw raw[kItems];
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      const int i = tid + u * kClsThreads;
      mykey[u] = 0ull;
      if (i < m) {
        mykey[u] = gkeys[i];
        ukeys[i] = mykey[u];
        raw[u] = decode_load(p, b, c, mykey[u]);
      }
    }
    if (tid == 0 && (m & 1)) ukeys[m] = 0ull;                      // pad to an even count for the 128-bit reads
    __syncthreads();
    RADET_DBG(1);
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      const int i = tid + u * kClsThreads;
      if (i < m) {
        int rank = 0;
        const ulonglong2* uk2 = reinterpret_cast<const ulonglong2*>(ukeys);
        const int n2 = (m + 1) >> 1;
#pragma unroll 4
        for (int j = 0; j < n2; ++j) {
          const ulonglong2 k2 = uk2[j];
          rank += (k2.x > mykey[u] ? 1 : 0) + (k2.y > mykey[u] ? 1 : 0);
        }
        float4 bx;
        float c_, v_;
        decode_finish(p, raw[u], bx, c_, v_);
        keys[rank] = mykey[u];
        box[rank] = bx;
        area[rank] = box_area_rn(bx);
        cs[rank] = c_;
        vs[rank] = v_;
      }
    }
    RADET_DBG(2);
    for (int i = m + tid; i < ((m + 31) & ~31) + 1 && i < kMaskItems; i += kClsThreads) {   // sentinel columns (+ row m)
      box[i] = make_float4(-1.f, -1.f, -1.f, -1.f);                 // clamped boxes are >= 0: never overlaps
      area[i] = 0.f;
    }
    __syncthreads();
    RADET_DBG(3);
    // IoU bit matrix, upper triangle: one warp per row i (round-robin), lane = column inside each 32-wide word.
    // nzrow collects which rows have any overlap at all: only those take part in the sequential scan below.
    unsigned* nzrow = reinterpret_cast<unsigned*>(seeds + kMaskItems);              // [W]
    unsigned* seedbits = nzrow + kMaskWords;                                        // [W]
    if (tid < kMaskWords) nzrow[tid] = 0u;
    __syncthreads();
    // Two adjacent rows per warp trip (iA, iA + 1 share the column loads and the diagonal word); lanes = columns.
    // Columns m .. 32 W - 1 hold a sentinel box that overlaps nothing (written with the sort above), so the loop has no
    // bounds tests; "column above the diagonal" is a mask on the first word; threshold decisions that the rcp-filtered
    // test cannot make are only flagged, and the (very rare) row pair with a flag is redone with the exact division.
    constexpr int kNW = kClsThreads / 32;
    const bool fast_ok = p.thr >= 1e-30f;                           // iou_gt_fast needs thr > 0
    const float thr_lo = p.thr * (1.f - 1e-6f), thr_hi = p.thr * (1.f + 1e-6f);
    for (int iA = 2 * wid; iA < m; iA += 2 * kNW) {
      const int iB = iA + 1;                                        // <= 511: the row exists in shared memory even when iB == m
      const float4 bA = box[iA], bB = box[iB];                      // iB == m reads the sentinel
      const float areaA = area[iA], areaB = area[iB];
      const int w0 = iA >> 5;                                       // == iB >> 5 (iA is even)
      unsigned* mA = mask + iA * W;
      unsigned* mB = mask + iB * W;
      if (lane < w0) {                                              // words below the diagonal
        mA[lane] = 0u;
        mB[lane] = 0u;
      }
      const unsigned aboveA = ~((2u << (iA & 31)) - 1u), aboveB = ~((2u << (iB & 31)) - 1u);   // columns > i in word w0
      unsigned anyA = 0u, anyB = 0u;
      bool redo = !fast_ok;
      if (fast_ok) {
        bool amb = false;
        const float4* bp = box + w0 * 32 + lane;
        const float* ap = area + w0 * 32 + lane;
#pragma unroll 2
        for (int wj = w0; wj < W; ++wj, bp += 32, ap += 32) {
          const float4 bj = *bp;
          const float aj = *ap;
          bool ambA, ambB;
          const bool hitA = iou_gt_fast(bA, areaA, bj, aj, thr_lo, thr_hi, ambA);            // vote_ext.cpp:169
          const bool hitB = iou_gt_fast(bB, areaB, bj, aj, thr_lo, thr_hi, ambB);
          amb = amb || ambA || ambB;
          unsigned wordA = __ballot_sync(kFull, hitA), wordB = __ballot_sync(kFull, hitB);
          if (wj == w0) {
            wordA &= aboveA;
            wordB &= aboveB;
          }
          if (lane == 0) {
            mA[wj] = wordA;
            mB[wj] = wordB;
          }
          anyA |= wordA;
          anyB |= wordB;
        }
        redo = __any_sync(kFull, amb);
      }
      if (redo) {                                                   // exact division for the whole row pair
        anyA = anyB = 0u;
        for (int wj = w0; wj < W; ++wj) {
          const int j = wj * 32 + lane;
          const float4 bj = box[j];
          const bool hitA = j > iA && j < m && iou_gt(bA, areaA, bj, p.thr);
          const bool hitB = j > iB && j < m && iB < m && iou_gt(bB, areaB, bj, p.thr);
          const unsigned wordA = __ballot_sync(kFull, hitA), wordB = __ballot_sync(kFull, hitB);
          if (lane == 0) {
            mA[wj] = wordA;
            mB[wj] = wordB;
          }
          anyA |= wordA;
          anyB |= wordB;
        }
      }
      if (lane == 0 && anyA) atomicOr(&nzrow[iA >> 5], 1u << (iA & 31));
      if (lane == 0 && anyB) atomicOr(&nzrow[iB >> 5], 1u << (iB & 31));
    }
    __syncthreads();
    RADET_DBG(4);
    // Sequential seed scan (warp 0, lanes = words of the alive set).  A box still alive when reached is a seed and
    // removes `row & alive`; boxes whose row is empty cannot remove anything, so only non-empty rows are visited.
    if (wid == 0) {
      unsigned alive = 0u;
      if (lane < W) {
        const int rem = m - lane * 32;
        alive = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
      }
      if (p.mode == RADET_NMS_GLOBAL_VOTE) {
        // vote_ext.cpp:257-263: once a label has been emitted every later seed of it is dropped -> only box 0 clusters
        alive = lane == 0 ? 1u : 0u;    // row 0 already is its member set (everything was alive)
      } else {
        const unsigned nz = lane < W ? nzrow[lane] : 0u;
        int i = -1;
        while (true) {
          const int t = i + 1, wi = t >> 5;
          unsigned wmask = lane > wi ? 0xffffffffu : 0u;
          if (lane == wi) wmask = ~((1u << (t & 31)) - 1u);
          const unsigned cand = alive & nz & wmask;
          const int mine = cand ? lane * 32 + __ffs((int)cand) - 1 : 0x7fffffff;
          const int nxt = __reduce_min_sync(kFull, mine);
          if (nxt == 0x7fffffff) break;
          i = nxt;
          const unsigned row = (lane < W && lane >= (i >> 5)) ? mask[i * W + lane] : 0u;
          if (lane < W) mask[i * W + lane] = row & alive;          // row i now holds the member set of cluster i
          alive &= ~row;
        }
      }
      if (lane < kMaskWords) seedbits[lane] = alive;               // seeds = everything still alive
      const int ns_ = __reduce_add_sync(kFull, __popc(alive));
      if (lane == 0) {
        s_nseed = ns_;
        s_base = atomicAdd(&p.img_seed_count[b], ns_);            // this class's slice of the image's seed list
      }
    }
    __syncthreads();
    RADET_DBG(5);
    const int ns = s_nseed;
    float* sout = p.seed_out + ((int64_t)b * p.img_cap + s_base) * 5;
    u64* skeys_out = p.seed_keys + (int64_t)b * p.img_cap + s_base;
    for (int i = tid; i < m; i += kClsThreads) {                    // seed list in index (= score) order
      const unsigned wbits = seedbits[i >> 5];
      if ((wbits >> (i & 31)) & 1u) {
        int r = __popc(wbits & ((1u << (i & 31)) - 1u));
        for (int w_ = 0; w_ < (i >> 5); ++w_) r += __popc(seedbits[w_]);
        seeds[r] = (short)i;
      }
    }
    __syncthreads();
    // vote (all clusters; 4 threads per cluster, one per coordinate)
    for (int t = tid; t < ns * 4; t += kClsThreads) {
      const int s = t >> 2, axis = t & 3;
      const int i = seeds[s];
      float v;
      if (p.mode == RADET_NMS_PLAIN) {
        v = reinterpret_cast<const float*>(&box[i])[axis];
      } else {
        const float4 bi = box[i];
        const float area_i = box_area_rn(bi);
        auto gs = [&](int j) -> float {
          float s_ = vs[j];
          if (p.iou_enable && j != i) {   // vote_ext.cpp:164-167: float exp() (glibc expf, <= 0.502 ulp) and a float multiply;
                                          // the correctly rounded value computed through double is what it returns bar near-ties
            const float d = __fsub_rn(1.f, iou_rn(bi, area_i, box[j]));
            const float e = __fdiv_rn(-__fmul_rn(d, d), p.sigma);
            s_ = __fmul_rn(s_, (float)exp((double)e));
          }
          return s_;
        };
        auto gx = [&](int j) -> float { return reinterpret_cast<const float*>(&box[j])[axis]; };
        // a row that had no overlap at all (nzrow bit clear) is a singleton cluster without looking at the matrix
        unsigned anym = 0u;
        if ((nzrow[i >> 5] >> (i & 31)) & 1u)
          for (int w_ = i >> 5; w_ < W; ++w_) anym |= mask[i * W + w_];
        if (anym == 0u) {
          // singleton cluster: the same operation sequence as vote_single_dim with n = 1 (vote_ext.cpp:8-35)
          const float s1 = vs[i], x = gx(i);
          const float mean = __fdiv_rn(__fadd_rn(0.f, __fmul_rn(s1, x)), __fadd_rn(0.f, s1));
          const float d = __fsub_rn(x, mean);
          const float sd = __fsqrt_rn(__fdiv_rn(__fadd_rn(0.f, __fmul_rn(__fmul_rn(s1, d), d)), __fadd_rn(0.f, s1)));
          const bool in = (__fsub_rn(mean, sd) <= x) && (x <= __fadd_rn(mean, sd));
          v = in ? __fdiv_rn(__fadd_rn(0.f, __fmul_rn(s1, x)), __fadd_rn(0.f, s1)) : __fdiv_rn(0.f, 0.f);
        } else {
          v = vote_axis_members(mask + i * W, W, i, gs, gx);
        }
      }
      sout[s * 5 + axis] = v;
      if (axis == 0) {
        sout[s * 5 + 4] = cs[i];       // cluster score = max over members = the seed's (vote_ext.cpp:196-197)
        skeys_out[s] = keys[i];
      }
    }
    __syncthreads();
    RADET_DBG(6);
    if (p.dbg && tid == 0) {
      p.dbg[(int64_t)bc * 16 + 8] = m;
      p.dbg[(int64_t)bc * 16 + 9] = ns;
    }
    return;
  }

  // ---------------- large-class fallback (m > 512): records in global memory, per-seed block-wide scan
  float4* box = p.g_box + (int64_t)bc * p.class_cap;
  float* vs = p.g_vs + (int64_t)bc * p.class_cap;
  int* owner = p.g_owner + (int64_t)bc * p.class_cap;
  int npad = 32;
  while (npad < m) npad <<= 1;
  // sort in place in the bin (padding lives beyond m only if npad <= class_cap; otherwise sort via odd-even fallback)
  if (npad <= p.class_cap) {
    for (int i = m + tid; i < npad; i += kClsThreads) gkeys[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(gkeys, npad);
  } else {
    // odd-even transposition sort (m steps); only reached when class_cap is not a power of two and m is close to it
    for (int step = 0; step < m; ++step) {
      for (int t = tid; 2 * t + (step & 1) + 1 < m; t += kClsThreads) {
        const int i = 2 * t + (step & 1);
        const u64 a = gkeys[i], b2 = gkeys[i + 1];
        if (a < b2) {
          gkeys[i] = b2;
          gkeys[i + 1] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < m; i += kClsThreads) {
    float4 bx;
    float c_, v_;
    decode_item(p, b, c, gkeys[i], bx, c_, v_);
    box[i] = bx;
    vs[i] = v_;
    owner[i] = -1;
  }
  __syncthreads();
  int ns = 0;
  for (int a = 0; a < m; ++a) {
    if (owner[a] != -1) continue;                          // block-uniform (written before the last barrier)
    if (p.mode == RADET_NMS_GLOBAL_VOTE && ns == 1) break;
    const float4 bi = box[a];
    const float area_i = box_area_rn(bi);
    for (int j = a + 1 + tid; j < m; j += kClsThreads) {
      if (owner[j] != -1) continue;
      const float iou = iou_rn(bi, area_i, box[j]);
      if (iou > p.thr) {
        owner[j] = a;
        if (p.iou_enable) {
          const float d = __fsub_rn(1.f, iou);
          const float e = __fdiv_rn(-__fmul_rn(d, d), p.sigma);
          vs[j] = __fmul_rn(vs[j], (float)exp((double)e));
        }
      }
    }
    if (tid == 0) owner[a] = a;
    ++ns;
    __syncthreads();
  }
  if (tid == 0) s_base = atomicAdd(&p.img_seed_count[b], ns);
  __syncthreads();
  float* sout = p.seed_out + ((int64_t)b * p.img_cap + s_base) * 5;
  u64* skeys_out = p.seed_keys + (int64_t)b * p.img_cap + s_base;
  // vote: thread per (cluster, axis), members found by scanning owners (O(m) per cluster; rare path)
  for (int t = tid; t < ns * 4; t += kClsThreads) {
    const int s = t >> 2, axis = t & 3;
    // locate the s-th seed: seeds are the boxes with owner == self, in order
    int i = -1, seen = 0;
    for (int a = 0; a < m; ++a) {
      if (owner[a] == a) {
        if (seen == s) {
          i = a;
          break;
        }
        ++seen;
      }
    }
    float v;
    if (p.mode == RADET_NMS_PLAIN) {
      v = reinterpret_cast<const float*>(&box[i])[axis];
    } else {
      float ss = 0.f, acc = 0.f;
      for (int j = i; j < m; ++j)
        if (owner[j] == i) {
          const float s_ = vs[j], x = reinterpret_cast<const float*>(&box[j])[axis];
          ss = __fadd_rn(ss, s_);
          acc = __fadd_rn(acc, __fmul_rn(s_, x));
        }
      const float mean = __fdiv_rn(acc, ss);
      float var = 0.f;
      for (int j = i; j < m; ++j)
        if (owner[j] == i) {
          const float d = __fsub_rn(reinterpret_cast<const float*>(&box[j])[axis], mean);
          var = __fadd_rn(var, __fmul_rn(__fmul_rn(vs[j], d), d));
        }
      const float sd = __fsqrt_rn(__fdiv_rn(var, ss));
      const float lo = __fsub_rn(mean, sd), hi = __fadd_rn(mean, sd);
      float fs = 0.f, fx = 0.f;
      for (int j = i; j < m; ++j)
        if (owner[j] == i) {
          const float x = reinterpret_cast<const float*>(&box[j])[axis];
          if (lo <= x && x <= hi) {
            fx = __fadd_rn(fx, __fmul_rn(vs[j], x));
            fs = __fadd_rn(fs, vs[j]);
          }
        }
      v = __fdiv_rn(fx, fs);
    }
    sout[s * 5 + axis] = v;
    if (axis == 0) {
      float4 bx;
      float c_, v_;
      decode_item(p, b, c, gkeys[i], bx, c_, v_);
      sout[s * 5 + 4] = c_;
      skeys_out[s] = gkeys[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------ 4. rank + emit
constexpr int kRankThreads = 512;
constexpr int kRankMaxOut = 4096;

struct RankParams {
  int C, img_cap, max_num;
  const float* seed_out;
  const u64* seed_keys;
  int* img_seed_count;  // re-armed here
  int* class_counts;    // re-armed here
  float* dets;          // [B][max_num][5]
  int64_t* labels;      // [B][max_num]
  int* num_dets;        // [B]
  long long* dbg;
};

constexpr int kRankNK = 8;   // seeds a thread keeps in registers (images with <= 4096 seeds read them once)

__global__ void __launch_bounds__(kRankThreads)
detect_rank_kernel(RankParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(16) SelectSmem s_sel;
  __shared__ int s_tot;
  const int b = blockIdx.x, tid = threadIdx.x;
#define RANK_DBG(k) do { if (p.dbg && tid == 0) p.dbg[(int64_t)(64 + b) * 16 + 10 + (k)] = clock64(); } while (0)
  RANK_DBG(0);
  u64* sel = reinterpret_cast<u64*>(smem_raw);          // [pad(max_num)] selected keys
  int* selidx = reinterpret_cast<int*>(sel + kRankMaxOut);  // position of each selected seed in the image's list
  const int S = min(p.img_seed_count[b], p.img_cap);
  __syncthreads();
  if (tid == 0) {
    p.img_seed_count[b] = 0;                               // re-arm for the next call
    s_tot = 0;
  }
  for (int c = tid; c < p.C; c += kRankThreads) p.class_counts[b * p.C + c] = 0;
  const int nkeep = min(S, p.max_num);
  if (tid == 0) p.num_dets[b] = nkeep;
  if (nkeep == 0) return;
  const u64* all = p.seed_keys + (int64_t)b * p.img_cap;
  KeyCache<kRankNK> kc;
  kc.load(all, S);
  u64 kth = 0ull;
  RANK_DBG(1);
  if (S > nkeep) kth = select_kth(kc, nkeep, &s_sel);
  __syncthreads();                                         // s_tot
  RANK_DBG(2);
  {
    int slot_of[kRankNK];
#pragma unroll
    for (int j = 0; j < kRankNK; ++j) slot_of[j] = (tid + j * kRankThreads < S && kc.r[j] >= kth) ? atomicAdd(&s_tot, 1) : 0;
#pragma unroll
    for (int j = 0; j < kRankNK; ++j)
      if (tid + j * kRankThreads < S && kc.r[j] >= kth) {
        sel[slot_of[j]] = kc.r[j];
        selidx[slot_of[j]] = tid + j * kRankThreads;
      }
  }
  for (int base = kRankNK * kRankThreads; base < S; base += kRankThreads) {
    const int i = base + tid;
    const u64 key = i < S ? all[i] : 0ull;
    const bool keep = i < S && key >= kth;
    if (keep) {
      const int slot = atomicAdd(&s_tot, 1);
      sel[slot] = key;
      selidx[slot] = i;
    }
  }
  __syncthreads();
  if (nkeep <= 256) {
    // descending order by counting: rank = #{selected keys above mine} (unique keys), broadcast shared-memory reads, no
    // barriers; two threads share a key and split the sweep
    const int r = tid >> 1, half = tid & 1;
    int rank = 0;
    u64 mine = 0ull;
    int mi = 0;
    if (r < nkeep) {
      mine = sel[r];
      mi = selidx[r];
      const int mid = (nkeep + 1) >> 1, j0 = half ? mid : 0, j1 = half ? nkeep : mid;
      for (int j = j0; j < j1; ++j) rank += sel[j] > mine ? 1 : 0;
    }
    rank += __shfl_xor_sync(kFull, rank, 1);
    __syncthreads();
    if (r < nkeep && half == 0) {
      sel[rank] = mine;
      selidx[rank] = mi;
    }
    __syncthreads();
  } else {
    int npad = 32;
    while (npad < nkeep) npad <<= 1;
    for (int i = nkeep + tid; i < npad; i += kRankThreads) {
      sel[i] = 0ull;
      selidx[i] = 0;
    }
    __syncthreads();
    // descending bitonic sort of (key, index) pairs
    for (int k = 2; k <= npad; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (npad >> 1); t += kRankThreads) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int ixj = i | j;
          const bool desc = (i & k) == 0;
          const u64 a = sel[i], bk = sel[ixj];
          if ((a < bk) == desc) {
            sel[i] = bk;
            sel[ixj] = a;
            const int ti = selidx[i];
            selidx[i] = selidx[ixj];
            selidx[ixj] = ti;
          }
        }
        __syncthreads();
      }
    }
  }
  for (int t = tid; t < nkeep * 5; t += kRankThreads) {
    const int r = t / 5, k = t - 5 * r;
    p.dets[((int64_t)b * p.max_num + r) * 5 + k] = p.seed_out[((int64_t)b * p.img_cap + selidx[r]) * 5 + k];
    if (k == 0) {
      // the class is part of the key's order field: ord = level << 27 | (point * C + class)
      const unsigned ord = 0xffffffffu - (unsigned)(sel[r] & 0xffffffffull);
      p.labels[(int64_t)b * p.max_num + r] = (int64_t)((ord & ((1u << kOrdLevelShift) - 1u)) % (unsigned)p.C);
    }
  }
}

}  // namespace radet

// ================================================================================================ C ABI
using namespace radet;

static int class_cap_of(const GridDev& g, int nms_pre) {
  // one class can hold at most one candidate per point and level, and at most nms_pre per level
  int64_t cap = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int64_t hw = (int64_t)g.h[l] * g.w[l];
    cap += (nms_pre > 0 && nms_pre < hw) ? nms_pre : hw;
  }
  return (int)(cap > (1 << 24) ? (1 << 24) : cap);
}

static int sel_plan(const GridDev& g, int B, int C, SelPlan* plan) {
  int blocks = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int hw = g.h[l] * g.w[l];
    if ((int64_t)hw * C >= (1ll << kOrdLevelShift)) return RADET_E_UNSUPPORTED;
    plan->bpl[l] = ((hw + 3) / 4 + kSelThreads - 1) / kSelThreads;
    plan->boff[l] = blocks;
    plan->coff[l] = (int64_t)C * g.off[l];
    blocks += plan->bpl[l];
  }
  for (int l = g.num_levels; l <= RADET_MAX_LEVELS; ++l) {
    plan->boff[l] = blocks;
    plan->coff[l] = (int64_t)C * g.off[g.num_levels];
  }
  for (int l = g.num_levels; l < RADET_MAX_LEVELS; ++l) plan->bpl[l] = 0;
  // class chunks: two rounds of kSelChunk planes per CTA.  The kernel is latency-bound; fewer, longer-lived CTAs cost the
  // other kernels in flight less than many short ones (measured: 7 chunks of 3 classes -> 3 chunks of 7 at C = 21).
  int nj_target = (C + 2 * kSelChunk - 1) / (2 * kSelChunk);
  if (nj_target < 1) nj_target = 1;
  int cc = (C + nj_target - 1) / nj_target;
  if (cc < 1) cc = 1;
  plan->cc = cc;
  plan->nj = (C + cc - 1) / cc;
  return RADET_OK;
}

// ------------------------------------------------------------------------------------------------ with_nms=False
// radet_head.py:165-169: every candidate that survived score_thr and the per-level top-k, un-suppressed, as rows
// [x1,y1,x2,y2 (decoded, clamped, rescaled), score*centerness, anchor x1,y1,x2,y2 (rescaled)] + its class.
// One CTA per (image, class) bin; rows are written class-major and by descending score*centerness inside a class
// (the reference's own order inside a level is whatever topk(sorted=False) returns).
struct EmitParams {
  ClsParams cls;      // grid, maps, C, class_cap, rescale, img_shapes, scale_factors, class_counts, bins (cs_mode = 0)
  float* rows;        // [B][img_cap][9]
  int64_t* labels;    // [B][img_cap]
  int* num;           // [B]
  int* done;          // [B] CTAs finished (re-armed here, like class_counts)
};

__global__ void __launch_bounds__(kClsThreads)
detect_emit_kernel(EmitParams e) {
  const ClsParams& p = e.cls;
  __shared__ int s_base, s_total;
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  if (tid < 32) {                                  // exclusive prefix of the class counts of this image
    int before = 0, total = 0;
    for (int k = tid; k < p.C; k += 32) {
      const int n = min(p.class_counts[b * p.C + k], p.class_cap);
      total += n;
      if (k < c) before += n;
    }
    before = __reduce_add_sync(kFull, before);
    total = __reduce_add_sync(kFull, total);
    if (tid == 0) {
      s_base = before;
      s_total = total;
    }
  }
  __syncthreads();
  const int m = min(p.class_counts[b * p.C + c], p.class_cap);
  const u64* keys = p.bins + ((int64_t)b * p.C + c) * p.class_cap;
  const GridDev& g = p.grid;
  for (int i = tid; i < m; i += kClsThreads) {
    const u64 key = keys[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) rank += keys[j] > key ? 1 : 0;
    float4 bx;
    float cs, vs;
    decode_item(p, b, c, key, bx, cs, vs);
    const unsigned ord = 0xffffffffu - (unsigned)(key & 0xffffffffull);
    const int l = (int)(ord >> kOrdLevelShift);
    const int q = (int)((ord & ((1u << kOrdLevelShift) - 1u)) / (unsigned)p.C);
    const int y = q / g.w[l], x = q - y * g.w[l];
    const float st = (float)g.stride[l], half = __fmul_rn(0.5f, __fmul_rn(g.anchor_scale, st));   // anchor_generator.py:142-185
    float4 an = make_float4(__fsub_rn((float)x * st, half), __fsub_rn((float)y * st, half), __fadd_rn((float)x * st, half),
                            __fadd_rn((float)y * st, half));
    if (p.rescale) {                                // radet_head.py:141-143
      const float4 sf = *reinterpret_cast<const float4*>(p.scale_factors + 4 * b);
      an = make_float4(__fdiv_rn(an.x, sf.x), __fdiv_rn(an.y, sf.y), __fdiv_rn(an.z, sf.z), __fdiv_rn(an.w, sf.w));
    }
    const int64_t row = (int64_t)b * p.img_cap + s_base + rank;
    float* o = e.rows + row * 9;
    o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
    o[4] = cs;
    o[5] = an.x; o[6] = an.y; o[7] = an.z; o[8] = an.w;
    e.labels[row] = c;
  }
  __syncthreads();
  if (tid == 0) {
    if (c == 0) e.num[b] = s_total;
    __threadfence();
    if (atomicAdd(&e.done[b], 1) == p.C - 1) {     // last class CTA of the image: re-arm the counters
      for (int k = 0; k < p.C; ++k) p.class_counts_rw[b * p.C + k] = 0;
      e.done[b] = 0;
    }
  }
}

struct DetWs {
  int* counts;        // [B][8]
  int* class_counts;  // [B][C]
  int* img_seed_count;  // [B]
  u64* cand;          // [B][C*P]
  u64* bins;          // [B][C][cap]
  u64* seed_keys;     // [B][img_cap]
  float* seed_out;    // [B][img_cap][5]
  float4* g_box;      // [B][C][cap]
  float* g_vs;
  int* g_owner;
  size_t total;
};

static int img_cap_of(const GridDev& g, int C, int nms_pre) {
  int64_t cap = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int64_t full = (int64_t)g.h[l] * g.w[l] * C;
    cap += (nms_pre > 0 && nms_pre < full) ? nms_pre : full;
  }
  return (int)(cap > (1 << 28) ? (1 << 28) : cap);
}

static DetWs det_ws_layout(unsigned char* base, const GridDev& g, int B, int C, int cap, int img_cap) {
  DetWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  // the three counter arrays first: they are what must be zero before the first call
  w.counts = reinterpret_cast<int*>(take((size_t)B * RADET_MAX_LEVELS * 4));
  w.class_counts = reinterpret_cast<int*>(take((size_t)B * C * 4));
  w.img_seed_count = reinterpret_cast<int*>(take((size_t)B * 4));
  w.cand = reinterpret_cast<u64*>(take((size_t)B * g.off[g.num_levels] * C * 8));
  w.bins = reinterpret_cast<u64*>(take((size_t)B * C * cap * 8));
  w.seed_keys = reinterpret_cast<u64*>(take((size_t)B * img_cap * 8));
  w.seed_out = reinterpret_cast<float*>(take((size_t)B * img_cap * 20));
  w.g_box = reinterpret_cast<float4*>(take((size_t)B * C * cap * 16));
  w.g_vs = reinterpret_cast<float*>(take((size_t)B * C * cap * 4));
  w.g_owner = reinterpret_cast<int*>(take((size_t)B * C * cap * 4));
  w.total = off;
  return w;
}

extern "C" size_t radet_get_bboxes_workspace_bytes(const radet_grid_t* grid, int32_t batch, int32_t num_classes,
                                                   const radet_detect_cfg_t* cfg) {
  GridDev g;
  if (make_grid_dev(grid, &g) != RADET_OK || batch <= 0 || num_classes <= 0 || !cfg) return 0;
  return det_ws_layout(nullptr, g, batch, num_classes, class_cap_of(g, cfg->nms_pre), img_cap_of(g, num_classes, cfg->nms_pre)).total;
}

extern "C" int radet_get_bboxes(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_maps_t* maps,
                                const int32_t* img_shapes, const float* scale_factors, const radet_detect_cfg_t* cfg,
                                float* dets, int64_t* labels, int32_t* num_dets, void* workspace, size_t workspace_bytes,
                                void* stream) {
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc != RADET_OK) return rc;
  if (batch <= 0 || num_classes <= 0 || !maps || !img_shapes || !cfg || !dets || !labels || !num_dets || !workspace) return RADET_E_BADARG;
  if (cfg->rescale && !scale_factors) return RADET_E_BADARG;
  if (cfg->max_per_img <= 0 || cfg->nms_mode < 0 || cfg->nms_mode > 2) return RADET_E_BADARG;
  if (cfg->max_per_img > kRankMaxOut || batch > 65535 || num_classes > 65535) return RADET_E_UNSUPPORTED;
  if (cfg->cluster_score_mode < 0 || cfg->cluster_score_mode > 2 || cfg->vote_score_mode < 0 || cfg->vote_score_mode > 2) return RADET_E_BADARG;
  if (workspace_bytes < radet_get_bboxes_workspace_bytes(grid, batch, num_classes, cfg) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return RADET_E_WORKSPACE;
  MapsDev md;
  for (int l = 0; l < RADET_MAX_LEVELS; ++l) {
    const bool on = l < g.num_levels;
    md.cls[l] = on ? maps->cls[l] : nullptr;
    md.bbox[l] = on ? maps->bbox[l] : nullptr;
    md.iou[l] = on ? maps->iou[l] : nullptr;
    if (on && (!md.cls[l] || !md.bbox[l] || !md.iou[l])) return RADET_E_BADARG;
    if (on && ((g.h[l] * g.w[l]) & 3) == 0 && (reinterpret_cast<uintptr_t>(md.cls[l]) & 15)) return RADET_E_BADARG;
  }
  SelPlan plan;
  rc = sel_plan(g, batch, num_classes, &plan);
  if (rc != RADET_OK) return rc;
  const int cap = class_cap_of(g, cfg->nms_pre);
  const int img_cap = img_cap_of(g, num_classes, cfg->nms_pre);
  DetWs w = det_ws_layout(static_cast<unsigned char*>(workspace), g, batch, num_classes, cap, img_cap);
  cudaStream_t st = (cudaStream_t)stream;
  // conservative logit prefilter: sigmoid(x) > thr  =>  x > logit(thr) - margin
  float x_lo = -INFINITY;
  const double thr = (double)cfg->score_thr;
  if (thr >= 1.0) x_lo = INFINITY;
  else if (thr > 0.0) {
    const double lg = log(thr / (1.0 - thr));
    x_lo = (float)(lg - 1e-3 * (1.0 + fabs(lg)));
  }
  detect_select_kernel<<<dim3(plan.boff[g.num_levels], batch, plan.nj), kSelThreads, 0, st>>>(g, plan, num_classes, md, cfg->score_thr,
                                                                                              x_lo, w.cand, w.counts);
  RADET_LAUNCH_CHECK();
  BinParams bp{};
  bp.grid = g;
  bp.maps = md;
  for (int l = 0; l <= RADET_MAX_LEVELS; ++l) bp.coff[l] = plan.coff[l];
  bp.C = num_classes;
  bp.nms_pre = cfg->nms_pre;
  bp.cs_mode = cfg->cluster_score_mode;
  bp.dbg = static_cast<long long*>(g_debug_buf);
  bp.class_cap = cap;
  bp.cand = w.cand;
  bp.counts = w.counts;
  bp.bins = w.bins;
  bp.class_counts = w.class_counts;
  detect_bin_kernel<<<dim3(g.num_levels, batch), kBinThreads, 0, st>>>(bp);
  RADET_LAUNCH_CHECK();
  ClsParams cp{};
  cp.grid = g;
  cp.maps = md;
  cp.C = num_classes;
  cp.class_cap = cap;
  cp.rescale = cfg->rescale;
  cp.cs_mode = cfg->cluster_score_mode;
  cp.vs_mode = cfg->vote_score_mode;
  cp.mode = cfg->nms_mode;
  cp.iou_enable = cfg->iou_enable;
  cp.thr = cfg->iou_threshold;
  cp.sigma = cfg->sigma;
  cp.img_shapes = img_shapes;
  cp.scale_factors = scale_factors;
  cp.class_counts = w.class_counts;
  cp.bins = w.bins;
  cp.img_cap = img_cap;
  cp.seed_out = w.seed_out;
  cp.seed_keys = w.seed_keys;
  cp.img_seed_count = w.img_seed_count;
  cp.dbg = static_cast<long long*>(g_debug_buf);
  cp.g_box = w.g_box;
  cp.g_vs = w.g_vs;
  cp.g_owner = w.g_owner;
  const size_t cls_smem = (size_t)kMaskItems * (8 + 16 + 4 + 4 + kMaskWords * 4 + 2 + 4) + 2 * kMaskWords * 4;
  cudaFuncSetAttribute(class_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cls_smem);
  cudaFuncSetAttribute(class_nms_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  class_nms_kernel<<<dim3(num_classes, batch), kClsThreads, cls_smem, st>>>(cp);
  RADET_LAUNCH_CHECK();
  RankParams rp{};
  rp.C = num_classes;
  rp.img_cap = img_cap;
  rp.max_num = cfg->max_per_img;
  rp.seed_out = w.seed_out;
  rp.seed_keys = w.seed_keys;
  rp.img_seed_count = w.img_seed_count;
  rp.class_counts = w.class_counts;
  rp.dets = dets;
  rp.labels = labels;
  rp.num_dets = num_dets;
  rp.dbg = static_cast<long long*>(g_debug_buf);
  const size_t rank_smem = (size_t)kRankMaxOut * 12;
  cudaFuncSetAttribute(detect_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rank_smem);
  detect_rank_kernel<<<batch, kRankThreads, rank_smem, st>>>(rp);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

// bbox2result (core/bbox/transforms.py:99-116), batched: per image, detections regrouped by class in their original
// (score) order.  One warp per image; rows <= 1024.
__global__ void bbox2result_kernel(const float* __restrict__ dets, const int64_t* __restrict__ labels,
                                   const int* __restrict__ num, int max_rows, int C, int xywh, float* __restrict__ out,
                                   int* __restrict__ class_offsets) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = min(num[b], max_rows);
  const float* d = dets + (int64_t)b * max_rows * 5;
  const int64_t* lb = labels + (int64_t)b * max_rows;
  float* o = out + (int64_t)b * max_rows * 5;
  int* off = class_offsets + (int64_t)b * (C + 1);
  // offsets: exclusive prefix of the class histogram (serial over C per lane-strided chunk is enough at these sizes)
  for (int c = lane; c <= C; c += 32) {
    int before = 0;
    for (int i = 0; i < n; ++i) before += (lb[i] >= 0 && lb[i] < c) ? 1 : 0;
    off[c] = before;
  }
  for (int i = lane; i < n; i += 32) {
    const int64_t c = lb[i];
    if (c < 0 || c >= C) continue;                 // labels outside [0, C) appear in no class list (labels == i never true)
    int pos = 0;
    for (int j = 0; j < n; ++j) pos += (lb[j] >= 0 && (lb[j] < c || (lb[j] == c && j < i))) ? 1 : 0;
    float x1 = d[i * 5 + 0], y1 = d[i * 5 + 1], x2 = d[i * 5 + 2], y2 = d[i * 5 + 3];
    if (xywh) {                                    // BOPDataset.xyxy2xywh (datasets/bop.py): [x1, y1, x2 - x1, y2 - y1]
      x2 = __fsub_rn(x2, x1);
      y2 = __fsub_rn(y2, y1);
    }
    o[pos * 5 + 0] = x1; o[pos * 5 + 1] = y1; o[pos * 5 + 2] = x2; o[pos * 5 + 3] = y2; o[pos * 5 + 4] = d[i * 5 + 4];
  }
}

extern "C" int radet_bbox2result(const float* dets, const int64_t* labels, const int32_t* num, int32_t batch, int32_t max_rows,
                                 int32_t num_classes, int32_t xywh, float* out, int32_t* class_offsets, void* stream) {
  if (batch == 0) return RADET_OK;
  if (batch < 0 || max_rows < 0 || num_classes <= 0 || !dets || !labels || !num || !out || !class_offsets) return RADET_E_BADARG;
  if (max_rows > 1024) return RADET_E_UNSUPPORTED;
  bbox2result_kernel<<<batch, 32, 0, (cudaStream_t)stream>>>(dets, labels, num, max_rows, num_classes, xywh, out, class_offsets);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

// radet_head.py:165-169 (with_nms=False): select -> per-level top-k / class bins -> emit
extern "C" int radet_get_candidates(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_maps_t* maps,
                                    const int32_t* img_shapes, const float* scale_factors, const radet_detect_cfg_t* cfg,
                                    float* rows, int64_t* labels, int32_t* num_rows, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc != RADET_OK) return rc;
  if (batch <= 0 || num_classes <= 0 || !maps || !img_shapes || !cfg || !rows || !labels || !num_rows || !workspace) return RADET_E_BADARG;
  if (cfg->rescale && !scale_factors) return RADET_E_BADARG;
  if (batch > 65535 || num_classes > 65535) return RADET_E_UNSUPPORTED;
  if (workspace_bytes < radet_get_bboxes_workspace_bytes(grid, batch, num_classes, cfg) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return RADET_E_WORKSPACE;
  MapsDev md;
  for (int l = 0; l < RADET_MAX_LEVELS; ++l) {
    const bool on = l < g.num_levels;
    md.cls[l] = on ? maps->cls[l] : nullptr;
    md.bbox[l] = on ? maps->bbox[l] : nullptr;
    md.iou[l] = on ? maps->iou[l] : nullptr;
    if (on && (!md.cls[l] || !md.bbox[l] || !md.iou[l])) return RADET_E_BADARG;
    if (on && ((g.h[l] * g.w[l]) & 3) == 0 && (reinterpret_cast<uintptr_t>(md.cls[l]) & 15)) return RADET_E_BADARG;
  }
  SelPlan plan;
  rc = sel_plan(g, batch, num_classes, &plan);
  if (rc != RADET_OK) return rc;
  const int cap = class_cap_of(g, cfg->nms_pre);
  const int img_cap = img_cap_of(g, num_classes, cfg->nms_pre);
  DetWs w = det_ws_layout(static_cast<unsigned char*>(workspace), g, batch, num_classes, cap, img_cap);
  cudaStream_t st = (cudaStream_t)stream;
  float x_lo = -INFINITY;
  const double thr = (double)cfg->score_thr;
  if (thr >= 1.0) x_lo = INFINITY;
  else if (thr > 0.0) {
    const double lg = log(thr / (1.0 - thr));
    x_lo = (float)(lg - 1e-3 * (1.0 + fabs(lg)));
  }
  detect_select_kernel<<<dim3(plan.boff[g.num_levels], batch, plan.nj), kSelThreads, 0, st>>>(g, plan, num_classes, md, cfg->score_thr,
                                                                                              x_lo, w.cand, w.counts);
  RADET_LAUNCH_CHECK();
  BinParams bp{};
  bp.grid = g;
  bp.maps = md;
  for (int l = 0; l <= RADET_MAX_LEVELS; ++l) bp.coff[l] = plan.coff[l];
  bp.C = num_classes;
  bp.nms_pre = cfg->nms_pre;
  bp.cs_mode = 0;                 // bins keyed by score * centerness
  bp.class_cap = cap;
  bp.cand = w.cand;
  bp.counts = w.counts;
  bp.bins = w.bins;
  bp.class_counts = w.class_counts;
  detect_bin_kernel<<<dim3(g.num_levels, batch), kBinThreads, 0, st>>>(bp);
  RADET_LAUNCH_CHECK();
  EmitParams ep{};
  ep.cls.grid = g;
  ep.cls.maps = md;
  ep.cls.C = num_classes;
  ep.cls.class_cap = cap;
  ep.cls.rescale = cfg->rescale;
  ep.cls.cs_mode = 0;
  ep.cls.vs_mode = 0;
  ep.cls.img_shapes = img_shapes;
  ep.cls.scale_factors = scale_factors;
  ep.cls.class_counts = w.class_counts;
  ep.cls.class_counts_rw = w.class_counts;
  ep.cls.bins = w.bins;
  ep.cls.img_cap = img_cap;
  ep.rows = rows;
  ep.labels = labels;
  ep.num = num_rows;
  ep.done = w.img_seed_count;
  detect_emit_kernel<<<dim3(num_classes, batch), kClsThreads, 0, st>>>(ep);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int64_t radet_candidates_capacity(const radet_grid_t* grid, int32_t num_classes, int32_t nms_pre) {
  GridDev g;
  if (make_grid_dev(grid, &g) != RADET_OK || num_classes <= 0) return 0;
  return img_cap_of(g, num_classes, nms_pre);
}
