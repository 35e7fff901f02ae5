// Inference: score threshold -> per-level top-k -> TBLR decode -> class-aware (vote-)NMS, on sm_100a.
//
// Reference semantics: ATSSHead.get_bboxes (atss_head.py:326-387) + RADetHead._get_bboxes_single
// (radet_head.py:55-169), which loops over images and levels in Python, syncs the host per level, copies four
// tensors per image to the CPU and runs the single-threaded O(n^2) vote_ext.cpp.  Here the whole batch is two
// launches, nothing leaves the device:
//
//   detect_select_kernel  HBM-bound scan of the class logits IN PLACE (NCHW, 128-bit streaming loads): sigmoid,
//                         strict `> score_thr`, survivors appended as (score, flat index) keys per (image, level).
//   nms_image_kernel      one CTA per image, working set in shared memory: exact per-level top-k (radix select),
//                         centerness gather, bitonic sort by cluster score, decode+clamp+rescale, grouping by label,
//                         one warp per (image, class) segment doing the greedy clustering with warp-wide IoU tests,
//                         seed ranking by block scan, then the sigma-filtered weighted box vote per kept cluster.
//
// All arithmetic that feeds a discrete decision or an output box uses explicit round-to-nearest intrinsics
// (__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn): no FMA contraction, same operation order as vote_ext.cpp, so keep
// sets and voted boxes are bit-exact with the reference.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace radet {

struct MapsDev {
  const float* cls[RADET_MAX_LEVELS];
  const float* bbox[RADET_MAX_LEVELS];
  const float* iou[RADET_MAX_LEVELS];
};

// torch's CUDA sigmoid: 1 / (1 + exp(-x)) in fp32 with IEEE division (radet_head.py:106-109 run on CUDA tensors)
__device__ __forceinline__ float sigmoid_rn(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

constexpr int kSelThreads = 256;
constexpr int kOrdLevelShift = 27;  // ord = level << 27 | flat (point*C + class)

struct SelTable {
  int uoff[RADET_MAX_LEVELS + 1];
  int upl[RADET_MAX_LEVELS];
  int64_t coff[RADET_MAX_LEVELS + 1];  // candidate-buffer offset of each level inside one image (= C * off[l])
};

__global__ void __launch_bounds__(kSelThreads)
detect_select_kernel(GridDev grid, SelTable tab, int B, int C, int cc, int nj, MapsDev maps, float thr, float x_lo,
                     unsigned long long* __restrict__ cand, int* __restrict__ counts) {
  const int U = tab.uoff[grid.num_levels];
  const int64_t t = (int64_t)blockIdx.x * kSelThreads + threadIdx.x;
  if (t >= (int64_t)U * nj) return;
  const int j = (int)(t / U), u = (int)(t - (int64_t)j * U);
  int l = 0;
#pragma unroll
  for (int k = 1; k < RADET_MAX_LEVELS; ++k) l += (k < grid.num_levels && u >= tab.uoff[k]) ? 1 : 0;
  const int ul = u - tab.uoff[l];
  const int b = ul / tab.upl[l];
  const int q0 = 4 * (ul - b * tab.upl[l]);
  const int hw = grid.h[l] * grid.w[l];
  const int nv = min(4, hw - q0);
  const bool vec = (hw & 3) == 0;
  const float* cp = maps.cls[l] + ((int64_t)b * C) * hw + q0;
  unsigned long long* out = cand + (int64_t)b * tab.coff[grid.num_levels] + tab.coff[l];
  int* cnt = counts + b * RADET_MAX_LEVELS + l;
  const int c0 = j * cc, c1 = min(C, c0 + cc);
#pragma unroll 4
  for (int c = c0; c < c1; ++c) {
    float xv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (vec) {
      const float4 v4 = ldg_stream4(cp + (int64_t)c * hw);
      xv[0] = v4.x; xv[1] = v4.y; xv[2] = v4.z; xv[3] = v4.w;
    } else {
      for (int i = 0; i < nv; ++i) xv[i] = cp[(int64_t)c * hw + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < nv && xv[i] > x_lo) {                 // cheap conservative prefilter on the logit
        const float s = sigmoid_rn(xv[i]);
        if (s > thr) {                              // radet_head.py:111 (strict)
          const unsigned flat = (unsigned)(q0 + i) * (unsigned)C + (unsigned)c;
          const int slot = atomicAdd(cnt, 1);
          out[slot] = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xffffffffu - flat);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ per-image NMS
constexpr int kNmsThreads = 1024;
constexpr int kNmsCap = 5120;       // shared-memory capacity (>= 5 levels x nms_pre 1000)
constexpr int kNmsCapPad = 8192;

template <typename IdxT>
struct NmsArrays {
  unsigned long long* keys;  // [pad]
  float4* box;               // [cap] sorted by cluster score
  float* cs;                 // [cap]
  int* lab;                  // [cap]
  IdxT* owner;               // [cap] -1 free, -2 dropped, else seed position
  IdxT* perm;                // [cap] label-grouped order -> score order
  IdxT* ipos;                // [cap] inverse of perm
  float* vs;                 // [cap] (global) vote score, possibly iou-weighted
  int* orig;                 // [cap] (global) row of the input list / ord
};

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const bool desc = (i & k) == 0;
        const unsigned long long a = keys[i], b = keys[ixj];
        if ((a < b) == desc) {
          keys[i] = b;
          keys[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
}

// vote_single_dim (vote_ext.cpp:8-35), fp32, sequential in member order, one rounding per operation
template <typename IdxT>
__device__ float vote_axis(const NmsArrays<IdxT>& A, int seed, int n, int axis) {
  const int lab = A.lab[seed];
  const int j0 = (int)A.ipos[seed];
  float ss = 0.f, acc = 0.f;
  for (int j = j0; j < n; ++j) {
    const int i = (int)A.perm[j];
    if (A.lab[i] != lab) break;
    if ((int)A.owner[i] != seed) continue;
    const float s = A.vs[i];
    const float x = reinterpret_cast<const float*>(&A.box[i])[axis];
    ss = __fadd_rn(ss, s);
    acc = __fadd_rn(acc, __fmul_rn(s, x));
  }
  const float mean = __fdiv_rn(acc, ss);
  float var = 0.f;
  for (int j = j0; j < n; ++j) {
    const int i = (int)A.perm[j];
    if (A.lab[i] != lab) break;
    if ((int)A.owner[i] != seed) continue;
    const float s = A.vs[i];
    const float x = reinterpret_cast<const float*>(&A.box[i])[axis];
    const float d = __fsub_rn(x, mean);
    var = __fadd_rn(var, __fmul_rn(__fmul_rn(s, d), d));
  }
  const float sd = __fsqrt_rn(__fdiv_rn(var, ss));
  const float lo = __fsub_rn(mean, sd), hi = __fadd_rn(mean, sd);
  float fs = 0.f, fx = 0.f;
  for (int j = j0; j < n; ++j) {
    const int i = (int)A.perm[j];
    if (A.lab[i] != lab) break;
    if ((int)A.owner[i] != seed) continue;
    const float x = reinterpret_cast<const float*>(&A.box[i])[axis];
    if (lo <= x && x <= hi) {
      const float s = A.vs[i];
      fx = __fadd_rn(fx, __fmul_rn(s, x));
      fs = __fadd_rn(fs, s);
    }
  }
  return __fdiv_rn(fx, fs);
}

struct NmsParams {
  // head source
  GridDev grid;
  SelTable tab;
  MapsDev maps;
  int C;
  const unsigned long long* cand;
  int* counts;
  const int* img_shapes;
  const float* scale_factors;
  int nms_pre, rescale, cs_mode, vs_mode;
  // list source
  const int* offsets;  // device copy of list offsets [batch+1]
  const float* in_boxes;
  const float* in_cs;
  const float* in_vs;
  const int64_t* in_labels;
  // common
  float thr, sigma;
  int iou_enable, mode, max_num;
  int cap;          // capacity of the per-image arrays
  int out_stride;   // rows of out_* per image (head) ; list: rows start at offsets[b]
  float* out_dets;
  int64_t* out_labels;
  int64_t* out_index;
  int* num_out;
  int64_t* instance_ids;
  int64_t* clusters_num;
  // global arrays (vs/orig always; everything when !kSmem)
  unsigned char* gws;
  size_t gws_per_image;
};

__host__ __device__ inline size_t nms_global_bytes(int cap, bool smem_variant) {
  size_t s = (size_t)cap * 8;  // vs + orig
  if (!smem_variant) {
    int pad = 32;
    while (pad < cap) pad <<= 1;
    s += (size_t)pad * 8 + (size_t)cap * (16 + 4 + 4 + 4 + 4 + 4);
  }
  return (s + 255) & ~size_t(255);
}

template <bool kHead, bool kSmem>
__global__ void __launch_bounds__(kNmsThreads)
nms_image_kernel(NmsParams p) {
  using IdxT = typename std::conditional<kSmem, short, int>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[34];
  __shared__ int s_n, s_nseg, s_hist[256], s_misc[4];
  __shared__ unsigned long long s_prefix;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cap = p.cap;
  NmsArrays<IdxT> A;
  unsigned char* gw = p.gws + (size_t)b * p.gws_per_image;
  A.vs = reinterpret_cast<float*>(gw);
  A.orig = reinterpret_cast<int*>(gw + (size_t)cap * 4);
  if (kSmem) {
    unsigned char* c = smem_raw;
    A.keys = reinterpret_cast<unsigned long long*>(c); c += (size_t)kNmsCapPad * 8;
    A.box = reinterpret_cast<float4*>(c); c += (size_t)kNmsCap * 16;
    A.cs = reinterpret_cast<float*>(c); c += (size_t)kNmsCap * 4;
    A.lab = reinterpret_cast<int*>(c); c += (size_t)kNmsCap * 4;
    A.owner = reinterpret_cast<IdxT*>(c); c += (size_t)kNmsCap * sizeof(IdxT);
    A.perm = reinterpret_cast<IdxT*>(c); c += (size_t)kNmsCap * sizeof(IdxT);
    A.ipos = reinterpret_cast<IdxT*>(c);
  } else {
    int pad = 32;
    while (pad < cap) pad <<= 1;
    unsigned char* c = gw + (size_t)cap * 8;
    A.keys = reinterpret_cast<unsigned long long*>(c); c += (size_t)pad * 8;
    A.box = reinterpret_cast<float4*>(c); c += (size_t)cap * 16;
    A.cs = reinterpret_cast<float*>(c); c += (size_t)cap * 4;
    A.lab = reinterpret_cast<int*>(c); c += (size_t)cap * 4;
    A.owner = reinterpret_cast<IdxT*>(c); c += (size_t)cap * 4;
    A.perm = reinterpret_cast<IdxT*>(c); c += (size_t)cap * 4;
    A.ipos = reinterpret_cast<IdxT*>(c);
  }
  if (tid == 0) s_n = 0;
  __syncthreads();

  // ---------------------------------------------------------------- A. gather items, build sort keys
  int n;
  if (kHead) {
    const GridDev& g = p.grid;
    for (int l = 0; l < g.num_levels; ++l) {
      const unsigned long long* src = p.cand + (int64_t)b * p.tab.coff[g.num_levels] + p.tab.coff[l];
      const int nl = p.counts[b * RADET_MAX_LEVELS + l];
      unsigned long long kth = 0ull;  // keep keys >= kth
      if (p.nms_pre > 0 && nl > p.nms_pre) {
        // exact k-th largest key by 8-bit radix select (keys are unique: score bits | ~flat index)
        int k = p.nms_pre;
        unsigned long long prefix = 0ull, pmask = 0ull;
        for (int shift = 56; shift >= 0; shift -= 8) {
          for (int i = tid; i < 256; i += kNmsThreads) s_hist[i] = 0;
          __syncthreads();
          for (int i = tid; i < nl; i += kNmsThreads) {
            const unsigned long long key = src[i];
            if ((key & pmask) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & 0xffull)], 1);
          }
          __syncthreads();
          if (tid == 0) {
            int acc = 0, bin = 255;
            for (; bin > 0; --bin) {
              if (acc + s_hist[bin] >= k) break;
              acc += s_hist[bin];
            }
            s_misc[0] = bin;
            s_misc[1] = k - acc;
          }
          __syncthreads();
          prefix |= (unsigned long long)s_misc[0] << shift;
          pmask |= 0xffull << shift;
          k = s_misc[1];
          __syncthreads();
        }
        kth = prefix;
      }
      const int hw = g.h[l] * g.w[l];
      for (int i = tid; i < nl; i += kNmsThreads) {
        const unsigned long long key = src[i];
        if (key < kth) continue;
        const float S = __uint_as_float((unsigned)(key >> 32));
        const unsigned flat = 0xffffffffu - (unsigned)(key & 0xffffffffull);
        const int q = (int)(flat / (unsigned)p.C);
        float cs = S;
        if (p.cs_mode != 1) {
          const float ctr = sigmoid_rn(p.maps.iou[l][(int64_t)b * hw + q]);      // radet_head.py:109
          cs = p.cs_mode == 0 ? __fmul_rn(S, ctr) : ctr;                          // vote_wrapper.py:14-21
        }
        const unsigned ord = ((unsigned)l << kOrdLevelShift) | flat;
        const int slot = atomicAdd(&s_n, 1);
        if (slot < cap) A.keys[slot] = ((unsigned long long)float_order_key(cs) << 32) | (unsigned long long)(0xffffffffu - ord);
      }
    }
    __syncthreads();
    n = min(s_n, cap);
  } else {
    const int r0 = p.offsets[b];
    n = min(p.offsets[b + 1] - r0, cap);
    for (int i = tid; i < n; i += kNmsThreads)
      A.keys[i] = ((unsigned long long)float_order_key(p.in_cs[r0 + i]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
  }
  int npad = 32;
  while (npad < n) npad <<= 1;
  for (int i = n + tid; i < npad; i += kNmsThreads) A.keys[i] = 0ull;
  __syncthreads();

  const int out_base = kHead ? b * p.out_stride : p.offsets[b];
  const int out_cap = kHead ? p.out_stride : (p.offsets[b + 1] - p.offsets[b]);
  if (n == 0) {
    if (tid == 0) p.num_out[b] = 0;
    if (kHead && tid == 0)
      for (int l = 0; l < p.grid.num_levels; ++l) p.counts[b * RADET_MAX_LEVELS + l] = 0;  // re-arm
    return;
  }

  // ---------------------------------------------------------------- B. sort by cluster score (desc), ties by order
  bitonic_sort_desc(A.keys, npad);

  // ---------------------------------------------------------------- C. records in score order
  for (int i = tid; i < n; i += kNmsThreads) {
    const unsigned long long key = A.keys[i];
    const unsigned ord = 0xffffffffu - (unsigned)(key & 0xffffffffull);
    float4 bx;
    float cs, vs;
    int lab;
    if (kHead) {
      const GridDev& g = p.grid;
      const int l = (int)(ord >> kOrdLevelShift);
      const unsigned flat = ord & ((1u << kOrdLevelShift) - 1u);
      const int q = (int)(flat / (unsigned)p.C);
      lab = (int)(flat - (unsigned)q * (unsigned)p.C);
      const int hw = g.h[l] * g.w[l];
      const int y = q / g.w[l], x = q - y * g.w[l];
      const float st = (float)g.stride[l];
      const float cx = (float)x * st, cy = (float)y * st;
      const float side = __fmul_rn(g.anchor_scale, st);
      const float* bp = p.maps.bbox[l] + (int64_t)b * 4 * hw + q;
      // tblr2bboxes (tblr_bbox_coder.py:154-171): (v * normalizer) * side, then centre -/+
      const float T = __fmul_rn(__fmul_rn(bp[0], g.nrm), side), Bt = __fmul_rn(__fmul_rn(bp[hw], g.nrm), side);
      const float L = __fmul_rn(__fmul_rn(bp[2 * hw], g.nrm), side), R = __fmul_rn(__fmul_rn(bp[3 * hw], g.nrm), side);
      const float H = (float)p.img_shapes[2 * b], W = (float)p.img_shapes[2 * b + 1];
      bx.x = fminf(fmaxf(__fsub_rn(cx, L), 0.f), W);
      bx.y = fminf(fmaxf(__fsub_rn(cy, T), 0.f), H);
      bx.z = fminf(fmaxf(__fadd_rn(cx, R), 0.f), W);
      bx.w = fminf(fmaxf(__fadd_rn(cy, Bt), 0.f), H);
      if (p.rescale) {                                                    // radet_head.py:141-143
        const float4 sf = *reinterpret_cast<const float4*>(p.scale_factors + 4 * b);
        bx.x = __fdiv_rn(bx.x, sf.x); bx.y = __fdiv_rn(bx.y, sf.y);
        bx.z = __fdiv_rn(bx.z, sf.z); bx.w = __fdiv_rn(bx.w, sf.w);
      }
      const float S = sigmoid_rn(p.maps.cls[l][((int64_t)b * p.C + lab) * hw + q]);
      const float ctr = sigmoid_rn(p.maps.iou[l][(int64_t)b * hw + q]);
      cs = p.cs_mode == 0 ? __fmul_rn(S, ctr) : (p.cs_mode == 1 ? S : ctr);
      vs = p.vs_mode == 0 ? __fmul_rn(S, ctr) : (p.vs_mode == 1 ? S : ctr);
      A.orig[i] = (int)ord;
    } else {
      const int r = p.offsets[b] + (int)ord;
      bx = *reinterpret_cast<const float4*>(p.in_boxes + 4 * (int64_t)r);
      cs = p.in_cs[r];
      vs = p.in_vs[r];
      lab = (int)p.in_labels[r];
      A.orig[i] = (int)ord;
    }
    A.box[i] = bx;
    A.cs[i] = cs;
    A.vs[i] = vs;
    A.lab[i] = lab;
    A.owner[i] = (IdxT)-1;
  }
  __syncthreads();

  // ---------------------------------------------------------------- D. group by label (stable in score order)
  for (int i = tid; i < npad; i += kNmsThreads)
    A.keys[i] = i < n ? (((unsigned long long)(0xffffffffu - ((unsigned)A.lab[i] ^ 0x80000000u)) << 32) |
                         (unsigned long long)(0xffffffffu - (unsigned)i))
                      : 0ull;
  __syncthreads();
  bitonic_sort_desc(A.keys, npad);
  for (int j = tid; j < n; j += kNmsThreads) {
    const int i = (int)(0xffffffffu - (unsigned)(A.keys[j] & 0xffffffffull));
    A.perm[j] = (IdxT)i;
    A.ipos[i] = (IdxT)j;
  }
  __syncthreads();
  // segment starts -> compacted into keys[] (reused as int list)
  int* seg_start = reinterpret_cast<int*>(A.keys);
  __syncthreads();
  int nseg = 0;
  for (int base = 0; base < n; base += kNmsThreads) {
    const int j = base + tid;
    int flag = 0;
    if (j < n) flag = (j == 0) || (A.lab[(int)A.perm[j]] != A.lab[(int)A.perm[j - 1]]);
    int total;
    const int pos = block_exclusive_scan(flag, s_scan, &total);
    // keys[] still holds sort output needed above only for perm (already extracted) -> safe to overwrite,
    // but perm extraction of other threads must be complete: guaranteed by the __syncthreads before this loop
    if (flag) seg_start[nseg + pos] = j;
    nseg += total;
  }
  __syncthreads();

  // ---------------------------------------------------------------- E. greedy clustering, one warp per segment
  for (int sg = wid; sg < nseg; sg += kNmsThreads / 32) {
    const int s0 = seg_start[sg], s1 = (sg + 1 < nseg) ? seg_start[sg + 1] : n;
    for (int a = s0; a < s1; ++a) {
      const int ia = (int)A.perm[a];
      if ((int)A.owner[ia] != -1) continue;                 // warp-uniform
      if (p.mode == RADET_NMS_GLOBAL_VOTE && a != s0) {     // vote_ext.cpp:257-263: label already emitted
        if (lane == 0) A.owner[ia] = (IdxT)-2;
        __syncwarp();
        continue;
      }
      if (lane == 0) A.owner[ia] = (IdxT)ia;
      const float4 bi = A.box[ia];
      const float area_i = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
      for (int jj = a + 1 + lane; jj < s1; jj += 32) {
        const int ij = (int)A.perm[jj];
        if ((int)A.owner[ij] != -1) continue;
        const float4 bj = A.box[ij];
        const float xl = fmaxf(bj.x, bi.x), yt = fmaxf(bj.y, bi.y), xr = fminf(bj.z, bi.z), yb = fminf(bj.w, bi.w);
        const float iw = fmaxf(0.f, __fsub_rn(xr, xl)), ih = fmaxf(0.f, __fsub_rn(yb, yt));
        const float inter = __fmul_rn(iw, ih);
        const float area_j = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_j, area_i), inter));  // vote_ext.cpp:162
        if (iou > p.thr) {                                                                // :169 (strict; NaN -> false)
          A.owner[ij] = (IdxT)ia;
          if (p.iou_enable) {  // :164-167, exp() evaluated in double as in the reference build
            const float d = __fsub_rn(1.f, iou);
            const float e = __fdiv_rn(-__fmul_rn(d, d), p.sigma);
            A.vs[ij] = (float)((double)A.vs[ij] * exp((double)e));
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- F. rank seeds in score order
  int nclu = 0;
  // slots are written into keys[] region (as int) beyond the segment list: nseg <= n, so offset by n ints
  int* slot = reinterpret_cast<int*>(A.keys) + n;
  for (int base = 0; base < n; base += kNmsThreads) {
    const int i = base + tid;
    const int flag = (i < n && (int)A.owner[i] == i) ? 1 : 0;
    int total;
    const int pos = block_exclusive_scan(flag, s_scan, &total);
    if (i < n) slot[i] = flag ? nclu + pos : -1;
    nclu += total;
  }
  __syncthreads();
  int nkeep = nclu;
  if (p.max_num > 0 && nkeep > p.max_num) nkeep = p.max_num;
  if (nkeep > out_cap) nkeep = out_cap;
  if (tid == 0) p.num_out[b] = nkeep;

  // ---------------------------------------------------------------- G. box voting for the kept clusters
  for (int wk = tid; wk < n * 4; wk += kNmsThreads) {
    const int i = wk >> 2, axis = wk & 3;
    const int sl = slot[i];
    if (sl < 0 || sl >= nkeep) continue;
    float v;
    if (p.mode == RADET_NMS_PLAIN) v = reinterpret_cast<const float*>(&A.box[i])[axis];
    else v = vote_axis<IdxT>(A, i, n, axis);
    float* o = p.out_dets + (int64_t)(out_base + sl) * 5;
    o[axis] = v;
    if (axis == 0) {
      o[4] = A.cs[i];  // max cluster score of the cluster = the seed's (vote_ext.cpp:196-197)
      p.out_labels[out_base + sl] = (int64_t)A.lab[i];
      if (p.out_index) p.out_index[out_base + sl] = (int64_t)A.orig[i];
    }
  }
  // ---------------------------------------------------------------- H. cluster ids / sizes (cluster_ext.cpp:4-87)
  if (!kHead && p.instance_ids) {
    const int r0 = p.offsets[b];
    for (int i = tid; i < n; i += kNmsThreads) {
      const int ow = (int)A.owner[i];
      p.instance_ids[r0 + A.orig[i]] = ow >= 0 ? (int64_t)slot[ow] : 0;
      if (p.clusters_num) p.clusters_num[r0 + A.orig[i]] = 0;
    }
    __syncthreads();
    if (p.clusters_num) {
      for (int i = tid; i < n; i += kNmsThreads) {
        const int ow = (int)A.owner[i];
        if (ow >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(&p.clusters_num[r0 + A.orig[ow]]), 1ull);
      }
    }
  }
  if (kHead) {
    __syncthreads();
    if (tid < p.grid.num_levels) p.counts[b * RADET_MAX_LEVELS + tid] = 0;  // re-arm the candidate counters
  }
}

}  // namespace radet

// ================================================================================================ C ABI
using namespace radet;

static size_t nms_smem_bytes() {
  return (size_t)kNmsCapPad * 8 + (size_t)kNmsCap * (16 + 4 + 4 + 2 + 2 + 2);
}

static int sel_plan(const GridDev& g, int B, int C, SelTable* tab, int* cc, int* nj) {
  int64_t u = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int hw = g.h[l] * g.w[l];
    if ((int64_t)hw * C >= (1ll << kOrdLevelShift)) return RADET_E_UNSUPPORTED;
    tab->uoff[l] = (int)u;
    tab->upl[l] = (hw + 3) / 4;
    tab->coff[l] = (int64_t)C * g.off[l];
    u += (int64_t)B * tab->upl[l];
    if (u > (1ll << 30)) return RADET_E_BADARG;
  }
  for (int l = g.num_levels; l <= RADET_MAX_LEVELS; ++l) {
    tab->uoff[l] = (int)u;
    tab->coff[l] = (int64_t)C * g.off[g.num_levels];
  }
  for (int l = g.num_levels; l < RADET_MAX_LEVELS; ++l) tab->upl[l] = 1;
  const int64_t target_threads = 148ll * 2048 * 2;
  int c = (int)((u * (int64_t)C + target_threads - 1) / target_threads);
  if (c < 1) c = 1;
  if (c > C) c = C;
  *cc = c;
  *nj = (C + c - 1) / c;
  return RADET_OK;
}

static int head_item_cap(const GridDev& g, int C, int nms_pre) {
  int64_t cap = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int64_t full = (int64_t)g.h[l] * g.w[l] * C;
    cap += (nms_pre > 0 && nms_pre < full) ? nms_pre : full;
  }
  return (int)(cap > (1 << 28) ? (1 << 28) : cap);
}

extern "C" size_t radet_get_bboxes_workspace_bytes(const radet_grid_t* grid, int32_t batch, int32_t num_classes,
                                                   const radet_detect_cfg_t* cfg) {
  GridDev g;
  if (make_grid_dev(grid, &g) != RADET_OK || batch <= 0 || num_classes <= 0 || !cfg) return 0;
  const int cap = head_item_cap(g, num_classes, cfg->nms_pre);
  const bool smem = cap <= kNmsCap;
  return align_up((size_t)batch * RADET_MAX_LEVELS * 4, 256) +
         align_up((size_t)batch * g.off[g.num_levels] * num_classes * 8, 256) + (size_t)batch * nms_global_bytes(cap, smem);
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
  if (cfg->cluster_score_mode < 0 || cfg->cluster_score_mode > 2 || cfg->vote_score_mode < 0 || cfg->vote_score_mode > 2) return RADET_E_BADARG;
  if (workspace_bytes < radet_get_bboxes_workspace_bytes(grid, batch, num_classes, cfg) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return RADET_E_WORKSPACE;
  NmsParams p{};
  SelTable tab;
  int cc, nj;
  rc = sel_plan(g, batch, num_classes, &tab, &cc, &nj);
  if (rc != RADET_OK) return rc;
  for (int l = 0; l < RADET_MAX_LEVELS; ++l) {
    const bool on = l < g.num_levels;
    p.maps.cls[l] = on ? maps->cls[l] : nullptr;
    p.maps.bbox[l] = on ? maps->bbox[l] : nullptr;
    p.maps.iou[l] = on ? maps->iou[l] : nullptr;
    if (on && (!p.maps.cls[l] || !p.maps.bbox[l] || !p.maps.iou[l])) return RADET_E_BADARG;
    if (on && ((g.h[l] * g.w[l]) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.maps.cls[l]) & 15)) return RADET_E_BADARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  int* counts = reinterpret_cast<int*>(ws);
  ws += align_up((size_t)batch * RADET_MAX_LEVELS * 4, 256);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws);
  ws += align_up((size_t)batch * g.off[g.num_levels] * num_classes * 8, 256);
  // conservative logit prefilter: sigmoid(x) > thr  =>  x > logit(thr) - margin
  float x_lo = -INFINITY;
  const double thr = (double)cfg->score_thr;
  if (thr >= 1.0) x_lo = INFINITY;
  else if (thr > 0.0) {
    const double lg = log(thr / (1.0 - thr));
    x_lo = (float)(lg - 1e-3 * (1.0 + fabs(lg)));
  }
  const int64_t sthreads = (int64_t)tab.uoff[g.num_levels] * nj;
  detect_select_kernel<<<(unsigned)((sthreads + kSelThreads - 1) / kSelThreads), kSelThreads, 0, st>>>(
      g, tab, batch, num_classes, cc, nj, p.maps, cfg->score_thr, x_lo, cand, counts);
  RADET_LAUNCH_CHECK();
  p.grid = g;
  p.tab = tab;
  p.C = num_classes;
  p.cand = cand;
  p.counts = counts;
  p.img_shapes = img_shapes;
  p.scale_factors = scale_factors;
  p.nms_pre = cfg->nms_pre;
  p.rescale = cfg->rescale;
  p.cs_mode = cfg->cluster_score_mode;
  p.vs_mode = cfg->vote_score_mode;
  p.thr = cfg->iou_threshold;
  p.sigma = cfg->sigma;
  p.iou_enable = cfg->iou_enable;
  p.mode = cfg->nms_mode;
  p.max_num = cfg->max_per_img;
  p.cap = head_item_cap(g, num_classes, cfg->nms_pre);
  p.out_stride = cfg->max_per_img;
  p.out_dets = dets;
  p.out_labels = labels;
  p.out_index = nullptr;
  p.num_out = num_dets;
  p.gws = ws;
  const bool smem = p.cap <= kNmsCap;
  p.gws_per_image = nms_global_bytes(p.cap, smem);
  if (smem) {
    cudaFuncSetAttribute(nms_image_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem_bytes());
    nms_image_kernel<true, true><<<batch, kNmsThreads, nms_smem_bytes(), st>>>(p);
  } else {
    nms_image_kernel<true, false><<<batch, kNmsThreads, 0, st>>>(p);
  }
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" size_t radet_vote_nms_workspace_bytes(int32_t batch, int64_t total_boxes, int64_t max_boxes_per_list) {
  if (batch <= 0 || total_boxes < 0 || max_boxes_per_list < 0) return 0;
  const int cap = (int)(max_boxes_per_list < 1 ? 1 : max_boxes_per_list);
  return align_up((size_t)(batch + 1) * 4, 256) + (size_t)batch * nms_global_bytes(cap, cap <= kNmsCap);
}

extern "C" int radet_vote_nms(int32_t batch, const int32_t* offsets_host, const float* boxes, const float* cluster_scores,
                              const float* vote_scores, const int64_t* labels, float iou_threshold, int32_t iou_enable,
                              float sigma, int32_t mode, int32_t max_num, float* out_dets, int64_t* out_labels,
                              int64_t* out_index, int32_t* num_out, int64_t* instance_ids, int64_t* clusters_num,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (batch == 0) return RADET_OK;
  if (batch < 0 || !offsets_host || !num_out || !workspace || mode < 0 || mode > 2) return RADET_E_BADARG;
  int64_t maxn = 0;
  for (int b = 0; b < batch; ++b) {
    const int64_t nb = (int64_t)offsets_host[b + 1] - offsets_host[b];
    if (nb < 0) return RADET_E_BADARG;
    maxn = nb > maxn ? nb : maxn;
  }
  const int64_t total = offsets_host[batch];
  if (total > 0 && (!boxes || !cluster_scores || !vote_scores || !labels || !out_dets || !out_labels)) return RADET_E_BADARG;
  if (maxn >= (1ll << 28)) return RADET_E_UNSUPPORTED;
  if (workspace_bytes < radet_vote_nms_workspace_bytes(batch, total, maxn) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return RADET_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  int* d_off = reinterpret_cast<int*>(ws);
  ws += align_up((size_t)(batch + 1) * 4, 256);
  cudaError_t ce = cudaMemcpyAsync(d_off, offsets_host, (size_t)(batch + 1) * 4, cudaMemcpyHostToDevice, st);
  if (ce != cudaSuccess) return (int)ce;
  NmsParams p{};
  p.offsets = d_off;
  p.in_boxes = boxes;
  p.in_cs = cluster_scores;
  p.in_vs = vote_scores;
  p.in_labels = labels;
  p.thr = iou_threshold;
  p.sigma = sigma;
  p.iou_enable = iou_enable;
  p.mode = mode;
  p.max_num = max_num;
  p.cap = (int)(maxn < 1 ? 1 : maxn);
  p.out_dets = out_dets;
  p.out_labels = out_labels;
  p.out_index = out_index;
  p.num_out = num_out;
  p.instance_ids = instance_ids;
  p.clusters_num = clusters_num;
  p.gws = ws;
  const bool smem = p.cap <= kNmsCap;
  p.gws_per_image = nms_global_bytes(p.cap, smem);
  if (smem) {
    cudaFuncSetAttribute(nms_image_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem_bytes());
    nms_image_kernel<false, true><<<batch, kNmsThreads, nms_smem_bytes(), st>>>(p);
  } else {
    nms_image_kernel<false, false><<<batch, kNmsThreads, 0, st>>>(p);
  }
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}
