// Vote-NMS family on explicit box lists (radet.ops.vote_nms / global_vote_nms / cluster_nms), sm_100a.
//
// Reference semantics: vote_ext.cpp:70-353 and cluster_ext.cpp:4-87 (single-threaded C++ on CPU tensors).
// One CTA per list: bitonic sort by cluster score (ties: lower row first), grouping by label, one warp per label
// segment doing the greedy clustering with warp-wide IoU tests, seed ranking by block scan, sigma-filtered weighted
// box vote per kept cluster.  Lists up to 5120 boxes live in shared memory, longer ones in the global workspace.
// All arithmetic that feeds a decision or an output uses explicit round-to-nearest intrinsics (no FMA contraction,
// reference operation order): keep sets, cluster ids and voted boxes are bit-exact with the reference.
// (The head's get_bboxes path uses the multi-kernel pipeline in detect.cu instead.)
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace radet {

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
  const int* offsets;  // device copy of list offsets [batch+1]
  const float* in_boxes;
  const float* in_cs;
  const float* in_vs;
  const int64_t* in_labels;
  float thr, sigma;
  int iou_enable, mode, max_num;
  int cap;  // capacity of the per-list arrays (longest list)
  float* out_dets;
  int64_t* out_labels;
  int64_t* out_index;
  int* num_out;
  int64_t* instance_ids;
  int64_t* clusters_num;
  unsigned char* gws;  // global arrays (vs/orig always; everything when !kSmem)
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

template <bool kSmem>
__global__ void __launch_bounds__(kNmsThreads)
nms_list_kernel(NmsParams p) {
  using IdxT = typename std::conditional<kSmem, short, int>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_scan[34];
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

  // ---------------------------------------------------------------- A. gather items, build sort keys
  int n;
  {
    const int r0 = p.offsets[b];
    n = min(p.offsets[b + 1] - r0, cap);
    for (int i = tid; i < n; i += kNmsThreads)
      A.keys[i] = ((unsigned long long)float_order_key(p.in_cs[r0 + i]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
  }
  int npad = 32;
  while (npad < n) npad <<= 1;
  for (int i = n + tid; i < npad; i += kNmsThreads) A.keys[i] = 0ull;
  __syncthreads();

  const int out_base = p.offsets[b];
  const int out_cap = p.offsets[b + 1] - p.offsets[b];
  if (n == 0) {
    if (tid == 0) p.num_out[b] = 0;
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
    {
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
      const int ow = (int)A.owner[ia];
      __syncwarp();                                         // every lane has read owner[ia] before lane 0 rewrites it
      if (ow != -1) continue;                               // warp-uniform
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
          if (p.iou_enable) {  // :164-167: float exp() and a float multiply (see detect.cu)
            const float d = __fsub_rn(1.f, iou);
            const float e = __fdiv_rn(-__fmul_rn(d, d), p.sigma);
            A.vs[ij] = __fmul_rn(A.vs[ij], (float)exp((double)e));
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
  if (p.instance_ids) {
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
}

}  // namespace radet

// ================================================================================================ C ABI
using namespace radet;

static size_t nms_smem_bytes() {
  return (size_t)kNmsCapPad * 8 + (size_t)kNmsCap * (16 + 4 + 4 + 2 + 2 + 2);
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
    cudaFuncSetAttribute(nms_list_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_smem_bytes());
    nms_list_kernel<true><<<batch, kNmsThreads, nms_smem_bytes(), st>>>(p);
  } else {
    nms_list_kernel<false><<<batch, kNmsThreads, 0, st>>>(p);
  }
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}
