// Target encode + fused head loss (forward and backward) on sm_100a.
//
// Reference semantics: RADetHead.get_targets / loss (radet/models/dense_heads/radet_head.py:173-392) and the loss
// modules it calls (focal_loss.py, iou_loss.py, cross_entropy_loss.py, tblr_bbox_coder.py, iou2d_calculator.py).
// The reference flattens NCHW -> (point, channel) with permute+reshape+cat, materialises labels / bbox targets /
// anchors and runs ~80 eager kernels forward plus autograd backward.  Here:
//
//   loss_pos_kernel      one pass over (image, point): the sparse positive terms (IoU target, GIoU, BCE), their gradients, and
//                        the normalisers num_pos = sum w, sum wq.  Deterministic two-level reduction (block partials in
//                        fixed order, last block finalises).  ~12 B/point of traffic.
//   loss_dense_w_kernel  the HBM-bound pass (default): one warp per item of 128 / 256 points streams the class planes IN PLACE
//                        (NCHW) through a TMA-fed shared-memory ring, computes sigmoid-focal loss and its gradient with the
//                        final normaliser already applied and writes the gradient once (128-bit streaming stores).  Labels
//                        and TBLR targets are rebuilt on the fly from points_to_gt_index (no label / target / anchor
//                        tensors).  Launched as the programmatic dependent of loss_pos_kernel: with the assignment's
//                        per-image weight sums it runs NEXT to it ("overlapped" order, radet_loss_fwd_bwd).
//   loss_dense_tma_kernel / loss_dense_kernel   the round-2 / round-1 predecessors (4-warp TMA tiles; register-pipelined
//                        loads): the latter serves plane sizes that are not multiples of 4, the former RADET_LOSS_IMPL=tma.
//
// Algorithmic traffic: (8C + 52) B/point  (C logits read + C grads written + 16 B bbox + 4 B iou read, 20 B grads
// written, 8 B index + 4 B weight read).
#include "loss_common.cuh"

namespace radet {

// ------------------------------------------------------------------------------------------------ targets
// AnchorGenerator.grid_anchors / valid_flags (anchor_generator.py:206-298) for one image: priors [P,4] and, optionally, the
// flags of the cells inside ceil(pad_shape / stride).  The hot path never materialises them; this serves get_anchors().
__global__ void grid_priors_kernel(GridDev grid, int pad_h, int pad_w, float4* __restrict__ anchors, uint8_t* __restrict__ flags) {
  const int P = grid.off[grid.num_levels];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int l = level_of(grid, p);
  const int q = p - grid.off[l];
  const int y = q / grid.w[l], x = q - y * grid.w[l];
  const int s = grid.stride[l];
  if (anchors) {
    const float cx = (float)(x * s), cy = (float)(y * s);
    const float half = 0.5f * (grid.anchor_scale * (float)s);
    anchors[p] = make_float4(cx - half, cy - half, cx + half, cy + half);
  }
  if (flags) {
    const int vh = min((pad_h + s - 1) / s, grid.h[l]), vw = min((pad_w + s - 1) / s, grid.w[l]);
    flags[p] = (y < vh && x < vw) ? 1 : 0;
  }
}

__global__ void get_targets_kernel(GridDev grid, int B, int C, const int* __restrict__ gt_offsets,
                                   const float* __restrict__ gt_bboxes, const int64_t* __restrict__ gt_labels,
                                   const int64_t* __restrict__ pidx, const float* __restrict__ pw,
                                   int64_t* __restrict__ labels, float4* __restrict__ targets,
                                   float* __restrict__ weights, float4* __restrict__ anchors) {
  const int P = grid.off[grid.num_levels];
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)B * P) return;
  const int b = (int)(t / P), p = (int)(t - (int64_t)b * P);
  const int l = level_of(grid, p);
  const int q = p - grid.off[l];
  const int hw = grid.h[l] * grid.w[l];
  const int y = q / grid.w[l], x = q - y * grid.w[l];
  const int s = grid.stride[l];
  const float cx = (float)(x * s), cy = (float)(y * s);
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  const int64_t idx = pidx[t];
  const int64_t row = (int64_t)B * grid.off[l] + (int64_t)b * hw + q;  // level-major, image-minor (radet_head.py:356-368)
  labels[row] = label_of(idx, G, gt_labels + g0, C);
  float4 tg = make_float4(0.f, 0.f, 0.f, 0.f);
  const float side = grid.anchor_scale * (float)s;
  if (G > 0 && idx > 0) {
    const int k = (int)((idx - 1) < (int64_t)(G - 1) ? (idx - 1) : (int64_t)(G - 1));
    const float4 gb = *reinterpret_cast<const float4*>(gt_bboxes + 4 * (int64_t)(g0 + k));
    const float nrm = grid.nrm;  // TBLRBBoxCoder(normalizer=1/8): python float 0.125
    tg.x = __fdiv_rn(__fdiv_rn(cy - gb.y, side), nrm);
    tg.y = __fdiv_rn(__fdiv_rn(gb.w - cy, side), nrm);
    tg.z = __fdiv_rn(__fdiv_rn(cx - gb.x, side), nrm);
    tg.w = __fdiv_rn(__fdiv_rn(gb.z - cx, side), nrm);
  }
  targets[row] = tg;
  weights[row] = pw[t];
  if (anchors) {
    const float half = 0.5f * side;
    anchors[row] = make_float4(cx - half, cy - half, cx + half, cy + half);
  }
}

// "Overlapped" order (radet_loss_fwd_bwd): loss_pos_kernel zero-fills the regression / IoU gradient planes itself and leaves
// the positives' un-normalised gradient values in a compact list; the dense kernel, running next to it, writes the
// normalised values into the planes at its end.  Unused (all null) in the classic order.
struct PosList {
  int* plist;                        // scratch [n]: flat (image, point) index of every positive of the batch (-1: skip) ...
  float* pvals;                      // scratch [5n]: ... and its five un-normalised gradient values (T, B, L, R, IoU logit)
};

constexpr int kPosGroups = kPosPerThread / 4;   // a thread scans kPosGroups groups of 4 consecutive points, kPosThreads * 4 apart

__global__ void __launch_bounds__(kPosThreads)
loss_pos_kernel(GridDev grid, int B, MapsDev maps, const int* __restrict__ gt_offsets,
                const float* __restrict__ gt_bboxes, const int64_t* __restrict__ pidx, const float* __restrict__ pw,
                radet_loss_cfg_t cfg, LossWs* __restrict__ ws, double* __restrict__ partials, GradsDev grads, PosList fin,
                unsigned long long* dbg) {
#define PDBG(k) do { if (dbg && threadIdx.x == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); dbg[((size_t)50000 + blockIdx.x) * 8 + (k)] = t__; } } while (0)
  PDBG(0);
  // Phase 1: every thread scans kPosPerThread points (groups of 4 consecutive ones) and the CTA compacts the (sparse,
  // spatially clustered) positives into shared memory.  Phase 2: one positive per thread -- IoU target, GIoU and BCE terms
  // and their gradients, evaluated ONCE here.
  //   Classic order: the un-normalised gradient values are parked in the gradient planes themselves at the positives' cells;
  //   the dense kernel (next in the stream) rescales those cells while it zero-fills the rest of the regression / IoU planes.
  //   Overlapped order (fin.plist): this kernel zero-fills the planes itself and the un-normalised values go to a compact
  //   list; the dense kernel writes the normalised values into the planes once both kernels' class-plane work is done.
  // CTAs are small (128 threads) so that a whole grid of them fits NEXT to the resident items of the dense kernel.
  __shared__ int s_scan[34];
  __shared__ int s_list[kPosThreads * kPosPerThread];
  __shared__ int s_base;
  // Programmatic dependent launch: a dense kernel that follows in the stream as a programmatic dependent may start now; it
  // executes griddepcontrol.wait before it touches anything written here.  Otherwise a no-op.
  griddep_launch_dependents();
  const bool want_grad = grads.bbox[0] != nullptr;
  const bool finish = fin.plist != nullptr;
  const int P = grid.off[grid.num_levels];
  const int64_t n = (int64_t)B * P;
  const int64_t blk0 = (int64_t)blockIdx.x * kPosThreads * kPosPerThread;
  unsigned flags = 0u;
#pragma unroll
  for (int g = 0; g < kPosGroups; ++g) {
    const int64_t t0 = blk0 + (int64_t)g * (kPosThreads * 4) + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t0 + k < n && pidx[t0 + k] >= 0) flags |= 1u << (4 * g + k);   // pos_inds incl. ignored points (radet_head.py:245-247)
    }
  }
  if (finish && want_grad) {
    // zero the regression / IoU gradient cells of this thread's points (plane sizes are multiples of 4 in this order, so the 4
    // points of a group lie in one plane and the stores are aligned); the positives are written by the dense kernel
#pragma unroll
    for (int g = 0; g < kPosGroups; ++g) {
      const int64_t t0 = blk0 + (int64_t)g * (kPosThreads * 4) + threadIdx.x * 4;
      if (t0 < n) {
        const int b = (int)(t0 / P), p = (int)(t0 - (int64_t)b * P);
        const int l = level_of(grid, p);
        const int q = p - grid.off[l];
        const int hw = grid.h[l] * grid.w[l];
        float* gb = grads.bbox[l] + (int64_t)b * 4 * hw + q;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        stg_stream4(gb, z);
        stg_stream4(gb + hw, z);
        stg_stream4(gb + 2 * hw, z);
        stg_stream4(gb + 3 * hw, z);
        stg_stream4(grads.iou[l] + (int64_t)b * hw + q, z);
      }
    }
  }
  int total;
  int pos = block_exclusive_scan(__popc(flags), s_scan, &total);
  PDBG(1);
  if (finish && want_grad && threadIdx.x == 0 && total > 0) s_base = (int)atomicAdd(&ws->pad[0], (unsigned)total);   // list slots of this CTA
  while (flags) {
    const int k = __ffs((int)flags) - 1;
    flags &= flags - 1u;
    s_list[pos++] = (k >> 2) * (kPosThreads * 4) + threadIdx.x * 4 + (k & 3);
  }
  __syncthreads();
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < total; i += kPosThreads) {
    const int64_t t = blk0 + s_list[i];
    const int64_t idx = pidx[t];
    const int b = (int)(t / P), p = (int)(t - (int64_t)b * P);
    const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
    int listed = -1;
    float pv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (G > 0) {
      const float w = pw[t];
      const int l = level_of(grid, p);
      const int q = p - grid.off[l];
      const int hw = grid.h[l] * grid.w[l];
      const int y = q / grid.w[l], x = q - y * grid.w[l];
      const float st = (float)grid.stride[l];
      const float s = grid.nrm * (grid.anchor_scale * st);
      const float cx = (float)x * st, cy = (float)y * st;
      const float* bp = maps.bbox[l] + (int64_t)b * 4 * hw + q;
      const float T = bp[0], Bt = bp[hw], L = bp[2 * hw], R = bp[3 * hw];
      const float xi = maps.iou[l][(int64_t)b * hw + q];
      float tT, tB, tL, tR;
      point_target(idx, G, gt_bboxes + 4 * (int64_t)g0, cx, cy, tT, tB, tL, tR);
      const BoxTerms bt = box_terms<true>(cx, cy, s, T, Bt, L, R, tT, tB, tL, tR, 1e-6f, cfg.eps);
      const float wq = fmaxf(bt.iou, 1e-12f) * w;           // radet_head.py:272
      acc[0] += w;
      acc[1] += wq;
      acc[2] += wq * (1.f - bt.giou);
      acc[3] += w * bce_logits(xi, bt.iou);
      acc[4] += (T + Bt) + (L + R);
      acc[5] += xi;
      pv[0] = -wq * bt.d[0];                                // d(1 - giou) = -d giou
      pv[1] = -wq * bt.d[1];
      pv[2] = -wq * bt.d[2];
      pv[3] = -wq * bt.d[3];
      pv[4] = w * (sigmoidf_(xi) - bt.iou);
      listed = (int)t;                                      // n < 2^31 in the dense-first order (radet_loss_fwd_bwd)
      if (want_grad && !finish) {
        float* gb = grads.bbox[l] + (int64_t)b * 4 * hw + q;
        gb[0] = pv[0];
        gb[hw] = pv[1];
        gb[2 * hw] = pv[2];
        gb[3 * hw] = pv[3];
        grads.iou[l][(int64_t)b * hw + q] = pv[4];
      }
    }
    if (want_grad && finish) {
      const int64_t o = (int64_t)s_base + i;
      fin.plist[o] = listed;
#pragma unroll
      for (int k = 0; k < 5; ++k) fin.pvals[o * 5 + k] = pv[k];
    }
  }
  PDBG(2);
  __shared__ double s_part[kPosThreads / 32][6];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool warp_any = __any_sync(kFull, threadIdx.x < total);   // most warps hold no positive: skip their shuffles
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float v = warp_any ? warp_sum(acc[k]) : 0.f;
    if (lane == 0) s_part[wid][k] = (double)v;
  }
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < kPosThreads / 32; ++w) v += s_part[w][threadIdx.x];
    partials[(int64_t)blockIdx.x * 6 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&ws->counter_pos, 1u) == gridDim.x - 1);
  __syncthreads();
  PDBG(3);
  if (!s_last) return;
  __threadfence();
  // last block: fixed-order reduction over block partials (deterministic)
  double tot[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kPosThreads) {
#pragma unroll
    for (int k = 0; k < 6; ++k) tot[k] += __ldcg(partials + (int64_t)i * 6 + k);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double v = warp_sum(tot[k]);
    if (lane == 0) s_part[wid][k] = v;
  }
  __syncthreads();
  __shared__ double s_fin[6];
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < kPosThreads / 32; ++w) v += s_part[w][threadIdx.x];
    s_fin[threadIdx.x] = v;
    ws->norm[threadIdx.x] = v;
    // norm[0], norm[1] are the NORMALISERS (a caller running the opt-in FCOS-style reduce_mean all-reduces them
    // between the two passes); norm[6], norm[7] keep the rank-local values (radet_head.py:254 is rank-local)
    if (threadIdx.x < 2) ws->norm[6 + threadIdx.x] = v;
  }
  if (threadIdx.x == 0) ws->counter_pos = 0u;  // re-arm for the next call on this workspace
  PDBG(4);
#undef PDBG
}

// ------------------------------------------------------------------------------------------------ dense pass
constexpr int kDenseThreads = 256;

struct DenseTable {
  int uoff[RADET_MAX_LEVELS + 1];  // unit (4-point group) offsets per level over the whole batch
  int upl[RADET_MAX_LEVELS];       // units per (image, level) plane
};

constexpr int kDG = 2;   // class planes per load group (pairs ping-pong between two register sets)

template <bool kGamma2>
__global__ void __launch_bounds__(kDenseThreads, 3)
loss_dense_kernel(GridDev grid, DenseTable tab, int B, int C, int cc /* channels per work item */, int nj, MapsDev maps,
                  GradsDev grads, const int* __restrict__ gt_offsets, const float* __restrict__ gt_bboxes,
                  const int64_t* __restrict__ gt_labels, const int64_t* __restrict__ pidx, const float* __restrict__ pw,
                  radet_loss_cfg_t cfg, const float* __restrict__ grad_scale, LossWs* __restrict__ ws,
                  double* __restrict__ partials, float* __restrict__ losses) {
  // One thread = one unit (4 consecutive points of one (image, level) plane) x the channel chunk blockIdx.y.
  // Channels 0..C-1 are the class logits (focal loss + gradient); channels C..C+3 are the T,B,L,R gradient planes and
  // C+4 the IoU-logit gradient plane (zero except at the sparse positives), so every output plane is written by the
  // same streaming loop.  The per-point setup (index -> label, weight) is done once per thread and amortised over
  // the whole chunk (all C+5 channels when the units alone fill the machine); 3 CTAs x 256 threads per SM with 4
  // independent 128-bit loads in flight per thread keep ~48 KB outstanding per SM, enough for HBM latency x bandwidth.
  const int P = grid.off[grid.num_levels];
  const int U = tab.uoff[grid.num_levels];
  const int CH = C + 5;
  const double num_pos = ws->norm[0], sum_wq = ws->norm[1];
  const bool has_pos = ws->norm[6] > 0.0;                                        // radet_head.py:261
  const float gs_cls = grad_scale ? grad_scale[0] : 1.f, gs_box = grad_scale ? grad_scale[1] : 1.f,
              gs_iou = grad_scale ? grad_scale[2] : 1.f;
  const float k_cls = gs_cls * cfg.w_cls / (float)(num_pos + (double)cfg.avg_extra);
  const float k_box = has_pos ? gs_box * cfg.w_bbox / (float)sum_wq : gs_box;
  const float k_iou = has_pos ? gs_iou * cfg.w_iou / (float)num_pos : gs_iou;
  const bool want_grad = grads.cls[0] != nullptr;
  const float gamma = cfg.gamma, alpha = cfg.alpha;
  float lsum = 0.f;
  // Balanced mapping: the U x nj work items are cut into gridDim.x equal contiguous slices (gridDim.x is a multiple of
  // the SM count), so every SM hosts the same number of equally loaded CTAs.
  const int64_t items = (int64_t)U * nj;
  const int64_t ipc = (items + gridDim.x - 1) / gridDim.x;
  const int64_t w_end = min(items, (int64_t)(blockIdx.x + 1) * ipc);
  for (int64_t wi = (int64_t)blockIdx.x * ipc + threadIdx.x; wi < w_end; wi += kDenseThreads) {
    const int j = (int)(wi / U), u = (int)(wi - (int64_t)j * U);
    const int c0 = j * cc, c1 = min(CH, c0 + cc);
    const int cls_end = min(c1, C);
    int l = 0;
#pragma unroll
    for (int k = 1; k < RADET_MAX_LEVELS; ++k) l += (k < grid.num_levels && u >= tab.uoff[k]) ? 1 : 0;
    const int ul = u - tab.uoff[l];
    const int b = ul / tab.upl[l];
    const int q0 = 4 * (ul - b * tab.upl[l]);
    const int hw = grid.h[l] * grid.w[l];
    const int nv = min(4, hw - q0);
    const bool vec = (hw & 3) == 0;
    const float* src = maps.cls[l] + ((int64_t)b * C + c0) * hw + q0;
    float* dst = want_grad ? grads.cls[l] + ((int64_t)b * C + c0) * hw + q0 : nullptr;
    const int ncls = cls_end - c0;                  // class planes of this thread (<= 0: regression chunk only)
    const int64_t pstride = hw;                     // plane stride in floats
    // Software pipeline over pairs of planes with two register sets (A, B) that ping-pong: no copies, no per-plane
    // predicates in the steady state, plane addresses advance by pointer increments.
    const int npairs = vec ? ncls / 2 : 0;          // full pairs; an odd last plane is handled after the loop
    float4 a0, a1, b0, b1;
    a0 = a1 = b0 = b1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (npairs > 0) {                               // in flight before the index -> label dependency chain
      a0 = ldg_stream4(src);
      a1 = ldg_stream4(src + pstride);
    }
    const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
    const int64_t pbase = (int64_t)b * P + grid.off[l] + q0;
    int idx[4];   // 1-based GT index fits 32 bits (G <= RADET_MAX_GT_PER_IMAGE)
    float w[4], kw[4];
    int lab[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      idx[i] = -1;
      w[i] = 0.f;
      if (i < nv) {
        const int64_t v = pidx[pbase + i];
        idx[i] = v < 0 ? -1 : (int)(v > (int64_t)G ? (int64_t)G : v);
        w[i] = pw[pbase + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lab[i] = (int)label_of(idx[i], G, gt_labels + g0, C) - c0;   // relative to the chunk
      kw[i] = k_cls * w[i];
    }
    auto plane = [&](const float4& xv, int cr, float* out) {
      float lo_, gr_;
      float4 gv;
      focal_elem<kGamma2>(xv.x, lab[0] == cr, gamma, alpha, lo_, gr_); lsum = fmaf(w[0], lo_, lsum); gv.x = kw[0] * gr_;
      focal_elem<kGamma2>(xv.y, lab[1] == cr, gamma, alpha, lo_, gr_); lsum = fmaf(w[1], lo_, lsum); gv.y = kw[1] * gr_;
      focal_elem<kGamma2>(xv.z, lab[2] == cr, gamma, alpha, lo_, gr_); lsum = fmaf(w[2], lo_, lsum); gv.z = kw[2] * gr_;
      focal_elem<kGamma2>(xv.w, lab[3] == cr, gamma, alpha, lo_, gr_); lsum = fmaf(w[3], lo_, lsum); gv.w = kw[3] * gr_;
      if (out) stg_stream4(out, gv);
    };
    {
      const float* sp_ = src + 2 * pstride;         // next pair to load
      float* dp_ = dst;                             // next pair to store
      int cr = 0;
      int left = npairs;
      while (left >= 2) {                           // two pairs per trip: compute A while B loads, then B while A loads
        b0 = ldg_stream4(sp_);
        b1 = ldg_stream4(sp_ + pstride);
        plane(a0, cr, dp_);
        plane(a1, cr + 1, dp_ ? dp_ + pstride : nullptr);
        if (left > 2) {
          a0 = ldg_stream4(sp_ + 2 * pstride);
          a1 = ldg_stream4(sp_ + 3 * pstride);
        }
        plane(b0, cr + 2, dp_ ? dp_ + 2 * pstride : nullptr);
        plane(b1, cr + 3, dp_ ? dp_ + 3 * pstride : nullptr);
        sp_ += 4 * pstride;
        if (dp_) dp_ += 4 * pstride;
        cr += 4;
        left -= 2;
      }
      if (left == 1) {                              // one full pair left (already in A)
        plane(a0, cr, dp_);
        plane(a1, cr + 1, dp_ ? dp_ + pstride : nullptr);
        if (dp_) dp_ += 2 * pstride;
        cr += 2;
      }
      if (vec && cr < ncls)                         // odd last plane
        plane(ldg_stream4(src + (int64_t)cr * pstride), cr, dst ? dst + (int64_t)cr * pstride : nullptr);
    }
    if (!vec) {                                     // planes whose size is not a multiple of 4: scalar path
      for (int cr = 0; cr < ncls; ++cr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {                // static indices keep lab/w/kw in registers
          if (i < nv) {
            float lo_, gr_;
            focal_elem<kGamma2>(src[(int64_t)cr * hw + i], lab[i] == cr, gamma, alpha, lo_, gr_);
            lsum = fmaf(w[i], lo_, lsum);
            if (dst) dst[(int64_t)cr * hw + i] = kw[i] * gr_;
          }
        }
      }
    }
    // gradient planes of the regression / IoU branches (channels C .. C+4): zero everywhere except at the positives,
    // whose un-normalised values loss_pos_kernel parked in these planes: rescale those cells in place
    if (want_grad && c1 > C) {
      const bool anypos = G > 0 && (idx[0] >= 0 || idx[1] >= 0 || idx[2] >= 0 || idx[3] >= 0);
      for (int ch = max(c0, C); ch < c1; ++ch) {
        const int kk = ch - C;  // 0..3: T,B,L,R ; 4: iou logit
        float* o = kk < 4 ? grads.bbox[l] + ((int64_t)b * 4 + kk) * hw + q0 : grads.iou[l] + (int64_t)b * hw + q0;
        const float kn = kk < 4 ? k_box : k_iou, gs = kk < 4 ? gs_box : gs_iou;
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
        if (anypos) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nv && idx[i] >= 0) gv[i] = has_pos ? kn * o[i] : gs;        // radet_head.py:280-281 when num_pos == 0
        }
        if (vec) {
          stg_stream4(o, make_float4(gv[0], gv[1], gv[2], gv[3]));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nv) o[i] = gv[i];
        }
      }
    }
  }
  // loss_cls = w_cls * sum / (num_pos + num_imgs): deterministic block partials, last block finalises
  __shared__ double s_part[kDenseThreads / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned nblocks = gridDim.x, bid = blockIdx.x;
  const double v = warp_sum((double)lsum);
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w_ = 0; w_ < kDenseThreads / 32; ++w_) s += s_part[w_];
    partials[bid] = s;
    __threadfence();
    s_last = (atomicAdd(&ws->counter_dense, 1u) == nblocks - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double tot = 0.0;
  for (int i = threadIdx.x; i < (int)nblocks; i += kDenseThreads) tot += partials[i];
  tot = warp_sum(tot);
  if (lane == 0) s_part[wid] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w_ = 0; w_ < kDenseThreads / 32; ++w_) s += s_part[w_];
    losses[0] = (float)((double)cfg.w_cls * s / (num_pos + (double)cfg.avg_extra));           // radet_head.py:256-259
    losses[1] = has_pos ? (float)((double)cfg.w_bbox * ws->norm[2] / sum_wq) : (float)ws->norm[4];   // :269-274 / :280
    losses[2] = has_pos ? (float)((double)cfg.w_iou * ws->norm[3] / num_pos) : (float)ws->norm[5];   // :275-278 / :281
    losses[3] = (float)num_pos;
    ws->counter_dense = 0u;
  }
}

// ---- TMA-pipelined variant (all planes 16-byte tileable: h*w % 4 == 0 on every level) ---------------------------
// One CTA = one tile of up to 512 consecutive points of one (image, level), one warp per 128 of them.  The C class
// planes of a warp's points stream through its private 8-stage shared-memory ring, filled by 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx; no registers tied up by loads); lanes read their float4 from the ring,
// compute, and store the gradient straight to global memory.  8 CTAs x 4 warps x 8 stages x 512 B = 128 KB of loads in
// flight per SM (HBM latency x bandwidth is ~45 KB per SM; the register-pipelined kernel tops out at 48 KB).
constexpr int kTilePts = 512;
constexpr int kTmaStages = 8;
struct TileTable {
  int tpl[RADET_MAX_LEVELS];        // tiles per (image, level)
  int toff[RADET_MAX_LEVELS + 1];   // tile offset of each level inside one image
};

constexpr int kTmaWarps = kTilePts / 128;             // one warp per 128 points of the tile (4 points per lane)
constexpr int kTmaThreads = kTmaWarps * 32;

template <bool kGamma2>
__global__ void __launch_bounds__(kTmaThreads, 8)
loss_dense_tma_kernel(GridDev grid, TileTable tt, int nch /* class chunks per tile */, int cc /* classes per chunk */, int B, int C,
                      MapsDev maps, GradsDev grads,
                      const int* __restrict__ gt_offsets, const int64_t* __restrict__ gt_labels,
                      const int64_t* __restrict__ pidx, const float* __restrict__ pw, radet_loss_cfg_t cfg,
                      const float* __restrict__ grad_scale, LossWs* __restrict__ ws, double* __restrict__ partials,
                      float* __restrict__ losses) {
  // Every warp runs its own pipeline over its 128 points: a private 8-stage ring (512 B per stage) with one mbarrier
  // per stage; lane 0 issues the bulk copy of plane c+8 as soon as the warp has consumed plane c.  Nothing couples the
  // warps of a CTA inside the plane loop (no block barrier, no producer warp that could fall behind).
  __shared__ __align__(128) float s_ring[kTmaWarps][kTmaStages][128];
  __shared__ __align__(8) uint64_t s_full[kTmaWarps][kTmaStages];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kTmaStages; ++k) mbar_init(&s_full[wid][k], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int P = grid.off[grid.num_levels];
  const int tiles_per_image = tt.toff[grid.num_levels];
  const int total_items = B * tiles_per_image * nch;   // (tile, class chunk); small batches split the classes to fill the SMs
  const double num_pos = ws->norm[0], sum_wq = ws->norm[1];
  const bool has_pos = ws->norm[6] > 0.0;                                        // radet_head.py:261
  const float gs_cls = grad_scale ? grad_scale[0] : 1.f, gs_box = grad_scale ? grad_scale[1] : 1.f,
              gs_iou = grad_scale ? grad_scale[2] : 1.f;
  const float k_cls = gs_cls * cfg.w_cls / (float)(num_pos + (double)cfg.avg_extra);
  const float k_box = has_pos ? gs_box * cfg.w_bbox / (float)sum_wq : gs_box;
  const float k_iou = has_pos ? gs_iou * cfg.w_iou / (float)num_pos : gs_iou;
  const bool want_grad = grads.cls[0] != nullptr;
  const float gamma = cfg.gamma, alpha = cfg.alpha;
  const float* ring = &s_ring[wid][0][0];
  const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(&s_full[wid][0]);   // shared-window addresses, computed once
  float lsum = 0.f;
  unsigned fills = 0;                                   // planes this warp has pushed through its ring so far
  for (int it = blockIdx.x; it < total_items; it += gridDim.x) {
    const int t = it / nch, ch = it - t * nch;
    const int c0 = ch * cc, cn = min(cc, C - c0);                    // planes c0 .. c0 + cn - 1
    const int b = t / tiles_per_image, r = t - b * tiles_per_image;
    int l = 0;
#pragma unroll
    for (int k = 1; k < RADET_MAX_LEVELS; ++k) l += (k < grid.num_levels && r >= tt.toff[k]) ? 1 : 0;
    const int hw = grid.h[l] * grid.w[l];
    const int q_warp = (r - tt.toff[l]) * kTilePts + wid * 128;      // first point of this warp
    const int npts = min(128, hw - q_warp);                          // multiple of 4; <= 0: nothing for this warp
    if (npts <= 0) continue;                                          // warp-uniform
    const uint32_t bytes = (uint32_t)npts * 4u;
    const float* nsrc = maps.cls[l] + ((int64_t)b * C + c0) * hw + q_warp;   // lane 0: next plane to request
    int issued = 0;
    if (lane == 0) {                                                  // prologue: the first planes
      const int n0 = min(kTmaStages, cn);
      for (; issued < n0; ++issued, nsrc += hw) {
        const unsigned st = (fills + issued) % kTmaStages;
        mbar_expect_tx_s(full_s + st * 8u, bytes);
        tma_bulk_g2s_s(ring_s + st * 512u, nsrc, bytes, full_s + st * 8u);
      }
    }
    // per-lane setup for its 4 points (index -> label, weight) while the ring fills
    const int q0 = q_warp + 4 * lane;
    const bool active = 4 * lane < npts;
    const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
    const int64_t pbase = (int64_t)b * P + grid.off[l] + q0;
    int idx[4], lab[4];
    float wt[4], wn[4];                                              // alpha w and (1 - alpha) w of the four points
    {
      longlong2 a01 = make_longlong2(-1, -1), a23 = make_longlong2(-1, -1);
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) {                                                   // P and every level size are multiples of 4: aligned vectors
        a01 = *reinterpret_cast<const longlong2*>(pidx + pbase);
        a23 = *reinterpret_cast<const longlong2*>(pidx + pbase + 2);
        wv = *reinterpret_cast<const float4*>(pw + pbase);
      }
      const int64_t v[4] = {a01.x, a01.y, a23.x, a23.y};
      const float w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        idx[i] = v[i] < 0 ? -1 : (int)(v[i] > (int64_t)G ? (int64_t)G : v[i]);
        lab[i] = (int)label_of(idx[i], G, gt_labels + g0, C) - c0;    // relative to the chunk
        wt[i] = alpha * w[i];
        wn[i] = (1.f - alpha) * w[i];
      }
    }
    float* dst = want_grad ? grads.cls[l] + ((int64_t)b * C + c0) * hw + q0 : nullptr;   // next plane to store
    const float* rp0 = ring + 4 * lane;
    for (int c = 0; c < cn; ++c, ++fills) {
      const unsigned st = fills % kTmaStages;
      mbar_wait_s(full_s + st * 8u, (fills / kTmaStages) & 1u);
      if (active) {
        const float4 xv = *reinterpret_cast<const float4*>(rp0 + st * 128);
        float4 gv;
        gv.x = focal_acc<kGamma2>(xv.x, lab[0] == c, wt[0], wn[0], k_cls, gamma, lsum);
        gv.y = focal_acc<kGamma2>(xv.y, lab[1] == c, wt[1], wn[1], k_cls, gamma, lsum);
        gv.z = focal_acc<kGamma2>(xv.z, lab[2] == c, wt[2], wn[2], k_cls, gamma, lsum);
        gv.w = focal_acc<kGamma2>(xv.w, lab[3] == c, wt[3], wn[3], k_cls, gamma, lsum);
        if (dst) {
          stg_stream4(dst, gv);
          dst += hw;
        }
      }
      // Refill the stage only AFTER the values have been consumed.  Issuing the copy right behind the LDS is not
      // enough: neither a barrier nor an mbarrier arrive waits for a shared-memory read still in flight, and under
      // load (other kernels' CTAs on the SM) the bulk copy was observed to overwrite the stage before the read.
      __syncwarp();
      if (lane == 0 && issued < cn) {
        mbar_expect_tx_s(full_s + st * 8u, bytes);
        tma_bulk_g2s_s(ring_s + st * 512u, nsrc, bytes, full_s + st * 8u);
        nsrc += hw;
        ++issued;
      }
    }
    // regression / IoU gradient planes: zero except at the positives parked by loss_pos_kernel (rescaled in place)
    if (want_grad && active && ch == 0) {
      const bool anypos = G > 0 && (idx[0] >= 0 || idx[1] >= 0 || idx[2] >= 0 || idx[3] >= 0);
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        float* o = kk < 4 ? grads.bbox[l] + ((int64_t)b * 4 + kk) * hw + q0 : grads.iou[l] + (int64_t)b * hw + q0;
        const float kn = kk < 4 ? k_box : k_iou, gs = kk < 4 ? gs_box : gs_iou;
        float gv[4] = {0.f, 0.f, 0.f, 0.f};
        if (anypos) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (idx[i] >= 0) gv[i] = has_pos ? kn * o[i] : gs;                  // radet_head.py:280-281 when num_pos == 0
        }
        stg_stream4(o, make_float4(gv[0], gv[1], gv[2], gv[3]));
      }
    }
  }
  // loss_cls = w_cls * sum / (num_pos + num_imgs): deterministic block partials, last block finalises
  __shared__ double s_part[kTmaThreads / 32];
  __shared__ bool s_last;
  const unsigned nblocks = gridDim.x, bid = blockIdx.x;
  const double v = warp_sum((double)lsum);
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (tid == 0) {
    double s_ = 0.0;
    for (int w_ = 0; w_ < kTmaThreads / 32; ++w_) s_ += s_part[w_];
    partials[bid] = s_;
    __threadfence();
    s_last = (atomicAdd(&ws->counter_dense, 1u) == nblocks - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double tot = 0.0;
  for (int i = tid; i < (int)nblocks; i += kTmaThreads) tot += partials[i];
  tot = warp_sum(tot);
  if (lane == 0) s_part[wid] = tot;
  __syncthreads();
  if (tid == 0) {
    double s_ = 0.0;
    for (int w_ = 0; w_ < kTmaThreads / 32; ++w_) s_ += s_part[w_];
    losses[0] = (float)((double)cfg.w_cls * s_ / (num_pos + (double)cfg.avg_extra));           // radet_head.py:256-259
    losses[1] = has_pos ? (float)((double)cfg.w_bbox * ws->norm[2] / sum_wq) : (float)ws->norm[4];   // :269-274 / :280
    losses[2] = has_pos ? (float)((double)cfg.w_iou * ws->norm[3] / num_pos) : (float)ws->norm[5];   // :275-278 / :281
    losses[3] = (float)num_pos;
    ws->counter_dense = 0u;
  }
}

// ---- warp-item variant of the dense pass (default where the TMA kernel applies) ---------------------------------------
// One CTA = ONE warp = one item: 256 consecutive points of one (image, level) x one chunk of classes.  A lane holds 8
// points (two float4 halves of the 256), so a stage of the ring (2 planes x 1 KB, six stages = 12 KB) feeds 16
// independent focal chains per lane, and every per-plane overhead (mbarrier wait, bulk-copy issue, pointer bumps) is
// paid once per 256 points instead of once per 128.  What makes it cheaper than loss_dense_tma_kernel per element:
//   * planes in which no lane of the warp holds its target class (all but ~8 % of them) take a path without the
//     target / non-target selects, in packed fp32x2 arithmetic (FMUL2 / FFMA2 / FADD2): ~13 instructions per element
//     instead of ~21;
//   * single-warp CTAs need no block barrier anywhere and balance over the SMs at a granularity of 256 points.
constexpr int kWStages = 6;
constexpr int kWPlanes = 2;
struct WTable {
  int gpl[RADET_MAX_LEVELS];        // 256-point groups per (image, level)
  int goff[RADET_MAX_LEVELS + 1];   // group offset of each level inside one image
};

// non-target elements only (z = x, coef = 1 - alpha), two at a time; wn = (1 - alpha) w, wnk = wn * k_cls
__device__ __forceinline__ float2 focal_fast2(float2 x, float2 wn, float2 wnk, float2& lsum) {
  const float2 t = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), make_float2(-1.4426950408889634f, -1.4426950408889634f));
  const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));                       // exp(-|x|)
  const float2 d = __fadd2_rn(e, make_float2(1.f, 1.f));
  const float2 inv = make_float2(rcp_approx(d.x), rcp_approx(d.y));                     // 1 / (1 + e)
  const float2 lg = make_float2(lg2_approx(inv.x), lg2_approx(inv.y));
  const float2 sp = __ffma2_rn(lg, make_float2(-0.6931471805599453f, -0.6931471805599453f),
                               make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));          // softplus(x)
  const float2 ei = __fmul2_rn(e, inv);
  const float2 s = make_float2(x.x >= 0.f ? inv.x : ei.x, x.y >= 0.f ? inv.y : ei.y);  // sigmoid(x)
  const float2 u = __fmul2_rn(s, s);
  lsum = __ffma2_rn(__fmul2_rn(wn, u), sp, lsum);
  const float2 h = __ffma2_rn(__ffma2_rn(s, make_float2(-2.f, -2.f), make_float2(2.f, 2.f)), sp, s);
  return __fmul2_rn(__fmul2_rn(wnk, u), h);
}

template <bool kGamma2, bool kHint, int kH /* float4 groups per lane and plane: an item is 128 * kH points */>
__global__ void __launch_bounds__(32)
loss_dense_w_kernel(GridDev grid, WTable tab, int nch, int cc, int B, int C, MapsDev maps, GradsDev grads,
                    const int* __restrict__ gt_offsets, const int64_t* __restrict__ gt_labels, const int64_t* __restrict__ pidx,
                    const float* __restrict__ pw, radet_loss_cfg_t cfg, const float* __restrict__ grad_scale,
                    LossWs* __restrict__ ws, double* __restrict__ partials, float* __restrict__ losses,
                    const double* __restrict__ num_pos_hint, PosList plist, unsigned long long* dbg) {
#define WDBG(k) do { if (dbg && threadIdx.x == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); dbg[(size_t)blockIdx.x * 8 + (k)] = t__; } } while (0)
  WDBG(0);
  constexpr int kWPts = 128 * kH;
  __shared__ __align__(128) float s_ring[kWStages][kWPlanes][kWPts];
  __shared__ __align__(8) uint64_t s_full[kWStages];
  const int lane = threadIdx.x;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kWStages; ++k) mbar_init(&s_full[k], 1);
    fence_barrier_init();
  }
  // The ring starts out zeroed: lanes beyond a short group read whatever their slot held, and with a zero weight that
  // contributes exactly 0 as long as it is finite.
  for (int i = lane; i < kWStages * kWPlanes * kWPts / 4; i += 32) reinterpret_cast<float4*>(&s_ring[0][0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t ring_s = smem_u32(&s_ring[0][0][0]), full_s = smem_u32(&s_full[0]);

  const int P = grid.off[grid.num_levels];
  const int gpi = tab.goff[grid.num_levels];
  const unsigned it = blockIdx.x;
  const int gg = (int)(it / (unsigned)nch), ch = (int)(it - (unsigned)gg * (unsigned)nch);
  const int c0 = ch * cc, cn = min(cc, C - c0);                        // planes c0 .. c0 + cn - 1 (cn <= 32)
  const int b = gg / gpi, r = gg - b * gpi;
  int l = 0;
#pragma unroll
  for (int k = 1; k < RADET_MAX_LEVELS; ++k) l += (k < grid.num_levels && r >= tab.goff[k]) ? 1 : 0;
  const int hw = grid.h[l] * grid.w[l];
  const int q_warp = (r - tab.goff[l]) * kWPts;
  const int npts = min(kWPts, hw - q_warp);                            // multiple of 4, > 0
  const uint32_t bytes = (uint32_t)npts * 4u;
  const int nst = (cn + kWPlanes - 1) / kWPlanes;
  const float* nsrc = maps.cls[l] + ((int64_t)b * C + c0) * hw + q_warp;   // lane 0: first plane of the next stage to request
  int issued = 0;
  auto issue_next = [&](unsigned sl_) {
    const int np = cn - issued * kWPlanes;
    const uint32_t bar = full_s + sl_ * 8u;
    const uint32_t sdst = ring_s + sl_ * (kWPlanes * kWPts * 4u);
    mbar_expect_tx_s(bar, bytes * (uint32_t)min(np, kWPlanes));
#pragma unroll
    for (int p_ = 0; p_ < kWPlanes; ++p_)
      if (p_ < np) tma_bulk_g2s_s(sdst + p_ * (kWPts * 4u), nsrc + (int64_t)p_ * hw, bytes, bar);
    nsrc += (int64_t)kWPlanes * hw;
    ++issued;
  };
  // Normalisers.  loss_pos_kernel precedes this kernel in the stream; launched as its programmatic dependent this kernel may
  // start while it is still running, and nothing it writes (ws->norm, the positives' gradient values) is read before
  // griddep_wait().  The class planes need num_pos only (the gradient scale k_cls):
  //   * kHint ("overlapped" order): num_pos is the fixed-order sum of the assignment's per-image weight sums (num_pos_hint),
  //     the wait sits BEHIND the class planes, loss_pos_kernel has zero-filled the regression / IoU planes and listed the
  //     positives' un-normalised gradients -- the items share the list and write the normalised values;
  //   * otherwise the wait sits behind the per-lane set-up (index / weight loads, labels, the ring's first bulk copies),
  //     and the items of class chunk 0 zero-fill the regression / IoU planes, rescaling the cells loss_pos_kernel parked.
  const float gs_cls = grad_scale ? grad_scale[0] : 1.f, gs_box = grad_scale ? grad_scale[1] : 1.f,
              gs_iou = grad_scale ? grad_scale[2] : 1.f;
  const bool want_grad = grads.cls[0] != nullptr;
  const float gamma = cfg.gamma, alpha = cfg.alpha;
  // per-lane setup: 8 points = half 0 at 4*lane, half 1 at 128 + 4*lane
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  const int64_t pb = (int64_t)b * P + grid.off[l] + q_warp;
  bool act[kH];
#pragma unroll
  for (int h_ = 0; h_ < kH; ++h_) act[h_] = 128 * h_ + 4 * lane < npts;
  int lab[4 * kH];
  float w8[4 * kH], wn[4 * kH], wnk[4 * kH];
  unsigned lmask = 0u, posmask = 0u;
  // The index / weight loads go out BEFORE the ring's first bulk copies: all warps of the grid start together, and 12 KB of
  // prefetch per warp in front of these 3 KB would delay every warp's set-up by the time the prefetch takes (measured:
  // 5.4 us from CTA start to the first plane with the copies first).
  longlong2 ia[kH], ib[kH];
  float4 wv2[kH];
#pragma unroll
  for (int h_ = 0; h_ < kH; ++h_) {
    ia[h_] = ib[h_] = make_longlong2(-1, -1);
    wv2[h_] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act[h_]) {
      const int64_t o = pb + 128 * h_ + 4 * lane;
      ia[h_] = *reinterpret_cast<const longlong2*>(pidx + o);
      ib[h_] = *reinterpret_cast<const longlong2*>(pidx + o + 2);
      wv2[h_] = *reinterpret_cast<const float4*>(pw + o);
    }
  }
  // the image's labels, one per lane (G <= 32): the index -> label step below is then a shuffle instead of a second
  // dependent round trip behind the index loads
  const int lab_lane = (G > 0 && G <= 32 && lane < G) ? (int)gt_labels[g0 + lane] : C;
  if (lane == 0) {
    const int n0 = min(kWStages, nst);
    for (int s_ = 0; s_ < n0; ++s_) issue_next((unsigned)s_);
  }
#pragma unroll
  for (int h_ = 0; h_ < kH; ++h_) {
    const int64_t v[4] = {ia[h_].x, ia[h_].y, ib[h_].x, ib[h_].y};
    const float w4[4] = {wv2[h_].x, wv2[h_].y, wv2[h_].z, wv2[h_].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = 4 * h_ + i;
      const int ix = v[i] < 0 ? -1 : (int)(v[i] > (int64_t)G ? (int64_t)G : v[i]);
      if (ix >= 0) posmask |= 1u << k;
      int lb;
      if (G <= 32) {                                                                            // warp-uniform
        int kq = ix - 1;                                                                        // label_of(): python indexing of gt_labels
        if (kq < 0) kq += G;
        if (kq >= G) kq = G - 1;
        lb = __shfl_sync(kFull, lab_lane, kq & 31);
        if (G <= 0 || ix < 0) lb = C;
      } else {
        lb = (int)label_of(ix, G, gt_labels + g0, C);
      }
      lab[k] = lb - c0;                                                                         // relative to the chunk
      if (lab[k] >= 0 && lab[k] < cn) lmask |= 1u << lab[k];
      w8[k] = w4[i];
      wn[k] = (1.f - alpha) * w4[i];
    }
  }
  if (!kHint) griddep_wait();           // a no-op unless this kernel was launched as a programmatic dependent
  const double num_pos = kHint ? warp_sum_array(num_pos_hint, (unsigned)B, lane) : ws->norm[0];
  const float k_cls = gs_cls * cfg.w_cls / (float)(num_pos + (double)cfg.avg_extra);
#pragma unroll
  for (int k = 0; k < 4 * kH; ++k) wnk[k] = wn[k] * k_cls;
  float* dst0 = want_grad ? grads.cls[l] + ((int64_t)b * C + c0) * hw + q_warp + 4 * lane : nullptr;   // next plane to store
  WDBG(1);
  float2 ls2 = make_float2(0.f, 0.f);
  float lsum = 0.f;
  unsigned slot = 0, par = 0;
  for (int s_ = 0; s_ < nst; ++s_) {
    mbar_wait_s(full_s + slot * 8u, par);
    if (s_ == 0) WDBG(2);
    if (s_ == 7) WDBG(3);
    const float* rp = &s_ring[slot][0][0] + 4 * lane;
#pragma unroll
    for (int p_ = 0; p_ < kWPlanes; ++p_) {
      const int c = s_ * kWPlanes + p_;
      if (c < cn) {                                                                            // warp-uniform
        float4 xv[kH], gv_[kH];
#pragma unroll
        for (int h_ = 0; h_ < kH; ++h_) xv[h_] = *reinterpret_cast<const float4*>(rp + p_ * kWPts + 128 * h_);
        if (kGamma2 && !__any_sync(kFull, (lmask >> c) & 1u)) {
          // no lane holds its target class in this plane: packed non-target path
#pragma unroll
          for (int h_ = 0; h_ < kH; ++h_) {
            const int k = 4 * h_;
            float2 t;
            t = focal_fast2(make_float2(xv[h_].x, xv[h_].y), make_float2(wn[k], wn[k + 1]), make_float2(wnk[k], wnk[k + 1]), ls2);
            gv_[h_].x = t.x; gv_[h_].y = t.y;
            t = focal_fast2(make_float2(xv[h_].z, xv[h_].w), make_float2(wn[k + 2], wn[k + 3]), make_float2(wnk[k + 2], wnk[k + 3]), ls2);
            gv_[h_].z = t.x; gv_[h_].w = t.y;
          }
        } else {
#pragma unroll
          for (int h_ = 0; h_ < kH; ++h_) {
            const int k = 4 * h_;
            gv_[h_].x = focal_acc<kGamma2>(xv[h_].x, lab[k] == c, alpha * w8[k], wn[k], k_cls, gamma, lsum);
            gv_[h_].y = focal_acc<kGamma2>(xv[h_].y, lab[k + 1] == c, alpha * w8[k + 1], wn[k + 1], k_cls, gamma, lsum);
            gv_[h_].z = focal_acc<kGamma2>(xv[h_].z, lab[k + 2] == c, alpha * w8[k + 2], wn[k + 2], k_cls, gamma, lsum);
            gv_[h_].w = focal_acc<kGamma2>(xv[h_].w, lab[k + 3] == c, alpha * w8[k + 3], wn[k + 3], k_cls, gamma, lsum);
          }
        }
        if (dst0) {
#pragma unroll
          for (int h_ = 0; h_ < kH; ++h_)
            if (act[h_]) stg_stream4(dst0 + 128 * h_, gv_[h_]);
          dst0 += hw;
        }
      }
    }
    // Refill the slot only AFTER its values have been consumed (the stores above depend on them): neither a barrier nor
    // an mbarrier arrive orders the bulk copy's write behind a shared-memory read that is still in flight.
    __syncwarp();
    if (lane == 0 && issued < nst) issue_next(slot);
    if (++slot == kWStages) {
      slot = 0;
      par ^= 1u;
    }
  }
  WDBG(4);
  if (kHint) griddep_wait();
  WDBG(7);
  const double sum_wq = ws->norm[1];
  const bool has_pos = ws->norm[6] > 0.0;                                                     // radet_head.py:261
  if (kHint) {
    // the positives' regression / IoU gradients: this item's share of loss_pos_kernel's list.  The first entry is loaded
    // speculatively, together with the list length and the normalisers (one round trip instead of two); whatever the
    // slot holds is ignored when it lies beyond the length.
    if (want_grad) {
      const int npos = (int)__ldcg(&ws->pad[0]);
      for (int i = (int)it * 32 + lane; ; i += (int)gridDim.x * 32) {
        const int il = min(i, B * P - 1);                   // the scratch has B * P slots (>= the list length)
        const int t = __ldcg(plist.plist + il);
        float v[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = __ldcg(plist.pvals + (int64_t)il * 5 + k);
        if (i >= npos) break;
        if (t >= 0) {
          const float k_box = has_pos ? gs_box * cfg.w_bbox / (float)sum_wq : gs_box;
          const float k_iou = has_pos ? gs_iou * cfg.w_iou / (float)num_pos : gs_iou;
          const int tb = t / P, tp = t - tb * P;
          const int tl = level_of(grid, tp);
          const int tq = tp - grid.off[tl];
          const int thw = grid.h[tl] * grid.w[tl];
          float* gb = grads.bbox[tl] + (int64_t)tb * 4 * thw + tq;
#pragma unroll
          for (int k = 0; k < 4; ++k) gb[(int64_t)k * thw] = has_pos ? k_box * v[k] : gs_box;   // radet_head.py:280 when num_pos == 0
          grads.iou[tl][(int64_t)tb * thw + tq] = has_pos ? k_iou * v[4] : gs_iou;               // :281
        }
        if (i + (int)gridDim.x * 32 >= npos) break;       // (keeps the next speculative slot inside the scratch)
      }
    }
  }
  // regression / IoU gradient planes: zero except at the positives parked by loss_pos_kernel (rescaled in place).  All the
  // parked values are loaded before the first store (the loads of one plane must not queue behind the stores of another).
  if (!kHint && want_grad && ch == 0) {
    const float k_box = has_pos ? gs_box * cfg.w_bbox / (float)sum_wq : gs_box;
    const float k_iou = has_pos ? gs_iou * cfg.w_iou / (float)num_pos : gs_iou;
    float gv[kH][5][4];
#pragma unroll
    for (int h_ = 0; h_ < kH; ++h_) {
      const int q0 = q_warp + 128 * h_ + 4 * lane;
      const unsigned pm = (act[h_] && G > 0) ? (posmask >> (4 * h_)) & 15u : 0u;
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        const float* o = kk < 4 ? grads.bbox[l] + ((int64_t)b * 4 + kk) * hw + q0 : grads.iou[l] + (int64_t)b * hw + q0;
        const float kn = kk < 4 ? k_box : k_iou, gs = kk < 4 ? gs_box : gs_iou;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          gv[h_][kk][i] = 0.f;
          if ((pm >> i) & 1u) gv[h_][kk][i] = has_pos ? kn * __ldcg(o + i) : gs;   // radet_head.py:280-281 when num_pos == 0
        }
      }
    }
#pragma unroll
    for (int h_ = 0; h_ < kH; ++h_) {
      if (!act[h_]) continue;
      const int q0 = q_warp + 128 * h_ + 4 * lane;
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        float* o = kk < 4 ? grads.bbox[l] + ((int64_t)b * 4 + kk) * hw + q0 : grads.iou[l] + (int64_t)b * hw + q0;
        stg_stream4(o, make_float4(gv[h_][kk][0], gv[h_][kk][1], gv[h_][kk][2], gv[h_][kk][3]));
      }
    }
  }
  WDBG(5);
  // loss_cls = w_cls * sum / (num_pos + num_imgs): one partial sum per item, the last CTA adds them in a fixed order
  const double v = warp_sum((double)lsum + ((double)ls2.x + (double)ls2.y));
  const unsigned nblocks = gridDim.x;
  unsigned last = 0;
  if (lane == 0) {
    partials[it] = v;
    __threadfence();
    last = (atomicAdd(&ws->counter_dense, 1u) == nblocks - 1) ? 1u : 0u;
  }
  last = __shfl_sync(kFull, last, 0);
  WDBG(6);
  if (!last) return;
  __threadfence();
  // The partial sums come in through the (now idle) ring: one bulk copy per 1536 of them instead of a chain of L2 round trips
  // (1632 items at the cfg5 shape were four dependent rounds of loads for one warp).  Fixed order: lane-strided inside a round,
  // xor tree over the lanes, rounds in sequence.
  asm volatile("fence.proxy.async;" ::: "memory");      // written through the generic proxy, read by the async proxy
  double tot = 0.0;
  constexpr unsigned kRound = kWStages * kWPlanes * kWPts / 2;
  for (unsigned base = 0; base < nblocks; base += kRound) {
    const unsigned cnt = min(kRound, nblocks - base);
    const uint32_t nb = (cnt * 8u + 15u) & ~15u;        // the workspace pads the slots to 256 bytes
    const uint32_t bar = full_s + slot * 8u;
    if (lane == 0) {
      mbar_expect_tx_s(bar, nb);
      tma_bulk_g2s_s(ring_s, partials + base, nb, bar);
    }
    mbar_wait_s(bar, par);
    const double* sp = reinterpret_cast<const double*>(&s_ring[0][0][0]);
    double a = 0.0;
    for (unsigned i = (unsigned)lane; i < cnt; i += 32u) a += sp[i];
    tot += warp_sum(a);                                 // every lane's reads are consumed before the next round's copy is issued
    if (++slot == kWStages) {
      slot = 0;
      par ^= 1u;
    }
  }
  if (lane == 0) {
    losses[0] = (float)((double)cfg.w_cls * tot / (num_pos + (double)cfg.avg_extra));           // radet_head.py:256-259
    losses[1] = has_pos ? (float)((double)cfg.w_bbox * ws->norm[2] / sum_wq) : (float)ws->norm[4];   // :269-274 / :280
    losses[2] = has_pos ? (float)((double)cfg.w_iou * ws->norm[3] / num_pos) : (float)ws->norm[5];   // :275-278 / :281
    losses[3] = (float)num_pos;
    ws->counter_dense = 0u;
    if (kHint) {
      ws->pad[0] = 0u;                  // every item has read the list length (their counter increments came after)
      // weight sums that do not belong to these weights (a caller's mistake) must not pass silently
      const double own = ws->norm[6];
      if (!(fabs(num_pos - own) <= 1e-5 * fmax(1.0, fabs(own)))) losses[0] = losses[1] = losses[2] = __int_as_float(0x7fc00000);
    }
  }
}

struct ScaleTable {
  float* ptr[3 * RADET_MAX_LEVELS];
  int64_t n[3 * RADET_MAX_LEVELS];
  int which[3 * RADET_MAX_LEVELS];
};

__global__ void scale_grads_kernel(ScaleTable tab, const float* __restrict__ upstream) {
  const float a = upstream[0], b = upstream[1], c = upstream[2];
  if (a == 1.f && b == 1.f && c == 1.f) return;  // the common case: loss.backward() on the plain sum
  float* g = tab.ptr[blockIdx.y];
  if (!g) return;
  const int64_t n = tab.n[blockIdx.y];
  const float k = upstream[tab.which[blockIdx.y]];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] *= k;
}

// ------------------------------------------------------------------------------------------------ standalone TBLR coder
// TBLRBBoxCoder.encode / decode on explicit prior lists (tblr_bbox_coder.py:71-114, 117-172), normalize_by_wh=True.
__global__ void tblr_encode_kernel(const float4* __restrict__ priors, const float4* __restrict__ gts, int64_t n, float nrm,
                                   float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = priors[i], g = gts[i];
  const float cx = __fdiv_rn(__fadd_rn(p.x, p.z), 2.f), cy = __fdiv_rn(__fadd_rn(p.y, p.w), 2.f);
  const float w = __fsub_rn(p.z, p.x), h = __fsub_rn(p.w, p.y);
  float4 o;
  o.x = __fdiv_rn(__fdiv_rn(__fsub_rn(cy, g.y), h), nrm);
  o.y = __fdiv_rn(__fdiv_rn(__fsub_rn(g.w, cy), h), nrm);
  o.z = __fdiv_rn(__fdiv_rn(__fsub_rn(cx, g.x), w), nrm);
  o.w = __fdiv_rn(__fdiv_rn(__fsub_rn(g.z, cx), w), nrm);
  out[i] = o;
}

__global__ void tblr_decode_kernel(const float4* __restrict__ priors, const float4* __restrict__ tblr, int64_t n, float nrm,
                                   int clip, float max_h, float max_w, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = priors[i], t = tblr[i];
  const float cx = __fdiv_rn(__fadd_rn(p.x, p.z), 2.f), cy = __fdiv_rn(__fadd_rn(p.y, p.w), 2.f);
  const float w = __fsub_rn(p.z, p.x), h = __fsub_rn(p.w, p.y);
  const float T = __fmul_rn(__fmul_rn(t.x, nrm), h), Bt = __fmul_rn(__fmul_rn(t.y, nrm), h);
  const float L = __fmul_rn(__fmul_rn(t.z, nrm), w), R = __fmul_rn(__fmul_rn(t.w, nrm), w);
  float4 o = make_float4(__fsub_rn(cx, L), __fsub_rn(cy, T), __fadd_rn(cx, R), __fadd_rn(cy, Bt));
  if (clip) {
    o.x = fminf(fmaxf(o.x, 0.f), max_w);
    o.y = fminf(fmaxf(o.y, 0.f), max_h);
    o.z = fminf(fmaxf(o.z, 0.f), max_w);
    o.w = fminf(fmaxf(o.w, 0.f), max_h);
  }
  out[i] = o;
}

// ------------------------------------------------------------------------------------------------ standalone losses
// mmcv 1.3.x sigmoid_focal_loss forward / backward arithmetic (source not in the reference tree):
//   p = 1/(1+exp(-x));  target class: -alpha (1-p)^g log(max(p,FLT_MIN));  other: -(1-alpha) p^g log(max(1-p,FLT_MIN))
__global__ void focal_loss_kernel(const float* __restrict__ pred, const int64_t* __restrict__ target, int64_t total, int C,
                                  float gamma, float alpha, float* __restrict__ loss, float* __restrict__ dpred) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t n = i / C;
  const int c = (int)(i - n * C);
  const float x = pred[i];
  const bool is_t = target[n] == (int64_t)c;
  const float p = 1.f / (1.f + expf(-x));
  const float kMin = 1.17549435e-38f;
  if (is_t) {
    const float lg = logf(fmaxf(p, kMin)), m = powf(1.f - p, gamma);
    if (loss) loss[i] = -alpha * m * lg;
    if (dpred) dpred[i] = -alpha * m * (1.f - p - gamma * p * lg);
  } else {
    const float lg = logf(fmaxf(1.f - p, kMin)), m = powf(p, gamma);
    if (loss) loss[i] = -(1.f - alpha) * m * lg;
    if (dpred) dpred[i] = -(1.f - alpha) * m * (gamma * (1.f - p) * lg - p);
  }
}

__global__ void giou_loss_kernel(const float4* __restrict__ pred, const float4* __restrict__ target, int64_t n, float eps,
                                 float* __restrict__ loss, float4* __restrict__ dpred) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pred[i], t = target[i];
  // box_terms takes (centre, T, B, L, R): with centre 0 and scale 1, x1 = -L, y1 = -T, x2 = R, y2 = B (negation is exact)
  if (dpred) {
    const BoxTerms bt = box_terms<true>(0.f, 0.f, 1.f, -p.y, p.w, -p.x, p.z, -t.y, t.w, -t.x, t.z, eps, eps);
    loss[i] = 1.f - bt.giou;
    // d(1-giou)/d(x1,y1,x2,y2) from d giou / d(T,B,L,R)
    dpred[i] = make_float4(bt.d[2], bt.d[0], -bt.d[3], -bt.d[1]);
  } else {
    const BoxTerms bt = box_terms<false>(0.f, 0.f, 1.f, -p.y, p.w, -p.x, p.z, -t.y, t.w, -t.x, t.z, eps, eps);
    loss[i] = 1.f - bt.giou;
  }
}

__global__ void bce_logits_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                  float* __restrict__ loss, float* __restrict__ dpred) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pred[i], z = target[i];
  if (loss) loss[i] = bce_logits(x, z);
  if (dpred) dpred[i] = sigmoidf_(x) - z;
}

}  // namespace radet

// ================================================================================================ C ABI
using namespace radet;

extern "C" int radet_get_targets(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const int32_t* gt_offsets,
                                 const float* gt_bboxes, const int64_t* gt_labels, const int64_t* points_to_gt_index,
                                 const float* points_weight, int64_t* labels, float* bbox_targets, float* weights,
                                 float* anchors, void* stream) {
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc != RADET_OK) return rc;
  if (batch == 0) return RADET_OK;
  if (batch < 0 || num_classes <= 0 || !gt_offsets || !points_to_gt_index || !points_weight || !labels || !bbox_targets || !weights)
    return RADET_E_BADARG;
  const int64_t n = (int64_t)batch * g.off[g.num_levels];
  get_targets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      g, batch, num_classes, gt_offsets, gt_bboxes, gt_labels, points_to_gt_index, points_weight, labels,
      reinterpret_cast<float4*>(bbox_targets), weights, reinterpret_cast<float4*>(anchors));
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_grid_priors(const radet_grid_t* grid, int32_t pad_h, int32_t pad_w, float* anchors, uint8_t* valid_flags,
                                 void* stream) {
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc != RADET_OK) return rc;
  if (!anchors && !valid_flags) return RADET_E_BADARG;
  if (valid_flags && (pad_h <= 0 || pad_w <= 0)) return RADET_E_BADARG;
  const int P = g.off[g.num_levels];
  grid_priors_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(g, pad_h, pad_w, reinterpret_cast<float4*>(anchors), valid_flags);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

constexpr int kSMs = 148;
constexpr int kDenseOcc = 3;   // resident CTAs per SM (launch bounds: 256 threads x <= 85 registers, no spills)

static int dense_plan(const GridDev& g, int B, int C, DenseTable* tab, int* cc, int* nj, int* blocks) {
  int64_t u = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int hw = g.h[l] * g.w[l];
    tab->uoff[l] = (int)u;
    tab->upl[l] = (hw + 3) / 4;
    u += (int64_t)B * tab->upl[l];
    if (u > (1ll << 30)) return RADET_E_BADARG;
  }
  for (int l = g.num_levels; l <= RADET_MAX_LEVELS; ++l) tab->uoff[l] = (int)u;
  for (int l = g.num_levels; l < RADET_MAX_LEVELS; ++l) tab->upl[l] = 1;
  // Work items = units x channel chunks.  When the units alone fill the machine, one chunk per unit (all C+5 channels
  // in one thread) amortises the per-point setup best and the grid is a multiple of the SM count cut into equal
  // slices (no partially filled last wave).  Small batches stay at one item per thread in ceil(items / 256) full
  // CTAs: the kernel is latency-bound there, and the fewest resident CTAs x time leaves the other SMs to whatever
  // else is in flight (neighbouring steps, the decode branch); the chunk is only split to keep a thread's serial
  // plane loop at <= kMaxChunk channels.
  const int CH = C + 5;
  constexpr int kMaxChunk = 48;
  const int64_t resident = (int64_t)kSMs * kDenseOcc * kDenseThreads;
  int best_c = CH;
  if (u < resident && CH > kMaxChunk) {
    const int n = (CH + kMaxChunk - 1) / kMaxChunk;
    best_c = ((CH + n - 1) / n + kDG - 1) / kDG * kDG;
  }
  *cc = best_c;
  *nj = (CH + best_c - 1) / best_c;
  const int64_t items = u * *nj;
  if (items <= resident) {
    *blocks = (int)((items + kDenseThreads - 1) / kDenseThreads);
    if (*blocks < 1) *blocks = 1;
  } else {
    *blocks = (int)(kSMs * kDenseOcc);
  }
  return RADET_OK;
}

// Tiles of the TMA-pipelined dense kernel; false when some level is not 16-byte tileable (h*w % 4 != 0) or the
// development override RADET_DENSE_IMPL=reg is set.
static bool tile_plan(const GridDev& g, int B, int C, const void* pidx, const void* pw, TileTable* tt, int* nch, int* cc, int* blocks) {
  if ((reinterpret_cast<uintptr_t>(pidx) | reinterpret_cast<uintptr_t>(pw)) & 15) return false;   // 128-bit index / weight loads
  int t = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int hw = g.h[l] * g.w[l];
    if (hw & 3) return false;
    tt->tpl[l] = (hw + kTilePts - 1) / kTilePts;
    tt->toff[l] = t;
    t += tt->tpl[l];
  }
  for (int l = g.num_levels; l <= RADET_MAX_LEVELS; ++l) tt->toff[l] = t;
  for (int l = g.num_levels; l < RADET_MAX_LEVELS; ++l) tt->tpl[l] = 0;
  const int64_t total = (int64_t)B * t;
  if (total * C > (1ll << 30)) return false;
  const int64_t slots = (int64_t)kSMs * 8;
  // Class chunks (RADET_DENSE_CTAS = target CTA count, development switch): a small batch has too few tiles to occupy
  // the SMs and every warp then walks all C planes one after the other; splitting the classes over `nch` CTAs per
  // tile shortens that chain (cfg2 alone: 22.3 -> 18.5 us at 3 chunks, 17.7 us at 5).  It is OFF by default: under the
  // step scheduler the kernel shares the SMs with other steps, and the extra CTAs (each re-reading the
  // tile's indices / weights) cost the neighbours more than the shorter chain gains (583 k -> 550 k / 524 k images/s).
  int64_t want = 0;
  if (const char* e = getenv("RADET_DENSE_CTAS")) {
    const long v = atol(e);
    if (v > 0) want = v;
  }
  int64_t n = total > 0 ? (want + total - 1) / total : 1;
  if (n > (C + 1) / 2) n = (C + 1) / 2;             // at least two planes per chunk
  if (n < 1) n = 1;
  *cc = (int)((C + n - 1) / n);
  *nch = (C + *cc - 1) / *cc;
  const int64_t items = total * *nch;
  *blocks = (int)(items < slots ? items : slots);
  return true;
}

// Items of the warp-item dense kernel; false when it does not apply (a plane size that is not a multiple of 4, unaligned
// index / weight arrays).
static bool w_plan(const GridDev& g, int B, int C, const void* pidx, const void* pw, int pts, WTable* tab, int* nch, int* cc, int64_t* items) {
  if ((reinterpret_cast<uintptr_t>(pidx) | reinterpret_cast<uintptr_t>(pw)) & 15) return false;   // 128-bit index / weight loads
  int t = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int hw = g.h[l] * g.w[l];
    if (hw & 3) return false;
    tab->gpl[l] = (hw + pts - 1) / pts;
    tab->goff[l] = t;
    t += tab->gpl[l];
  }
  for (int l = g.num_levels; l <= RADET_MAX_LEVELS; ++l) tab->goff[l] = t;
  for (int l = g.num_levels; l < RADET_MAX_LEVELS; ++l) tab->gpl[l] = 0;
  const int64_t groups = (int64_t)B * t;
  // class chunks: at most 32 classes per item (one bit per class in the lanes' target masks); more chunks only when the
  // groups alone would leave most SMs without an item
  int64_t n = (C + 31) / 32;
  if (groups * n < kSMs) {
    n = (kSMs + groups - 1) / groups;
    if (n > (C + 3) / 4) n = (C + 3) / 4;             // at least four planes per chunk
  }
  if (const char* e = getenv("RADET_DENSE_CHUNKS")) { // development override
    const long v = atol(e);
    if (v >= n && v <= (C + 3) / 4) n = v;
  }
  if (n < 1) n = 1;
  *cc = (int)((C + n - 1) / n);
  if (*cc > 32) *cc = 32;
  *nch = (C + *cc - 1) / *cc;
  *items = groups * *nch;
  return *items < (1ll << 31);
}
static int64_t w_items_bound(const GridDev& g, int B, int C) {      // partial-sum slots the workspace reserves
  int64_t t = 0;
  for (int l = 0; l < g.num_levels; ++l) t += (g.h[l] * g.w[l] + 127) / 128;
  const int64_t n = ((C + 31) / 32) > ((C + 3) / 4) ? (C + 31) / 32 : (C + 3) / 4;
  return (int64_t)B * t * n;
}

// workspace: LossWs | loss_pos partials | dense partials (two-launch path) | fused-kernel partials | positives list
static void loss_ws_layout(const GridDev& g, int batch, int num_classes, int dblk, size_t* off_pos, size_t* off_dense, size_t* off_fused,
                           size_t* off_list, size_t* total) {
  const int64_t n = (int64_t)batch * g.off[g.num_levels];
  const int64_t pos_blocks = (n + kPosThreads * kPosPerThread - 1) / (kPosThreads * kPosPerThread);
  int64_t dense_blocks = dblk > (int)(kSMs * 8) ? dblk : (int)(kSMs * 8);          // any two-launch dense kernel's partial-sum slots
  const int64_t wb = w_items_bound(g, batch, num_classes);
  if (wb > dense_blocks) dense_blocks = wb;
  size_t o = align_up(sizeof(LossWs), 256);
  *off_pos = o;
  o += align_up((size_t)pos_blocks * 6 * 8, 256);
  *off_dense = o;
  o += align_up((size_t)dense_blocks * 8, 256);
  *off_fused = o;
  o += fused_part_bytes(g, batch, num_classes);
  *off_list = o;                                   // overlapped order: flat index + five gradient values of every positive
  o += align_up((size_t)n * 4, 256) + align_up((size_t)n * 20, 256);
  *total = o;
}

extern "C" size_t radet_loss_workspace_bytes(const radet_grid_t* grid, int32_t batch, int32_t num_classes) {
  GridDev g;
  if (make_grid_dev(grid, &g) != RADET_OK || batch <= 0 || num_classes <= 0) return 0;
  DenseTable tab;
  int cc, nj, dblk;
  if (dense_plan(g, batch, num_classes, &tab, &cc, &nj, &dblk) != RADET_OK) return 0;
  size_t a, b, c, d, total;
  loss_ws_layout(g, batch, num_classes, dblk, &a, &b, &c, &d, &total);
  return total;
}

extern "C" int radet_loss_fwd_bwd(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_maps_t* maps,
                                  const int32_t* gt_offsets, const float* gt_bboxes, const int64_t* gt_labels,
                                  const int64_t* points_to_gt_index, const float* points_weight, const radet_loss_cfg_t* cfg,
                                  const float* grad_scale, const radet_grad_maps_t* grads, float* losses, int32_t phases,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc != RADET_OK) return rc;
  if (batch <= 0 || num_classes <= 0 || !maps || !gt_offsets || !points_to_gt_index || !points_weight || !cfg || !losses || !workspace)
    return RADET_E_BADARG;
  if (workspace_bytes < radet_loss_workspace_bytes(grid, batch, num_classes) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return RADET_E_WORKSPACE;
  MapsDev md;
  GradsDev gd;
  for (int l = 0; l < RADET_MAX_LEVELS; ++l) {
    const bool on = l < g.num_levels;
    md.cls[l] = on ? maps->cls[l] : nullptr;
    md.bbox[l] = on ? maps->bbox[l] : nullptr;
    md.iou[l] = on ? maps->iou[l] : nullptr;
    gd.cls[l] = (on && grads) ? grads->cls[l] : nullptr;
    gd.bbox[l] = (on && grads) ? grads->bbox[l] : nullptr;
    gd.iou[l] = (on && grads) ? grads->iou[l] : nullptr;
    if (on && (!md.cls[l] || !md.bbox[l] || !md.iou[l])) return RADET_E_BADARG;
    if (on && grads && (!gd.cls[l] || !gd.bbox[l] || !gd.iou[l])) return RADET_E_BADARG;
    // 128-bit plane accesses need 16-byte aligned bases when h*w % 4 == 0
    if (on && ((g.h[l] * g.w[l]) & 3) == 0) {
      if ((reinterpret_cast<uintptr_t>(md.cls[l]) & 15) || (grads && ((reinterpret_cast<uintptr_t>(gd.cls[l]) & 15) ||
          (reinterpret_cast<uintptr_t>(gd.bbox[l]) & 15) || (reinterpret_cast<uintptr_t>(gd.iou[l]) & 15))))
        return RADET_E_BADARG;
    }
  }
  DenseTable tab;
  int cc, nj, dblk;
  rc = dense_plan(g, batch, num_classes, &tab, &cc, &nj, &dblk);
  if (rc != RADET_OK) return rc;
  // RADET_LOSS_IMPL: "fused" = the experimental single-launch kernel (loss_fused.cu); "reg" = the register-pipelined dense
  // kernel even where the TMA one applies.  Default: loss_pos_kernel + loss_dense_tma_kernel (see DESIGN.md section 4 for the
  // measurements behind that choice).
  const char* impl = getenv("RADET_LOSS_IMPL");
  TileTable tt;
  int tma_blocks = 0, nch = 1, tcc = 1;
  const bool use_tma = tile_plan(g, batch, num_classes, points_to_gt_index, points_weight, &tt, &nch, &tcc, &tma_blocks);
  if (use_tma && tma_blocks > dblk) dblk = tma_blocks;       // partial-sum slots (workspace_bytes sizes for the max)
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  LossWs* ws = reinterpret_cast<LossWs*>(wsb);
  size_t off_pos, off_dense, off_fused, off_list, total_ws;
  loss_ws_layout(g, batch, num_classes, dblk, &off_pos, &off_dense, &off_fused, &off_list, &total_ws);
  const int64_t n = (int64_t)batch * g.off[g.num_levels];
  const int64_t pos_blocks = (n + kPosThreads * kPosPerThread - 1) / (kPosThreads * kPosPerThread);
  double* pos_part = reinterpret_cast<double*>(wsb + off_pos);
  double* dense_part = reinterpret_cast<double*>(wsb + off_dense);
  cudaStream_t st = (cudaStream_t)stream;
  const bool hybrid = impl && impl[0] == 'h';   // "hybrid": loss_pos_kernel, then the ticketed streaming kernel for the dense pass
  if (hybrid && phases == RADET_LOSS_PHASE_ALL) {
    loss_pos_kernel<<<(unsigned)pos_blocks, kPosThreads, 0, st>>>(g, batch, md, gt_offsets, gt_bboxes, points_to_gt_index,
                                                                  points_weight, *cfg, ws, pos_part, gd, PosList{}, static_cast<unsigned long long*>(g_debug_buf));
    RADET_LAUNCH_CHECK();
    rc = launch_loss_fused(g, batch, num_classes, md, gd, gt_offsets, gt_bboxes, gt_labels, points_to_gt_index, points_weight, *cfg,
                           grad_scale, ws, wsb + off_fused, false, true, losses, nullptr, st);
    if (rc != RADET_E_UNSUPPORTED) return rc;
    phases = RADET_LOSS_PHASE_DENSE;
  }
  if (impl && impl[0] == 'f') {
    // single launch: normalisers, positive terms, dense pass and the three losses (or only one of the two phases).
    // Shapes it does not take (a plane size that is not a multiple of 4, unaligned index / weight arrays) fall through.
    const bool with_p1 = (phases & RADET_LOSS_PHASE_NORMALIZERS) != 0, with_items = (phases & RADET_LOSS_PHASE_DENSE) != 0;
    rc = launch_loss_fused(g, batch, num_classes, md, gd, gt_offsets, gt_bboxes, gt_labels, points_to_gt_index, points_weight, *cfg,
                           grad_scale, ws, wsb + off_fused, with_p1, with_items, losses, cfg->weight_sums, st);
    if (rc != RADET_E_UNSUPPORTED) return rc;
  }
  // ---- "overlapped" order (the default whenever the caller hands over the assignment's per-image weight sums): the class
  // planes need no result of loss_pos_kernel then.  loss_pos_kernel is launched first and the dense kernel as its
  // programmatic dependent: the items stream their class planes NEXT to it and wait for its completion only before the
  // sparse regression / IoU gradients and the final sums.  loss_pos_kernel zero-fills those planes and lists the positives.
  WTable wtab;
  int wnch = 1, wcc = 1;
  int64_t witems = 0;
  static const bool pdl_on = !(getenv("RADET_LOSS_PDL") && getenv("RADET_LOSS_PDL")[0] == '0');
  // Item size: 256 points (8 per lane) halves every per-plane overhead; 128 points doubles the warps per SM sub-partition
  // (the streaming loop is latency-bound per warp).  RADET_DENSE_PTS = 128 | 256 overrides the choice.
  int wpts = 256;
  {
    int64_t g256 = 0;
    for (int l = 0; l < g.num_levels; ++l) g256 += (g.h[l] * g.w[l] + 255) / 256;
    if (g256 * batch <= 8 * (int64_t)kSMs) wpts = 128;     // cfg2 / cfg3: 224 items of 256 points; cfg5: 1632
  }
  if (const char* e = getenv("RADET_DENSE_PTS")) wpts = atoi(e) == 128 ? 128 : 256;
  const bool w_ok = !(impl && (impl[0] == 'r' || impl[0] == 't' || impl[0] == 'f')) &&
                    w_plan(g, batch, num_classes, points_to_gt_index, points_weight, wpts, &wtab, &wnch, &wcc, &witems);
  // Kernels that are meant to share an SM must agree on its shared-memory carve-out: an SM does not change the split while
  // CTAs are resident, so items of the dense kernel (12 KB each) could not join loss_pos CTAs running under a small carve-out.
  static const bool carve_set = [] {
    cudaFuncSetAttribute(loss_pos_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    for (auto k : {loss_dense_w_kernel<true, true, 2>, loss_dense_w_kernel<false, true, 2>, loss_dense_w_kernel<true, false, 2>,
                   loss_dense_w_kernel<false, false, 2>, loss_dense_w_kernel<true, true, 1>, loss_dense_w_kernel<false, true, 1>,
                   loss_dense_w_kernel<true, false, 1>, loss_dense_w_kernel<false, false, 1>})
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  (void)carve_set;
  if (w_ok && pdl_on && phases == RADET_LOSS_PHASE_ALL && cfg->weight_sums && n < (1ll << 31)) {
    PosList pl{reinterpret_cast<int*>(wsb + off_list), reinterpret_cast<float*>(wsb + off_list + align_up((size_t)n * 4, 256))};
    loss_pos_kernel<<<(unsigned)pos_blocks, kPosThreads, 0, st>>>(g, batch, md, gt_offsets, gt_bboxes, points_to_gt_index,
                                                                  points_weight, *cfg, ws, pos_part, gd, pl,
                                                                  static_cast<unsigned long long*>(g_debug_buf));
    RADET_LAUNCH_CHECK();
    auto kern = cfg->gamma == 2.0f ? (wpts == 128 ? loss_dense_w_kernel<true, true, 1> : loss_dense_w_kernel<true, true, 2>)
                                   : (wpts == 128 ? loss_dense_w_kernel<false, true, 1> : loss_dense_w_kernel<false, true, 2>);
    cudaError_t le = launch_after_trigger(kern, dim3((unsigned)witems), dim3(32), 0, st, true, g, wtab, wnch, wcc, batch, num_classes,
                                          md, gd, gt_offsets, gt_labels, points_to_gt_index, points_weight, *cfg, grad_scale, ws, dense_part,
                                          losses, cfg->weight_sums, pl, static_cast<unsigned long long*>(g_debug_buf));
    if (le != cudaSuccess) return (int)le;
    RADET_LAUNCH_CHECK();
    return RADET_OK;
  }
  if (phases & RADET_LOSS_PHASE_NORMALIZERS) {
    loss_pos_kernel<<<(unsigned)pos_blocks, kPosThreads, 0, st>>>(g, batch, md, gt_offsets, gt_bboxes, points_to_gt_index,
                                                                  points_weight, *cfg, ws, pos_part, gd, PosList{}, static_cast<unsigned long long*>(g_debug_buf));
    RADET_LAUNCH_CHECK();
  }
  if (!(phases & RADET_LOSS_PHASE_DENSE)) return RADET_OK;
  if (!(impl && (impl[0] == 'r' || impl[0] == 't'))) {          // default: the warp-item kernel ("tma" / "reg": the older two)
    // classic order: loss_pos_kernel ran above.  Small grids launch as its programmatic dependent (prologue next to it);
    // large ones do not -- items that trickle in while loss_pos CTAs still hold registers are placed unevenly over the SMs,
    // and the imbalance costs more than the overlap gains (cfg5: 40.4 -> 42.5 us; cfg2: 27.8 -> 25.1 us).
    if (w_ok) {
      const bool prog = (phases == RADET_LOSS_PHASE_ALL) && pdl_on && witems <= 4 * (int64_t)kSMs;
      auto kern = cfg->gamma == 2.0f ? (wpts == 128 ? loss_dense_w_kernel<true, false, 1> : loss_dense_w_kernel<true, false, 2>)
                                     : (wpts == 128 ? loss_dense_w_kernel<false, false, 1> : loss_dense_w_kernel<false, false, 2>);
      cudaError_t le = launch_after_trigger(kern, dim3((unsigned)witems), dim3(32), 0, st, prog, g, wtab, wnch, wcc, batch, num_classes,
                                            md, gd, gt_offsets, gt_labels, points_to_gt_index, points_weight, *cfg, grad_scale, ws, dense_part,
                                            losses, (const double*)nullptr, PosList{}, static_cast<unsigned long long*>(g_debug_buf));
      if (le != cudaSuccess) return (int)le;
      RADET_LAUNCH_CHECK();
      return RADET_OK;
    }
  }
  if (use_tma && !(impl && impl[0] == 'r')) {
    if (cfg->gamma == 2.0f)
      loss_dense_tma_kernel<true><<<(unsigned)tma_blocks, kTmaThreads, 0, st>>>(g, tt, nch, tcc, batch, num_classes, md, gd, gt_offsets, gt_labels,
                                                                                  points_to_gt_index, points_weight, *cfg, grad_scale,
                                                                                  ws, dense_part, losses);
    else
      loss_dense_tma_kernel<false><<<(unsigned)tma_blocks, kTmaThreads, 0, st>>>(g, tt, nch, tcc, batch, num_classes, md, gd, gt_offsets, gt_labels,
                                                                                   points_to_gt_index, points_weight, *cfg, grad_scale,
                                                                                   ws, dense_part, losses);
    RADET_LAUNCH_CHECK();
    return RADET_OK;
  }
  const unsigned dblocks = (unsigned)dblk;
  if (cfg->gamma == 2.0f)
    loss_dense_kernel<true><<<dblocks, kDenseThreads, 0, st>>>(g, tab, batch, num_classes, cc, nj, md, gd, gt_offsets, gt_bboxes,
                                                               gt_labels, points_to_gt_index, points_weight, *cfg, grad_scale,
                                                               ws, dense_part, losses);
  else
    loss_dense_kernel<false><<<dblocks, kDenseThreads, 0, st>>>(g, tab, batch, num_classes, cc, nj, md, gd, gt_offsets, gt_bboxes,
                                                                gt_labels, points_to_gt_index, points_weight, *cfg, grad_scale,
                                                                ws, dense_part, losses);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_scale_grads(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_grad_maps_t* grads,
                                 const float* upstream, void* stream) {
  GridDev g;
  int rc = make_grid_dev(grid, &g);
  if (rc != RADET_OK) return rc;
  if (batch <= 0 || num_classes <= 0 || !grads || !upstream) return RADET_E_BADARG;
  ScaleTable tab{};
  int64_t nmax = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int64_t hw = (int64_t)g.h[l] * g.w[l];
    float* ptr[3] = {grads->cls[l], grads->bbox[l], grads->iou[l]};
    const int64_t n[3] = {batch * hw * num_classes, batch * hw * 4, batch * hw};
    for (int k = 0; k < 3; ++k) {
      if (!ptr[k]) return RADET_E_BADARG;
      tab.ptr[3 * l + k] = ptr[k];
      tab.n[3 * l + k] = n[k];
      tab.which[3 * l + k] = k;
      nmax = n[k] > nmax ? n[k] : nmax;
    }
  }
  const unsigned bx = (unsigned)((nmax + 255) / 256 > 16 ? 16 : (nmax + 255) / 256);   // the usual call exits at once: keep the grid small
  scale_grads_kernel<<<dim3(bx, 3 * g.num_levels), 256, 0, (cudaStream_t)stream>>>(tab, upstream);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_tblr_encode(const float* priors, const float* gt_bboxes, int64_t n, float normalizer, float* out,
                                 void* stream) {
  if (n == 0) return RADET_OK;
  if (n < 0 || !priors || !gt_bboxes || !out || !(normalizer > 0.f)) return RADET_E_BADARG;
  tblr_encode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(priors), reinterpret_cast<const float4*>(gt_bboxes), n, normalizer,
      reinterpret_cast<float4*>(out));
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_tblr_decode(const float* priors, const float* tblr, int64_t n, float normalizer, int32_t clip,
                                 float max_h, float max_w, float* out, void* stream) {
  if (n == 0) return RADET_OK;
  if (n < 0 || !priors || !tblr || !out || !(normalizer > 0.f)) return RADET_E_BADARG;
  tblr_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(priors), reinterpret_cast<const float4*>(tblr), n, normalizer, clip, max_h, max_w,
      reinterpret_cast<float4*>(out));
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_sigmoid_focal_loss(const float* pred, const int64_t* target, int64_t n, int32_t num_classes, float gamma,
                                        float alpha, float* loss, float* dloss_dpred, void* stream) {
  if (n == 0) return RADET_OK;
  if (n < 0 || num_classes <= 0 || !pred || !target || (!loss && !dloss_dpred)) return RADET_E_BADARG;
  const int64_t total = n * num_classes;
  focal_loss_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, target, total, num_classes, gamma, alpha,
                                                                                      loss, dloss_dpred);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_giou_loss(const float* pred, const float* target, int64_t n, float eps, float* loss, float* dloss_dpred,
                               void* stream) {
  if (n == 0) return RADET_OK;
  if (n < 0 || !pred || !target || !loss) return RADET_E_BADARG;
  if ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(dloss_dpred)) & 15)
    return RADET_E_BADARG;
  giou_loss_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(pred), reinterpret_cast<const float4*>(target), n, eps, loss,
      reinterpret_cast<float4*>(dloss_dpred));
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_bce_with_logits(const float* pred, const float* target, int64_t n, float* loss, float* dloss_dpred,
                                     void* stream) {
  if (n == 0) return RADET_OK;
  if (n < 0 || !pred || !target || (!loss && !dloss_dpred)) return RADET_E_BADARG;
  bce_logits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred, target, n, loss, dloss_dpred);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}
