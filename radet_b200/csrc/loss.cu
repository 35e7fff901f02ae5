// Target encode + fused head loss (forward and backward) on sm_100a.
//
// Reference semantics: RADetHead.get_targets / loss (radet/models/dense_heads/radet_head.py:173-392) and the loss
// modules it calls (focal_loss.py, iou_loss.py, cross_entropy_loss.py, tblr_bbox_coder.py, iou2d_calculator.py).
// The reference flattens NCHW -> (point, channel) with permute+reshape+cat, materialises labels / bbox targets /
// anchors and runs ~80 eager kernels forward plus autograd backward.  Here:
//
//   loss_pos_kernel    one pass over (image, point): the sparse positive terms (IoU target, GIoU, BCE) and the
//                      normalisers num_pos = sum w, sum wq.  Deterministic two-level reduction (block partials in
//                      fixed order, last block finalises).  ~12 B/point of traffic.
//   loss_dense_kernel  the HBM-bound pass: reads every logit once IN PLACE (NCHW planes, 128-bit streaming loads,
//                      4 consecutive points per thread), computes sigmoid-focal loss and its gradient with the final
//                      normaliser already applied, writes the gradient once (128-bit stores); the threads of class
//                      chunk 0 also emit the bbox / iou gradients of their 4 points.  Labels and TBLR targets are
//                      rebuilt on the fly from points_to_gt_index (no label / target / anchor tensors).
//
// Algorithmic traffic: (8C + 52) B/point  (C logits read + C grads written + 16 B bbox + 4 B iou read, 20 B grads
// written, 8 B index + 4 B weight read).
#include "common.cuh"

namespace radet {

struct MapsDev {
  const float* cls[RADET_MAX_LEVELS];
  const float* bbox[RADET_MAX_LEVELS];
  const float* iou[RADET_MAX_LEVELS];
};
struct GradsDev {
  float* cls[RADET_MAX_LEVELS];
  float* bbox[RADET_MAX_LEVELS];
  float* iou[RADET_MAX_LEVELS];
};

// ------------------------------------------------------------------------------------------------ targets
// radet_head.py:373-392 + tblr_bbox_coder.py:71-114.  label: C for idx<0; gt_labels[idx-1] with python negative
// indexing for idx==0 (last GT).  target: ((d / (scale*stride)) / 0.125), order T,B,L,R.
__device__ __forceinline__ int64_t label_of(int64_t idx, int G, const int64_t* __restrict__ gt_labels, int C) {
  if (G <= 0 || idx < 0) return C;
  int64_t k = idx - 1;
  if (k < 0) k += G;
  if (k >= G) k = G - 1;
  return gt_labels[k];
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

// ------------------------------------------------------------------------------------------------ box terms
struct BoxTerms {
  float iou, giou;
  float d[4];  // d giou / d (T, B, L, R)
};

// split of torch.max / torch.min gradients at ties (0.5 each), clamp(min=0) passes the gradient at 0
__device__ __forceinline__ float sel_gt(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

template <bool kGrad>
__device__ __forceinline__ BoxTerms box_terms(float cx, float cy, float s, float T, float Bt, float L, float R,
                                              float tT, float tB, float tL, float tR, float eps_iou, float eps_giou) {
  // decode (tblr_bbox_coder.py:154-166): loc = (v*normalizer)*side = v*s with s = normalizer*side (= stride for
  // the shipped 1/8 x 8*stride, where both scalings are exact); no clamp in training
  const float px1 = cx - L * s, py1 = cy - T * s, px2 = cx + R * s, py2 = cy + Bt * s;
  const float tx1 = cx - tL, ty1 = cy - tT, tx2 = cx + tR, ty2 = cy + tB;
  const float wp = px2 - px1, hp = py2 - py1;
  const float area_p = wp * hp, area_t = (tx2 - tx1) * (ty2 - ty1);
  const float ltx = fmaxf(px1, tx1), lty = fmaxf(py1, ty1), rbx = fminf(px2, tx2), rby = fminf(py2, ty2);
  const float iw_raw = rbx - ltx, ih_raw = rby - lty;
  const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
  const float ov = iw * ih;
  const float union_raw = area_p + area_t - ov;
  BoxTerms o;
  // IoU target: bbox_overlaps(..., eps=1e-6) (radet_head.py:267)
  o.iou = ov / fmaxf(union_raw, eps_iou);
  // GIoU: bbox_overlaps(mode='giou', eps=GIoULoss.eps) (iou_loss.py:96)
  const float uni = fmaxf(union_raw, eps_giou);
  const float iou_g = ov / uni;
  const float elx = fminf(px1, tx1), ely = fminf(py1, ty1), erx = fmaxf(px2, tx2), ery = fmaxf(py2, ty2);
  const float ew_raw = erx - elx, eh_raw = ery - ely;
  const float ew = fmaxf(ew_raw, 0.f), eh = fmaxf(eh_raw, 0.f);
  const float ea_raw = ew * eh;
  const float ea = fmaxf(ea_raw, eps_giou);
  o.giou = iou_g - (ea - uni) / ea;
  if (kGrad) {
    const float pw_ = iw_raw >= 0.f ? 1.f : 0.f, ph_ = ih_raw >= 0.f ? 1.f : 0.f;
    // d ov / d (px1, py1, px2, py2)
    const float dov[4] = {-pw_ * sel_gt(px1, tx1) * ih, -ph_ * sel_gt(py1, ty1) * iw, pw_ * sel_gt(tx2, px2) * ih,
                          ph_ * sel_gt(ty2, py2) * iw};
    const float dap[4] = {-hp, -wp, hp, wp};
    const float up = sel_gt(union_raw, eps_giou);
    const float pew = ew_raw >= 0.f ? 1.f : 0.f, peh = eh_raw >= 0.f ? 1.f : 0.f;
    const float ep = sel_gt(ea_raw, eps_giou);
    const float dea[4] = {-ep * pew * sel_gt(tx1, px1) * eh, -ep * peh * sel_gt(ty1, py1) * ew,
                          ep * pew * sel_gt(px2, tx2) * eh, ep * peh * sel_gt(py2, ty2) * ew};
    const float inv_u = 1.f / uni, inv_e = 1.f / ea;
    const float u_over_e = uni * inv_e;
    float dz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float dun = up * (dap[k] - dov[k]);
      const float diou = (dov[k] - iou_g * dun) * inv_u;
      dz[k] = diou + (dun - u_over_e * dea[k]) * inv_e;
    }
    // chain to (T,B,L,R): py1 = cy - T s, py2 = cy + B s, px1 = cx - L s, px2 = cx + R s
    o.d[0] = -s * dz[1];
    o.d[1] = s * dz[3];
    o.d[2] = -s * dz[0];
    o.d[3] = s * dz[2];
  }
  return o;
}

__device__ __forceinline__ void point_target(int64_t idx, int G, const float* __restrict__ gtb, float cx, float cy,
                                             float& tT, float& tB, float& tL, float& tR) {
  // decoded target distances in pixels: encode/decode scalings are exact powers of two, so decode(encode(d)) = d
  tT = tB = tL = tR = 0.f;
  if (idx > 0) {
    const int k = (int)((idx - 1) < (int64_t)(G - 1) ? (idx - 1) : (int64_t)(G - 1));
    const float4 gb = *reinterpret_cast<const float4*>(gtb + 4 * (int64_t)k);
    tT = cy - gb.y;
    tB = gb.w - cy;
    tL = cx - gb.x;
    tR = gb.z - cx;
  }
}

__device__ __forceinline__ float bce_logits(float x, float z) {  // torch: max(x,0) - x*z + log1p(exp(-|x|))
  return fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// workspace layout (doubles): [0..7] final sums / normalisers, then block partials
constexpr int kPosThreads = 256;
constexpr int kNormSlots = 8;   // S0 num_pos, S1 sum wq, S2 sum wq(1-giou), S3 sum w*bce, S4 sum pred, S5 sum iou logit
struct LossWs {
  double norm[kNormSlots];
  unsigned int counter_pos, counter_dense;
  unsigned int pad[2];
};

__global__ void __launch_bounds__(kPosThreads)
loss_pos_kernel(GridDev grid, int B, MapsDev maps, const int* __restrict__ gt_offsets,
                const float* __restrict__ gt_bboxes, const int64_t* __restrict__ pidx, const float* __restrict__ pw,
                radet_loss_cfg_t cfg, LossWs* __restrict__ ws, double* __restrict__ partials) {
  const int P = grid.off[grid.num_levels];
  const int64_t t = (int64_t)blockIdx.x * kPosThreads + threadIdx.x;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (t < (int64_t)B * P) {
    const int b = (int)(t / P), p = (int)(t - (int64_t)b * P);
    const int64_t idx = pidx[t];
    const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
    if (idx >= 0 && G > 0) {  // pos_inds: 0 <= label < C, includes ignored points (radet_head.py:245-247)
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
      const BoxTerms bt = box_terms<false>(cx, cy, s, T, Bt, L, R, tT, tB, tL, tR, 1e-6f, cfg.eps);
      const float wq = fmaxf(bt.iou, 1e-12f) * w;           // radet_head.py:272
      acc[0] = w;
      acc[1] = wq;
      acc[2] = wq * (1.f - bt.giou);
      acc[3] = w * bce_logits(xi, bt.iou);
      acc[4] = (T + Bt) + (L + R);
      acc[5] = xi;
    }
  }
  __shared__ double s_part[kPosThreads / 32][6];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double v = warp_sum((double)acc[k]);
    if (lane == 0) s_part[wid][k] = v;
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
  if (!s_last) return;
  __threadfence();
  // last block: fixed-order reduction over block partials (deterministic)
  double tot[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kPosThreads) {
#pragma unroll
    for (int k = 0; k < 6; ++k) tot[k] += partials[(int64_t)i * 6 + k];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double v = warp_sum(tot[k]);
    if (lane == 0) s_part[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S[6];
    for (int k = 0; k < 6; ++k) {
      S[k] = 0.0;
      for (int w = 0; w < kPosThreads / 32; ++w) S[k] += s_part[w][k];
      ws->norm[k] = S[k];
    }
    // norm[0], norm[1] are the NORMALISERS (a caller running the opt-in FCOS-style reduce_mean all-reduces them
    // between the two passes); norm[6], norm[7] keep the rank-local values (radet_head.py:254 is rank-local)
    ws->norm[6] = S[0];
    ws->norm[7] = S[1];
    ws->counter_pos = 0u;  // re-arm for the next call on this workspace
  }
}

// ------------------------------------------------------------------------------------------------ dense pass
constexpr int kDenseThreads = 256;

struct DenseTable {
  int uoff[RADET_MAX_LEVELS + 1];  // unit (4-point group) offsets per level over the whole batch
  int upl[RADET_MAX_LEVELS];       // units per (image, level) plane
};

// Sigmoid focal loss and its derivative for one logit (mmcv sigmoid_focal_loss semantics; restated from
// focal_loss.py:10-41 because the mmcv op is not in the reference tree).  Tolerance parity (not bit parity), so the
// transcendental part is kept short: one ex2, one rcp and a degree-8 polynomial
//     log1p(e) = e * P(e),  P = near-minimax fit of log1p(t)/t on [0,1]  (relative error 2.5e-7 in fp32)
// -> ~35 instructions per element, which keeps the kernel HBM-bound instead of ALU-bound at large batch.
template <bool kGamma2>
__device__ __forceinline__ void focal_elem(float x, bool is_t, float gamma, float alpha, float& loss, float& grad) {
  const float e = __expf(-fabsf(x));        // in (0, 1]
  float P = 0.00512610236f;
  P = fmaf(P, e, -0.0290740654f);
  P = fmaf(P, e, 0.0775160864f);
  P = fmaf(P, e, -0.136022478f);
  P = fmaf(P, e, 0.190768808f);
  P = fmaf(P, e, -0.248353988f);
  P = fmaf(P, e, 0.333181202f);
  P = fmaf(P, e, -0.499994457f);
  P = fmaf(P, e, 0.99999994f);
  const float l1p = e * P;                  // log1p(exp(-|x|))
  const float sp_x = fmaxf(x, 0.f) + l1p;   // softplus(x)  = -log(1-p)
  const float sp_nx = sp_x - x;             // softplus(-x) = -log(p)
  const float inv = __frcp_rn(1.0f + e);
  const float p = x >= 0.f ? inv : e * inv;
  const float q = x >= 0.f ? e * inv : inv;
  if (is_t) {
    const float m = kGamma2 ? q * q : __expf(-gamma * sp_x);      // (1-p)^gamma
    loss = alpha * m * sp_nx;
    grad = -alpha * m * (q + gamma * p * sp_nx);
  } else {
    const float m = kGamma2 ? p * p : __expf(-gamma * sp_nx);     // p^gamma
    loss = (1.f - alpha) * m * sp_x;
    grad = (1.f - alpha) * m * (p + gamma * q * sp_x);
  }
}

template <bool kGamma2>
__global__ void __launch_bounds__(kDenseThreads)
loss_dense_kernel(GridDev grid, DenseTable tab, int B, int C, int cc /* classes per thread */, int nj, MapsDev maps,
                  GradsDev grads, const int* __restrict__ gt_offsets, const float* __restrict__ gt_bboxes,
                  const int64_t* __restrict__ gt_labels, const int64_t* __restrict__ pidx, const float* __restrict__ pw,
                  radet_loss_cfg_t cfg, const float* __restrict__ grad_scale, LossWs* __restrict__ ws,
                  double* __restrict__ partials, float* __restrict__ losses) {
  const int P = grid.off[grid.num_levels];
  const int U = tab.uoff[grid.num_levels];
  const int64_t t = (int64_t)blockIdx.x * kDenseThreads + threadIdx.x;
  const double num_pos = ws->norm[0], sum_wq = ws->norm[1];
  const bool has_pos = ws->norm[6] > 0.0;                                        // radet_head.py:261
  const float gs_cls = grad_scale ? grad_scale[0] : 1.f, gs_box = grad_scale ? grad_scale[1] : 1.f,
              gs_iou = grad_scale ? grad_scale[2] : 1.f;
  const float k_cls = gs_cls * cfg.w_cls / (float)(num_pos + (double)cfg.avg_extra);
  const float k_box = has_pos ? gs_box * cfg.w_bbox / (float)sum_wq : gs_box;
  const float k_iou = has_pos ? gs_iou * cfg.w_iou / (float)num_pos : gs_iou;
  const bool want_grad = grads.cls[0] != nullptr;
  float lsum = 0.f;
  if (t < (int64_t)U * nj) {
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
    const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
    const int64_t pbase = (int64_t)b * P + grid.off[l] + q0;
    const int c0 = j * cc, c1 = min(C, c0 + cc);
    const float* cp = maps.cls[l] + ((int64_t)b * C) * hw + q0;
    float* gp = want_grad ? grads.cls[l] + ((int64_t)b * C) * hw + q0 : nullptr;
    // logits of the first class group go in flight BEFORE the index -> label dependency chain is resolved;
    // afterwards the next group is prefetched while the current one is being computed (register double buffer)
    constexpr int DG = 4;
    float nx[DG][4];
    auto load_group = [&](int cb) {
#pragma unroll
      for (int k = 0; k < DG; ++k) {
        nx[k][0] = nx[k][1] = nx[k][2] = nx[k][3] = 0.f;
        if (cb + k < c1) {
          if (vec) {
            const float4 v4 = ldg_stream4(cp + (int64_t)(cb + k) * hw);
            nx[k][0] = v4.x; nx[k][1] = v4.y; nx[k][2] = v4.z; nx[k][3] = v4.w;
          } else {
            for (int i = 0; i < nv; ++i) nx[k][i] = cp[(int64_t)(cb + k) * hw + i];
          }
        }
      }
    };
    load_group(c0);
    int64_t idx[4];
    float w[4];
    int lab[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      idx[i] = -1;
      w[i] = 0.f;
      if (i < nv) {
        idx[i] = pidx[pbase + i];
        w[i] = pw[pbase + i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) lab[i] = (int)label_of(idx[i], G, gt_labels + g0, C);
    for (int cb = c0; cb < c1; cb += DG) {
      float xv[DG][4];
#pragma unroll
      for (int k = 0; k < DG; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[k][i] = nx[k][i];
      if (cb + DG < c1) load_group(cb + DG);
#pragma unroll
      for (int k = 0; k < DG; ++k) {
        const int c = cb + k;
        if (c < c1) {
          float gv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float lo_, gr_;
            focal_elem<kGamma2>(xv[k][i], lab[i] == c, cfg.gamma, cfg.alpha, lo_, gr_);
            lsum += w[i] * lo_;          // w = 0 for padding lanes
            gv[i] = k_cls * w[i] * gr_;
          }
          if (want_grad) {
            if (vec) {
              stg_stream4(gp + (int64_t)c * hw, make_float4(gv[0], gv[1], gv[2], gv[3]));
            } else {
              for (int i = 0; i < nv; ++i) gp[(int64_t)c * hw + i] = gv[i];
            }
          }
        }
      }
    }
    if (j == 0 && want_grad) {
      // bbox / iou gradients of these 4 points (zero for negatives)
      float gb[4][4], gi[4];
      const float st = (float)grid.stride[l];
      const float s = grid.nrm * (grid.anchor_scale * st);
      const float* bp = maps.bbox[l] + (int64_t)b * 4 * hw + q0;
      const float* ip = maps.iou[l] + (int64_t)b * hw + q0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        gb[0][i] = gb[1][i] = gb[2][i] = gb[3][i] = 0.f;
        gi[i] = 0.f;
        if (i < nv && idx[i] >= 0 && G > 0) {
          if (has_pos) {
            const int q = q0 + i;
            const int y = q / grid.w[l], x = q - y * grid.w[l];
            const float cx = (float)x * st, cy = (float)y * st;
            const float T = bp[i], Bt = bp[hw + i], L = bp[2 * hw + i], R = bp[3 * hw + i];
            const float xi = ip[i];
            float tT, tB, tL, tR;
            point_target(idx[i], G, gt_bboxes + 4 * (int64_t)g0, cx, cy, tT, tB, tL, tR);
            const BoxTerms bt = box_terms<true>(cx, cy, s, T, Bt, L, R, tT, tB, tL, tR, 1e-6f, cfg.eps);
            const float wq = fmaxf(bt.iou, 1e-12f) * w[i];
            const float kb = -k_box * wq;  // d(1-giou) = -dgiou
#pragma unroll
            for (int k = 0; k < 4; ++k) gb[k][i] = kb * bt.d[k];
            gi[i] = k_iou * w[i] * (sigmoidf_(xi) - bt.iou);
          } else {  // radet_head.py:280-281: loss = sum of the positive predictions
#pragma unroll
            for (int k = 0; k < 4; ++k) gb[k][i] = gs_box;
            gi[i] = gs_iou;
          }
        }
      }
      float* gbp = grads.bbox[l] + (int64_t)b * 4 * hw + q0;
      float* gip = grads.iou[l] + (int64_t)b * hw + q0;
      if (vec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) stg_stream4(gbp + (int64_t)k * hw, make_float4(gb[k][0], gb[k][1], gb[k][2], gb[k][3]));
        stg_stream4(gip, make_float4(gi[0], gi[1], gi[2], gi[3]));
      } else {
        for (int i = 0; i < nv; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) gbp[(int64_t)k * hw + i] = gb[k][i];
          gip[i] = gi[i];
        }
      }
    }
  }
  // loss_cls = w_cls * sum / (num_pos + num_imgs): deterministic block partials, last block finalises
  __shared__ double s_part[kDenseThreads / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double v = warp_sum((double)lsum);
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w_ = 0; w_ < kDenseThreads / 32; ++w_) s += s_part[w_];
    partials[blockIdx.x] = s;
    __threadfence();
    s_last = (atomicAdd(&ws->counter_dense, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double tot = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kDenseThreads) tot += partials[i];
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

static int dense_plan(const GridDev& g, int B, int C, DenseTable* tab, int* cc, int* nj) {
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
  // classes per thread: every thread re-derives the labels of its 4 points, so a thread should own several classes
  // (>= 4, one register-buffered load group); beyond that, keep >= ~2 CTAs of 256 threads per SM in flight
  const int64_t target_threads = 148ll * 256 * 2;
  int c = (int)((u * (int64_t)C + target_threads - 1) / target_threads);
  if (c < 4) c = 4;
  if (c > C) c = C;
  *cc = c;
  *nj = (C + c - 1) / c;
  return RADET_OK;
}

extern "C" size_t radet_loss_workspace_bytes(const radet_grid_t* grid, int32_t batch, int32_t num_classes) {
  GridDev g;
  if (make_grid_dev(grid, &g) != RADET_OK || batch <= 0 || num_classes <= 0) return 0;
  DenseTable tab;
  int cc, nj;
  if (dense_plan(g, batch, num_classes, &tab, &cc, &nj) != RADET_OK) return 0;
  const int64_t n = (int64_t)batch * g.off[g.num_levels];
  const int64_t pos_blocks = (n + kPosThreads - 1) / kPosThreads;
  const int64_t dense_blocks = ((int64_t)tab.uoff[g.num_levels] * num_classes + kDenseThreads - 1) / kDenseThreads;  // upper bound (cc=1)
  return align_up(sizeof(LossWs), 256) + align_up((size_t)pos_blocks * 6 * 8, 256) + align_up((size_t)dense_blocks * 8, 256);
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
  int cc, nj;
  rc = dense_plan(g, batch, num_classes, &tab, &cc, &nj);
  if (rc != RADET_OK) return rc;
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  LossWs* ws = reinterpret_cast<LossWs*>(wsb);
  const int64_t n = (int64_t)batch * g.off[g.num_levels];
  const int64_t pos_blocks = (n + kPosThreads - 1) / kPosThreads;
  double* pos_part = reinterpret_cast<double*>(wsb + align_up(sizeof(LossWs), 256));
  double* dense_part = reinterpret_cast<double*>(wsb + align_up(sizeof(LossWs), 256) + align_up((size_t)pos_blocks * 6 * 8, 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (phases & RADET_LOSS_PHASE_NORMALIZERS) {
    loss_pos_kernel<<<(unsigned)pos_blocks, kPosThreads, 0, st>>>(g, batch, md, gt_offsets, gt_bboxes, points_to_gt_index,
                                                                  points_weight, *cfg, ws, pos_part);
    RADET_LAUNCH_CHECK();
  }
  if (!(phases & RADET_LOSS_PHASE_DENSE)) return RADET_OK;
  const int64_t dthreads = (int64_t)tab.uoff[g.num_levels] * nj;
  const unsigned dblocks = (unsigned)((dthreads + kDenseThreads - 1) / kDenseThreads);
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
  const unsigned bx = (unsigned)((nmax + 255) / 256 > 74 ? 74 : (nmax + 255) / 256);
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
