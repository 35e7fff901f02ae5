// Device helpers shared by the loss kernels (loss.cu: target encode, two-launch path, standalone losses;
// loss_fused.cu: the single-launch fused forward+backward).
#pragma once
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

// radet_head.py:373-392 + tblr_bbox_coder.py:71-114.  label: C for idx<0; gt_labels[idx-1] with python negative
// indexing for idx==0 (last GT).  target: ((d / (scale*stride)) / 0.125), order T,B,L,R.
__device__ __forceinline__ int64_t label_of(int64_t idx, int G, const int64_t* __restrict__ gt_labels, int C) {
  if (G <= 0 || idx < 0) return C;
  int64_t k = idx - 1;
  if (k < 0) k += G;
  if (k >= G) k = G - 1;
  return gt_labels[k];
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
constexpr int kPosThreads = 128;
constexpr int kPosPerThread = 8;
constexpr int kNormSlots = 8;   // S0 num_pos, S1 sum wq, S2 sum wq(1-giou), S3 sum w*bce, S4 sum pred, S5 sum iou logit
struct LossWs {
  double norm[kNormSlots];
  unsigned int counter_pos, counter_dense;
  unsigned int pad[2];
};
static_assert(sizeof(LossWs) <= 256, "LossWs must fit its 256-byte slot at the head of the workspace");

// loss_fused.cu
size_t fused_part_bytes(const GridDev& g, int B, int C);   // control block + partial sums
int launch_loss_fused(const GridDev& g, int B, int C, const MapsDev& md, const GradsDev& gd, const int* gt_offsets, const float* gt_bboxes,
                      const int64_t* gt_labels, const int64_t* pidx, const float* pw, const radet_loss_cfg_t& cfg,
                      const float* grad_scale, LossWs* ws, unsigned char* parts, bool with_phase1, bool with_items, float* losses,
                      const double* num_pos_hint, cudaStream_t st);

// Sigmoid focal loss and its derivative for one logit (mmcv sigmoid_focal_loss semantics; restated from
// focal_loss.py:10-41 because the mmcv op is not in the reference tree).  Tolerance parity (not bit parity), so the
// transcendental part is kept to ~28 instructions per element (at 70+ the kernel is issue-bound, not HBM-bound):
// one ex2.approx, one rcp.approx and one lg2.approx:
//     e = exp(-|z|),  inv = 1/(1+e),  log1p(e) = -ln(inv).
// The target case is folded into the non-target one by symmetry: with z = -x,
//     FL_target(x) = alpha * sigmoid(z)^gamma * softplus(z),  dFL_target/dx = -d/dz[...],
// so a single expression  m = coef * s^gamma,  loss = m * softplus(z),  dloss/dz = m * (s + gamma * (1-s) * softplus(z))
// with s = sigmoid(z) serves both (coef = alpha for the target class, 1-alpha otherwise).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool kGamma2>
__device__ __forceinline__ void focal_elem(float x, bool is_t, float gamma, float alpha, float& loss, float& grad) {
  const float z = is_t ? -x : x;
  const float e = ex2_approx(fabsf(z) * -1.4426950408889634f);   // exp(-|z|) in (0, 1]
  const float inv = rcp_approx(1.0f + e);                         // 1/(1+e) in [0.5, 1)
  // softplus(z) = max(z,0) + log1p(e) = max(z,0) - ln(inv): one lg2.approx (absolute error ~1e-7 on a term that is
  // only tiny where the whole element is negligible: its gradient carries the factor s^2 ~ e^2)
  const float sp = fmaf(-0.6931471805599453f, lg2_approx(inv), fmaxf(z, 0.f));
  const float s = z >= 0.f ? inv : e * inv;       // sigmoid(z)
  const float coef = is_t ? alpha : 1.f - alpha;
  const float m = coef * (kGamma2 ? s * s : ex2_approx(gamma * -1.4426950408889634f * (sp - z)));   // s^gamma = exp(-gamma*softplus(-z))
  loss = m * sp;
  const float g = m * fmaf(kGamma2 ? fmaf(-2.f, s, 2.f) : gamma * (1.f - s), sp, s);
  grad = is_t ? -g : g;
}

// shared-memory mbarrier / bulk-copy helpers on precomputed 32-bit shared addresses (no cvta in the loop)
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(phase)
      : "memory");
}

// Sigmoid focal loss of one logit, accumulated, and its gradient (see focal_elem in loss_common.cuh for the derivation).
// wt = alpha * w, wn = (1 - alpha) * w of the point; the sign that the target class puts on the gradient travels in A.
template <bool kGamma2>
__device__ __forceinline__ float focal_acc(float x, bool is_t, float wt, float wn, float k_cls, float gamma, float& lsum) {
  const float z = is_t ? -x : x;
  const float wl = is_t ? -wt : wn;
  const float e = ex2_approx(fabsf(z) * -1.4426950408889634f);   // exp(-|z|) in (0, 1]
  const float inv = rcp_approx(1.0f + e);                         // 1/(1+e) in [0.5, 1)
  const float sp = fmaf(-0.6931471805599453f, lg2_approx(inv), fmaxf(z, 0.f));   // softplus(z)
  const float s = z >= 0.f ? inv : e * inv;                       // sigmoid(z)
  const float u = kGamma2 ? s * s : ex2_approx(gamma * -1.4426950408889634f * (sp - z));   // s^gamma
  const float A = wl * u;
  lsum = fmaf(fabsf(A), sp, lsum);
  const float h = fmaf(kGamma2 ? fmaf(-2.f, s, 2.f) : gamma * (1.f - s), sp, s);
  return (A * h) * k_cls;
}

// Fixed-order partial sum of p[first], p[first + stride], ...: up to 16 loads are issued before the first addition
// (one L2 round trip for the usual few hundred to few thousand partial sums); the combination order is fixed, so the
// value depends only on (n, first, stride).
__device__ __forceinline__ double strided_sum(const double* p, unsigned n, unsigned first, unsigned stride) {
  double tot = 0.0;
  for (unsigned i = first; i < n; i += 16 * stride) {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = (i + k * stride < n) ? __ldcg(p + i + k * stride) : 0.0;
#pragma unroll
    for (int w_ = 1; w_ < 16; w_ <<= 1) {
#pragma unroll
      for (int k = 0; k < 16; k += 2 * w_) a[k] += a[k + w_];
    }
    tot += a[0];
  }
  return tot;
}
// ... by one warp (lane-strided, then an xor tree): the same result whichever warp evaluates it
__device__ __forceinline__ double warp_sum_array(const double* p, unsigned n, int lane) {
  return warp_sum(strided_sum(p, n, (unsigned)lane, 32u));
}

}  // namespace radet
