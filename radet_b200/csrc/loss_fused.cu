// Single-launch fused head loss (forward + backward) on sm_100a: RADetHead.loss (radet_head.py:173-288).
//
// One kernel of independent WARPS that draw work from a ticket counter:
//
//   tickets 0 .. n1-1      phase 1, one 512-point chunk each: (a) sum of the weights of the "positive" points
//                          (idx >= 0, radet_head.py:245-254) -> num_pos, published behind flag_a as soon as the last
//                          chunk has reported; (b) the sparse positive terms (IoU target, GIoU, BCE and their
//                          gradients, parked un-normalised in the gradient planes) -> sum wq etc., behind flag_b.
//   tickets n1 ..          phase 2 items = (image, level, 128-point group, class chunk): the HBM-bound pass.  The
//                          class planes of the group stream through the warp's private shared-memory ring, filled by
//                          1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), CG planes per stage; lanes
//                          hold 4 consecutive points (one float4 per plane), rebuild labels from points_to_gt_index,
//                          evaluate the sigmoid focal loss and its gradient and store the gradient with the final
//                          normaliser applied.  The item of class chunk 0 also writes the regression / IoU gradient
//                          planes (zero except at the positives parked by phase 1).
//
// Why tickets: an item has to wait for flag_a (it needs 1 / (num_pos + num_imgs) before it can store a gradient).
// A warp only ever waits for phase-1 chunks, and every phase-1 chunk was drawn by a warp that is already running (it
// holds the ticket) and never waits itself -- so the wait cannot deadlock, whatever else shares the GPU and however
// many CTAs of this grid are resident.  While it waits, an item's first stages are already in flight.
//
// Determinism: every chunk / item writes its partial sum to its own slot; the last reporter adds the slots in a fixed
// order (lane-strided, then an xor tree), so the result does not depend on which warp ran what.
//
// Algorithmic traffic: (8C + 52) B/point: C logits read + C gradients written, 16 + 4 B predictions read,
// 20 B gradients written, 8 B index + 4 B weight read.
#include "loss_common.cuh"

namespace radet {

constexpr int kFWarps = 8;
constexpr int kFThreads = kFWarps * 32;
constexpr int kFOcc = 2;              // CTAs per SM: 16 warps with up to 128 registers per thread and a 12 KB ring each
constexpr int kChunkPts = 512;        // phase-1 chunk: 32 lanes x 4 points x 4 iterations, all loads in flight at once
constexpr int kChunkIters = kChunkPts / 128;

// Control block of one launch, in the workspace behind the partial sums.  Every hot word has its own 128-byte line: the
// ticket counter takes one atomic per CTA, the report counters one per phase-1 chunk, and the two "normalisers are
// ready" flags are REPLICATED per SM (flag and values in one line; a warp polls the copy of the SM it runs on), so the
// thousands of warps that wait for num_pos at the start do not hammer the L2 slice the reporters' atomics go to.
constexpr int kFlagCopies = 160;
struct __align__(128) FlagLine {
  unsigned flag, pad;
  double v[3];
  unsigned fill[24];
};
struct __align__(128) Counter {
  unsigned v;
  unsigned fill[31];
};
struct FusedCtl {
  Counter ticket, cnt_a, cnt_b, done;
  FlagLine fa[kFlagCopies];     // v[0] = num_pos (the normaliser), v[1] = rank-local num_pos
  FlagLine fb[kFlagCopies];     // v[0] = sum wq
};
static_assert(sizeof(FlagLine) == 128 && sizeof(Counter) == 128, "one line each");

struct FusedPlan {
  int gpl[RADET_MAX_LEVELS];        // 128-point groups per (image, level)
  int goff[RADET_MAX_LEVELS + 1];   // group offset of each level inside one image
  int cc, nj;                       // classes per item, class chunks per group
  int n1;                           // phase-1 chunks (0: the normalisers are already in the workspace)
  int items;                        // 0: phase 1 only (the caller all-reduces the normalisers before the items run)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait for a phase-1 flag (lane 0 polls).  The wait is bounded by construction (see the header); the timeout only turns a
// logic error into a trap instead of a hung device.
__device__ __forceinline__ unsigned flag_copy_of_this_sm() {
  unsigned s;
  asm("mov.u32 %0, %%smid;" : "=r"(s));
  return s % kFlagCopies;
}
__device__ __forceinline__ void wait_flag(const unsigned* flag) {
  unsigned polls = 0;
  while (ld_acquire_u32(flag) == 0u) {
    __nanosleep(400);
    if (++polls > (1u << 22)) __trap();     // seconds: a logic error, not a wait
  }
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {   // after a __threadfence(): fence + relaxed store = release
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <bool kGamma2, int CG, int D>
__global__ void __launch_bounds__(kFThreads, kFOcc)
loss_fused_kernel(GridDev grid, FusedPlan plan, int B, int C, MapsDev maps, GradsDev grads, const int* __restrict__ gt_offsets,
                  const float* __restrict__ gt_bboxes, const int64_t* __restrict__ gt_labels, const int64_t* __restrict__ pidx,
                  const float* __restrict__ pw, radet_loss_cfg_t cfg, const float* __restrict__ grad_scale, LossWs* ws,
                  FusedCtl* ctl, double* part_a, double* part_b, double* part_c, float* __restrict__ losses,
                  const double* __restrict__ num_pos_hint, unsigned long long* dbg) {
#define FDBG(tk, k) do { if (dbg && lane == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); dbg[(size_t)(tk) * 8 + (k)] = t__; } } while (0)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ unsigned s_base, s_done;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#define FDBG_CTA(k) do { if (dbg && tid == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); dbg[(size_t)(190000 + blockIdx.x) * 8 + (k)] = t__; } } while (0)
  FDBG_CTA(0);
  float* ring = reinterpret_cast<float*>(smem_raw) + wid * (D * CG * 128);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kFWarps * D * CG * 512) + wid * D;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < D; ++k) mbar_init(&full[k], 1);
    fence_barrier_init();
  }
  // The ring starts out zeroed: lanes beyond a short group read whatever their slot held before, and with a zero weight
  // that contributes exactly 0 as long as it is finite (stale logits or these zeros).
  for (int i = lane; i < D * CG * 32; i += 32) reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(full);

  const int P = grid.off[grid.num_levels];
  const int64_t n = (int64_t)B * P;
  const bool n32 = n < (1ll << 31);                      // image of a point with a 32-bit division in the usual case
  auto image_of = [&](int64_t pt) -> int { return n32 ? (int)((unsigned)pt / (unsigned)P) : (int)(pt / P); };
  const int gpi = plan.goff[grid.num_levels];            // groups per image
  const int nj = plan.nj, cc = plan.cc;
  const bool want_grad = grads.cls[0] != nullptr;
  const unsigned n1 = (unsigned)plan.n1;
  const unsigned n_groups = (unsigned)B * (unsigned)gpi;
  const unsigned n_cls = plan.items ? n_groups * (unsigned)nj : 0u;           // class items
  const unsigned n_head = plan.items ? n_groups * (unsigned)(nj - 1) : 0u;    // ... of which chunks 0 .. nj-2 come first
  const unsigned n_tickets = n1 + n_cls;
  const float gamma = cfg.gamma, alpha = cfg.alpha;
  const float gs_cls = grad_scale ? grad_scale[0] : 1.f, gs_box = grad_scale ? grad_scale[1] : 1.f,
              gs_iou = grad_scale ? grad_scale[2] : 1.f;
  bool have_a = false, have_b = false, has_pos = false;
  double s_num_pos = 0.0;
  float k_cls = 0.f, k_box = 0.f, k_iou = 0.f;
  unsigned slot = 0, par = 0;                            // next ring slot of this warp and its mbarrier phase parity
  const unsigned sm_copy = flag_copy_of_this_sm();

  // num_pos, published by the last phase-1 chunk (or left in the workspace by an earlier launch)
  auto need_a = [&]() {
    if (have_a) return;
    double num_pos, num_pos_local;
    if (num_pos_hint) {                                  // per-image sums handed over by the producer of the assignment
      num_pos = num_pos_local = warp_sum_array(num_pos_hint, (unsigned)B, lane);
    } else if (n1) {                                     // every chunk has reported: add the partial sums (fixed order)
      if (lane == 0) wait_flag(&ctl->fa[sm_copy].flag);
      __syncwarp();
      num_pos = num_pos_local = warp_sum_array(part_a, n1, lane);
    } else {
      num_pos = ws->norm[0];
      num_pos_local = ws->norm[6];
    }
    has_pos = num_pos_local > 0.0;                                       // radet_head.py:261
    s_num_pos = num_pos;
    k_cls = gs_cls * cfg.w_cls / (float)(num_pos + (double)cfg.avg_extra);
    have_a = true;
  };

  // The first ticket of every warp comes from ONE atomic per CTA (consecutive tickets: neighbouring chunks / items);
  // this is the only CTA-wide synchronisation of the kernel.  Every later ticket is drawn by the warp itself, one ticket
  // ahead, so that the atomic's round trip hides behind the work of the current ticket.
  if (tid == 0) {
    s_base = atomicAdd(&ctl->ticket.v, (unsigned)kFWarps);
    s_done = 0u;
  }
  __syncthreads();
  FDBG_CTA(1);
  unsigned t = s_base + (unsigned)wid;
  while (t < n_tickets) {
    unsigned t_next = 0;
    if (lane == 0) t_next = atomicAdd(&ctl->ticket.v, 1u);
    FDBG(t, 0);
    if (t < n1) {
      // ======================================================================================== phase 1, chunk t
      const int64_t base = (int64_t)t * kChunkPts;
      // (a) num_pos = sum of the weights of the points with idx >= 0 in images that have ground truth
      longlong2 i01[kChunkIters], i23[kChunkIters];
      float4 w4[kChunkIters];
      int gcnt[kChunkIters];
#pragma unroll
      for (int it = 0; it < kChunkIters; ++it) {
        const int64_t pt = base + it * 128 + 4 * lane;
        i01[it] = i23[it] = make_longlong2(-1, -1);
        w4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        gcnt[it] = 0;
        if (pt < n) {                                    // n is a multiple of 4: whole vectors only
          i01[it] = *reinterpret_cast<const longlong2*>(pidx + pt);
          i23[it] = *reinterpret_cast<const longlong2*>(pidx + pt + 2);
          w4[it] = *reinterpret_cast<const float4*>(pw + pt);
          const int b = image_of(pt);
          gcnt[it] = gt_offsets[b + 1] - gt_offsets[b];
        }
      }
      unsigned mask = 0u;
      double s0 = 0.0;
#pragma unroll
      for (int it = 0; it < kChunkIters; ++it) {
        if (gcnt[it] > 0) {
          if (i01[it].x >= 0) { mask |= 1u << (4 * it + 0); s0 += (double)w4[it].x; }
          if (i01[it].y >= 0) { mask |= 1u << (4 * it + 1); s0 += (double)w4[it].y; }
          if (i23[it].x >= 0) { mask |= 1u << (4 * it + 2); s0 += (double)w4[it].z; }
          if (i23[it].y >= 0) { mask |= 1u << (4 * it + 3); s0 += (double)w4[it].w; }
        }
      }
      s0 = warp_sum(s0);
      FDBG(t, 1);
      unsigned old = 0;
      if (lane == 0) {
        part_a[t] = s0;
        __threadfence();
        old = atomicAdd(&ctl->cnt_a.v, 1u);
      }
      old = __shfl_sync(kFull, old, 0);
      FDBG(t, 2);
      if (old == n1 - 1) {                               // last chunk to report: raise the per-SM flags (ONE fence, then the
        __threadfence();                                 // stores); every waiter adds the partial sums itself
        for (int i = lane; i < kFlagCopies; i += 32) st_relaxed_u32(&ctl->fa[i].flag, 1u);
        FDBG(t, 3);
      }
      // (b) the positives of the chunk, compacted through the (idle) ring: one positive per lane and round
      unsigned short* s_list = reinterpret_cast<unsigned short*>(ring);
      const int cnt = __popc(mask);
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
      }
      const int total = __shfl_sync(kFull, inc, 31);
      {
        int pos = inc - cnt;
        unsigned m = mask;
        while (m) {
          const int k = __ffs((int)m) - 1;
          m &= m - 1u;
          s_list[pos++] = (unsigned short)((k >> 2) * 128 + 4 * lane + (k & 3));
        }
      }
      __syncwarp();
      float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      for (int i = lane; i < total; i += 32) {
        const int64_t pt = base + s_list[i];
        const int64_t idx = pidx[pt];
        const int b = image_of(pt), p = (int)(pt - (int64_t)b * P);
        const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;          // > 0 (checked in (a))
        const float w = pw[pt];
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
        const float wq = fmaxf(bt.iou, 1e-12f) * w;                        // radet_head.py:272
        acc[0] += wq;
        acc[1] += wq * (1.f - bt.giou);
        acc[2] += w * bce_logits(xi, bt.iou);
        acc[3] += (T + Bt) + (L + R);
        acc[4] += xi;
        if (want_grad) {                                                    // parked un-normalised; rescaled by the group's box item
          float* gb = grads.bbox[l] + (int64_t)b * 4 * hw + q;
          gb[0] = -wq * bt.d[0];                                            // d(1 - giou) = -d giou
          gb[hw] = -wq * bt.d[1];
          gb[2 * hw] = -wq * bt.d[2];
          gb[3 * hw] = -wq * bt.d[3];
          grads.iou[l][(int64_t)b * hw + q] = w * (sigmoidf_(xi) - bt.iou);
        }
      }
      FDBG(t, 4);
      __syncwarp();                                      // the list lives in the ring: done with it before the next item
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy use of the ring before later bulk copies into it
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) part_b[(size_t)k * n1 + t] = (double)v;
      }
      old = 0;
      if (lane == 0) {
        __threadfence();
        old = atomicAdd(&ctl->cnt_b.v, 1u);
      }
      old = __shfl_sync(kFull, old, 0);
      if (old == n1 - 1) {
        __threadfence();
        for (int i = lane; i < kFlagCopies; i += 32) st_relaxed_u32(&ctl->fb[i].flag, 1u);
        FDBG(t, 6);
      }
      FDBG(t, 5);
    } else {
      // ======================================================================================== phase 2, one class item
      // Ticket order: chunks 0 .. nj-2 of every group first (group-major), then the LAST chunk of every group.  The
      // last-chunk items also write the group's regression / IoU gradient planes, for which they need sum wq -- by the
      // time their tickets come up, phase 1 has long finished.
      const unsigned it_ = t - n1;
      int gg, j;
      if (it_ < n_head) {
        gg = (int)(it_ / (unsigned)(nj - 1));
        j = (int)(it_ - (unsigned)gg * (unsigned)(nj - 1));
      } else {
        gg = (int)(it_ - n_head);
        j = nj - 1;
      }
      const unsigned item = (unsigned)gg * (unsigned)nj + (unsigned)j;        // slot of its partial sum
      const int b = gg / gpi, r = gg - b * gpi;
      int l = 0;
#pragma unroll
      for (int k = 1; k < RADET_MAX_LEVELS; ++k) l += (k < grid.num_levels && r >= plan.goff[k]) ? 1 : 0;
      const int hw = grid.h[l] * grid.w[l];
      const int q_warp = (r - plan.goff[l]) * 128;
      const int npts = min(128, hw - q_warp);                                // multiple of 4, > 0
      const int q0 = q_warp + 4 * lane;
      const bool active = 4 * lane < npts;
      const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
      const int64_t pbase = (int64_t)b * P + grid.off[l] + q0;
      longlong2 a01 = make_longlong2(-1, -1), a23 = make_longlong2(-1, -1);
      if (active) {
        a01 = *reinterpret_cast<const longlong2*>(pidx + pbase);
        a23 = *reinterpret_cast<const longlong2*>(pidx + pbase + 2);
      }
      {
        // ---- class item: planes c0 .. c0 + ncls - 1 of the group
        const uint32_t bytes = (uint32_t)npts * 4u;
        const int c0 = j * cc, ncls = min(cc, C - c0);
        const int nst = (ncls + CG - 1) / CG;
        const float* nsrc = maps.cls[l] + ((int64_t)b * C + c0) * hw + q_warp;   // first plane of the next stage to request
        int issued = 0;
        auto issue_next = [&](unsigned sl_) {                                // lane 0: the next stage of this item into ring slot sl_
          const int np = ncls - issued * CG;                                 // planes left; a full stage has CG
          const uint32_t bar = full_s + sl_ * 8u;
          const uint32_t sdst = ring_s + sl_ * (CG * 512u);
          mbar_expect_tx_s(bar, bytes * (uint32_t)min(np, CG));
#pragma unroll
          for (int p_ = 0; p_ < CG; ++p_)
            if (p_ < np) tma_bulk_g2s_s(sdst + p_ * 512u, nsrc + (int64_t)p_ * hw, bytes, bar);
          nsrc += (int64_t)CG * hw;
          ++issued;
        };
        auto fill_ring = [&](int upto) {                                     // lane 0: request stages until `upto` are in flight
          unsigned sl = slot + (unsigned)issued;
          if (sl >= (unsigned)D) sl -= (unsigned)D;
          while (issued < upto) {
            issue_next(sl);
            sl = (sl + 1 == (unsigned)D) ? 0u : sl + 1;
          }
        };
        // Until num_pos is known only ONE stage is requested: the phase-1 chunks that every item waits for read their
        // indices and weights through the same memory system, and a full ring per warp in front of them delays them.
        if (lane == 0) fill_ring(min(have_a ? D : 1, nst));
        float w[4];
        {
          float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (active) wv = *reinterpret_cast<const float4*>(pw + pbase);
          w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
        }
        int lab[4];
        float wt[4], wn[4];
        {
          const int64_t v[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ix = v[i] < 0 ? -1 : (int)(v[i] > (int64_t)G ? (int64_t)G : v[i]);
            lab[i] = (int)label_of(ix, G, gt_labels + g0, C) - c0;           // relative to the chunk
            wt[i] = alpha * w[i];
            wn[i] = (1.f - alpha) * w[i];
          }
        }
        if (!have_a) {
          need_a();
          if (lane == 0) fill_ring(min(D, nst));
        }
        FDBG(t, 1);
        float* dst = want_grad && active ? grads.cls[l] + ((int64_t)b * C + c0) * hw + q0 : nullptr;   // next plane to store
        float lsum = 0.f;
        for (int s = 0; s < nst; ++s) {
          mbar_wait_s(full_s + slot * 8u, par);
          const float4* rp = reinterpret_cast<const float4*>(ring + slot * (CG * 128)) + lane;
#pragma unroll
          for (int p_ = 0; p_ < CG; ++p_) {
            const int c = s * CG + p_;
            if (c < ncls) {                                                  // warp-uniform: the last stage of a chunk may be short
              const float4 xv = rp[p_ * 32];
              float4 gv;
              gv.x = focal_acc<kGamma2>(xv.x, lab[0] == c, wt[0], wn[0], k_cls, gamma, lsum);
              gv.y = focal_acc<kGamma2>(xv.y, lab[1] == c, wt[1], wn[1], k_cls, gamma, lsum);
              gv.z = focal_acc<kGamma2>(xv.z, lab[2] == c, wt[2], wn[2], k_cls, gamma, lsum);
              gv.w = focal_acc<kGamma2>(xv.w, lab[3] == c, wt[3], wn[3], k_cls, gamma, lsum);
              if (dst) {
                __stcs(reinterpret_cast<float4*>(dst), gv);
                dst += hw;
              }
            }
          }
          // Refill the slot only AFTER its values have been consumed (the stores above depend on them): neither a barrier
          // nor an mbarrier arrive orders the bulk copy's write behind a shared-memory read that is still in flight.
          __syncwarp();
          if (lane == 0 && issued < nst) issue_next(slot);
          if (++slot == D) {
            slot = 0;
            par ^= 1u;
          }
        }
        FDBG(t, 2);
        const double ls = warp_sum((double)lsum);
        if (lane == 0) part_c[item] = ls;
        if (want_grad && j == nj - 1) {
          // ---- regression / IoU gradient planes of the group: zero except at the positives parked by phase 1, which
          // are rescaled in place
          if (!have_b) {
            double sum_wq;
            if (n1) {
              if (lane == 0) wait_flag(&ctl->fb[sm_copy].flag);
              __syncwarp();
              sum_wq = warp_sum_array(part_b, n1, lane);                   // part_b[0][*] = sum wq per chunk
            } else {
              sum_wq = ws->norm[1];
            }
            k_box = has_pos ? gs_box * cfg.w_bbox / (float)sum_wq : gs_box;
            k_iou = has_pos ? gs_iou * cfg.w_iou / (float)s_num_pos : gs_iou;
            have_b = true;
          }
          if (active) {
            const bool pos4[4] = {a01.x >= 0, a01.y >= 0, a23.x >= 0, a23.y >= 0};   // idx >= 0 (radet_head.py:245-247)
            const bool anypos = G > 0 && (pos4[0] || pos4[1] || pos4[2] || pos4[3]);
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) {
              float* o = kk < 4 ? grads.bbox[l] + ((int64_t)b * 4 + kk) * hw + q0 : grads.iou[l] + (int64_t)b * hw + q0;
              const float kn = kk < 4 ? k_box : k_iou, gs = kk < 4 ? gs_box : gs_iou;
              float gvv[4] = {0.f, 0.f, 0.f, 0.f};
              if (anypos) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (pos4[i]) gvv[i] = has_pos ? kn * __ldcg(o + i) : gs;    // radet_head.py:280-281 when num_pos == 0
              }
              __stcs(reinterpret_cast<float4*>(o), make_float4(gvv[0], gvv[1], gvv[2], gvv[3]));
            }
          }
        }
        FDBG(t, 3);
      }
    }
    t = __shfl_sync(kFull, t_next, 0);
  }

  // ---- exit: the warps of a CTA join, the last CTA to leave adds the partial sums in a fixed order with all its warps,
  // writes the losses and the normalisers, and re-arms the control block
  __syncthreads();
  if (tid == 0) {
    __threadfence();                                     // the CTA's writes (ordered before by the barrier) become visible
    s_done = (atomicAdd(&ctl->done.v, 1u) == gridDim.x - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_done) return;
  __threadfence();
  if (dbg && tid == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); dbg[(size_t)(199990) * 8 + 0] = t__; }
  __shared__ double s_fin[kFWarps], s_cls[kFWarps];
  // class-loss items: thread-strided over the whole CTA (its size is a compile-time constant, so the order is fixed);
  // then warps 1..6 add one phase-1 array each (num_pos, sum wq, sum wq (1 - giou), sum w bce, the two plain sums)
  {
    const double c = warp_sum(strided_sum(part_c, n_cls, (unsigned)tid, (unsigned)kFThreads));
    if (lane == 0) s_cls[wid] = c;
  }
  double v = 0.0;
  if (wid >= 1 && wid <= 6 && n1) v = wid == 1 ? warp_sum_array(part_a, n1, lane) : warp_sum_array(part_b + (size_t)(wid - 2) * n1, n1, lane);
  if (wid == 7 && num_pos_hint) v = warp_sum_array(num_pos_hint, (unsigned)B, lane);
  if (lane == 0) s_fin[wid] = v;
  __syncthreads();
  if (tid == 0) {
    double cls_sum = 0.0;
    for (int k = 0; k < kFWarps; ++k) cls_sum += s_cls[k];
    s_fin[0] = cls_sum;
    if (n1) {
      if (num_pos_hint) s_fin[1] = s_fin[7];             // the value every item used
      ws->norm[0] = s_fin[1];                            // the normaliser (a caller may all-reduce it between two launches)
      ws->norm[6] = s_fin[1];                            // rank-local num_pos (radet_head.py:254,261)
      for (int k = 0; k < 5; ++k) ws->norm[1 + k] = s_fin[2 + k];
      ws->norm[7] = s_fin[2];
    }
    if (n_cls) {
      const double num_pos = ws->norm[0], sum_wq = ws->norm[1];
      const bool hp = ws->norm[6] > 0.0;
      losses[0] = (float)((double)cfg.w_cls * s_fin[0] / (num_pos + (double)cfg.avg_extra));                 // radet_head.py:256-259
      losses[1] = hp ? (float)((double)cfg.w_bbox * ws->norm[2] / sum_wq) : (float)ws->norm[4];               // :269-274 / :280
      losses[2] = hp ? (float)((double)cfg.w_iou * ws->norm[3] / num_pos) : (float)ws->norm[5];               // :275-278 / :281
      losses[3] = (float)num_pos;
    }
    ctl->ticket.v = 0u;
    ctl->cnt_a.v = 0u;
    ctl->cnt_b.v = 0u;
    ctl->done.v = 0u;
  }
  if (n1) {
    for (int i = tid; i < kFlagCopies; i += kFThreads) {
      ctl->fa[i].flag = 0u;
      ctl->fb[i].flag = 0u;
    }
  }
  if (dbg && tid == 0) { unsigned long long t__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); dbg[(size_t)(199990) * 8 + 1] = t__; }
}

// ------------------------------------------------------------------------------------------------ host side
constexpr int kSMsF = 148;

bool fused_plan(const GridDev& g, int B, int C, int CG, bool with_phase1, bool with_items, FusedPlan* plan, int* blocks) {
  int t = 0;
  for (int l = 0; l < g.num_levels; ++l) {
    const int hw = g.h[l] * g.w[l];
    if (hw & 3) return false;                       // every plane must be 16-byte tileable
    plan->gpl[l] = (hw + 127) / 128;
    plan->goff[l] = t;
    t += plan->gpl[l];
  }
  for (int l = g.num_levels; l <= RADET_MAX_LEVELS; ++l) plan->goff[l] = t;
  for (int l = g.num_levels; l < RADET_MAX_LEVELS; ++l) plan->gpl[l] = 0;
  const int64_t groups = (int64_t)B * t;
  const int64_t n = (int64_t)B * g.off[g.num_levels];
  const int64_t n1 = with_phase1 ? (n + kChunkPts - 1) / kChunkPts : 0;
  if (groups * C + n1 >= (1ll << 31)) return false;
  // Class chunks per group: enough items for the warps to balance dynamically (about `ipw` items per resident warp), as
  // few as that allows (every item re-reads its group's indices / weights and pays a fixed set-up).  A small batch ends
  // up with short items on every SM (latency-bound regime), a large one with 10..16 planes per item.
  const int64_t workers = (int64_t)kSMsF * kFOcc * kFWarps;
  double ipw = 3.0;
  if (const char* e = getenv("RADET_FUSED_IPW")) {        // development override
    const double v = atof(e);
    if (v > 0.0) ipw = v;
  }
  int64_t nj = (int64_t)(ipw * (double)workers / (double)(groups > 0 ? groups : 1) + 0.5);
  const int max_nj = (C + CG - 1) / CG;
  if (nj > max_nj) nj = max_nj;
  if (nj < 1) nj = 1;
  int cc = (int)((C + nj - 1) / nj);
  cc = (cc + CG - 1) / CG * CG;
  plan->cc = cc;
  plan->nj = (C + cc - 1) / cc;
  plan->n1 = (int)n1;
  plan->items = with_items ? 1 : 0;
  const int64_t tickets = n1 + (with_items ? groups * plan->nj : 0);
  int64_t ctas = (tickets + kFWarps - 1) / kFWarps;
  const int64_t slots = (int64_t)kSMsF * kFOcc;
  if (ctas > slots) ctas = slots;
  *blocks = (int)(ctas < 1 ? 1 : ctas);
  return true;
}

size_t fused_part_bytes(const GridDev& g, int B, int C) {
  // control block + part_a [n1] + part_b [5][n1] + part_c [class items <= groups * C] doubles
  int64_t t = 0;
  for (int l = 0; l < g.num_levels; ++l) t += (g.h[l] * g.w[l] + 127) / 128;
  const int64_t n1 = ((int64_t)B * g.off[g.num_levels] + kChunkPts - 1) / kChunkPts;
  return align_up(sizeof(FusedCtl), 256) + align_up((size_t)n1 * 8, 256) + align_up((size_t)n1 * 5 * 8, 256) + align_up((size_t)B * t * C * 8, 256);
}

template <bool kGamma2, int CG, int D>
static int launch_variant(const GridDev& g, const FusedPlan& plan, int blocks, int B, int C, const MapsDev& md, const GradsDev& gd,
                          const int* gt_offsets, const float* gt_bboxes, const int64_t* gt_labels, const int64_t* pidx, const float* pw,
                          const radet_loss_cfg_t& cfg, const float* grad_scale, LossWs* ws, FusedCtl* ctl, double* pa, double* pb,
                          double* pc, float* losses, const double* hint, cudaStream_t st) {
  const size_t smem = (size_t)kFWarps * D * CG * 512 + (size_t)kFWarps * D * 8;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(loss_fused_kernel<kGamma2, CG, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  loss_fused_kernel<kGamma2, CG, D><<<(unsigned)blocks, kFThreads, smem, st>>>(g, plan, B, C, md, gd, gt_offsets, gt_bboxes, gt_labels,
                                                                              pidx, pw, cfg, grad_scale, ws, ctl, pa, pb, pc, losses, hint,
                                                                              static_cast<unsigned long long*>(g_debug_buf));
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

int fused_cg() {
  static const int cg = [] {
    const char* e = getenv("RADET_FUSED_CG");      // development override: 2 or 4 planes per stage
    return (e && e[0] == '2') ? 2 : 4;
  }();
  return cg;
}

// Launches the fused kernel; returns RADET_E_UNSUPPORTED when the shape is not eligible (caller falls back).
int launch_loss_fused(const GridDev& g, int B, int C, const MapsDev& md, const GradsDev& gd, const int* gt_offsets, const float* gt_bboxes,
                      const int64_t* gt_labels, const int64_t* pidx, const float* pw, const radet_loss_cfg_t& cfg,
                      const float* grad_scale, LossWs* ws, unsigned char* parts, bool with_phase1, bool with_items, float* losses,
                      const double* num_pos_hint, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(pidx) & 15) || (reinterpret_cast<uintptr_t>(pw) & 15)) return RADET_E_UNSUPPORTED;
  const int CG = fused_cg();
  FusedPlan plan;
  int blocks = 0;
  if (!fused_plan(g, B, C, CG, with_phase1, with_items, &plan, &blocks)) return RADET_E_UNSUPPORTED;
  const int64_t n1 = ((int64_t)B * g.off[g.num_levels] + kChunkPts - 1) / kChunkPts;
  FusedCtl* ctl = reinterpret_cast<FusedCtl*>(parts);
  parts += align_up(sizeof(FusedCtl), 256);
  double* pa = reinterpret_cast<double*>(parts);
  double* pb = reinterpret_cast<double*>(parts + align_up((size_t)n1 * 8, 256));
  double* pc = reinterpret_cast<double*>(parts + align_up((size_t)n1 * 8, 256) + align_up((size_t)n1 * 5 * 8, 256));
  const bool g2 = cfg.gamma == 2.0f;
#define RADET_FUSED(G2, CG_, D_) \
  return launch_variant<G2, CG_, D_>(g, plan, blocks, B, C, md, gd, gt_offsets, gt_bboxes, gt_labels, pidx, pw, cfg, grad_scale, ws, ctl, pa, pb, pc, losses, (with_phase1 && with_items) ? num_pos_hint : nullptr, st)
  if (CG == 2) {
    if (g2) RADET_FUSED(true, 2, 12);
    RADET_FUSED(false, 2, 12);
  }
  if (g2) RADET_FUSED(true, 4, 6);
  RADET_FUSED(false, 4, 6);
#undef RADET_FUSED
}

}  // namespace radet
