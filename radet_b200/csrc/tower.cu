// Head-tower epilogues on sm_100a: what sits between the cuDNN convolutions of the head and the hot path.
//
// Reference: every tower layer is ConvModule(conv3x3, GN(32), ReLU) (radet/models/dense_heads/atss_head.py:52-87,
// forward_single :118-145) and the regression branch ends in relu(Scale(conv)) (atss_head.py:141-143, radet_head.py:27-30).
// In the reference these are separate eager kernels (group-norm statistics, normalise, ReLU; multiply, ReLU) with their
// autograd backwards: 5 passes over a feature map forward and 7 backward per tower layer.  Here:
//
//   gn_relu_fwd_kernel     one CTA per (image, group): statistics and normalise + affine + ReLU in one launch.  The group
//                          (C/G channels x H*W, contiguous in NCHW) is read once from HBM and a second time out of L2:
//                          8 B/element of HBM traffic (read x, write y).  60x80 level, B=8, 256 channels: 21.5 us (3.7 TB/s;
//                          torch's native_group_norm + relu: 78 us).
//   gn_relu_bwd_kernel     one CTA per (image, group): ReLU mask recomputed from x (y is not read); teams of warps reduce the
//                          channels' two sums concurrently (one block barrier per round), per-(image, channel) dgamma /
//                          dbeta partial sums in a fixed order (the host adds the N rows), dx in a second pass out of L2:
//                          12 B/element (read dy, x; write dx).  Same shape: 38.9 us (3.0 TB/s; torch: 78 us).
//   A thread-block-cluster variant (one CTA per channel, partial sums exchanged through distributed shared memory, slices
//   held in registers) was built and measured: 30.6 / 70.4 us at the same shape and a ~20 us floor at the small levels --
//   slower than one CTA per group at every level of this head, so it was dropped.
//   scale_relu_*_kernel    y = relu(s * x) and its backward (dx, fixed-order partial sums of ds).
//
// The sigmoid / score-threshold epilogue of the classification branch needs no kernel: detect_select_kernel consumes the
// raw logits (no sigmoid tensor is ever materialised).  The convolutions themselves stay cuDNN (not the product).
#include "common.cuh"

namespace radet {

constexpr int kGnThreads = 512;

__device__ __forceinline__ double block_sum_d(double v, double* scratch /* [kGnThreads / 32] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();                       // scratch reuse
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kGnThreads / 32; ++w) t += scratch[w];   // same order in every thread
  return t;
}

// torch.nn.GroupNorm(G, C, eps) + ReLU.  x, y: [N, C, HW]; gamma, beta: [C] (or nullptr = 1 / 0); mean, rstd: [N, G].
__global__ void __launch_bounds__(kGnThreads)
gn_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int C, int HW,
                   int G, float eps, float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  __shared__ double s_red[kGnThreads / 32];
  const int ng = blockIdx.x, g = ng % G;
  const int cpg = C / G;
  const int64_t m = (int64_t)cpg * HW;
  const float* xg = x + (int64_t)ng * m;
  float* yg = y + (int64_t)ng * m;
  const bool vec = (m & 3) == 0 && (HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(xg) | reinterpret_cast<uintptr_t>(yg)) & 15) == 0;
  // pass 1: sums of (x - K) and (x - K)^2 with K = the group's first element (a shift near the mean keeps the fp32
  // per-thread sums of squares from cancelling); thread partials are combined in fp64 in a fixed order
  const float K = xg[0];
  float s1 = 0.f, s2 = 0.f;
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xg);
    for (int64_t i = threadIdx.x; i < m / 4; i += kGnThreads) {
      const float4 v = x4[i];
      const float a = v.x - K, b = v.y - K, c = v.z - K, d = v.w - K;
      s1 += (a + b) + (c + d);
      s2 = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, s2))));
    }
  } else {
    for (int64_t i = threadIdx.x; i < m; i += kGnThreads) {
      const float a = xg[i] - K;
      s1 += a;
      s2 = fmaf(a, a, s2);
    }
  }
  const double S1 = block_sum_d((double)s1, s_red), S2 = block_sum_d((double)s2, s_red);
  const double mu_s = S1 / (double)m;                         // mean of the shifted values
  const double var = fmax(S2 / (double)m - mu_s * mu_s, 0.0);    // biased variance, as torch
  const float mu = (float)((double)K + mu_s);
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  if (threadIdx.x == 0) {
    mean[ng] = mu;
    rstd[ng] = r;
  }
  // pass 2 (the group is in L2): y = relu((x - mu) * r * gamma_c + beta_c)
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xg);
    float4* y4 = reinterpret_cast<float4*>(yg);
    const int hw4 = HW / 4;
    for (int64_t i = threadIdx.x; i < m / 4; i += kGnThreads) {
      const int c = g * cpg + (int)(i / hw4);
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float a = r * ga, b = fmaf(-mu, a, be);
      const float4 v = x4[i];
      float4 o;
      o.x = fmaxf(fmaf(v.x, a, b), 0.f);
      o.y = fmaxf(fmaf(v.y, a, b), 0.f);
      o.z = fmaxf(fmaf(v.z, a, b), 0.f);
      o.w = fmaxf(fmaf(v.w, a, b), 0.f);
      y4[i] = o;
    }
  } else {
    for (int64_t i = threadIdx.x; i < m; i += kGnThreads) {
      const int c = g * cpg + (int)(i / HW);
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float a = r * ga, b = fmaf(-mu, a, be);
      yg[i] = fmaxf(fmaf(xg[i], a, b), 0.f);
    }
  }
}

// Backward of the above.  With xh = (x - mu) r, z = gamma xh + beta, dz = dy [z > 0]:
//   dgamma_c = sum_{n,hw} dz xh,  dbeta_c = sum_{n,hw} dz,
//   dx = r (gamma_c dz - (xh A + B) / m),  A = sum_c gamma_c sum_hw dz xh,  B = sum_c gamma_c sum_hw dz   (per image, group)
// dgamma_nc / dbeta_nc: [N, C] per-image partial sums (fixed order; the caller adds the N rows).
__global__ void __launch_bounds__(kGnThreads)
gn_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd, int C, int HW, int G,
                   float* __restrict__ dx, float* __restrict__ dgamma_nc, float* __restrict__ dbeta_nc) {
  constexpr int kNW = kGnThreads / 32;
  __shared__ double s_w1[kNW], s_w2[kNW];       // per-warp partial sums of the channel being reduced
  __shared__ double s_c1[64], s_c2[64];         // per-channel sums of the group (cpg <= 64, checked on the host)
  const int ng = blockIdx.x, g = ng % G, n = ng / G;
  const int cpg = C / G;
  const int64_t m = (int64_t)cpg * HW;
  const float* xg = x + (int64_t)ng * m;
  const float* dyg = dy + (int64_t)ng * m;
  float* dxg = dx + (int64_t)ng * m;
  const float mu = mean[ng], r = rstd[ng];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool vec = (HW & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(xg) | reinterpret_cast<uintptr_t>(dyg) | reinterpret_cast<uintptr_t>(dxg)) & 15) == 0;
  // pass 1: per channel, sum dz xh and sum dz.  The warps split into teams of wpc = kNW / min(cpg, kNW) warps per channel
  // (16 warps, 8 channels: two warps stride over each plane), rounds of kNW / wpc channels: one block barrier per round.
  const int par = cpg < kNW ? cpg : kNW;        // channels reduced concurrently
  const int wpc = kNW / par;                    // warps per channel (warps beyond par * wpc idle in pass 1)
  for (int c0 = 0; c0 < cpg; c0 += par) {
    const int cl = c0 + wid / wpc, sub = wid % wpc;
    float a1 = 0.f, a2 = 0.f;
    if (wid < par * wpc && cl < cpg) {
      const int c = g * cpg + cl;
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float* xc = xg + (int64_t)cl * HW;
      const float* dc = dyg + (int64_t)cl * HW;
      if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(xc);
        const float4* d4 = reinterpret_cast<const float4*>(dc);
        for (int i = sub * 32 + lane; i < HW / 4; i += wpc * 32) {
          const float4 xv = x4[i], dv = d4[i];
          const float h0 = (xv.x - mu) * r, h1 = (xv.y - mu) * r, h2 = (xv.z - mu) * r, h3 = (xv.w - mu) * r;
          const float z0 = fmaf(ga, h0, be) > 0.f ? dv.x : 0.f, z1 = fmaf(ga, h1, be) > 0.f ? dv.y : 0.f;
          const float z2 = fmaf(ga, h2, be) > 0.f ? dv.z : 0.f, z3 = fmaf(ga, h3, be) > 0.f ? dv.w : 0.f;
          a1 = fmaf(z0, h0, fmaf(z1, h1, fmaf(z2, h2, fmaf(z3, h3, a1))));
          a2 += (z0 + z1) + (z2 + z3);
        }
      } else {
        for (int i = sub * 32 + lane; i < HW; i += wpc * 32) {
          const float xh = (xc[i] - mu) * r;
          const float dz = fmaf(ga, xh, be) > 0.f ? dc[i] : 0.f;
          a1 = fmaf(dz, xh, a1);
          a2 += dz;
        }
      }
    }
    const double w1 = warp_sum((double)a1), w2 = warp_sum((double)a2);
    __syncthreads();                             // s_w1 / s_w2 of the previous round have been consumed
    if (lane == 0) {
      s_w1[wid] = w1;
      s_w2[wid] = w2;
    }
    __syncthreads();
    if (threadIdx.x < par && c0 + threadIdx.x < cpg) {
      double c1 = 0.0, c2 = 0.0;
      for (int q = 0; q < wpc; ++q) {            // fixed order
        c1 += s_w1[threadIdx.x * wpc + q];
        c2 += s_w2[threadIdx.x * wpc + q];
      }
      const int cl2 = c0 + threadIdx.x, c = g * cpg + cl2;
      s_c1[cl2] = c1;
      s_c2[cl2] = c2;
      if (dgamma_nc) dgamma_nc[(int64_t)n * C + c] = (float)c1;
      if (dbeta_nc) dbeta_nc[(int64_t)n * C + c] = (float)c2;
    }
  }
  __syncthreads();
  double A = 0.0, Bq = 0.0;
  for (int cl = 0; cl < cpg; ++cl) {
    const double ga = gamma ? (double)gamma[g * cpg + cl] : 1.0;
    A += ga * s_c1[cl];
    Bq += ga * s_c2[cl];
  }
  const float Am = (float)(A / (double)m), Bm = (float)(Bq / (double)m);
  // pass 2 (dy and x of the group are in L2)
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(xg);
    const float4* d4 = reinterpret_cast<const float4*>(dyg);
    float4* o4 = reinterpret_cast<float4*>(dxg);
    const int hw4 = HW / 4;
    for (int64_t i = threadIdx.x; i < m / 4; i += kGnThreads) {
      const int c = g * cpg + (int)(i / hw4);
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float4 xv = x4[i], dv = d4[i];
      const float h0 = (xv.x - mu) * r, h1 = (xv.y - mu) * r, h2 = (xv.z - mu) * r, h3 = (xv.w - mu) * r;
      const float z0 = fmaf(ga, h0, be) > 0.f ? dv.x : 0.f, z1 = fmaf(ga, h1, be) > 0.f ? dv.y : 0.f;
      const float z2 = fmaf(ga, h2, be) > 0.f ? dv.z : 0.f, z3 = fmaf(ga, h3, be) > 0.f ? dv.w : 0.f;
      float4 o;
      o.x = r * fmaf(ga, z0, -fmaf(h0, Am, Bm));
      o.y = r * fmaf(ga, z1, -fmaf(h1, Am, Bm));
      o.z = r * fmaf(ga, z2, -fmaf(h2, Am, Bm));
      o.w = r * fmaf(ga, z3, -fmaf(h3, Am, Bm));
      o4[i] = o;
    }
  } else {
    for (int64_t i = threadIdx.x; i < m; i += kGnThreads) {
      const int c = g * cpg + (int)(i / HW);
      const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
      const float xh = (xg[i] - mu) * r;
      const float dz = fmaf(ga, xh, be) > 0.f ? dyg[i] : 0.f;
      dxg[i] = r * (fmaf(ga, dz, -fmaf(xh, Am, Bm)));
    }
  }
}

// y = relu(s * x)   (Scale + ReLU of the regression branch, atss_head.py:141-143 + radet_head.py:29)
__global__ void scale_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ s, int64_t n, float* __restrict__ y) {
  const float sc = s[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaxf(sc * x[i], 0.f);
}

// dx = dy s [s x > 0];  ds_partials[block] = sum over the block's elements of dy x [s x > 0]  (fixed order; the caller adds them)
__global__ void __launch_bounds__(256)
scale_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ s, int64_t n,
                      float* __restrict__ dx, double* __restrict__ ds_partials) {
  __shared__ double s_w[8];
  const float sc = s[0];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float d = sc * xv > 0.f ? dy[i] : 0.f;
    dx[i] = d * sc;
    acc += (double)(d * xv);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    ds_partials[blockIdx.x] = t;
  }
}

}  // namespace radet

using namespace radet;

extern "C" int radet_gn_relu_forward(const float* x, const float* gamma, const float* beta, int32_t n, int32_t c, int32_t hw,
                                     int32_t groups, float eps, float* y, float* mean, float* rstd, void* stream) {
  if (n == 0) return RADET_OK;
  if (!x || !y || !mean || !rstd || n < 0 || c <= 0 || hw <= 0 || groups <= 0 || c % groups != 0 || !(eps >= 0.f)) return RADET_E_BADARG;
  if ((int64_t)n * groups > 0x7fffffffll) return RADET_E_BADARG;
  gn_relu_fwd_kernel<<<(unsigned)(n * groups), kGnThreads, 0, (cudaStream_t)stream>>>(x, gamma, beta, c, hw, groups, eps, y, mean, rstd);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_gn_relu_backward(const float* dy, const float* x, const float* gamma, const float* beta, const float* mean,
                                      const float* rstd, int32_t n, int32_t c, int32_t hw, int32_t groups, float* dx,
                                      float* dgamma_nc, float* dbeta_nc, void* stream) {
  if (n == 0) return RADET_OK;
  if (!dy || !x || !mean || !rstd || !dx || n < 0 || c <= 0 || hw <= 0 || groups <= 0 || c % groups != 0) return RADET_E_BADARG;
  if (c / groups > 64) return RADET_E_UNSUPPORTED;
  if ((int64_t)n * groups > 0x7fffffffll) return RADET_E_BADARG;
  gn_relu_bwd_kernel<<<(unsigned)(n * groups), kGnThreads, 0, (cudaStream_t)stream>>>(dy, x, gamma, beta, mean, rstd, c, hw, groups, dx,
                                                                                     dgamma_nc, dbeta_nc);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

static unsigned scale_relu_blocks(int64_t n) {
  const int64_t b = (n + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 4 * 148 ? 4 * 148 : b));
}

extern "C" int32_t radet_scale_relu_partials(int64_t n) { return n <= 0 ? 0 : (int32_t)scale_relu_blocks(n); }

extern "C" int radet_scale_relu_forward(const float* x, const float* scale, int64_t n, float* y, void* stream) {
  if (n == 0) return RADET_OK;
  if (!x || !scale || !y || n < 0) return RADET_E_BADARG;
  scale_relu_fwd_kernel<<<scale_relu_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, scale, n, y);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}

extern "C" int radet_scale_relu_backward(const float* dy, const float* x, const float* scale, int64_t n, float* dx, double* dscale_partials,
                                         void* stream) {
  if (n == 0) return RADET_OK;
  if (!dy || !x || !scale || !dx || !dscale_partials || n < 0) return RADET_E_BADARG;
  scale_relu_bwd_kernel<<<scale_relu_blocks(n), 256, 0, (cudaStream_t)stream>>>(dy, x, scale, n, dx, dscale_partials);
  RADET_LAUNCH_CHECK();
  return RADET_OK;
}
