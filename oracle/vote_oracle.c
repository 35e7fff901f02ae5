/* TEST INFRASTRUCTURE — plain-C restatement (oracle) of RADet's CPU vote-NMS.
 *
 * Follows /root/reference/radet/ops/vote/vote_ext.cpp:8-35 (vote_single_dim),
 * :70-207 (vote_nms), :210-353 (global_vote_nms) and
 * /root/reference/radet/ops/cluster/cluster_ext.cpp:4-87 (cluster ids / sizes).
 * All arithmetic is IEEE binary32, round-to-nearest, one rounding per operation:
 * compile with -ffp-contract=off (no FMA), as the reference build (plain g++ -O)
 * does on x86-64 where float_t == float.
 *
 * The sort (torch::sort descending, vote_ext.cpp:78) is done by the caller and
 * passed in as `order` so that tie-breaking is explicit.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
 * may load this library; it is never on the product path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static float vote_axis(const float *s, const float *x, int64_t n) {
  float ssum = 0.f, acc = 0.f;
  for (int64_t i = 0; i < n; ++i) {
    ssum += s[i];
    acc += s[i] * x[i];
  }
  const float mean = acc / ssum;
  float var = 0.f;
  for (int64_t i = 0; i < n; ++i) var += s[i] * (x[i] - mean) * (x[i] - mean);
  const float sd = (float)sqrt(var / ssum);
  float fs = 0.f, fx = 0.f;
  for (int64_t i = 0; i < n; ++i) {
    if ((mean - sd <= x[i]) & (x[i] <= mean + sd)) {
      fx += s[i] * x[i];
      fs += s[i];
    }
  }
  return fx / fs;
}

/* returns number of clusters k; out_* have capacity n */
int64_t oracle_vote_nms(const float *boxes, const float *cluster_scores, const float *vote_scores,
                        const int64_t *labels, const int64_t *order, int64_t n, float thr, int global_mode,
                        int iou_enable, float sigma, float *out_boxes, int64_t *out_labels, float *out_scores,
                        int64_t *instance_ids, int64_t *clusters_num) {
  if (n <= 0) return 0;
  unsigned char *gone = (unsigned char *)calloc((size_t)n, 1);
  float *ms = (float *)malloc(sizeof(float) * (size_t)n * 5); /* member vote score + 4 coords, SoA */
  float *mx[4] = {ms + n, ms + 2 * n, ms + 3 * n, ms + 4 * n};
  int64_t *seen_labels = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  int64_t n_seen = 0, k = 0;
  for (int64_t i = 0; i < n; ++i) {
    instance_ids[i] = 0;
    clusters_num[i] = 0;
  }
  for (int64_t a = 0; a < n; ++a) {
    const int64_t i = order[a];
    if (gone[i]) continue;
    const int64_t li = labels[i];
    if (global_mode) {
      int dup = 0;
      for (int64_t q = 0; q < n_seen; ++q) dup |= (seen_labels[q] == li);
      if (dup) {
        gone[i] = 1;
        continue;
      }
      seen_labels[n_seen++] = li;
    }
    const float *bi = boxes + 4 * i;
    const float area_i = (bi[2] - bi[0]) * (bi[3] - bi[1]);
    gone[i] = 1;
    int64_t m = 0;
    float best = cluster_scores[i];
    ms[m] = vote_scores[i];
    for (int d = 0; d < 4; ++d) mx[d][m] = bi[d];
    ++m;
    instance_ids[i] = k;
    for (int64_t b = a + 1; b < n; ++b) {
      const int64_t j = order[b];
      if (labels[j] != li || gone[j]) continue;
      const float *bj = boxes + 4 * j;
      const float xl = fmaxf(bj[0], bi[0]), yt = fmaxf(bj[1], bi[1]);
      const float xr = fminf(bj[2], bi[2]), yb = fminf(bj[3], bi[3]);
      const float iw = fmaxf(0.f, xr - xl), ih = fmaxf(0.f, yb - yt);
      const float inter = iw * ih;
      const float area_j = (bj[2] - bj[0]) * (bj[3] - bj[1]);
      const float iou = inter / (area_j + area_i - inter);
      float vj = vote_scores[j];
      if (iou_enable) { /* vote_ext.cpp:164-167: with <torch/extension.h> included first, the unqualified exp() on a
                           float resolves to the float overload (verified against the compiled reference, see
                           tests/test_oracle_golden.py::test_c_oracle_equals_the_compiled_reference_ops) */
        const float fac = expf(-(1 - iou) * (1 - iou) / sigma);
        vj = vj * fac;
      }
      if (iou > thr) {
        gone[j] = 1;
        instance_ids[j] = k;
        ms[m] = vj;
        for (int d = 0; d < 4; ++d) mx[d][m] = bj[d];
        if (cluster_scores[j] > best) best = cluster_scores[j];
        ++m;
      }
    }
    for (int d = 0; d < 4; ++d) out_boxes[4 * k + d] = vote_axis(ms, mx[d], m);
    out_labels[k] = li;
    out_scores[k] = best;
    clusters_num[i] = m;
    ++k;
  }
  free(gone);
  free(ms);
  free(seen_labels);
  return k;
}
