/* radet_b200 — C ABI of the B200-native (sm_100a) RADet dense-head hot path.
 *
 * Scope: visibility-guided sample assignment -> target encode -> fused head loss
 * (forward + backward) -> score-threshold / top-k / TBLR decode -> class-aware
 * vote-NMS.  One shared library (libradet_b200.so), plain pointers and sizes, no
 * torch / pybind types.  Every entry point
 *   - takes DEVICE pointers unless the parameter name ends in `_host`,
 *   - never allocates, never synchronises, only enqueues on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream),
 *   - returns 0 on success, a negative RADET_E_* for argument errors detected on
 *     the host, or a positive cudaError_t from the launch,
 *   - is re-entrant and thread-safe (no global state).
 * Buffers (inputs, outputs, workspace) are owned by the caller; workspace sizes
 * come from the matching *_workspace_bytes() query and need 256-byte alignment.
 *
 * The reference (YangHai-1218/RADet) has no C ABI: its native boundary is three
 * pybind11/libtorch CPU extensions plus Python.  Each function below names the
 * reference interface it replaces (paths relative to the reference root).
 */
#ifndef RADET_B200_H_
#define RADET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADET_MAX_LEVELS 8
#define RADET_MAX_GT_PER_IMAGE 256   /* per image; masks are bit sets of 32-bit words */
#define RADET_MAX_POSITIVE_NUM 32    /* LabelAssignment.positive_num (config: 10) */
/* bits of radet_assign's `balance_sample` argument (a plain 0 / 1 keeps its original meaning) */
#define RADET_ASSIGN_BALANCE 1        /* LabelAssignment.balance_sample (label_assignment.py:109-116) */
#define RADET_ASSIGN_ADAPT_K 2        /* adapt_positive_num: per-GT positive_num = adapt_cal_k(...) (:88-95, :104-105) */
#define RADET_ASSIGN_WEIGHT_BY_PRO 4  /* multiply_samplepro_for_weight: weight *= sample probability (:127-128) */
#define RADET_MT_STATE_WORDS 625     /* 624 key words + position, numpy legacy MT19937 */

enum {
  RADET_OK = 0,
  RADET_E_BADARG = -1,      /* null pointer / non-positive size / inconsistent grid */
  RADET_E_TOO_MANY_GT = -2, /* more than RADET_MAX_GT_PER_IMAGE boxes in one image */
  RADET_E_WORKSPACE = -3,   /* workspace too small or misaligned */
  RADET_E_UNSUPPORTED = -4  /* option outside the implemented surface */
};

/* Prior grid of the head: what AnchorGenerator(ratios=[1], octave_base_scale, scales_per_octave=1,
 * strides) + LabelAssignment.regress_ranges describe (core/anchor/anchor_generator.py:206-271,
 * datasets/pipelines/label_assignment.py:30-52,136-148).  Level l has level_h[l] x level_w[l] cells;
 * cell (y,x) has centre (x*stride, y*stride) and a square prior of side anchor_scale*stride.
 * Point order everywhere: level-major, then row-major (x fastest). */
typedef struct {
  int32_t num_levels;
  int32_t level_h[RADET_MAX_LEVELS];
  int32_t level_w[RADET_MAX_LEVELS];
  int32_t stride[RADET_MAX_LEVELS];
  float range_lo[RADET_MAX_LEVELS]; /* regress range, inclusive on both ends */
  float range_hi[RADET_MAX_LEVELS];
  float anchor_scale;    /* octave_base_scale (8): prior side = anchor_scale * stride */
  float tblr_normalizer; /* TBLRBBoxCoder.normalizer (1/8), core/bbox/coder/tblr_bbox_coder.py:25-27 */
} radet_grid_t;

/* NCHW float32 head outputs, one pointer per level (radet_head.py:27-30):
 * cls[l]: [B, C, h_l, w_l] logits; bbox[l]: [B, 4, h_l, w_l] (T,B,L,R, post-ReLU); iou[l]: [B, 1, h_l, w_l]. */
typedef struct {
  const float* cls[RADET_MAX_LEVELS];
  const float* bbox[RADET_MAX_LEVELS];
  const float* iou[RADET_MAX_LEVELS];
} radet_maps_t;

typedef struct {
  float* cls[RADET_MAX_LEVELS];
  float* bbox[RADET_MAX_LEVELS];
  float* iou[RADET_MAX_LEVELS];
} radet_grad_maps_t;

/* ------------------------------------------------------------------------------------------------
 * Library / build info. */
const char* radet_version(void);
/* number of kernel launches enqueued by this library in this process so far (bench.py's gpu_launches) */
uint64_t radet_launch_count(void);
int64_t radet_num_points(const radet_grid_t* grid);
/* Measurement aid (bench.py): holds `stream` until *flag != 0 (flag: int32 in pinned, device-mapped host memory; the
 * host opens the gate with a plain store) or until max_wait_ns have passed.  Work enqueued behind it starts the moment
 * the gate opens, so host launch latency stays outside a CUDA-event timing window.  Not counted by radet_launch_count. */
int radet_stream_gate(const int32_t* flag, int64_t max_wait_ns, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Visible-mask hand-off (replaces the [G,H,W] uint8 `distance_maps.to_ndarray()` input of
 * LabelAssignment.__call__, label_assignment.py:151-152, and cal_sample_pro's gather :78-86).
 * The assignment only ever reads pixel (y*step, x*step) with step = gcd(strides); this packs exactly
 * those samples, 1 bit each (mask != 0), into bits[g][gy][gx/32] (row pitch = ceil(grid_w/32) words).
 *   src: uint8 [num_gt, src_h, src_w]; sample (gy,gx) is src[g][gy*step][gx*step]
 *        (pass the full-resolution masks with step=8, or the pre-sampled grid with step=1).
 *   status (optional, int32[1]): bit 0 is OR-ed in when a mask holds two different non-zero sample
 *        values (real-valued distance maps are outside the implemented surface). */
int radet_pack_masks(const uint8_t* src, int64_t num_gt, int32_t src_h, int32_t src_w, int32_t step,
                     int32_t grid_h, int32_t grid_w, uint32_t* bits, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Numpy-legacy MT19937 on the device: out[b][0..n) = what `np.random.seed(seeds[b]);
 * np.random.random_sample(n)` returns (53-bit doubles).  Used by tests and by the seeded mode below. */
int radet_mt19937_uniforms(const uint32_t* seeds, int32_t batch, int32_t n, double* out, void* stream);
/* mt_states[b] (625 words: key[624], pos) = the legacy state right after `np.random.seed(seeds[b])` (first block
 * already regenerated, pos = 0).  The 624-step seeding recurrence is sequential but depends on nothing: enqueue it
 * early / on a side stream and hand the states to radet_assign(mt_states=...). */
int radet_mt19937_seed(const uint32_t* seeds, int32_t batch, uint32_t* mt_states, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Visibility-guided sample assignment for a batch of images.
 * Replaces LabelAssignment.__call__ (label_assignment.py:136-201: generate_candidate_cell :57-76,
 * cal_sample_pro :78-86, area-sorted min-area claiming :156-197, random_sample :97-131) for the
 * configuration surface ambiguous_sample='min_area', random_sample_by_distance=True,
 * adapt_positive_num=False, multiply_samplepro_for_weight=False, binary masks.
 *
 *   gt_offsets   int32 [B+1]   image b owns GT rows gt_offsets[b] .. gt_offsets[b+1]
 *   gt_bboxes    f32   [Gtot,4] x1,y1,x2,y2
 *   mask_bits    u32   [Gtot, mask_h, ceil(mask_w/32)] from radet_pack_masks (mask_step = sample pitch in px)
 *   RNG (numpy global-state compatible), exactly one of:
 *     uniforms   f64   [B, n_uniform]  pre-drawn random_sample() stream per image, or
 *     seeds      u32   [B]             np.random.seed(seeds[b]) right before image b, or
 *     mt_states  u32   [B, 625]        full legacy state (key[624], pos); updated in place to the
 *                                      state after the call (so the host RNG can be advanced exactly)
 *   positive_num, balance_sample: LabelAssignment(positive_num, balance_sample); balance_sample also carries the
 *                      RADET_ASSIGN_* option bits.  With RADET_ASSIGN_ADAPT_K the per-GT draw count is
 *                      int(positive_num * sum_l ratio_l * exp((max(w,h) - size_l) / (2 size_l)) + 0.5) over the levels of the
 *                      GT's remaining candidates (float32 exp as numpy's); it must stay <= 32.
 * Outputs:
 *   points_to_gt_index int64 [B,P]  1-based GT index; -1 negative; 0 ignore   (label_assignment.py:198)
 *   points_weight      f32   [B,P]                                           (label_assignment.py:199)
 *   consumed           int32 [B]    doubles drawn from the stream; -1 if `uniforms` ran out, -2 if an adaptive
 *                                   positive_num exceeded 32 (outputs invalid in both cases)
 *   weight_sums        f64   [B]    optional (NULL = skip): sum of points_weight over the points with index >= 0 of each
 *                                   image (0 for an image without ground truth) -- radet_loss_cfg_t.weight_sums
 * workspace: radet_assign_workspace_bytes(). */
size_t radet_assign_workspace_bytes(const radet_grid_t* grid, int32_t batch);
int radet_assign(const radet_grid_t* grid, int32_t batch, const int32_t* gt_offsets, const int32_t* gt_offsets_host,
                 const float* gt_bboxes, const uint32_t* mask_bits, int32_t mask_h, int32_t mask_w, int32_t mask_step,
                 const double* uniforms, int32_t n_uniform, const uint32_t* seeds, uint32_t* mt_states,
                 int32_t positive_num, int32_t balance_sample, int64_t* points_to_gt_index, float* points_weight,
                 int32_t* consumed, double* weight_sums, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AnchorGenerator.grid_anchors / valid_flags for ONE image (core/anchor/anchor_generator.py:206-271, 273-298;
 * ratios=[1], one square prior of side anchor_scale * stride centred on (x*s, y*s) per cell).  The kernels of the hot
 * path compute priors in closed form; this entry serves AnchorHead.get_anchors (anchor_head.py:142-170).
 *   anchors     f32 [P,4]   optional   x1,y1,x2,y2, level-major / row-major
 *   valid_flags u8  [P]     optional   1 for cells inside ceil(pad_h / stride) x ceil(pad_w / stride) */
int radet_grid_priors(const radet_grid_t* grid, int32_t pad_h, int32_t pad_w, float* anchors, uint8_t* valid_flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Target gather + TBLR encode.  Replaces RADetHead.get_targets / _get_target_single
 * (models/dense_heads/radet_head.py:290-369, 373-392) and TBLRBBoxCoder.encode
 * (core/bbox/coder/tblr_bbox_coder.py:29-46, 71-114): target = (distance / prior side) / normalizer.
 * Outputs are the reference's per-level concatenation, level-major / image-minor: level l starts at
 * row B*sum_{k<l} h_k*w_k and holds B blocks of h_l*w_l rows.  anchors may be NULL.
 *   labels int64 [B*P] (num_classes = background; ignored points take the LAST GT's label, radet_head.py:390)
 *   bbox_targets f32 [B*P,4] (T,B,L,R)/stride;  weights f32 [B*P];  anchors f32 [B*P,4]. */
int radet_get_targets(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const int32_t* gt_offsets,
                      const float* gt_bboxes, const int64_t* gt_labels, const int64_t* points_to_gt_index,
                      const float* points_weight, int64_t* labels, float* bbox_targets, float* weights,
                      float* anchors, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused head loss, forward + backward.  Replaces RADetHead.loss (radet_head.py:173-288) with
 * FocalLoss (models/losses/focal_loss.py:44-87,91-157 -> mmcv.ops.sigmoid_focal_loss),
 * TBLRBBoxCoder.decode (tblr_bbox_coder.py:117-172), bbox_overlaps aligned iou/giou
 * (core/bbox/iou_calculators/iou2d_calculator.py:107-159), GIoULoss (models/losses/iou_loss.py:82-98,
 * 319-354), CrossEntropyLoss(use_sigmoid) (models/losses/cross_entropy_loss.py:58-91) and
 * weight_reduce_loss (models/losses/utils.py:26-52), plus the autograd backward of all of them.
 * Reads the NCHW maps in place (no permute/cat), builds targets on the fly from the assignment.
 *   losses f32[4]: loss_cls, loss_bbox, loss_iou, num_pos  (num_pos is rank-local, radet_head.py:254)
 *   grads (optional, may be NULL): d(loss_cls+loss_bbox+loss_iou)/d(map), same NCHW layout, every
 *        element written.  If grad_scale (f32[3], device, optional) is given, the three terms are
 *        scaled by it (upstream gradients of the three returned losses).
 *   avg_extra: added to num_pos in loss_cls's avg_factor (the reference passes num_imgs, :259).
 *   workspace: radet_loss_workspace_bytes(); its first 256 bytes must be zero before the FIRST call
 *        (the kernels re-arm their completion counters, so one cudaMemset at allocation time is enough). */
typedef struct {
  float gamma, alpha;                 /* FocalLoss */
  float w_cls, w_bbox, w_iou;         /* loss_weight of the three terms (1, 2, 1) */
  float eps;                          /* GIoULoss.eps (1e-6); iou target uses bbox_overlaps default 1e-6 */
  float avg_extra;                    /* num_imgs */
  /* Optional (NULL = compute it): per-image sum of points_weight over the points with points_to_gt_index >= 0 in images
   * that have ground truth, double[batch] on the device -- what radet_assign(weight_sums=...) writes next to the
   * assignment.  num_pos (radet_head.py:254) is the sum of these; with it the dense pass does not have to wait for the
   * reduction over the index / weight arrays: the dense kernel is launched as a programmatic dependent of the sparse one and
   * streams its class planes next to it.  Must belong to the arrays passed: the kernel compares the sum with its own
   * reduction at the end and writes NaN into the three losses when they disagree (relative 1e-5).  Ignored unless
   * phases == 3 and every level's h*w is a multiple of 4. */
  const double* weight_sums;
} radet_loss_cfg_t;
/* phases: RADET_LOSS_PHASE_NORMALIZERS computes num_pos / sum(wq) into workspace doubles [0] and [1];
 * RADET_LOSS_PHASE_DENSE consumes them.  Pass both (3) for the reference behaviour (rank-local normalisers,
 * radet_head.py:254).  A caller that wants the FCOS/ATSS-style reduce_mean (atss_head.py:278,296; opt-in, NOT what
 * RADetHead does) runs phase 1, all-reduces those two doubles over NCCL, then runs phase 2. */
enum { RADET_LOSS_PHASE_NORMALIZERS = 1, RADET_LOSS_PHASE_DENSE = 2, RADET_LOSS_PHASE_ALL = 3 };
size_t radet_loss_workspace_bytes(const radet_grid_t* grid, int32_t batch, int32_t num_classes);
int radet_loss_fwd_bwd(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_maps_t* maps,
                       const int32_t* gt_offsets, const float* gt_bboxes, const int64_t* gt_labels,
                       const int64_t* points_to_gt_index, const float* points_weight, const radet_loss_cfg_t* cfg,
                       const float* grad_scale, const radet_grad_maps_t* grads, float* losses, int32_t phases,
                       void* workspace, size_t workspace_bytes, void* stream);
/* grads *= upstream (f32[3] on device: d/dloss_cls, d/dloss_bbox, d/dloss_iou); exits early when all are 1. */
int radet_scale_grads(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_grad_maps_t* grads,
                      const float* upstream, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Standalone TBLR coder on explicit prior lists [n,4] (x1,y1,x2,y2), normalize_by_wh=True.
 * Replaces TBLRBBoxCoder.encode / .decode (core/bbox/coder/tblr_bbox_coder.py:29-68 -> bboxes2tblr :71-114,
 * tblr2bboxes :117-172).  clip != 0 clamps x to [0,max_w], y to [0,max_h] (:167-171). */
int radet_tblr_encode(const float* priors, const float* gt_bboxes, int64_t n, float normalizer, float* out, void* stream);
int radet_tblr_decode(const float* priors, const float* tblr, int64_t n, float normalizer, int32_t clip, float max_h,
                      float max_w, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Standalone element-wise losses (the LOSSES registry entries called outside the fused head).  Each call writes the
 * un-reduced, un-weighted loss and, when dloss_dpred != NULL, its derivative w.r.t. pred in the same launch; weighting
 * and reduction (losses/utils.py:26-52 weight_reduce_loss) stay with the caller.
 *   radet_sigmoid_focal_loss  FocalLoss -> sigmoid_focal_loss (models/losses/focal_loss.py:44-87): mmcv.ops
 *       sigmoid_focal_loss(pred [n,C], target int64 [n], gamma, alpha, None, 'none'); a target outside [0,C) is an
 *       all-negative row.  loss / dloss_dpred: [n,C].
 *   radet_giou_loss           GIoULoss -> giou_loss (models/losses/iou_loss.py:82-98): 1 - bbox_overlaps(pred, target,
 *       mode='giou', is_aligned=True, eps) on (x1,y1,x2,y2) rows.  loss [n], dloss_dpred [n,4].
 *   radet_bce_with_logits     CrossEntropyLoss(use_sigmoid=True) -> binary_cross_entropy
 *       (models/losses/cross_entropy_loss.py:58-91) with pred.dim() == label.dim():
 *       F.binary_cross_entropy_with_logits(pred, label.float(), reduction='none').  loss / dloss_dpred: [n]. */
int radet_sigmoid_focal_loss(const float* pred, const int64_t* target, int64_t n, int32_t num_classes, float gamma, float alpha,
                             float* loss, float* dloss_dpred, void* stream);
int radet_giou_loss(const float* pred, const float* target, int64_t n, float eps, float* loss, float* dloss_dpred, void* stream);
int radet_bce_with_logits(const float* pred, const float* target, int64_t n, float* loss, float* dloss_dpred, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Vote-NMS family on explicit box lists.  Replaces the pybind modules
 *   vote_ext.vote_nms / vote_ext.global_vote_nms (ops/vote/vote_ext.cpp:70-207, 210-353, 358-361) and
 *   cluster_ext.cluster_nms (ops/cluster/cluster_ext.cpp:4-87, 90-92),
 * batched over `batch` independent lists (list i owns rows offsets[i] .. offsets[i+1]).
 * Sort: descending cluster score, ties by lower row index (torch::sort observed stable).
 * IEEE fp32, no FMA contraction, member sums in descending-score order: bit-exact with the reference.
 *   mode: 0 vote_nms, 1 global_vote_nms, 2 plain class-aware NMS (keep = seeds, boxes not voted)
 * Outputs per list i (rows offsets[i].. of the out_* arrays, up to max_out[i] = list length or max_num):
 *   out_dets f32 [.,5] (x1,y1,x2,y2,score), out_labels int64, out_index int64 (row of the seed box),
 *   num_out int32 [batch].  Optional: instance_ids int64 [n], clusters_num int64 [n] (cluster_ext layout). */
enum { RADET_NMS_VOTE = 0, RADET_NMS_GLOBAL_VOTE = 1, RADET_NMS_PLAIN = 2 };
size_t radet_vote_nms_workspace_bytes(int32_t batch, int64_t total_boxes, int64_t max_boxes_per_list);
int radet_vote_nms(int32_t batch, const int32_t* offsets_host, const float* boxes, const float* cluster_scores,
                   const float* vote_scores, const int64_t* labels, float iou_threshold, int32_t iou_enable, float sigma,
                   int32_t mode, int32_t max_num, float* out_dets, int64_t* out_labels, int64_t* out_index,
                   int32_t* num_out, int64_t* instance_ids, int64_t* clusters_num, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Inference: score threshold + per-level top-k + TBLR decode + class-aware (vote-)NMS for a batch.
 * Replaces ATSSHead.get_bboxes (models/dense_heads/atss_head.py:326-387) + RADetHead._get_bboxes_single
 * (radet_head.py:55-169) + radet.ops.vote_nms / global_vote_nms wrappers (ops/vote/vote_wrapper.py:7-83)
 * + the batched_nms branch (radet_head.py:159-163).
 *   img_shapes   int32 [B,2]  (h, w) used for the decode clamp (radet_head.py:130-131)
 *   scale_factors f32  [B,4]  boxes /= scale_factor when rescale != 0 (radet_head.py:141-143)
 *   score mode: how cluster / vote scores are formed from cls score S and centerness c
 *               (vote_wrapper.py:14-30): 0 = S*c (list/tuple), 1 = S ('cls'), 2 = c ('iou')
 * Outputs: dets f32 [B, max_per_img, 5], labels int64 [B, max_per_img], num_dets int32 [B].
 * workspace: radet_get_bboxes_workspace_bytes(); it must be zero-filled once before the FIRST call with a given
 *        (grid, batch, classes, nms_pre) — the kernels re-arm their counters afterwards. */
typedef struct {
  float score_thr;
  int32_t nms_pre;        /* <=0: no per-level limit */
  int32_t max_per_img;    /* >0 */
  int32_t nms_mode;       /* RADET_NMS_* */
  float iou_threshold;
  int32_t cluster_score_mode, vote_score_mode;
  int32_t iou_enable;
  float sigma;
  int32_t rescale;
} radet_detect_cfg_t;
size_t radet_get_bboxes_workspace_bytes(const radet_grid_t* grid, int32_t batch, int32_t num_classes,
                                        const radet_detect_cfg_t* cfg);
int radet_get_bboxes(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_maps_t* maps,
                     const int32_t* img_shapes, const float* scale_factors, const radet_detect_cfg_t* cfg,
                     float* dets, int64_t* labels, int32_t* num_dets, void* workspace, size_t workspace_bytes,
                     void* stream);

/* with_nms=False branch of _get_bboxes_single (radet_head.py:165-169): every candidate that passed score_thr and the
 * per-level top-k, un-suppressed.  rows f32 [batch][cap][9] = decoded box (clamped to img_shape, divided by
 * scale_factor when cfg->rescale), score * centerness, prior box (same rescale); labels int64 [batch][cap];
 * num_rows int32 [batch].  cap = radet_candidates_capacity(); rows of one image are ordered by class, then by
 * descending score * centerness (the reference's order inside a level is topk(sorted=False)'s).
 * Workspace: radet_get_bboxes_workspace_bytes().  nms_* fields of cfg are ignored. */
int64_t radet_candidates_capacity(const radet_grid_t* grid, int32_t num_classes, int32_t nms_pre);
int radet_get_candidates(const radet_grid_t* grid, int32_t batch, int32_t num_classes, const radet_maps_t* maps,
                         const int32_t* img_shapes, const float* scale_factors, const radet_detect_cfg_t* cfg, float* rows,
                         int64_t* labels, int32_t* num_rows, void* workspace, size_t workspace_bytes, void* stream);

/* Post-NMS formatting, batched on the device.  Replaces bbox2result (core/bbox/transforms.py:99-116): per image the
 * detections regrouped by class, original (score) order kept inside a class.  out f32 [batch][max_rows][5] holds the
 * class-sorted rows, class_offsets int32 [batch][num_classes+1] the row range of each class (rows with a label
 * outside [0,num_classes) are dropped, as `labels == i` never selects them).  xywh != 0 additionally rewrites the box
 * as (x1, y1, x2-x1, y2-y1) = BOPDataset.xyxy2xywh used by _bop_det2json (datasets/bop.py:99-118).  max_rows <= 1024. */
int radet_bbox2result(const float* dets, const int64_t* labels, const int32_t* num, int32_t batch, int32_t max_rows,
                      int32_t num_classes, int32_t xywh, float* out, int32_t* class_offsets, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Head-tower epilogues: what sits between the head's cuDNN convolutions and the path above.
 *
 * radet_gn_relu_forward / _backward replace GroupNorm(groups) + ReLU of every tower layer, ConvModule(conv3x3,
 * norm_cfg=GN(32), act=ReLU) (models/dense_heads/atss_head.py:52-87, forward_single :133-138): statistics, normalise +
 * affine and the ReLU in one launch each way (torch: F.group_norm + F.relu and their autograd backwards).
 *   x, y, dy, dx   f32 [n, c, hw]  NCHW planes, contiguous
 *   gamma, beta    f32 [c]         GroupNorm.weight / .bias (NULL = 1 / 0)
 *   mean, rstd     f32 [n, groups] written by forward (biased variance, eps inside the root), read by backward
 *   dgamma_nc, dbeta_nc  f32 [n, c]  optional: per-image sums; GroupNorm.weight.grad / .bias.grad are their sums over n
 * backward recomputes the ReLU mask from x (y is not read); c / groups <= 64.
 *
 * radet_scale_relu_forward / _backward replace relu(Scale(x)) on the regression branch (atss_head.py:141-143 mmcv Scale,
 * radet_head.py:27-30 F.relu): y = max(scale[0] * x, 0); dx = dy * scale * [scale x > 0]; dscale_partials f64
 * [radet_scale_relu_partials(n)] = fixed-order partial sums of dy * x * [scale x > 0] (Scale.scale.grad is their sum).
 *
 * The sigmoid / score-threshold epilogue of the classification branch needs no entry: radet_get_bboxes consumes raw logits. */
int radet_gn_relu_forward(const float* x, const float* gamma, const float* beta, int32_t n, int32_t c, int32_t hw, int32_t groups,
                          float eps, float* y, float* mean, float* rstd, void* stream);
int radet_gn_relu_backward(const float* dy, const float* x, const float* gamma, const float* beta, const float* mean,
                           const float* rstd, int32_t n, int32_t c, int32_t hw, int32_t groups, float* dx, float* dgamma_nc,
                           float* dbeta_nc, void* stream);
int32_t radet_scale_relu_partials(int64_t n);
int radet_scale_relu_forward(const float* x, const float* scale, int64_t n, float* y, void* stream);
int radet_scale_relu_backward(const float* dy, const float* x, const float* scale, int64_t n, float* dx, double* dscale_partials,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADET_B200_H_ */
