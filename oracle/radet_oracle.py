"""TEST INFRASTRUCTURE — CPU restatement (oracle) of RADet's dense-head hot path.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product (``radet_b200``) never does and
fails loudly when its CUDA library is missing.

Parity pinning: the reference holds NO golden vectors or tests for this path
(SURVEY.md §4).  The oracle is therefore pinned against outputs of the
reference itself, executed in the build container through ``oracle/ref_shim.py``
and committed as fixtures under ``tests/golden/`` (generator:
``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` replays them.
The two mmcv 1.3.18 ops on the path (``sigmoid_focal_loss``, ``batched_nms``)
are not in the reference tree; they are restated from the reference's own
``py_sigmoid_focal_loss`` and from SURVEY.md Appendix B ("parity unpinned" for
those two third-party kernels: tolerance parity only).

Every function cites the reference file:line it follows (paths relative to
``/root/reference``).  numpy for integer/index work, plain-C (``vote_oracle.c``)
for the O(n^2) NMS loops, torch-CPU autograd only for the floating-point loss.
"""
import ctypes
import math
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

STRIDES = (8, 16, 32, 64, 128)
REGRESS_RANGES = ((-1.0, 64.0), (64.0, 128.0), (128.0, 256.0), (256.0, 512.0), (512.0, 1e8))
EPS_PRO = 1e-8   # label_assignment.py:13


# --------------------------------------------------------------------------- priors
def level_shapes(H, W, strides=STRIDES):
    # label_assignment.py:138
    return [(math.ceil(H / s), math.ceil(W / s)) for s in strides]


def grid_points(H, W, strides=STRIDES):
    """Centres of the square priors, level-major then row-major (x fastest).

    anchor_generator.py:142-185 (base anchor centred at 0 because center_offset=0),
    :232-271 (shift = (x*stride, y*stride)).  Returns cx, cy (f32), stride, level id.
    """
    cxs, cys, ss, ls = [], [], [], []
    for lvl, ((h, w), s) in enumerate(zip(level_shapes(H, W, strides), strides)):
        ys, xs = np.divmod(np.arange(h * w), w)
        cxs.append((xs * s).astype(np.float32))
        cys.append((ys * s).astype(np.float32))
        ss.append(np.full(h * w, s, np.float32))
        ls.append(np.full(h * w, lvl, np.int64))
    return np.concatenate(cxs), np.concatenate(cys), np.concatenate(ss), np.concatenate(ls)


def grid_anchors(H, W, strides=STRIDES, octave_base_scale=8):
    """[P,4] f32 square anchors of side octave_base_scale*stride (anchor_generator.py:206-271)."""
    cx, cy, s, _ = grid_points(H, W, strides)
    half = np.float32(0.5) * (s * np.float32(octave_base_scale))
    return np.stack([cx - half, cy - half, cx + half, cy + half], 1).astype(np.float32)


# --------------------------------------------------------------------------- assignment
class UniformStream:
    """Sequential consumer of pre-drawn doubles (what MT19937 ``random_sample`` would hand out)."""

    def __init__(self, u: np.ndarray):
        self.u = np.asarray(u, np.float64)
        self.pos = 0

    def random_sample(self, n):
        if self.pos + n > self.u.size:
            raise OverflowError("uniform stream exhausted")
        out = self.u[self.pos:self.pos + n]
        self.pos += n
        return out


class GlobalNumpyStream:
    """Draws from numpy's global legacy RNG exactly as np.random.choice does (label_assignment.py:112,119)."""

    def __init__(self):
        self.pos = 0

    def random_sample(self, n):
        self.pos += n
        return np.random.random_sample(n)


def legacy_choice(stream, p_f32: np.ndarray, size: int, replace: bool) -> np.ndarray:
    """numpy legacy ``RandomState.choice(a=n, size, p, replace)`` (numpy/random/mtrand.pyx, `choice`).

    Call sites: label_assignment.py:112 (replace=True) and :119 (replace=False).
    """
    p = np.array(p_f32, dtype=np.float64)
    if replace:
        cdf = p.cumsum()
        cdf /= cdf[-1]
        x = stream.random_sample(size)
        return cdf.searchsorted(x, side="right").astype(np.int64)
    n_uniq = 0
    found = np.zeros(size, np.int64)
    while n_uniq < size:
        x = stream.random_sample(size - n_uniq)
        if n_uniq > 0:
            p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, unique_indices = np.unique(new, return_index=True)
        unique_indices.sort()
        new = new.take(unique_indices)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found


def candidate_flags(gt_bboxes: np.ndarray, H: int, W: int, strides=STRIDES, regress_ranges=REGRESS_RANGES):
    """[P,G] bool: inside-box (min side > 0.01) AND level range inclusive (label_assignment.py:57-76)."""
    cx, cy, _, lvl = grid_points(H, W, strides)
    rr = np.asarray(regress_ranges, np.float32)[lvl]               # [P,2]
    g = np.asarray(gt_bboxes, np.float32)
    left = cx[:, None] - g[None, :, 0]
    right = g[None, :, 2] - cx[:, None]
    top = cy[:, None] - g[None, :, 1]
    bottom = g[None, :, 3] - cy[:, None]
    t = np.stack([left, top, right, bottom], -1)
    mn = t.min(-1)
    mx = t.max(-1)
    return (mn > np.float32(0.01)) & (mx >= rr[:, None, 0]) & (mx <= rr[:, None, 1])


def sample_cells(masks: np.ndarray, H: int, W: int, strides=STRIDES, grid_step: Optional[int] = None):
    """[P,G] f32 mask value at (int(cy), int(cx)) (label_assignment.py:78-86).

    ``masks`` is either the full-resolution [G,H,W] array (grid_step=None) or the
    [G,ceil(H/step),ceil(W/step)] sub-grid holding exactly the pixels (y*step, x*step).
    """
    cx, cy, _, _ = grid_points(H, W, strides)
    xs, ys = cx.astype(np.int64), cy.astype(np.int64)
    if grid_step is not None:
        xs, ys = xs // grid_step, ys // grid_step
    return np.asarray(masks)[:, ys, xs].T.astype(np.float32)


def adapt_cal_k(candidate_anchor_sizes, object_size, positive_num):
    """LabelAssignment.adapt_cal_k (label_assignment.py:88-95); sizes float32, object_size float32 scalar."""
    size_lvl, num_lvl = np.unique(candidate_anchor_sizes, return_counts=True)
    ratio = num_lvl / candidate_anchor_sizes.shape[0]
    dk = np.exp((object_size - size_lvl) / (2 * size_lvl))
    dk = (ratio * dk).sum()
    return int(positive_num * dk + 0.5)


def assign_image(gt_bboxes, masks, H, W, stream, strides=STRIDES, regress_ranges=REGRESS_RANGES,
                 positive_num=10, neg_threshold=0.2, balance_sample=True, grid_step=None, adapt_positive_num=False,
                 multiply_samplepro_for_weight=False, octave_base_scale=8):
    """LabelAssignment.__call__ (label_assignment.py:136-201) + random_sample (:97-131), config defaults
    (ambiguous_sample='min_area', random_sample_by_distance=True); adapt_positive_num (:88-95, :104-107) and
    multiply_samplepro_for_weight (:127-128) are the two optional switches of the constructor.

    Returns points_to_gt_index int64 [P] (1-based; -1 negative, 0 ignore), points_weight f32 [P],
    and the number of uniforms consumed.
    """
    gt_bboxes = np.asarray(gt_bboxes, np.float32).reshape(-1, 4)
    G = gt_bboxes.shape[0]
    P = sum(h * w for h, w in level_shapes(H, W, strides))
    idx = np.full(P, -1, np.int64)
    wts = np.ones(P, np.float32)
    start = stream.pos
    if G == 0:
        return idx, wts, 0
    cand = candidate_flags(gt_bboxes, H, W, strides, regress_ranges)
    cell = sample_cells(masks, H, W, strides, grid_step)
    anchor_sizes = np.concatenate([np.full(h * w, np.float32(octave_base_scale * s), np.float32)
                                   for (h, w), s in zip(level_shapes(H, W, strides), strides)])            # :149
    areas = (gt_bboxes[:, 2] - gt_bboxes[:, 0]) * (gt_bboxes[:, 3] - gt_bboxes[:, 1])   # f32, :156
    order = sorted(range(G), key=lambda k: areas[k])                                    # stable, :170
    for g in order:
        R = np.nonzero(cand[:, g])[0]
        R = R[idx[R] == -1]                                                             # :177-179
        if R.size == 0:
            continue                                                                    # :182-183 (no RNG use)
        pro = cell[R, g].clip(min=EPS_PRO)                                              # :186-187 (f32)
        nonneg = pro > (neg_threshold * np.max(pro))                                    # :98
        N = R[nonneg]
        pro_n = pro[nonneg]
        p = pro_n / np.sum(pro_n)                                                       # :103 (f32)
        n = N.size
        k = positive_num
        if adapt_positive_num:                                                          # :104-105
            wh = (gt_bboxes[g, 2] - gt_bboxes[g, 0], gt_bboxes[g, 3] - gt_bboxes[g, 1])
            k = adapt_cal_k(anchor_sizes[R], max(wh[0], wh[1]), positive_num)
        if n < k:
            if balance_sample:
                chosen = legacy_choice(stream, p, k, True)                              # :112
            else:
                chosen = np.arange(n)                                                   # :116
        else:
            chosen = legacy_choice(stream, p, k, False)                                 # :119
        pos, cnt = np.unique(chosen, return_counts=True)                                # :125
        sampled = np.zeros(n, bool)
        sampled[chosen] = True
        weight = cnt.astype(np.float32)
        if multiply_samplepro_for_weight:                                               # :127-128
            weight *= pro_n[pos]
        idx[N[pos]] = g + 1                                                             # :193
        idx[N[~sampled]] = 0                                                            # :194
        wts[N[pos]] = weight                                                            # :195
        wts[N[~sampled]] = 0.0                                                          # :196
    return idx, wts, stream.pos - start


def assign_image_seeded(gt_bboxes, masks, H, W, seed, **kw):
    """np.random.seed(seed) immediately before the call (SURVEY §8c determinism note)."""
    rs = np.random.RandomState(seed)

    class _S:
        pos = 0

        def random_sample(self, n):
            self.pos += n
            return rs.random_sample(n)

    return assign_image(gt_bboxes, masks, H, W, _S(), **kw)


# --------------------------------------------------------------------------- targets
def targets_image(gt_bboxes, gt_labels, idx, num_classes, H, W, strides=STRIDES):
    """RADetHead._get_target_single (radet_head.py:373-392) + TBLR encode (tblr_bbox_coder.py:71-114).

    labels[idx>-1] = gt_labels[idx-1] — idx==0 wraps to the LAST GT (python negative index, :390).
    bbox_targets[idx>0] = ((d / side) / 0.125) with side = 8*stride, order T,B,L,R.
    """
    gt_bboxes = np.asarray(gt_bboxes, np.float32).reshape(-1, 4)
    gt_labels = np.asarray(gt_labels, np.int64)
    P = idx.shape[0]
    labels = np.full(P, num_classes, np.int64)
    tg = np.zeros((P, 4), np.float32)
    if gt_labels.shape[0] == 0:
        return labels, tg
    nn_ = idx > -1
    labels[nn_] = gt_labels[idx[nn_] - 1]
    pos = idx > 0
    cx, cy, s, _ = grid_points(H, W, strides)
    b = gt_bboxes[idx[pos] - 1]
    side = (s[pos] * np.float32(8.0))
    d = np.stack([cy[pos] - b[:, 1], b[:, 3] - cy[pos], cx[pos] - b[:, 0], b[:, 2] - cx[pos]], 1).astype(np.float32)
    tg[pos] = (d / side[:, None]) / np.float32(0.125)
    return labels, tg


def get_targets(gt_bboxes_list, gt_labels_list, idx_list, w_list, num_classes, H, W, strides=STRIDES):
    """RADetHead.get_targets (radet_head.py:290-369): per-level concat, level-major / image-minor."""
    shapes = level_shapes(H, W, strides)
    nl = [h * w for h, w in shapes]
    anchors = grid_anchors(H, W, strides)
    per_img = [targets_image(b, l, i, num_classes, H, W, strides) for b, l, i in zip(gt_bboxes_list, gt_labels_list, idx_list)]
    labels, tgs, wts, ancs = [], [], [], []
    off = 0
    for n in nl:
        labels.append(np.concatenate([p[0][off:off + n] for p in per_img]))
        tgs.append(np.concatenate([p[1][off:off + n] for p in per_img]))
        wts.append(np.concatenate([np.asarray(w, np.float32)[off:off + n] for w in w_list]))
        ancs.append(np.concatenate([anchors[off:off + n] for _ in per_img]))
        off += n
    return labels, tgs, wts, ancs


# --------------------------------------------------------------------------- loss (floating point: torch CPU)
def head_loss(cls_maps, bbox_maps, iou_maps, gt_bboxes_list, gt_labels_list, idx_list, w_list, num_classes,
              H, W, strides=STRIDES, gamma=2.0, alpha=0.25, w_cls=1.0, w_bbox=2.0, w_iou=1.0, eps=1e-6,
              dtype="float32", need_grad=True, normalizers=None):
    """RADetHead.loss (radet_head.py:173-288) with FocalLoss (focal_loss.py:10-41 py version, mmcv op absent),
    GIoULoss (iou_loss.py:82-98,319-354; bbox_overlaps iou2d_calculator.py:107-159) and
    CrossEntropyLoss(use_sigmoid) (cross_entropy_loss.py:58-91).

    Inputs are numpy NCHW maps per level.  Returns dict of python floats and (if need_grad) the gradients of
    (loss_cls + loss_bbox + loss_iou) w.r.t. every map, as numpy arrays in the same NCHW layout.

    normalizers=(num_pos, sum_wq): the opt-in FCOS/ATSS-style variant in which the two normalisers are the reduce_mean
    over the ranks (core/utils/dist_utils.py:63-69, atss_head.py:278,296) while the `num_pos > 0` branch (:261) stays
    rank-local; None = RADetHead's own behaviour (rank-local, :254-259).  `sum_wq` (local) is part of the result.
    """
    import torch
    import torch.nn.functional as F

    td = getattr(torch, dtype)
    B = cls_maps[0].shape[0]
    C = num_classes
    cls_t = [torch.tensor(np.asarray(m), dtype=td, requires_grad=need_grad) for m in cls_maps]
    box_t = [torch.tensor(np.asarray(m), dtype=td, requires_grad=need_grad) for m in bbox_maps]
    iou_t = [torch.tensor(np.asarray(m), dtype=td, requires_grad=need_grad) for m in iou_maps]
    labels, tgs, wts, ancs = get_targets(gt_bboxes_list, gt_labels_list, idx_list, w_list, C, H, W, strides)
    flat_cls = torch.cat([m.permute(0, 2, 3, 1).reshape(-1, C) for m in cls_t])           # :222-236
    flat_box = torch.cat([m.permute(0, 2, 3, 1).reshape(-1, 4) for m in box_t])
    flat_iou = torch.cat([m.permute(0, 2, 3, 1).reshape(-1) for m in iou_t])
    labels = torch.from_numpy(np.concatenate(labels))
    tgs = torch.from_numpy(np.concatenate(tgs)).to(td)
    wts = torch.from_numpy(np.concatenate(wts)).to(td)
    ancs = torch.from_numpy(np.concatenate(ancs)).to(td)
    pos = ((labels >= 0) & (labels < C)).nonzero().reshape(-1)                             # :245-247
    pw = wts[pos]
    num_pos = pw.sum()                                                                     # :254
    # focal, one-hot with label C = all-zero row
    onehot = F.one_hot(labels, C + 1)[:, :C].to(td)
    ps = flat_cls.sigmoid()
    pt = (1 - ps) * onehot + ps * (1 - onehot)
    fw = (alpha * onehot + (1 - alpha) * (1 - onehot)) * pt.pow(gamma)
    fl = F.binary_cross_entropy_with_logits(flat_cls, onehot, reduction="none") * fw
    n_norm = num_pos if normalizers is None else torch.tensor(float(normalizers[0]), dtype=td)
    loss_cls = w_cls * (fl * wts.view(-1, 1)).sum() / (n_norm + B)                         # :256-259

    def decode(anc, tblr):                                                                 # tblr_bbox_coder.py:117-172
        loc = tblr * 0.125
        ctr = (anc[:, 0:2] + anc[:, 2:4]) / 2
        wh = anc[:, 2:4] - anc[:, 0:2]
        t, b, l, r = loc[:, 0] * wh[:, 1], loc[:, 1] * wh[:, 1], loc[:, 2] * wh[:, 0], loc[:, 3] * wh[:, 0]
        return torch.stack([ctr[:, 0] - l, ctr[:, 1] - t, ctr[:, 0] + r, ctr[:, 1] + b], 1)

    def overlaps(b1, b2, mode):                                                            # iou2d_calculator.py:107-159
        a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
        a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
        lt = torch.max(b1[:, :2], b2[:, :2])
        rb = torch.min(b1[:, 2:], b2[:, 2:])
        wh = (rb - lt).clamp(min=0)
        ov = wh[:, 0] * wh[:, 1]
        e = a1.new_tensor([eps])
        union = torch.max(a1 + a2 - ov, e)
        ious = ov / union
        if mode == "iou":
            return ious
        elt = torch.min(b1[:, :2], b2[:, :2])
        erb = torch.max(b1[:, 2:], b2[:, 2:])
        ewh = (erb - elt).clamp(min=0)
        ea = torch.max(ewh[:, 0] * ewh[:, 1], e)
        return ious - (ea - union) / ea

    if float(num_pos) > 0:                                                                 # :261
        dp = decode(ancs[pos], flat_box[pos])
        dt = decode(ancs[pos], tgs[pos])
        iou_tg = overlaps(dp, dt, "iou").detach()                                          # :267
        wq = iou_tg.clamp(min=1e-12) * pw                                                  # :272
        sum_wq = float(wq.sum())
        q_norm = wq.sum() if normalizers is None else torch.tensor(float(normalizers[1]), dtype=td)
        loss_bbox = w_bbox * ((1 - overlaps(dp, dt, "giou")) * wq).sum() / q_norm          # :269-274
        bce = F.binary_cross_entropy_with_logits(flat_iou[pos], iou_tg, reduction="none")
        loss_iou = w_iou * (bce * pw).sum() / (pw.sum() if normalizers is None else n_norm)   # :275-278
    else:
        sum_wq = 0.0
        loss_bbox = flat_box[pos].sum()                                                    # :280-281
        loss_iou = flat_iou[pos].sum()
    out = dict(loss_cls=float(loss_cls.detach()), loss_bbox=float(loss_bbox.detach()), loss_iou=float(loss_iou.detach()),
               num_pos=float(num_pos), sum_wq=sum_wq)
    if need_grad:
        (loss_cls + loss_bbox + loss_iou).backward()
        z = lambda t: np.zeros(t.shape, np.dtype(dtype)) if t.grad is None else t.grad.numpy()
        out["grad_cls"] = [z(t) for t in cls_t]
        out["grad_bbox"] = [z(t) for t in box_t]
        out["grad_iou"] = [z(t) for t in iou_t]
    return out


# --------------------------------------------------------------------------- standalone TBLR coder
def tblr_encode(priors, gts, normalizer=0.125):
    """bboxes2tblr (tblr_bbox_coder.py:71-114), normalize_by_wh=True, scalar normalizer; float32, one rounding per op."""
    f = np.float32
    p, g = np.asarray(priors, f), np.asarray(gts, f)
    cx, cy = (p[:, 0] + p[:, 2]) / f(2), (p[:, 1] + p[:, 3]) / f(2)
    w, h = p[:, 2] - p[:, 0], p[:, 3] - p[:, 1]
    loc = np.stack([(cy - g[:, 1]) / h, (g[:, 3] - cy) / h, (cx - g[:, 0]) / w, (g[:, 2] - cx) / w], 1).astype(f)
    return (loc / f(normalizer)).astype(f)


def tblr_decode(priors, tblr, normalizer=0.125, max_shape=None, clip_border=True):
    """tblr2bboxes (tblr_bbox_coder.py:117-172)."""
    f = np.float32
    p, t = np.asarray(priors, f), np.asarray(tblr, f)
    loc = (t * f(normalizer)).astype(f)
    cx, cy = (p[:, 0] + p[:, 2]) / f(2), (p[:, 1] + p[:, 3]) / f(2)
    w, h = p[:, 2] - p[:, 0], p[:, 3] - p[:, 1]
    top, bot, left, right = loc[:, 0] * h, loc[:, 1] * h, loc[:, 2] * w, loc[:, 3] * w
    out = np.stack([cx - left, cy - top, cx + right, cy + bot], 1).astype(f)
    if clip_border and max_shape is not None:
        out[:, 0::2] = np.clip(out[:, 0::2], 0, f(max_shape[1]))
        out[:, 1::2] = np.clip(out[:, 1::2], 0, f(max_shape[0]))
    return out


# --------------------------------------------------------------------------- standalone LOSSES modules
def weight_reduce(loss, weight=None, reduction="mean", avg_factor=None):
    """losses/utils.py:26-52."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss
    if reduction == "mean":
        return loss.sum() / avg_factor
    if reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def standalone_loss(kind, pred, target, weight=None, reduction="mean", avg_factor=None, loss_weight=1.0, gamma=2.0, alpha=0.25,
                    eps=1e-6, dtype="float64"):
    """FocalLoss (focal_loss.py:10-41 py version + :72-86 weight reshape), GIoULoss (iou_loss.py:82-98, 327-354) and
    CrossEntropyLoss(use_sigmoid=True) (cross_entropy_loss.py:40-91) called on their own.  numpy in, (loss, grad of
    loss.sum() w.r.t. pred) out; evaluated in `dtype` (float64 gives the tolerance-free reference value)."""
    import torch
    import torch.nn.functional as F

    td = getattr(torch, dtype)
    x = torch.tensor(np.asarray(pred), dtype=td, requires_grad=True)
    w = None if weight is None else torch.tensor(np.asarray(weight), dtype=td)
    if kind == "focal":
        C = x.shape[1]
        onehot = F.one_hot(torch.from_numpy(np.asarray(target, np.int64)), C + 1)[:, :C].to(td)
        ps = x.sigmoid()
        pt = (1 - ps) * onehot + ps * (1 - onehot)
        loss = F.binary_cross_entropy_with_logits(x, onehot, reduction="none") * (alpha * onehot + (1 - alpha) * (1 - onehot)) * pt.pow(gamma)
        if w is not None and w.shape != loss.shape:
            w = w.view(-1, 1) if w.shape[0] == loss.shape[0] else w.view(loss.shape[0], -1)
    elif kind == "giou":
        t = torch.tensor(np.asarray(target), dtype=td)
        if w is not None and not bool((w > 0).any()):
            out = (x * w).sum()
            out.backward()
            return out.detach().numpy(), x.grad.numpy()
        if w is not None and w.dim() > 1:
            w = w.mean(-1)
        a1 = (x[:, 2] - x[:, 0]) * (x[:, 3] - x[:, 1])
        a2 = (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
        wh = (torch.min(x[:, 2:], t[:, 2:]) - torch.max(x[:, :2], t[:, :2])).clamp(min=0)
        ov = wh[:, 0] * wh[:, 1]
        e = a1.new_tensor([eps])
        union = torch.max(a1 + a2 - ov, e)
        ewh = (torch.max(x[:, 2:], t[:, 2:]) - torch.min(x[:, :2], t[:, :2])).clamp(min=0)
        ea = torch.max(ewh[:, 0] * ewh[:, 1], e)
        loss = 1 - (ov / union - (ea - union) / ea)
    elif kind == "bce":
        t = torch.from_numpy(np.asarray(target))
        if x.dim() != t.dim():
            C = x.shape[-1]
            onehot = torch.zeros((t.shape[0], C), dtype=td)
            rows = ((t >= 0) & (t < C)).nonzero().reshape(-1)
            onehot[rows, t[rows]] = 1
            t = onehot
            if w is not None:
                w = w.view(-1, 1).expand(w.shape[0], C)
        loss = F.binary_cross_entropy_with_logits(x, t.to(td), reduction="none")
    else:
        raise ValueError(kind)
    out = loss_weight * weight_reduce(loss, w, reduction, avg_factor)
    out.sum().backward()
    return out.detach().numpy(), x.grad.numpy()


# --------------------------------------------------------------------------- decode + candidate selection
def _sigmoid_f32(x):
    """``Tensor.sigmoid()`` of radet_head.py:106-109.

    torch's sigmoid is not one function: the CPU kernel mixes a SIMD (Sleef) path and a scalar path depending on
    memory alignment / chunking, and the CUDA kernel (what the reference really runs at test time, the maps being
    CUDA tensors) computes 1/(1+expf(-x)).  They differ by <=1 ulp on some elements.  The default here is torch-CPU;
    the GPU parity tests install the CUDA flavour through ``set_sigmoid`` so that "bit-exact" means bit-exact against
    the reference run the way it is deployed.
    """
    if _SIGMOID_IMPL is not None:
        return np.asarray(_SIGMOID_IMPL(np.ascontiguousarray(x, np.float32)), np.float32)
    import torch

    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).sigmoid().numpy()


_SIGMOID_IMPL = None


def set_sigmoid(fn):
    """Install the sigmoid flavour (callable np.f32 array -> np.f32 array) or None for torch-CPU."""
    global _SIGMOID_IMPL
    _SIGMOID_IMPL = fn


def select_candidates(cls_maps, bbox_maps, iou_maps, img_shape, scale_factor, score_thr, nms_pre, strides=STRIDES,
                      rescale=True):
    """RADetHead._get_bboxes_single up to the NMS call (radet_head.py:96-146) for ONE image.

    cls_maps[l]: [C,h,w]; bbox_maps[l]: [4,h,w] (T,B,L,R, post-ReLU); iou_maps[l]: [1,h,w].
    Candidate order: level-major, then ascending flat (point*C + class) index — the reference's
    ``topk(sorted=False)`` order is implementation-defined; only ties in the sort key depend on it.
    Returns boxes [n,4], scores [n], ctr [n], labels [n] (i64), anchors [n,4].
    """
    Himg, Wimg = img_shape[0], img_shape[1]
    boxes, scores, ctrs, cats, ancs = [], [], [], [], []
    for lvl, s in enumerate(strides):
        C, h, w = cls_maps[lvl].shape
        sc = _sigmoid_f32(np.transpose(cls_maps[lvl], (1, 2, 0)).reshape(-1, C))             # :106-107
        bp = np.transpose(bbox_maps[lvl], (1, 2, 0)).reshape(-1, 4).astype(np.float32)
        ct = _sigmoid_f32(np.transpose(iou_maps[lvl], (1, 2, 0)).reshape(-1))
        cand = sc > np.float32(score_thr)                                                     # :111 strict
        flat = np.nonzero(cand.reshape(-1))[0]
        k = min(nms_pre, flat.size) if nms_pre > 0 else flat.size
        if k == 0:
            continue
        v = sc.reshape(-1)[flat]
        if k < flat.size:
            # top-k over (point,class) pairs of this level (:122); ties -> lower flat index
            o = np.lexsort((flat, -v.astype(np.float64)))[:k]
            flat = np.sort(flat[o])
            v = sc.reshape(-1)[flat]
        pt, cl = np.divmod(flat, C)
        ys, xs = np.divmod(pt, w)
        cx, cy = (xs * s).astype(np.float32), (ys * s).astype(np.float32)
        side = np.float32(8 * s)
        loc = bp[pt] * np.float32(0.125)                                                      # tblr_bbox_coder.py:154-160
        t, b, l, r = loc[:, 0] * side, loc[:, 1] * side, loc[:, 2] * side, loc[:, 3] * side
        bx = np.stack([cx - l, cy - t, cx + r, cy + b], 1).astype(np.float32)
        bx[:, 0::2] = bx[:, 0::2].clip(0, Wimg)                                               # :167-171
        bx[:, 1::2] = bx[:, 1::2].clip(0, Himg)
        half = np.float32(0.5) * side
        boxes.append(bx)
        scores.append(v.astype(np.float32))
        ctrs.append(ct[pt])
        cats.append(cl.astype(np.int64))
        ancs.append(np.stack([cx - half, cy - half, cx + half, cy + half], 1).astype(np.float32))
    if not boxes:
        z = np.zeros((0, 4), np.float32)
        return z, np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.int64), z
    boxes = np.concatenate(boxes)
    ancs = np.concatenate(ancs)
    if rescale:
        sf = np.asarray(scale_factor, np.float32)
        boxes = boxes / sf                                                                    # :141-143
        ancs = ancs / sf
    return boxes, np.concatenate(scores), np.concatenate(ctrs), np.concatenate(cats), ancs


# --------------------------------------------------------------------------- vote NMS (python restatement; small n only)
def _vote_dim(s, x):
    """vote_single_dim (vote_ext.cpp:8-35): f32, no FMA, sequential sums in member order."""
    f = np.float32
    ss, vx = f(0), f(0)
    for si, xi in zip(s, x):
        ss = f(ss + si)
        vx = f(vx + f(si * xi))
    vx = f(vx / ss)
    sg = f(0)
    for si, xi in zip(s, x):
        sg = f(sg + f(f(si * f(xi - vx)) * f(xi - vx)))
    sg = f(np.sqrt(f(sg / ss)))
    fs, fx = f(0), f(0)
    lo, hi = f(vx - sg), f(vx + sg)
    for si, xi in zip(s, x):
        if lo <= xi <= hi:
            fx = f(fx + f(si * xi))
            fs = f(fs + si)
    with np.errstate(invalid="ignore", divide="ignore"):
        return f(fx / fs)


def stable_desc_order(scores):
    """torch::sort(scores, 0, descending) (vote_ext.cpp:78) — ties: lower index first (observed stable)."""
    return np.argsort(-np.asarray(scores, np.float32).astype(np.float64), kind="stable")


def vote_nms_py(boxes, cluster_scores, vote_scores, labels, thr, global_mode=False, return_clusters=False):
    """vote_nms (vote_ext.cpp:70-207) / global_vote_nms (:210-353), iou_enable=False. Pure python: small n."""
    f = np.float32
    boxes = np.asarray(boxes, f)
    cs, vs = np.asarray(cluster_scores, f), np.asarray(vote_scores, f)
    n = cs.shape[0]
    order = stable_desc_order(cs)
    sup = np.zeros(n, bool)
    seen = set()
    out_b, out_l, out_s, clusters = [], [], [], []
    thr = f(thr)
    for ii in range(n):
        i = order[ii]
        if sup[i]:
            continue
        if global_mode and labels[i] in seen:                                   # :257-263
            sup[i] = True
            continue
        seen.add(labels[i])
        sup[i] = True
        mem = [i]
        ai = f(f(boxes[i, 2] - boxes[i, 0]) * f(boxes[i, 3] - boxes[i, 1]))
        for jj in range(ii + 1, n):
            j = order[jj]
            if labels[j] != labels[i] or sup[j]:
                continue
            iw = max(f(0), f(min(boxes[j, 2], boxes[i, 2]) - max(boxes[j, 0], boxes[i, 0])))
            ih = max(f(0), f(min(boxes[j, 3], boxes[i, 3]) - max(boxes[j, 1], boxes[i, 1])))
            inter = f(iw * ih)
            aj = f(f(boxes[j, 2] - boxes[j, 0]) * f(boxes[j, 3] - boxes[j, 1]))
            with np.errstate(invalid="ignore", divide="ignore"):
                iou = f(inter / f(f(aj + ai) - inter))
            if iou > thr:
                sup[j] = True
                mem.append(j)
        s = vs[mem]
        out_b.append([_vote_dim(s, boxes[mem, d]) for d in range(4)])
        out_l.append(labels[i])
        out_s.append(max(cs[mem]))
        clusters.append(mem)
    res = (np.asarray(out_b, f).reshape(-1, 4), np.asarray(out_l, np.int64), np.asarray(out_s, f))
    return res + (clusters,) if return_clusters else res


# --------------------------------------------------------------------------- vote NMS (plain C restatement; any n)
_HERE = os.path.dirname(os.path.abspath(__file__))
_C_LIB = None


def build_c_oracle(force=False):
    """gcc -O2 -ffp-contract=off oracle/vote_oracle.c -> oracle/_build/libvote_oracle.so"""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libvote_oracle.so")
    src = os.path.join(_HERE, "vote_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return so


def _clib():
    global _C_LIB
    if _C_LIB is None:
        _C_LIB = ctypes.CDLL(build_c_oracle())
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        _C_LIB.oracle_vote_nms.restype = ctypes.c_int64
        _C_LIB.oracle_vote_nms.argtypes = [fp, fp, fp, ip, ip, ctypes.c_int64, ctypes.c_float, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_float, fp, ip, fp, ip, ip]
    return _C_LIB


def vote_nms_c(boxes, cluster_scores, vote_scores, labels, thr, global_mode=False, iou_enable=False, sigma=0.025):
    """C restatement of vote_ext.cpp:70-353 + cluster_ext.cpp:4-87 (instance ids / cluster sizes as by-products).

    Returns voted boxes [k,4], labels [k], scores [k], instance_ids [n], clusters_num [n].
    """
    lib = _clib()
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    cs = np.ascontiguousarray(cluster_scores, np.float32)
    vs = np.ascontiguousarray(vote_scores, np.float32)
    lb = np.ascontiguousarray(labels, np.int64)
    n = cs.shape[0]
    order = np.ascontiguousarray(stable_desc_order(cs), np.int64)
    ob = np.zeros((max(n, 1), 4), np.float32)
    ol = np.zeros(max(n, 1), np.int64)
    os_ = np.zeros(max(n, 1), np.float32)
    inst = np.zeros(max(n, 1), np.int64)
    cnum = np.zeros(max(n, 1), np.int64)
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
    P = lambda a, t: a.ctypes.data_as(t)
    k = lib.oracle_vote_nms(P(boxes, fp), P(cs, fp), P(vs, fp), P(lb, ip), P(order, ip), n, ctypes.c_float(thr),
                            int(global_mode), int(iou_enable), ctypes.c_float(sigma),
                            P(ob, fp), P(ol, ip), P(os_, fp), P(inst, ip), P(cnum, ip))
    return ob[:k], ol[:k], os_[:k], inst[:n], cnum[:n]


def vote_nms_wrapper(bboxes, cls_scores, labels, nms_cfg, score_factor=None, max_num=0, global_mode=False):
    """radet.ops.vote_nms / global_vote_nms python wrappers (vote_wrapper.py:7-43, 47-83)."""
    cfg = dict(nms_cfg)
    thr = cfg.pop("iou_threshold", 0.6)
    cst = cfg.pop("cluster_score", "cls")
    vst = cfg.pop("vote_score", "iou")
    iou_enable = cfg.pop("iou_enable", False)
    sigma = cfg.pop("sigma", 0.025)
    f = np.float32

    def pick(t):
        if isinstance(t, (list, tuple)):
            return (np.asarray(cls_scores, f) * np.asarray(score_factor, f)).astype(f)
        if t == "cls":
            return np.asarray(cls_scores, f)
        if t == "iou":
            return np.asarray(score_factor, f)
        raise RuntimeError(f"Unexpected score type:{t}")

    b, l, s, _, _ = vote_nms_c(bboxes, pick(cst), pick(vst), labels, thr, global_mode, iou_enable, sigma)
    dets = np.concatenate([b, s[:, None]], 1).astype(f)
    if max_num > 0:
        dets, l = dets[:max_num], l[:max_num]
    return dets, l


def greedy_nms_keep(boxes, scores, labels, thr):
    """Class-aware greedy NMS keep-set (mmcv batched_nms semantics per SURVEY Appendix B, with the a13 IoU rule).
    The keep set equals the seed set of vote_nms."""
    _, _, _, inst, cnum = vote_nms_c(boxes, scores, scores, labels, thr)
    order = stable_desc_order(scores)
    keep = [i for i in order if cnum[i] > 0]
    return np.asarray(keep, np.int64)


def get_bboxes_image(cls_maps, bbox_maps, iou_maps, img_shape, scale_factor, score_thr=0.05, nms_pre=1000,
                     max_per_img=100, nms_cfg=None, rescale=True, strides=STRIDES):
    """RADetHead._get_bboxes_single (radet_head.py:55-169), nms.type in {'vote','global_vote','nms'}."""
    nms_cfg = dict(nms_cfg or dict(type="vote", iou_threshold=0.65, cluster_score=["cls", "iou"],
                                   vote_score=["iou", "cls"], iou_enable=False))
    typ = nms_cfg.pop("type", "vote")
    boxes, sc, ctr, cats, ancs = select_candidates(cls_maps, bbox_maps, iou_maps, img_shape, scale_factor, score_thr,
                                                   nms_pre, strides, rescale)
    if boxes.shape[0] == 0:
        return np.zeros((0, 5), np.float32), np.zeros((0,), np.int64)
    if typ in ("vote", "global_vote"):
        return vote_nms_wrapper(boxes, sc, cats, nms_cfg, score_factor=ctr, max_num=max_per_img,
                                global_mode=(typ == "global_vote"))
    s = (sc * ctr).astype(np.float32)
    keep = greedy_nms_keep(boxes, s, cats, nms_cfg.get("iou_threshold", 0.5))
    if max_per_img > 0:
        keep = keep[:max_per_img]
    return np.concatenate([boxes[keep], s[keep, None]], 1), cats[keep]
