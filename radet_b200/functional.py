"""Device-level Python API over the C ABI: torch tensors in, torch tensors out, everything enqueued on torch's
current CUDA stream.  torch is used for device memory, streams and (optionally) torch.distributed only; all
arithmetic of the hot path happens in libradet_b200.so.  No CPU / PyTorch fallback exists: without a CUDA device
and the compiled library these functions raise.
"""
import ctypes
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import RadetError, check

STRIDES = (8, 16, 32, 64, 128)
REGRESS_RANGES = ((-1.0, 64.0), (64.0, 128.0), (128.0, 256.0), (256.0, 512.0), (512.0, 1e8))


class Geometry:
    """Prior grid of the head + assignment ranges (AnchorGenerator(ratios=[1], octave_base_scale, scales_per_octave=1),
    TBLRBBoxCoder(normalizer), LabelAssignment(regress_ranges)); see include/radet_b200.h radet_grid_t."""

    def __init__(self, strides=STRIDES, regress_ranges=REGRESS_RANGES, anchor_scale=8.0, tblr_normalizer=0.125):
        self.strides = tuple(int(s) for s in strides)
        self.regress_ranges = tuple((float(a), float(b)) for a, b in regress_ranges)
        self.anchor_scale = float(anchor_scale)
        self.tblr_normalizer = float(tblr_normalizer)
        self.mask_step = int(np.gcd.reduce(np.asarray(self.strides)))
        self._grids = {}

    def level_shapes(self, H, W):
        return tuple((math.ceil(H / s), math.ceil(W / s)) for s in self.strides)   # label_assignment.py:138

    def grid(self, level_shapes):
        key = tuple(tuple(int(v) for v in hw) for hw in level_shapes)
        g = self._grids.get(key)
        if g is None:
            g = _lib.make_grid(key, self.strides, self.regress_ranges, self.anchor_scale, self.tblr_normalizer)
            self._grids[key] = g
        return g

    @staticmethod
    def num_points(level_shapes):
        return int(sum(int(h) * int(w) for h, w in level_shapes))


def _require_cuda(t: torch.Tensor, name: str):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RadetError(f"{name} must be a CUDA tensor (radet_b200 has no CPU path)")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_WS = {}


def _workspace(key, nbytes, device):
    """Zero-initialised, cached per configuration AND per stream (the kernels re-arm their counters; calls on the same
    stream are ordered, calls enqueued on different streams may overlap and must not share scratch memory)."""
    k = (key, str(device), torch.cuda.current_stream(device).cuda_stream)
    t = _WS.get(k)
    if t is None or t.numel() < nbytes:
        t = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS[k] = t
    return t


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def offsets_of(counts: Sequence[int], device) -> Tuple[np.ndarray, torch.Tensor]:
    off = np.zeros(len(counts) + 1, np.int32)
    np.cumsum(np.asarray(counts, np.int64), out=off[1:])
    return off, torch.from_numpy(off).to(device, non_blocking=True)


# ------------------------------------------------------------------------------------------------ masks
def pack_masks(src: torch.Tensor, step: int, grid_h: int, grid_w: int, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 [G,h,w] (full-res with step=mask_step, or the pre-sampled grid with step=1) -> u32 bits [G,grid_h,ceil(grid_w/32)]."""
    _require_cuda(src, "masks")
    if src.dtype not in (torch.uint8, torch.bool):
        raise RadetError("masks must be uint8 / bool (real-valued distance maps are outside the implemented surface)")
    src = src.contiguous().view(torch.uint8) if src.dtype == torch.bool else src.contiguous()
    G = src.shape[0]
    pitch = (grid_w + 31) // 32
    bits = torch.empty((G, grid_h, pitch), dtype=torch.int32, device=src.device)
    if G:
        check(_lib.load().radet_pack_masks(_ptr(src), G, src.shape[1], src.shape[2], step, grid_h, grid_w, _ptr(bits),
                                           _ptr(status), _stream()), "radet_pack_masks")
    return bits


def mt19937_uniforms(seeds: torch.Tensor, n: int) -> torch.Tensor:
    _require_cuda(seeds, "seeds")
    seeds = seeds.to(torch.int64).to(torch.int32) if seeds.dtype != torch.int32 else seeds
    out = torch.empty((seeds.numel(), n), dtype=torch.float64, device=seeds.device)
    check(_lib.load().radet_mt19937_uniforms(_ptr(seeds), seeds.numel(), n, _ptr(out), _stream()), "radet_mt19937_uniforms")
    return out


def seed_states(seeds: torch.Tensor) -> torch.Tensor:
    """[B,625] legacy MT19937 states of np.random.seed(seeds[b]) (enqueue early / on a side stream)."""
    _require_cuda(seeds, "seeds")
    if seeds.dtype != torch.int32:
        seeds = seeds.to(torch.int32)
    st = torch.empty((seeds.numel(), _lib.MT_STATE_WORDS), dtype=torch.int32, device=seeds.device)
    check(_lib.load().radet_mt19937_seed(_ptr(seeds), seeds.numel(), _ptr(st), _stream()), "radet_mt19937_seed")
    return st


# ------------------------------------------------------------------------------------------------ assignment
def assign(geom: Geometry, level_shapes, gt_counts: Sequence[int], gt_bboxes: torch.Tensor, mask_bits: torch.Tensor,
           mask_hw: Tuple[int, int], *, uniforms: Optional[torch.Tensor] = None, seeds: Optional[torch.Tensor] = None,
           mt_states: Optional[torch.Tensor] = None, positive_num: int = 10, balance_sample: bool = True,
           gt_offsets: Optional[Tuple[np.ndarray, torch.Tensor]] = None, out=None, adapt_positive_num: bool = False,
           multiply_samplepro_for_weight: bool = False, weight_sums: Optional[torch.Tensor] = None):
    """Batched LabelAssignment (label_assignment.py:136-201).  Returns points_to_gt_index int64 [B,P],
    points_weight f32 [B,P], consumed int32 [B] (written into `out` = (idx, w, consumed) when given; -1 = the uniforms
    ran out, -2 = an adaptive positive_num above 32).  weight_sums (optional float64 [B], written): per-image sum of the
    weights of the points with index >= 0 -- hand it to loss_fwd_bwd(weight_sums=...) together with idx / w."""
    _require_cuda(gt_bboxes, "gt_bboxes")
    dev = gt_bboxes.device
    off_h, off_d = gt_offsets if gt_offsets is not None else offsets_of(gt_counts, dev)
    B = int(off_h.shape[0]) - 1
    grid = geom.grid(level_shapes)
    P = geom.num_points(level_shapes)
    if out is not None:
        idx, w, consumed = out
        if tuple(idx.shape) != (B, P) or idx.dtype != torch.int64 or tuple(w.shape) != (B, P) or w.dtype != torch.float32 \
                or consumed.numel() != B or consumed.dtype != torch.int32 or not (idx.is_contiguous() and w.is_contiguous()):
            raise RadetError("assign(out=...): need contiguous idx int64 [B,P], w float32 [B,P], consumed int32 [B]")
    else:
        idx = torch.empty((B, P), dtype=torch.int64, device=dev)
        w = torch.empty((B, P), dtype=torch.float32, device=dev)
        consumed = torch.empty((B,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    nws = lib.radet_assign_workspace_bytes(ctypes.byref(grid), B)
    ws = _workspace(("assign", B, P), nws, dev)
    n_uniform = 0
    if uniforms is not None:
        uniforms = uniforms.contiguous()
        n_uniform = uniforms.shape[1]
    if seeds is not None and seeds.dtype != torch.int32:
        seeds = seeds.to(torch.int32)
    if weight_sums is not None and (weight_sums.dtype != torch.float64 or weight_sums.numel() != B or not weight_sums.is_cuda):
        raise RadetError("assign(weight_sums=...): need a CUDA float64 [B] tensor")
    check(lib.radet_assign(ctypes.byref(grid), B, _ptr(off_d), off_h.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                           _ptr(gt_bboxes.contiguous()), _ptr(mask_bits), mask_hw[0], mask_hw[1], geom.mask_step,
                           _ptr(uniforms), n_uniform, _ptr(seeds), _ptr(mt_states), positive_num,
                           int(bool(balance_sample)) | (2 if adapt_positive_num else 0) | (4 if multiply_samplepro_for_weight else 0),
                           _ptr(idx), _ptr(w), _ptr(consumed), _ptr(weight_sums), _ptr(ws), ws.numel(), _stream()), "radet_assign")
    return idx, w, consumed


# ------------------------------------------------------------------------------------------------ priors
def grid_priors(geom: Geometry, level_shapes, device, pad_shape=None, want_anchors=True):
    """AnchorGenerator.grid_anchors / valid_flags of one image on the device: (anchors f32 [P,4] or None, flags u8 [P] or None)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RadetError("grid_priors runs on a CUDA device (radet_b200 has no CPU path)")
    grid = geom.grid(level_shapes)
    P = geom.num_points(level_shapes)
    with torch.cuda.device(device):
        anc = torch.empty((P, 4), dtype=torch.float32, device=device) if want_anchors else None
        flg = torch.empty((P,), dtype=torch.uint8, device=device) if pad_shape is not None else None
        ph, pw_ = (int(pad_shape[0]), int(pad_shape[1])) if pad_shape is not None else (0, 0)
        check(_lib.load().radet_grid_priors(ctypes.byref(grid), ph, pw_, _ptr(anc), _ptr(flg), _stream()), "radet_grid_priors")
    return anc, flg


def release_workspaces():
    """Drop the cached per-(configuration, stream) workspaces (they are kept for the life of the process otherwise: a stream
    that goes away leaves its scratch behind)."""
    _WS.clear()


# ------------------------------------------------------------------------------------------------ targets
def get_targets(geom: Geometry, level_shapes, num_classes: int, gt_counts, gt_bboxes, gt_labels, idx, w,
                with_anchors=True, gt_offsets=None):
    """RADetHead.get_targets (radet_head.py:290-369): flat level-major/image-minor tensors
    (labels [B*P], bbox_targets [B*P,4], weights [B*P], anchors [B*P,4])."""
    _require_cuda(idx, "points_to_gt_index")
    dev = idx.device
    B, P = idx.shape
    grid = geom.grid(level_shapes)
    off_h, off_d = gt_offsets if gt_offsets is not None else offsets_of(gt_counts, dev)
    labels = torch.empty((B * P,), dtype=torch.int64, device=dev)
    tg = torch.empty((B * P, 4), dtype=torch.float32, device=dev)
    wt = torch.empty((B * P,), dtype=torch.float32, device=dev)
    anc = torch.empty((B * P, 4), dtype=torch.float32, device=dev) if with_anchors else None
    check(_lib.load().radet_get_targets(ctypes.byref(grid), B, num_classes, _ptr(off_d), _ptr(gt_bboxes.contiguous()),
                                        _ptr(gt_labels.contiguous()), _ptr(idx.contiguous()), _ptr(w.contiguous()),
                                        _ptr(labels), _ptr(tg), _ptr(wt), _ptr(anc), _stream()), "radet_get_targets")
    return labels, tg, wt, anc


# ------------------------------------------------------------------------------------------------ loss
class LossConfig:
    def __init__(self, gamma=2.0, alpha=0.25, w_cls=1.0, w_bbox=2.0, w_iou=1.0, eps=1e-6):
        self.gamma, self.alpha, self.w_cls, self.w_bbox, self.w_iou, self.eps = gamma, alpha, w_cls, w_bbox, w_iou, eps

    def c_struct(self, avg_extra, weight_sums=None):
        c = _lib.LossCfg()
        c.gamma, c.alpha, c.w_cls, c.w_bbox, c.w_iou, c.eps, c.avg_extra = (self.gamma, self.alpha, self.w_cls, self.w_bbox,
                                                                              self.w_iou, self.eps, float(avg_extra))
        c.weight_sums = None if weight_sums is None else weight_sums.data_ptr()
        return c


def _check_maps(cls, bbox, iou, level_shapes, B, C):
    for l, (c, b, o) in enumerate(zip(cls, bbox, iou)):
        h, w = level_shapes[l]
        for t, ch, nm in ((c, C, "cls_scores"), (b, 4, "bbox_preds"), (o, 1, "iou_preds")):
            _require_cuda(t, nm)
            if tuple(t.shape) != (B, ch, h, w) or t.dtype != torch.float32:
                raise RadetError(f"{nm}[{l}] must be float32 [{B},{ch},{h},{w}], got {t.dtype} {tuple(t.shape)}")


def loss_fwd_bwd(geom: Geometry, num_classes: int, cls, bbox, iou, gt_counts, gt_bboxes, gt_labels, idx, w,
                 cfg: LossConfig, want_grads=True, grad_scale: Optional[torch.Tensor] = None, gt_offsets=None,
                 sync_group=None, weight_sums: Optional[torch.Tensor] = None):
    """Fused RADetHead.loss forward+backward (radet_head.py:173-288).  Returns losses f32[4]
    (loss_cls, loss_bbox, loss_iou, num_pos) and (grad_cls, grad_bbox, grad_iou) lists (or None).

    sync_group: opt-in FCOS/ATSS-style reduce_mean of the two normalisers over that process group (NOT the
    reference behaviour of RADetHead, which keeps them rank-local).
    weight_sums: optional float64 [B] written by assign(weight_sums=...) for these idx / w: the dense pass then streams its
    class planes NEXT to the sparse positive pass instead of behind it (radet_loss_fwd_bwd, "overlapped" order)."""
    B = cls[0].shape[0]
    level_shapes = tuple(tuple(t.shape[-2:]) for t in cls)
    _check_maps(cls, bbox, iou, level_shapes, B, num_classes)
    dev = cls[0].device
    cls = [t.contiguous() for t in cls]
    bbox = [t.contiguous() for t in bbox]
    iou = [t.contiguous() for t in iou]
    grid = geom.grid(level_shapes)
    P = geom.num_points(level_shapes)
    if tuple(idx.shape) != (B, P) or tuple(w.shape) != (B, P):
        raise RadetError(f"points_to_gt_index / points_weight must be [{B},{P}] for these feature maps "
                         f"(got {tuple(idx.shape)}): LabelAssignment and the head disagree on the level sizes")
    off_h, off_d = gt_offsets if gt_offsets is not None else offsets_of(gt_counts, dev)
    lib = _lib.load()
    maps = _lib.make_maps([t.data_ptr() for t in cls], [t.data_ptr() for t in bbox], [t.data_ptr() for t in iou])
    grads = None
    gmaps = None
    if want_grads:
        # one allocation, every segment 16-byte aligned
        sizes = [t.numel() for t in cls] + [t.numel() for t in bbox] + [t.numel() for t in iou]
        starts, tot = [], 0
        for s in sizes:
            starts.append(tot)
            tot += (s + 3) // 4 * 4
        flat = torch.empty((tot,), dtype=torch.float32, device=dev)
        views = [flat[st:st + s].view(t.shape) for st, s, t in zip(starts, sizes, cls + bbox + iou)]
        L = len(cls)
        grads = (views[:L], views[L:2 * L], views[2 * L:])
        gmaps = _lib.make_maps([t.data_ptr() for t in grads[0]], [t.data_ptr() for t in grads[1]], [t.data_ptr() for t in grads[2]])
    losses = torch.empty((4,), dtype=torch.float32, device=dev)
    nws = lib.radet_loss_workspace_bytes(ctypes.byref(grid), B, num_classes)
    ws = _workspace(("loss", B, P, num_classes), nws, dev)
    if weight_sums is not None and (weight_sums.dtype != torch.float64 or weight_sums.numel() != B or not weight_sums.is_cuda):
        raise RadetError("loss_fwd_bwd(weight_sums=...): need a CUDA float64 [B] tensor")
    ccfg = cfg.c_struct(avg_extra=B, weight_sums=weight_sums if sync_group is None else None)

    def run(phases):
        check(lib.radet_loss_fwd_bwd(ctypes.byref(grid), B, num_classes, ctypes.byref(maps), _ptr(off_d),
                                     _ptr(gt_bboxes.contiguous()), _ptr(gt_labels.contiguous()), _ptr(idx.contiguous()),
                                     _ptr(w.contiguous()), ctypes.byref(ccfg), _ptr(grad_scale),
                                     ctypes.byref(gmaps) if gmaps is not None else None, _ptr(losses), phases, _ptr(ws),
                                     ws.numel(), _stream()), "radet_loss_fwd_bwd")

    if sync_group is None:
        run(3)
    else:
        from .sharding import reduce_mean_

        run(1)
        reduce_mean_(ws[:16].view(torch.float64), sync_group)   # num_pos, sum(wq): the two normalisers, over NCCL
        run(2)
    return losses, grads


def scale_grads(geom: Geometry, num_classes: int, grads, upstream: torch.Tensor):
    cls, bbox, iou = grads
    B = cls[0].shape[0]
    level_shapes = tuple(tuple(t.shape[-2:]) for t in cls)
    grid = geom.grid(level_shapes)
    gmaps = _lib.make_maps([t.data_ptr() for t in cls], [t.data_ptr() for t in bbox], [t.data_ptr() for t in iou])
    check(_lib.load().radet_scale_grads(ctypes.byref(grid), B, num_classes, ctypes.byref(gmaps), _ptr(upstream), _stream()),
          "radet_scale_grads")


class HeadLossFunction(torch.autograd.Function):
    """Autograd node of the fused loss: gradients are produced by the forward kernel and only rescaled in backward."""

    @staticmethod
    def forward(ctx, geom, num_classes, cfg, gt_counts, gt_bboxes, gt_labels, idx, w, sync_group, weight_sums, nlev, *maps):
        cls, bbox, iou = list(maps[:nlev]), list(maps[nlev:2 * nlev]), list(maps[2 * nlev:])
        need = any(t.requires_grad for t in maps)
        losses, grads = loss_fwd_bwd(geom, num_classes, cls, bbox, iou, gt_counts, gt_bboxes, gt_labels, idx, w, cfg,
                                     want_grads=need, sync_group=sync_group, weight_sums=weight_sums)
        ctx.geom, ctx.num_classes, ctx.grads, ctx.nlev = geom, num_classes, grads, nlev
        ctx.applied = None            # upstream factors already folded into ctx.grads (after the first backward)
        return losses[0], losses[1], losses[2], losses[3]

    @staticmethod
    def backward(ctx, g_cls, g_bbox, g_iou, _g_np):
        grads = ctx.grads
        if grads is None:
            return (None,) * (11 + 3 * ctx.nlev)
        z = lambda g, ref: torch.zeros((), dtype=torch.float32, device=ref.device) if g is None else g.reshape(()).float()
        ref = grads[0][0]
        up = torch.stack([z(g_cls, ref), z(g_bbox, ref), z(g_iou, ref)])
        if ctx.applied is None:       # first backward: rescale the forward's buffers in place (no extra memory, no copy)
            scale_grads(ctx.geom, ctx.num_classes, grads, up)   # exits early on the device when upstream == (1,1,1)
            ctx.applied = up
            return (None,) * 11 + tuple(grads[0]) + tuple(grads[1]) + tuple(grads[2])
        # re-entrant use (retain_graph=True, autograd.grad twice, checkpointing): the buffers already carry the first call's
        # factors and were handed out; return fresh tensors scaled by the ratio instead of compounding in place
        ratio = up / ctx.applied
        out = [[t * ratio[k] for t in grads[k]] for k in range(3)]
        return (None,) * 11 + tuple(out[0]) + tuple(out[1]) + tuple(out[2])


def head_loss(geom, num_classes, cfg, cls, bbox, iou, gt_counts, gt_bboxes, gt_labels, idx, w, sync_group=None, weight_sums=None):
    out = HeadLossFunction.apply(geom, num_classes, cfg, gt_counts, gt_bboxes, gt_labels, idx, w, sync_group, weight_sums, len(cls),
                                 *cls, *bbox, *iou)
    return dict(loss_cls=out[0], loss_bbox=out[1], loss_iou=out[2]), out[3]


# ------------------------------------------------------------------------------------------------ standalone coder
def tblr_encode(priors: torch.Tensor, gts: torch.Tensor, normalizer: float) -> torch.Tensor:
    _require_cuda(priors, "bboxes")
    priors, gts = priors.contiguous().float(), gts.contiguous().float()
    out = torch.empty_like(priors)
    check(_lib.load().radet_tblr_encode(_ptr(priors), _ptr(gts), priors.shape[0], normalizer, _ptr(out), _stream()), "radet_tblr_encode")
    return out


def tblr_decode(priors: torch.Tensor, tblr: torch.Tensor, normalizer: float, max_shape=None, clip_border=True) -> torch.Tensor:
    _require_cuda(priors, "bboxes")
    priors, tblr = priors.contiguous().float(), tblr.contiguous().float()
    out = torch.empty_like(priors)
    clip = int(bool(clip_border and max_shape is not None))
    mh, mw = (float(max_shape[0]), float(max_shape[1])) if clip else (0.0, 0.0)
    check(_lib.load().radet_tblr_decode(_ptr(priors), _ptr(tblr), priors.shape[0], normalizer, clip, mh, mw, _ptr(out), _stream()),
          "radet_tblr_decode")
    return out


# ------------------------------------------------------------------------------------------------ standalone losses
class _ElementwiseLoss(torch.autograd.Function):
    """loss = kernel(pred, target) element-wise; the same launch writes d loss / d pred, backward is one multiply."""

    @staticmethod
    def forward(ctx, pred, target, kind, a, b):
        _require_cuda(pred, "pred")
        p = pred.detach().contiguous().float()
        lib = _lib.load()
        want = pred.requires_grad
        if kind == "focal":
            t = target.detach().contiguous().to(torch.int64)
            loss = torch.empty_like(p)
            d = torch.empty_like(p) if want else None
            check(lib.radet_sigmoid_focal_loss(_ptr(p), _ptr(t), p.shape[0], p.shape[1], float(a), float(b), _ptr(loss), _ptr(d),
                                               _stream()), "radet_sigmoid_focal_loss")
        elif kind == "giou":
            t = target.detach().contiguous().float()
            loss = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
            d = torch.empty_like(p) if want else None
            check(lib.radet_giou_loss(_ptr(p), _ptr(t), p.shape[0], float(a), _ptr(loss), _ptr(d), _stream()), "radet_giou_loss")
        else:
            t = target.detach().contiguous().float()
            loss = torch.empty_like(p)
            d = torch.empty_like(p) if want else None
            check(lib.radet_bce_with_logits(_ptr(p), _ptr(t), p.numel(), _ptr(loss), _ptr(d), _stream()), "radet_bce_with_logits")
        ctx.save_for_backward(d)
        ctx.kind = kind
        return loss

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        if ctx.kind == "giou":
            g = g.unsqueeze(-1)
        return g * d, None, None, None, None


def sigmoid_focal_loss_elementwise(pred: torch.Tensor, target: torch.Tensor, gamma: float, alpha: float) -> torch.Tensor:
    """[N,C] un-reduced sigmoid focal loss (mmcv.ops.sigmoid_focal_loss(..., 'none'), focal_loss.py:70-71)."""
    if pred.dim() != 2 or target.dim() != 1 or target.shape[0] != pred.shape[0]:
        raise RadetError("sigmoid_focal_loss expects pred [N,C] and target [N]")
    return _ElementwiseLoss.apply(pred, target, "focal", gamma, alpha)


def giou_loss_elementwise(pred: torch.Tensor, target: torch.Tensor, eps: float) -> torch.Tensor:
    """[n] un-reduced 1 - GIoU of aligned (x1,y1,x2,y2) rows (iou_loss.py:82-98)."""
    if pred.dim() != 2 or pred.shape[1] != 4 or target.shape != pred.shape:
        raise RadetError("giou_loss expects pred and target of shape [n,4]")
    return _ElementwiseLoss.apply(pred, target, "giou", eps, 0.0)


def bce_with_logits_elementwise(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Un-reduced binary cross entropy with logits, same shape as pred (cross_entropy_loss.py:83-85)."""
    if target.shape != pred.shape:
        raise RadetError("binary_cross_entropy expects label of the same shape as pred")
    return _ElementwiseLoss.apply(pred, target, "bce", 0.0, 0.0)


# ------------------------------------------------------------------------------------------------ NMS family
_SCORE_MODE = {"cls": 1, "iou": 2}


def score_mode(t):
    """vote_wrapper.py:14-30: list/tuple -> cls*iou, 'cls', 'iou'."""
    if isinstance(t, (list, tuple)):
        return 0
    if t in _SCORE_MODE:
        return _SCORE_MODE[t]
    raise RuntimeError(f"Unexpected score type:{t}")


def vote_nms_lists(counts: Sequence[int], boxes, cluster_scores, vote_scores, labels, iou_threshold, mode=_lib.NMS_VOTE,
                   iou_enable=False, sigma=0.025, max_num=0, want_clusters=False):
    """Batched vote_ext.vote_nms / global_vote_nms / cluster_ext.cluster_nms on explicit lists (device tensors).
    Returns dets [n,5] (rows of list i start at offsets[i]), labels [n], index [n], num_out [batch], (inst, cnum)."""
    _require_cuda(boxes, "bboxes")
    dev = boxes.device
    batch = len(counts)
    off = np.zeros(batch + 1, np.int32)
    np.cumsum(np.asarray(counts, np.int64), out=off[1:])
    n = int(off[-1])
    lib = _lib.load()
    dets = torch.empty((n, 5), dtype=torch.float32, device=dev)
    olab = torch.empty((n,), dtype=torch.int64, device=dev)
    oidx = torch.empty((n,), dtype=torch.int64, device=dev)
    num = torch.empty((batch,), dtype=torch.int32, device=dev)
    inst = torch.empty((n,), dtype=torch.int64, device=dev) if want_clusters else None
    cnum = torch.empty((n,), dtype=torch.int64, device=dev) if want_clusters else None
    maxn = int(max(counts)) if batch else 0
    nws = lib.radet_vote_nms_workspace_bytes(batch, n, maxn)
    ws = _workspace(("nms", batch, maxn), nws, dev)
    check(lib.radet_vote_nms(batch, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _ptr(boxes.contiguous().float()),
                             _ptr(cluster_scores.contiguous().float()), _ptr(vote_scores.contiguous().float()),
                             _ptr(labels.contiguous().to(torch.int64)), float(iou_threshold), int(bool(iou_enable)), float(sigma),
                             mode, int(max_num), _ptr(dets), _ptr(olab), _ptr(oidx), _ptr(num), _ptr(inst), _ptr(cnum),
                             _ptr(ws), ws.numel(), _stream()), "radet_vote_nms")
    return dets, olab, oidx, num, inst, cnum, off


# ------------------------------------------------------------------------------------------------ inference
class DetectConfig:
    def __init__(self, score_thr=0.05, nms_pre=1000, max_per_img=100, nms_type="vote", iou_threshold=0.6,
                 cluster_score="cls", vote_score="iou", iou_enable=False, sigma=0.025):
        self.score_thr, self.nms_pre, self.max_per_img = float(score_thr), int(nms_pre), int(max_per_img)
        self.nms_mode = {"vote": _lib.NMS_VOTE, "global_vote": _lib.NMS_GLOBAL_VOTE}.get(nms_type, _lib.NMS_PLAIN)
        self.iou_threshold = float(iou_threshold)
        if self.nms_mode == _lib.NMS_PLAIN:   # radet_head.py:160: batched_nms(boxes, cls*centerness, ...)
            self.cs_mode = self.vs_mode = 0
        else:
            self.cs_mode, self.vs_mode = score_mode(cluster_score), score_mode(vote_score)
        self.iou_enable, self.sigma = bool(iou_enable), float(sigma)

    @classmethod
    def from_test_cfg(cls, cfg):
        """cfg: the head's test_cfg (configs/bop/r50_ycbv_pbr.py:70-80); nms keys as read by vote_wrapper.py:8-13
        (note the shipped configs spell `sima`, which the reference ignores too)."""
        get = cfg.get if hasattr(cfg, "get") else lambda k, d=None: getattr(cfg, k, d)
        nms = dict(get("nms"))
        return cls(score_thr=get("score_thr"), nms_pre=get("nms_pre", -1), max_per_img=get("max_per_img"),
                   nms_type=nms.get("type", "nms"),
                   iou_threshold=nms.get("iou_threshold", 0.6 if nms.get("type") in ("vote", "global_vote") else 0.5),
                   cluster_score=nms.get("cluster_score", "cls"), vote_score=nms.get("vote_score", "iou"),
                   iou_enable=nms.get("iou_enable", False), sigma=nms.get("sigma", 0.025))

    def with_max_per_img(self, n):
        import copy
        c = copy.copy(self)
        c.max_per_img = int(n)
        return c

    def c_struct(self, rescale):
        c = _lib.DetectCfg()
        c.score_thr, c.nms_pre, c.max_per_img, c.nms_mode = self.score_thr, self.nms_pre, self.max_per_img, self.nms_mode
        c.iou_threshold, c.cluster_score_mode, c.vote_score_mode = self.iou_threshold, self.cs_mode, self.vs_mode
        c.iou_enable, c.sigma, c.rescale = int(self.iou_enable), self.sigma, int(bool(rescale))
        return c


def _check_image_info(img_shapes, scale_factors, B, rescale):
    _require_cuda(img_shapes, "img_shapes")
    if img_shapes.dtype != torch.int32 or tuple(img_shapes.shape) != (B, 2) or not img_shapes.is_contiguous():
        raise RadetError(f"img_shapes must be a contiguous int32 [{B},2] (h,w) tensor, got {img_shapes.dtype} {tuple(img_shapes.shape)}")
    if scale_factors is None:
        if rescale:
            raise RadetError("rescale=True needs scale_factors")
        return
    _require_cuda(scale_factors, "scale_factors")
    if scale_factors.dtype != torch.float32 or tuple(scale_factors.shape) != (B, 4) or not scale_factors.is_contiguous():
        raise RadetError(f"scale_factors must be a contiguous float32 [{B},4] tensor, got {scale_factors.dtype} {tuple(scale_factors.shape)}")


def get_bboxes(geom: Geometry, num_classes: int, cls, bbox, iou, img_shapes: torch.Tensor, scale_factors: torch.Tensor,
               cfg: DetectConfig, rescale=False):
    """Batched ATSSHead.get_bboxes + RADetHead._get_bboxes_single (device in, device out).
    img_shapes int32 [B,2] (h,w), scale_factors f32 [B,4] (device).  Returns dets [B,max,5], labels [B,max], num [B]."""
    B = cls[0].shape[0]
    level_shapes = tuple(tuple(t.shape[-2:]) for t in cls)
    _check_maps(cls, bbox, iou, level_shapes, B, num_classes)
    _check_image_info(img_shapes, scale_factors, B, rescale)
    if cfg.max_per_img <= 0:     # the reference reads max_num <= 0 as "no cap" (vote_wrapper.py:39-42)
        cap = int(_lib.load().radet_candidates_capacity(ctypes.byref(geom.grid(level_shapes)), num_classes, cfg.nms_pre))
        cfg = cfg.with_max_per_img(min(cap, 4096))
    dev = cls[0].device
    cls = [t.contiguous() for t in cls]
    bbox = [t.contiguous() for t in bbox]
    iou = [t.contiguous() for t in iou]
    grid = geom.grid(level_shapes)
    lib = _lib.load()
    maps = _lib.make_maps([t.data_ptr() for t in cls], [t.data_ptr() for t in bbox], [t.data_ptr() for t in iou])
    ccfg = cfg.c_struct(rescale)
    dets = torch.empty((B, cfg.max_per_img, 5), dtype=torch.float32, device=dev)
    labels = torch.empty((B, cfg.max_per_img), dtype=torch.int64, device=dev)
    num = torch.empty((B,), dtype=torch.int32, device=dev)
    nws = lib.radet_get_bboxes_workspace_bytes(ctypes.byref(grid), B, num_classes, ctypes.byref(ccfg))
    ws = _workspace(("det", B, level_shapes, num_classes, cfg.nms_pre), nws, dev)
    check(lib.radet_get_bboxes(ctypes.byref(grid), B, num_classes, ctypes.byref(maps), _ptr(img_shapes), _ptr(scale_factors),
                               ctypes.byref(ccfg), _ptr(dets), _ptr(labels), _ptr(num), _ptr(ws), ws.numel(), _stream()),
          "radet_get_bboxes")
    return dets, labels, num


def get_candidates(geom: Geometry, num_classes: int, cls, bbox, iou, img_shapes: torch.Tensor, scale_factors: torch.Tensor,
                   cfg: DetectConfig, rescale=False):
    """with_nms=False branch (radet_head.py:165-169), batched: rows [B,cap,9] = (decoded box, score*centerness, prior box),
    labels [B,cap], num [B]; rows of an image ordered by class, then by descending score*centerness."""
    B = cls[0].shape[0]
    level_shapes = tuple(tuple(t.shape[-2:]) for t in cls)
    _check_maps(cls, bbox, iou, level_shapes, B, num_classes)
    _check_image_info(img_shapes, scale_factors, B, rescale)
    dev = cls[0].device
    cls = [t.contiguous() for t in cls]
    bbox = [t.contiguous() for t in bbox]
    iou = [t.contiguous() for t in iou]
    grid = geom.grid(level_shapes)
    lib = _lib.load()
    maps = _lib.make_maps([t.data_ptr() for t in cls], [t.data_ptr() for t in bbox], [t.data_ptr() for t in iou])
    ccfg = cfg.c_struct(rescale)
    cap = int(lib.radet_candidates_capacity(ctypes.byref(grid), num_classes, cfg.nms_pre))
    rows = torch.empty((B, cap, 9), dtype=torch.float32, device=dev)
    labels = torch.empty((B, cap), dtype=torch.int64, device=dev)
    num = torch.empty((B,), dtype=torch.int32, device=dev)
    nws = lib.radet_get_bboxes_workspace_bytes(ctypes.byref(grid), B, num_classes, ctypes.byref(ccfg))
    ws = _workspace(("det", B, level_shapes, num_classes, cfg.nms_pre), nws, dev)
    check(lib.radet_get_candidates(ctypes.byref(grid), B, num_classes, ctypes.byref(maps), _ptr(img_shapes), _ptr(scale_factors),
                                   ctypes.byref(ccfg), _ptr(rows), _ptr(labels), _ptr(num), _ptr(ws), ws.numel(), _stream()),
          "radet_get_candidates")
    return rows, labels, num


def bbox2result_batch(dets: torch.Tensor, labels: torch.Tensor, num: torch.Tensor, num_classes: int, xywh: bool = False):
    """Batched bbox2result (core/bbox/transforms.py:99-116) on the device: dets [B,max,5], labels [B,max], num [B] ->
    class-sorted rows [B,max,5] and class offsets int32 [B,C+1] (one D2H copy then serves every per-class slice)."""
    _require_cuda(dets, "dets")
    B, mx = int(dets.shape[0]), int(dets.shape[1])
    dets = dets.contiguous().float()
    labels = labels.contiguous().to(torch.int64)
    num = num.contiguous().to(torch.int32)
    out = torch.empty_like(dets)
    off = torch.empty((B, num_classes + 1), dtype=torch.int32, device=dets.device)
    check(_lib.load().radet_bbox2result(_ptr(dets), _ptr(labels), _ptr(num), B, mx, num_classes, int(bool(xywh)), _ptr(out), _ptr(off),
                                        _stream()), "radet_bbox2result")
    return out, off


# ------------------------------------------------------------------------------------------------ head-tower epilogues (f4)
class _GnRelu(torch.autograd.Function):
    """GroupNorm + ReLU of a tower layer (atss_head.py:52-87) in one launch each way (csrc/tower.cu)."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps):
        _require_cuda(x, "x")
        if x.dtype != torch.float32 or x.dim() < 2:
            raise RadetError("gn_relu: float32 [N, C, ...] expected")
        x = x.contiguous()
        N, C = int(x.shape[0]), int(x.shape[1])
        hw = x.numel() // max(1, N * C)
        if C % groups:
            raise RadetError(f"gn_relu: {C} channels do not split into {groups} groups")
        w = None if weight is None else weight.detach().contiguous().float()
        b = None if bias is None else bias.detach().contiguous().float()
        y = torch.empty_like(x)
        mean = torch.empty((N, groups), dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        with torch.cuda.device(x.device):
            check(_lib.load().radet_gn_relu_forward(_ptr(x), _ptr(w), _ptr(b), N, C, hw, groups, float(eps), _ptr(y), _ptr(mean), _ptr(rstd),
                                                    _stream()), "radet_gn_relu_forward")
        ctx.save_for_backward(x, w, b, mean, rstd)
        ctx.groups, ctx.has = groups, (weight is not None, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, mean, rstd = ctx.saved_tensors
        N, C = int(x.shape[0]), int(x.shape[1])
        hw = x.numel() // max(1, N * C)
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        need_w, need_b = ctx.has[0] and ctx.needs_input_grad[1], ctx.has[1] and ctx.needs_input_grad[2]
        dg = torch.empty((N, C), dtype=torch.float32, device=x.device) if need_w else None
        db = torch.empty((N, C), dtype=torch.float32, device=x.device) if need_b else None
        with torch.cuda.device(x.device):
            check(_lib.load().radet_gn_relu_backward(_ptr(dy), _ptr(x), _ptr(w), _ptr(b), _ptr(mean), _ptr(rstd), N, C, hw, ctx.groups,
                                                     _ptr(dx), _ptr(dg), _ptr(db), _stream()), "radet_gn_relu_backward")
        return dx, (dg.sum(0) if need_w else None), (db.sum(0) if need_b else None), None, None


def gn_relu(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], num_groups: int, eps: float = 1e-5):
    """relu(group_norm(x, num_groups, weight, bias, eps)) fused (forward and backward one launch each)."""
    return _GnRelu.apply(x, weight, bias, int(num_groups), float(eps))


class _ScaleRelu(torch.autograd.Function):
    """relu(scale * x): mmcv Scale + F.relu of the regression branch (atss_head.py:141-143, radet_head.py:27-30)."""

    @staticmethod
    def forward(ctx, x, scale):
        _require_cuda(x, "x")
        if x.dtype != torch.float32:
            raise RadetError("scale_relu: float32 expected")
        x = x.contiguous()
        s = scale.detach().reshape(1).contiguous().float()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(_lib.load().radet_scale_relu_forward(_ptr(x), _ptr(s), x.numel(), _ptr(y), _stream()), "radet_scale_relu_forward")
        ctx.save_for_backward(x, s)
        ctx.scale_shape = scale.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, s = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        lib = _lib.load()
        part = torch.empty((max(1, lib.radet_scale_relu_partials(x.numel())),), dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.radet_scale_relu_backward(_ptr(dy), _ptr(x), _ptr(s), x.numel(), _ptr(dx), _ptr(part), _stream()),
                  "radet_scale_relu_backward")
        ds = part.sum().float().reshape(ctx.scale_shape) if x.numel() else torch.zeros(ctx.scale_shape, device=x.device)
        return dx, ds


def scale_relu(x: torch.Tensor, scale: torch.Tensor):
    """relu(scale * x) fused, scale a one-element parameter (mmcv Scale)."""
    return _ScaleRelu.apply(x, scale)
