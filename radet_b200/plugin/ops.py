"""`radet.ops` mirror: vote_nms, global_vote_nms, cluster_nms with the reference's signatures.

Reference: radet/ops/__init__.py:1-12, ops/vote/vote_wrapper.py:7-43,47-83, ops/cluster/cluster_wrapper.py:6-22.
The reference wrappers were fed `.cpu()` tensors by the head (radet_head.py:150-158) and ran single-threaded C++.
These accept CUDA tensors (zero copies) or CPU tensors / numpy arrays (staged to the current CUDA device and the
result brought back to where the inputs lived, so existing call sites keep working).  There is no CPU implementation.
"""
import numpy as np
import torch

from .. import _lib
from .. import functional as F


def _stage(x, dtype=None):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected a torch.Tensor or numpy array, got {type(x)}")
    was_cuda = x.is_cuda
    if not was_cuda:
        x = x.cuda(non_blocking=True)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x, was_cuda


def _scores(cls_scores, score_factor, typ):
    mode = F.score_mode(typ)   # raises RuntimeError("Unexpected ... score type") like vote_wrapper.py:21,30
    if mode == 0:
        return cls_scores * score_factor
    return cls_scores if mode == 1 else score_factor


def _vote(bboxes, cls_scores, labels, nms_cfg, score_factor, max_num, mode):
    cfg = nms_cfg.copy()
    thr = cfg.pop("iou_threshold", 0.6)
    cst = cfg.pop("cluster_score", "cls")
    vst = cfg.pop("vote_score", "iou")
    iou_enable = cfg.pop("iou_enable", False)
    sigma = cfg.pop("sigma", 0.025)
    boxes, on_gpu = _stage(bboxes, torch.float32)
    cls_s, _ = _stage(cls_scores, torch.float32)
    lab, _ = _stage(labels, torch.int64)
    sf = None
    if score_factor is not None:
        sf, _ = _stage(score_factor, torch.float32)
    cs = _scores(cls_s, sf, cst)       # one fp32 multiply, same rounding as the reference's torch op
    vs = _scores(cls_s, sf, vst)
    n = boxes.shape[0]
    dets, olab, _, num, _, _, _ = F.vote_nms_lists([n], boxes.reshape(-1, 4), cs, vs, lab, thr, mode=mode, iou_enable=iou_enable,
                                                   sigma=sigma, max_num=max_num if max_num > 0 else 0)
    k = int(num.item())
    dets, olab = dets[:k], olab[:k]
    if not on_gpu:
        dets, olab = dets.cpu(), olab.cpu()
    return dets, olab


def vote_nms(bboxes, cls_scores, labels, nms_cfg, score_factor=None, max_num=0):
    """radet.ops.vote_nms (vote_wrapper.py:7-43) -> (dets [k,5], labels [k])."""
    return _vote(bboxes, cls_scores, labels, nms_cfg, score_factor, max_num, _lib.NMS_VOTE)


def global_vote_nms(bboxes, cls_scores, labels, nms_cfg, score_factor=None, max_num=0):
    """radet.ops.global_vote_nms (vote_wrapper.py:47-83): at most one detection per class."""
    return _vote(bboxes, cls_scores, labels, nms_cfg, score_factor, max_num, _lib.NMS_GLOBAL_VOTE)


def cluster_nms(bboxes, scores, categories, iou_threshold=0.65):
    """radet.ops.cluster_nms (cluster_wrapper.py:6-22) -> (instance_ids [n] i64, clusters_num [n] i64)."""
    boxes, on_gpu = _stage(bboxes, torch.float32)
    sc, _ = _stage(scores, torch.float32)
    lab, _ = _stage(categories, torch.int64)
    n = boxes.shape[0]
    _, _, _, _, inst, cnum, _ = F.vote_nms_lists([n], boxes.reshape(-1, 4), sc, sc, lab, iou_threshold, mode=_lib.NMS_PLAIN,
                                                 want_clusters=True)
    if not on_gpu:
        inst, cnum = inst.cpu(), cnum.cpu()
    return inst, cnum


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    """mmcv.ops.batched_nms stand-in for the `else` branch of radet_head.py:159-163 (class-aware greedy NMS, strict
    `iou > thr`, no +1): returns (dets [k,5] sorted by score desc, keep [k]).  The keep set equals vote_nms's seed set."""
    cfg = dict(nms_cfg)
    cfg.pop("type", None)
    thr = cfg.pop("iou_threshold", 0.5)
    b, on_gpu = _stage(boxes, torch.float32)
    s, _ = _stage(scores, torch.float32)
    l, _ = _stage(idxs, torch.int64)
    if class_agnostic:
        l = torch.zeros_like(l)
    n = b.shape[0]
    dets, _, keep, num, _, _, _ = F.vote_nms_lists([n], b.reshape(-1, 4), s, s, l, thr, mode=_lib.NMS_PLAIN)
    k = int(num.item())
    dets, keep = dets[:k], keep[:k]
    if not on_gpu:
        dets, keep = dets.cpu(), keep.cpu()
    return dets, keep


def bbox2result(bboxes, labels, num_classes):
    """bbox2result (core/bbox/transforms.py:99-116): `[bboxes[labels == i, :] for i in range(num_classes)]` as numpy
    arrays, computed as one class-sort on the device and one copy back (CPU tensors / numpy are staged like vote_nms)."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes)]
    b = _stage(bboxes, torch.float32)[0].reshape(1, -1, 5)
    l = _stage(labels, torch.int64)[0].reshape(1, -1)
    n = torch.tensor([b.shape[1]], dtype=torch.int32, device=b.device)
    return bbox2result_batch(b, l, n, num_classes)[0]


def bbox2result_batch(dets, labels, num, num_classes, xywh=False):
    """Batched form for the [B,max,5] / [B,max] / [B] outputs of `functional.get_bboxes` / `GraphedHotPath`:
    list over images of the reference's per-class list.  xywh=True gives BOPDataset.xyxy2xywh boxes (bop.py:99-118)."""
    out, off = F.bbox2result_batch(dets, labels, num, num_classes, xywh)
    out, off = out.cpu().numpy(), off.cpu().numpy()
    return [[out[i, off[i, c]:off[i, c + 1]] for c in range(num_classes)] for i in range(out.shape[0])]
