"""LOSSES mirrors named by the configs (configs/bop/r50_ycbv_pbr.py:46-55).

Inside RADetHead.loss the three modules are folded into one fused kernel pair (csrc/loss.cu); the modules carry the
hyper-parameters (gamma, alpha, eps, loss_weight) exactly as the reference's do.  Called on their own they run the
element-wise CUDA kernels (radet_sigmoid_focal_loss / radet_giou_loss / radet_bce_with_logits: loss and derivative in
one launch) and weight / reduce the way losses/utils.py does:
  FocalLoss        models/losses/focal_loss.py:91-157
  GIoULoss         models/losses/iou_loss.py:319-354
  CrossEntropyLoss models/losses/cross_entropy_loss.py:128-201 (use_sigmoid=True only)
"""
import torch
import torch.nn as nn

from .. import functional as F
from .registry import LOSSES


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """losses/utils.py:26-52 (element-wise weight, then mean / sum / sum-over-avg_factor)."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        if reduction == 'mean':
            loss = loss.mean()
        elif reduction == 'sum':
            loss = loss.sum()
        elif reduction != 'none':
            raise ValueError(f"{reduction} is not a valid value for reduction")
    elif reduction == 'mean':
        loss = loss.sum() / avg_factor
    elif reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


@LOSSES.register_module()
class FocalLoss(nn.Module):
    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, 'Only sigmoid focal loss supported now.'
        self.use_sigmoid = use_sigmoid
        self.gamma = gamma
        self.alpha = alpha
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        """focal_loss.py:122-157 -> sigmoid_focal_loss :44-87 (standalone use; RADetHead.loss runs the fused kernel)."""
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        loss = F.sigmoid_focal_loss_elementwise(pred, target, self.gamma, self.alpha)
        if weight is not None:
            if weight.shape != loss.shape:
                if weight.size(0) == loss.size(0):
                    weight = weight.view(-1, 1)
                else:
                    assert weight.numel() == loss.numel()
                    weight = weight.view(loss.size(0), -1)
            assert weight.ndim == loss.ndim
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


@LOSSES.register_module()
class GIoULoss(nn.Module):
    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.eps = eps
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        """iou_loss.py:327-354 -> giou_loss :82-98."""
        if weight is not None and not torch.any(weight > 0):
            return (pred * weight).sum()  # 0
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        loss = F.giou_loss_elementwise(pred, target, self.eps)
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction, avg_factor)


@LOSSES.register_module()
class CrossEntropyLoss(nn.Module):
    def __init__(self, use_sigmoid=False, use_mask=False, reduction='mean', class_weight=None, loss_weight=1.0):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        if not use_sigmoid or class_weight is not None:
            raise NotImplementedError("radet_b200 implements CrossEntropyLoss(use_sigmoid=True) without class_weight "
                                      "(the IoU-prediction branch of RADetHead)")
        self.use_sigmoid = use_sigmoid
        self.use_mask = use_mask
        self.reduction = reduction
        self.loss_weight = loss_weight
        self.class_weight = class_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        """cross_entropy_loss.py:163-201 -> binary_cross_entropy :58-91."""
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override if reduction_override else self.reduction
        if cls_score.dim() != label.dim():      # _expand_onehot_labels, cross_entropy_loss.py:40-55
            C = cls_score.size(-1)
            valid = (label >= 0) & (label < C)
            onehot = torch.zeros((label.size(0), C), dtype=cls_score.dtype, device=cls_score.device)
            rows = torch.nonzero(valid, as_tuple=False).squeeze(1)
            if rows.numel() > 0:
                onehot[rows, label[rows]] = 1
            if weight is not None:
                weight = weight.view(-1, 1).expand(weight.size(0), C)
            label = onehot
        if weight is not None:
            weight = weight.float()
        loss = F.bce_with_logits_elementwise(cls_score, label.float())
        return self.loss_weight * weight_reduce_loss(loss, weight, reduction=reduction, avg_factor=avg_factor)
