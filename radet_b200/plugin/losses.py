"""LOSSES mirrors named by the configs (configs/bop/r50_ycbv_pbr.py:46-55).

Inside RADetHead.loss the three modules are folded into one fused kernel pair (csrc/loss.cu); the modules carry the
hyper-parameters (gamma, alpha, eps, loss_weight) exactly as the reference's do:
  FocalLoss        models/losses/focal_loss.py:91-157
  GIoULoss         models/losses/iou_loss.py:319-354
  CrossEntropyLoss models/losses/cross_entropy_loss.py:128-201 (use_sigmoid=True only)
"""
import torch.nn as nn

from .registry import LOSSES

_STANDALONE = ("standalone {0}.forward is not part of the accelerated path in this round: RADetHead.loss runs the fused "
               "CUDA kernel (radet_loss_fwd_bwd); there is no PyTorch fallback")


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
        raise NotImplementedError(_STANDALONE.format("FocalLoss"))


@LOSSES.register_module()
class GIoULoss(nn.Module):
    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.eps = eps
        self.reduction = reduction
        self.loss_weight = loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        raise NotImplementedError(_STANDALONE.format("GIoULoss"))


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
        raise NotImplementedError(_STANDALONE.format("CrossEntropyLoss"))
