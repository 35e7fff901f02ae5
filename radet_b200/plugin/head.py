"""HEADS mirror: RADetHead with the mmdet dense-head API (forward / forward_train / loss / get_targets / get_bboxes).

Reference: radet/models/dense_heads/radet_head.py:16-392 (RADetHead) over atss_head.py:26-145,326-387 (ATSSHead),
anchor_head.py:33-93,142-170 (AnchorHead) and base_dense_head.py:22-59.  The conv towers stay torch/cuDNN modules with
the reference's parameter names (cls_convs.*, reg_convs.*, atss_cls, atss_reg, atss_centerness, scales.*) so released
checkpoints load; everything after the last conv (targets, loss fwd+bwd, decode, NMS) runs in libradet_b200.so.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as TF

from .. import functional as F
from .._lib import RadetError
from .registry import HEADS, ConfigDict, build_anchor_generator, build_bbox_coder, build_loss


class Scale(nn.Module):
    """mmcv.cnn.Scale: learnable scalar (parameter name `scale`)."""

    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


class ConvModule(nn.Module):
    """mmcv.cnn.ConvModule(conv_cfg=None, norm_cfg=GN): conv(bias=False) + GN + ReLU, sub-module names `conv`, `gn`."""

    def __init__(self, cin, cout, k, stride=1, padding=0, norm_cfg=None):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=norm_cfg is None)
        self.gn = None
        if norm_cfg is not None:
            if norm_cfg.get("type") != "GN":
                raise NotImplementedError("RADetHead towers use GroupNorm (atss_head.py:32)")
            self.gn = nn.GroupNorm(norm_cfg["num_groups"], cout)
            for p in self.gn.parameters():
                p.requires_grad = norm_cfg.get("requires_grad", True)

    def forward(self, x):
        x = self.conv(x)
        if self.gn is not None:
            if x.is_cuda and x.dtype == torch.float32 and x.shape[1] // self.gn.num_groups <= 64:
                # GroupNorm + ReLU in one launch each way (csrc/tower.cu); the convolution stays cuDNN
                return F.gn_relu(x, self.gn.weight, self.gn.bias, self.gn.num_groups, self.gn.eps)
            x = self.gn(x)
        return TF.relu(x, inplace=True)


def multi_apply(func, *args, **kwargs):
    """core/utils/misc.py:7-26."""
    from functools import partial

    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


@HEADS.register_module()
class RADetHead(nn.Module):
    def __init__(self, num_classes, in_channels, strides=(8, 16, 32, 64, 128), stacked_convs=4, conv_cfg=None,
                 quality='centerness', norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
                 loss_centerness=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0), feat_channels=256,
                 anchor_generator=dict(type='AnchorGenerator', ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                       strides=[8, 16, 32, 64, 128]),
                 bbox_coder=dict(type='TBLRBBoxCoder', normalizer=1 / 8), reg_decoded_bbox=False,
                 loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                 loss_bbox=dict(type='GIoULoss', loss_weight=2.0), train_cfg=None, test_cfg=None,
                 regress_ranges=F.REGRESS_RANGES, sync_num_pos=False, results_device='reference',
                 label_assignment=dict(positive_num=10, neg_threshold=0.2, adapt_positive_num=False, balance_sample=True)):
        super().__init__()
        if conv_cfg is not None:
            raise NotImplementedError("conv_cfg (DCN etc.) is outside the implemented surface")
        self.strides = strides
        self.stacked_convs = stacked_convs
        self.conv_cfg = conv_cfg
        self.norm_cfg = norm_cfg
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.feat_channels = feat_channels
        self.use_sigmoid_cls = loss_cls.get('use_sigmoid', False)
        if not self.use_sigmoid_cls or loss_cls['type'] != 'FocalLoss':
            raise NotImplementedError("RADetHead.loss is fused for sigmoid FocalLoss (configs/bop/*.py)")
        if loss_bbox['type'] != 'GIoULoss' or loss_centerness['type'] != 'CrossEntropyLoss':
            raise NotImplementedError("RADetHead.loss is fused for GIoULoss + CrossEntropyLoss(use_sigmoid=True)")
        self.sampling = False
        self.cls_out_channels = num_classes
        if self.cls_out_channels <= 0:
            raise ValueError(f'num_classes={num_classes} is too small')
        self.reg_decoded_bbox = reg_decoded_bbox
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)
        self.loss_centerness = build_loss(loss_centerness)
        self.loss_iou = self.loss_centerness                       # radet_head.py:25
        self.quality = quality
        self.train_cfg = train_cfg   # train_cfg.assigner (MaxIoUAssigner) is constructed but never used by RADetHead
        self.test_cfg = ConfigDict(test_cfg) if isinstance(test_cfg, dict) and not isinstance(test_cfg, ConfigDict) else test_cfg
        self.fp16_enabled = False
        self.anchor_generator = build_anchor_generator(anchor_generator)
        self.num_anchors = self.anchor_generator.num_base_anchors[0]
        if tuple(s[0] for s in self.anchor_generator.strides) != tuple(strides):
            raise RadetError("anchor_generator.strides and head strides differ")
        self.geom = F.Geometry(strides, regress_ranges, self.anchor_generator.anchor_scale, self.bbox_coder.normalizer)
        self.loss_cfg = F.LossConfig(gamma=self.loss_cls.gamma, alpha=self.loss_cls.alpha, w_cls=self.loss_cls.loss_weight,
                                     w_bbox=self.loss_bbox.loss_weight, w_iou=self.loss_iou.loss_weight, eps=self.loss_bbox.eps)
        self.sync_num_pos = sync_num_pos       # opt-in FCOS-style reduce_mean; False = reference behaviour
        self.results_device = results_device   # 'reference': vote branches return CPU tensors like radet_head.py:150-158
        # the assignment the head runs itself when the loader hands over mask grids (PackVisibleMaskGrid) instead of
        # assigned indices; the defaults are the shipped pipeline's (configs/base/datasets/bop_detection.py:19-32)
        self.label_assignment_cfg = dict(label_assignment)
        self._assigner = None
        self._init_layers()

    # ------------------------------------------------------------------ layers (atss_head.py:52-98)
    def _init_layers(self):
        self.relu = nn.ReLU(inplace=True)
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        for i in range(self.stacked_convs):
            chn = self.in_channels if i == 0 else self.feat_channels
            self.cls_convs.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1, norm_cfg=self.norm_cfg))
            self.reg_convs.append(ConvModule(chn, self.feat_channels, 3, stride=1, padding=1, norm_cfg=self.norm_cfg))
        self.atss_cls = nn.Conv2d(self.feat_channels, self.num_anchors * self.cls_out_channels, 3, padding=1)
        self.atss_reg = nn.Conv2d(self.feat_channels, self.num_anchors * 4, 3, padding=1)
        self.atss_centerness = nn.Conv2d(self.feat_channels, self.num_anchors * 1, 3, padding=1)
        self.scales = nn.ModuleList([Scale(1.0) for _ in self.anchor_generator.strides])

    def init_weights(self):
        for m in list(self.cls_convs) + list(self.reg_convs):
            nn.init.normal_(m.conv.weight, std=0.01)
        bias_cls = float(-np.log((1 - 0.01) / 0.01))
        for m, b in ((self.atss_cls, bias_cls), (self.atss_reg, 0.0), (self.atss_centerness, 0.0)):
            nn.init.normal_(m.weight, std=0.01)
            nn.init.constant_(m.bias, b)

    # ------------------------------------------------------------------ forward (atss_head.py:100-145, radet_head.py:27-30)
    def forward(self, feats):
        return multi_apply(self.forward_single, feats, self.scales)

    def forward_single(self, x, scale):
        cls_feat = x
        reg_feat = x
        for cls_conv in self.cls_convs:
            cls_feat = cls_conv(cls_feat)
        for reg_conv in self.reg_convs:
            reg_feat = reg_conv(reg_feat)
        cls_score = self.atss_cls(cls_feat)
        reg = self.atss_reg(reg_feat)
        if reg.is_cuda and reg.dtype == torch.float32:      # Scale + ReLU epilogue in one launch (csrc/tower.cu)
            bbox_pred = F.scale_relu(reg, scale.scale)
        else:
            bbox_pred = TF.relu(scale(reg).float())
        iou_pred = self.atss_centerness(reg_feat)
        return cls_score, bbox_pred, iou_pred

    def assigner(self):
        if self._assigner is None:
            from .pipelines import LabelAssignment
            self._assigner = LabelAssignment(strides=self.strides, regress_ranges=self.geom.regress_ranges,
                                             anchor_generator_cfg=dict(type='AnchorGenerator', ratios=[1.0],
                                                                       octave_base_scale=int(self.geom.anchor_scale),
                                                                       scales_per_octave=1, strides=list(self.strides)),
                                             **self.label_assignment_cfg)
        return self._assigner

    def assign_from_mask_grids(self, img_metas, gt_bboxes, mask_grids, seeds):
        """The loader handed over PackVisibleMaskGrid's sample grids: run LabelAssignment for the whole batch here, on the
        training GPU (label_assignment.py:136-201; one launch pair for all images instead of ~25 ms of numpy per image in
        a worker).  seeds: per-image int tensors / ints (np.random.seed(seed) right before each image)."""
        shapes = [(int(m['img_shape'][0]), int(m['img_shape'][1])) for m in img_metas]
        dev = self.atss_cls.weight.device
        sd = torch.as_tensor([int(torch.as_tensor(s).reshape(-1)[0]) for s in seeds], dtype=torch.int64).to(dev, non_blocking=True)
        with torch.cuda.device(dev):
            wsum = torch.empty((len(shapes),), dtype=torch.float64, device=dev)
            idx, w, consumed = self.assigner().assign_batch(shapes, gt_bboxes, mask_grids, seeds=sd, weight_sums=wsum)
        # The per-image weight sums travel with the index tensor through the reference's loss() signature: with them the dense
        # loss kernel does not wait for num_pos (functional.loss_fwd_bwd, weight_sums).
        idx._radet_weight_sums = wsum
        return idx, w

    def forward_train(self, x, img_metas, gt_bboxes, gt_labels=None, points_to_gt_index=None, points_weight=None,
                      gt_bboxes_ignore=None, proposal_cfg=None, **kwargs):
        from .pipelines import is_mask_grid_handoff
        if is_mask_grid_handoff(points_to_gt_index):          # PackVisibleMaskGrid in the pipeline: assign here, batched
            points_to_gt_index, points_weight = self.assign_from_mask_grids(img_metas, gt_bboxes, points_to_gt_index, points_weight)
        elif points_to_gt_index is None and gt_labels is not None:
            raise RadetError("RADetHead.forward_train needs points_to_gt_index / points_weight: either the reference's assigned "
                             "indices (LabelAssignment in the pipeline) or PackVisibleMaskGrid's mask grids + seeds under the same keys")
        outs = self(x)
        if gt_labels is None:
            loss_inputs = outs + (gt_bboxes, img_metas)
        else:
            loss_inputs = outs + (gt_bboxes, gt_labels, points_to_gt_index, points_weight, img_metas)
        losses = self.loss(*loss_inputs, gt_bboxes_ignore=gt_bboxes_ignore)
        if proposal_cfg is None:
            return losses
        proposal_list = self.get_bboxes(*outs, img_metas, cfg=proposal_cfg)
        return losses, proposal_list

    def get_anchors(self, featmap_sizes, img_metas, device='cuda'):
        """anchor_head.py:142-170 (materialised priors; the fused kernels do not need them)."""
        mla = self.anchor_generator.grid_anchors(featmap_sizes, device)
        anchor_list = [mla for _ in range(len(img_metas))]
        valid = [self.anchor_generator.valid_flags(featmap_sizes, m['pad_shape'], device) for m in img_metas]
        return anchor_list, valid

    # ------------------------------------------------------------------ batch staging
    @staticmethod
    def _cat_gt(gt_bboxes, gt_labels, dev):
        counts = [int(b.shape[0]) for b in gt_bboxes]
        if sum(counts):
            boxes = torch.cat([b.reshape(-1, 4) for b in gt_bboxes]).to(device=dev, dtype=torch.float32)
            labels = torch.cat([l.reshape(-1) for l in gt_labels]).to(device=dev, dtype=torch.int64)
        else:
            boxes = torch.zeros((0, 4), dtype=torch.float32, device=dev)
            labels = torch.zeros((0,), dtype=torch.int64, device=dev)
        return counts, boxes, labels

    @staticmethod
    def _stack_assignment(points_to_gt_index, points_weight, dev):
        idx = points_to_gt_index if isinstance(points_to_gt_index, torch.Tensor) else torch.stack(list(points_to_gt_index))
        w = points_weight if isinstance(points_weight, torch.Tensor) else torch.stack(list(points_weight))
        return idx.to(device=dev, dtype=torch.int64), w.to(device=dev, dtype=torch.float32)

    # ------------------------------------------------------------------ loss (radet_head.py:173-288)
    def loss(self, cls_scores, bbox_preds, iou_preds, gt_bboxes, gt_labels, points_to_gt_index, points_weight, img_metas,
             gt_bboxes_ignore=None):
        assert len(cls_scores) == len(bbox_preds) == len(iou_preds)
        num_imgs = cls_scores[0].size(0)
        assert num_imgs == len(img_metas) == len(points_to_gt_index) == len(points_weight) == len(gt_labels) == len(gt_bboxes)
        dev = cls_scores[0].device
        counts, boxes, labels = self._cat_gt(gt_bboxes, gt_labels, dev)
        idx, w = self._stack_assignment(points_to_gt_index, points_weight, dev)
        self.num_level_anchors = [int(t.shape[-2] * t.shape[-1]) for t in cls_scores]     # radet_head.py:327-328
        group = None
        if self.sync_num_pos and torch.distributed.is_available() and torch.distributed.is_initialized():
            group = torch.distributed.group.WORLD
        wsum = getattr(points_to_gt_index, "_radet_weight_sums", None) if idx is points_to_gt_index else None
        losses, _ = F.head_loss(self.geom, self.num_classes, self.loss_cfg, list(cls_scores), list(bbox_preds), list(iou_preds),
                                counts, boxes, labels, idx, w, sync_group=group, weight_sums=wsum)
        return losses

    # ------------------------------------------------------------------ get_targets (radet_head.py:290-369)
    def get_targets(self, anchors_list, cls_scores, bbox_preds, gt_bboxes_list, gt_labels_list, points_to_gt_index_list,
                    points_weight_list, image_metas):
        num_imgs = len(anchors_list)
        assert len(anchors_list) == len(image_metas) == len(points_to_gt_index_list) \
            == len(points_weight_list) == len(gt_labels_list) == len(gt_bboxes_list)
        num_levels = len(anchors_list[0])
        assert num_levels == len(bbox_preds) == len(cls_scores)
        num_level_anchors = [anchors.size(0) for anchors in anchors_list[0]]
        self.num_level_anchors = num_level_anchors
        level_shapes = tuple(tuple(t.shape[-2:]) for t in cls_scores)
        assert [h * w for h, w in level_shapes] == num_level_anchors
        dev = cls_scores[0].device
        counts, boxes, labels = self._cat_gt(gt_bboxes_list, gt_labels_list, dev)
        idx, w = self._stack_assignment(points_to_gt_index_list, points_weight_list, dev)
        lab, tg, wt, anc = F.get_targets(self.geom, level_shapes, self.num_classes, counts, boxes, labels, idx, w)
        sizes = [num_imgs * n for n in num_level_anchors]
        return list(lab.split(sizes)), list(tg.split(sizes)), list(wt.split(sizes)), list(anc.split(sizes))

    # ------------------------------------------------------------------ single image / TTA surface
    def _get_bboxes_single(self, cls_scores, bbox_preds, centernesses, mlvl_anchors, img_shape, scale_factor, cfg, rescale=False,
                           with_nms=True):
        """RADetHead._get_bboxes_single (radet_head.py:55-169) for ONE image: per-level maps [C,h,w] / [4,h,w] / [1,h,w].
        `mlvl_anchors` is accepted for signature parity and ignored (the priors are closed-form in the kernels)."""
        meta = dict(img_shape=tuple(img_shape), scale_factor=scale_factor)
        return self.get_bboxes([t.unsqueeze(0) for t in cls_scores], [t.unsqueeze(0) for t in bbox_preds],
                               [t.unsqueeze(0) for t in centernesses], [meta], cfg=cfg, rescale=rescale, with_nms=with_nms)[0]

    def aug_test_bboxes(self, feats, img_metas, rescale=False):
        """dense_test_mixins.py:38-97.  In the reference this path cannot run for RADetHead: its with_nms=False rows are
        [n, 9] (box, score, prior; radet_head.py:165-169), while `merge_aug_bboxes` -> `bbox_mapping_back` views its input as
        (-1, 4) boxes and `bbox_flip` asserts `shape[-1] % 4 == 0` (core/bbox/transforms.py:5-17,46-58).  The per-augmentation candidates are available from
        `get_bboxes(..., with_nms=False)`; merging them is left to the caller."""
        raise NotImplementedError("test-time augmentation is not functional for RADetHead in the reference either "
                                  "(bbox_mapping_back rejects the [n,9] candidate rows); use get_bboxes(with_nms=False) per "
                                  "augmentation and merge the rows yourself")

    def aug_test(self, feats, img_metas, rescale=False):
        return self.aug_test_bboxes(feats, img_metas, rescale=rescale)

    # ------------------------------------------------------------------ get_bboxes (atss_head.py:326-387, radet_head.py:55-169)
    def get_bboxes(self, cls_scores, bbox_preds, centernesses, img_metas, cfg=None, rescale=False, with_nms=True):
        cfg = self.test_cfg if cfg is None else cfg
        assert len(cls_scores) == len(bbox_preds)
        dcfg = F.DetectConfig.from_test_cfg(cfg)
        dev = cls_scores[0].device
        B = len(img_metas)
        assert B == cls_scores[0].shape[0], "one img_meta per image of the batch (atss_head.py:369 loops over img_metas)"
        shp = np.asarray([[m['img_shape'][0], m['img_shape'][1]] for m in img_metas], np.int32)
        sf = np.asarray([np.broadcast_to(np.asarray(m.get('scale_factor', 1.0), np.float32), (4,)) for m in img_metas], np.float32)
        shp_d = torch.from_numpy(shp).to(dev, non_blocking=True)
        sf_d = torch.from_numpy(sf).to(dev, non_blocking=True)
        if not with_nms:    # radet_head.py:165-169: [boxes, score*centerness, anchors] rows + categories, no suppression
            rows, cats, num = F.get_candidates(self.geom, self.num_classes, [t.detach() for t in cls_scores],
                                               [t.detach() for t in bbox_preds], [t.detach() for t in centernesses], shp_d,
                                               sf_d, dcfg, rescale=rescale)
            out = []
            for b, k in enumerate(num.cpu().tolist()):
                if k == 0:  # radet_head.py:137-138
                    out.append((torch.empty((0, 5)), torch.empty((0, 1), dtype=torch.int)))
                else:
                    out.append((rows[b, :k], cats[b, :k]))
            return out
        dets, labels, num = F.get_bboxes(self.geom, self.num_classes, [t.detach() for t in cls_scores],
                                         [t.detach() for t in bbox_preds], [t.detach() for t in centernesses], shp_d, sf_d,
                                         dcfg, rescale=rescale)
        to_cpu = self.results_device == 'reference' and dcfg.nms_mode != 2   # vote branches return CPU tensors
        if to_cpu:
            dets, labels = dets.cpu(), labels.cpu()
        num_h = num.cpu().tolist()
        out = []
        for b in range(B):
            k = num_h[b]
            if k == 0:   # radet_head.py:137-138
                out.append((torch.empty((0, 5)), torch.empty((0, 1), dtype=torch.int)))
            else:
                out.append((dets[b, :k], labels[b, :k]))
        return out
