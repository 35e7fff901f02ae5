"""Host-side mirror of the reference's plugin surface for the dense-head hot path.

Same registry names, constructor arguments, method signatures and error behaviour as RADet's own classes, so that
configs/bop/*.py build unchanged; the work is done by libradet_b200.so through radet_b200.functional.
"""
from . import ops
from .coders import AnchorGenerator, TBLRBBoxCoder
from .graphed import GraphedHotPath
from .head import RADetHead
from .losses import CrossEntropyLoss, FocalLoss, GIoULoss
from .pipelines import LabelAssignment, PackVisibleMaskGrid
from .registry import (ANCHOR_GENERATORS, BBOX_CODERS, HEADS, LOSSES, PIPELINES, ConfigDict, Registry, build_anchor_generator,
                       build_bbox_coder, build_from_cfg, build_head, build_loss)

__all__ = ['RADetHead', 'LabelAssignment', 'PackVisibleMaskGrid', 'GraphedHotPath', 'TBLRBBoxCoder', 'AnchorGenerator', 'FocalLoss', 'GIoULoss', 'CrossEntropyLoss',
           'ops', 'HEADS', 'LOSSES', 'BBOX_CODERS', 'ANCHOR_GENERATORS', 'PIPELINES', 'Registry', 'build_from_cfg', 'build_head',
           'build_loss', 'build_bbox_coder', 'build_anchor_generator', 'ConfigDict', 'install_into_reference']


def install_into_reference():
    """Register the B200 implementations into the reference's OWN registries (needs `radet` + mmcv importable):
    after this call `configs/bop/r50_ycbv_pbr.py` builds RADetHead / LabelAssignment from this package unchanged and
    `radet.ops.vote_nms` & co. resolve to the CUDA versions.  See INTEGRATION.md."""
    import radet.ops as rops
    from radet.datasets.builder import PIPELINES as R_PIPELINES
    from radet.models.builder import HEADS as R_HEADS

    R_HEADS.register_module(name='RADetHead', force=True, module=RADetHead)
    # `LabelAssignment` keeps its name and contract (one image per call, numpy global RNG parity) but needs a process that
    # may use CUDA: workers_per_gpu=0 or spawned workers; it refuses to run in a forked worker.  With the shipped
    # workers_per_gpu >= 4 replace the step by `PackVisibleMaskGrid` in train_pipeline: the assignment
    # then runs batched in RADetHead.forward_train.
    R_PIPELINES.register_module(name='LabelAssignment', force=True, module=LabelAssignment)
    R_PIPELINES.register_module(name='PackVisibleMaskGrid', force=True, module=PackVisibleMaskGrid)
    rops.vote_nms = ops.vote_nms
    rops.global_vote_nms = ops.global_vote_nms
    rops.cluster_nms = ops.cluster_nms
    return True
