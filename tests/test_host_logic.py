"""CPU: host-side mirror of the reference plugin surface (registries, config parsing, constructor contracts)."""
import numpy as np
import pytest
import torch

from radet_b200 import functional as F
from radet_b200 import plugin as P
from radet_b200 import sharding
from radet_b200 import synthetic as syn

# configs/bop/r50_ycbv_pbr.py:30-56 and :70-80, verbatim
HEAD_CFG = dict(
    type='RADetHead', num_classes=21, in_channels=256, stacked_convs=4, feat_channels=256, strides=[8, 16, 32, 64, 128],
    anchor_generator=dict(type='AnchorGenerator', ratios=[1.0], octave_base_scale=8, scales_per_octave=1, strides=[8, 16, 32, 64, 128]),
    bbox_coder=dict(type='TBLRBBoxCoder', normalizer=1 / 8),
    loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
    loss_bbox=dict(type='GIoULoss', loss_weight=2.0),
    loss_centerness=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0))
TEST_CFG = dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                nms=dict(type='vote', iou_threshold=0.65, cluster_score=['cls', 'iou'], vote_score=['iou', 'cls'], iou_enable=False,
                         sima=0.025), max_per_img=100)
# configs/base/datasets/bop_detection.py:19-32
ASSIGN_CFG = dict(type='LabelAssignment',
                  anchor_generator_cfg=dict(type='AnchorGenerator', ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                            strides=[8, 16, 32, 64, 128]),
                  neg_threshold=0.2, positive_num=10, adapt_positive_num=False, balance_sample=True)


def test_reference_config_builds_unchanged():
    head = P.build_head(dict(HEAD_CFG, train_cfg=None, test_cfg=TEST_CFG))
    assert head.num_classes == 21 and head.cls_out_channels == 21 and head.num_anchors == 1
    assert head.loss_iou is head.loss_centerness
    assert head.loss_cfg.w_bbox == 2.0 and head.loss_cfg.gamma == 2.0 and head.loss_cfg.alpha == 0.25
    assert head.geom.anchor_scale == 8.0 and head.geom.tblr_normalizer == 0.125
    la = P.build_from_cfg(ASSIGN_CFG, P.PIPELINES)
    assert la.positive_num == 10 and la.balance_sample is True and la.geom.mask_step == 8


def test_checkpoint_parameter_names_match_the_reference():
    """atss_head.py:52-87: released radet_*.pth must load (SURVEY §5 checkpoint row)."""
    head = P.build_head(dict(HEAD_CFG))
    keys = set(head.state_dict())
    want = {f"{t}_convs.{i}.{n}" for t in ("cls", "reg") for i in range(4) for n in ("conv.weight", "gn.weight", "gn.bias")}
    want |= {f"atss_{n}.{p}" for n in ("cls", "reg", "centerness") for p in ("weight", "bias")}
    want |= {f"scales.{i}.scale" for i in range(5)}
    assert keys == want
    assert head.atss_cls.weight.shape == (21, 256, 3, 3) and head.atss_reg.weight.shape == (4, 256, 3, 3)


def test_forward_shapes_and_relu():
    head = P.build_head(dict(HEAD_CFG, in_channels=16, feat_channels=32, stacked_convs=1,
                             norm_cfg=dict(type='GN', num_groups=4, requires_grad=True)))
    feats = [torch.randn(2, 16, h, w) for h, w in ((8, 10), (4, 5), (2, 3), (1, 2), (1, 1))]
    cls, box, iou = head(feats)
    assert [tuple(t.shape) for t in cls] == [(2, 21, h, w) for h, w in ((8, 10), (4, 5), (2, 3), (1, 2), (1, 1))]
    assert all(t.shape[1] == 4 and bool((t >= 0).all()) for t in box)      # radet_head.py:29
    assert all(t.shape[1] == 1 for t in iou)


def test_detect_config_reads_the_reference_test_cfg():
    d = F.DetectConfig.from_test_cfg(P.ConfigDict(TEST_CFG))
    assert (d.score_thr, d.nms_pre, d.max_per_img) == (0.05, 1000, 100)
    assert d.nms_mode == 0 and d.cs_mode == 0 and d.vs_mode == 0 and d.iou_enable is False
    assert abs(d.iou_threshold - 0.65) < 1e-12 and d.sigma == 0.025      # `sima` typo is ignored, like vote_wrapper.py:13
    g = F.DetectConfig.from_test_cfg(dict(TEST_CFG, nms=dict(type='global_vote', iou_threshold=0.5, cluster_score='cls', vote_score='iou')))
    assert g.nms_mode == 1 and g.cs_mode == 1 and g.vs_mode == 2
    n = F.DetectConfig.from_test_cfg(dict(TEST_CFG, nms=dict(type='nms', iou_threshold=0.6)))
    assert n.nms_mode == 2 and n.cs_mode == 0
    with pytest.raises(RuntimeError, match="Unexpected"):
        F.score_mode("bogus")                                              # vote_wrapper.py:21,30


def test_unsupported_surface_fails_loudly():
    with pytest.raises(NotImplementedError):
        P.build_from_cfg(dict(ASSIGN_CFG, random_sample_by_distance=False), P.PIPELINES)
    la = P.build_from_cfg(dict(ASSIGN_CFG, adapt_positive_num=True, multiply_samplepro_for_weight=True), P.PIPELINES)
    assert la.adapt_positive_num and la.multiply_sample_pro_for_weight        # both switches are on the device now
    with pytest.raises(NotImplementedError):
        P.build_from_cfg(dict(ASSIGN_CFG, ambiguous_sample='max_dis'), P.PIPELINES)     # crashes in the reference too
    with pytest.raises(NotImplementedError):
        P.build_anchor_generator(dict(type='AnchorGenerator', ratios=[0.5, 1.0], octave_base_scale=8, scales_per_octave=1, strides=[8]))
    with pytest.raises(NotImplementedError):
        P.build_head(dict(HEAD_CFG, loss_bbox=dict(type='GIoULoss', loss_weight=2.0), loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True)))
    with pytest.raises(KeyError):
        P.build_from_cfg(dict(type='NoSuchThing'), P.HEADS)


def test_anchor_generator_docstring_kat():
    # anchor_generator.py:40-55 (scale 9 instead of the config's 8, as in the docstring)
    ag = P.build_anchor_generator(dict(type='AnchorGenerator', strides=[16], ratios=[1.], scales=[1.], base_sizes=None))
    ag.scales = torch.Tensor([9. / 16])   # side 9 at stride 16
    a = ag.grid_anchors([(2, 2)], device='cpu')[0]
    assert torch.equal(a, torch.tensor([[-4.5, -4.5, 4.5, 4.5], [11.5, -4.5, 20.5, 4.5], [-4.5, 11.5, 4.5, 20.5], [11.5, 11.5, 20.5, 20.5]]))
    flags = ag.valid_flags([(2, 2)], (16, 32, 3), device='cpu')[0]
    assert flags.tolist() == [True, True, False, False]


def test_geometry_and_synthetic_shapes():
    g = F.Geometry()
    assert g.level_shapes(480, 640) == ((60, 80), (30, 40), (15, 20), (8, 10), (4, 5))
    assert g.level_shapes(960, 1280) == ((120, 160), (60, 80), (30, 40), (15, 20), (8, 10))
    assert g.num_points(g.level_shapes(480, 640)) == 6400 and g.num_points(g.level_shapes(960, 1280)) == 25580
    wl = syn.WORKLOADS["cfg2"]
    a, b = syn.make_batch(wl, 2), syn.make_batch(wl, 2)
    assert all(np.array_equal(x.gt_bboxes, y.gt_bboxes) and np.array_equal(x.masks, y.masks) for x, y in zip(a, b))   # seeded
    assert syn.sample_grid(a[0].masks).shape[1:] == (60, 80)


def test_image_sharding():
    assert [sharding.image_range(r, 4, 8) for r in range(4)] == [(0, 8), (8, 16), (16, 24), (24, 32)]
    with pytest.raises(ValueError):
        sharding.image_range(4, 4, 8)
    t = torch.tensor([3.0, 5.0])
    assert sharding.reduce_mean_(t) is t and t.tolist() == [3.0, 5.0]      # no process group: identity


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/radet"), reason="reference tree only exists in the build container")
def test_install_into_reference_registries():
    """the B200 classes replace the reference's registry entries and radet.ops functions."""
    from oracle import ref_shim

    ref_shim.install()
    import radet.ops as rops
    from radet.datasets.builder import PIPELINES as R_PIPELINES
    from radet.models.builder import HEADS as R_HEADS

    ref_head = R_HEADS.get("RADetHead")
    assert ref_head is not None and ref_head is not P.RADetHead
    P.install_into_reference()
    try:
        assert R_HEADS.get("RADetHead") is P.RADetHead
        assert R_PIPELINES.get("LabelAssignment") is P.LabelAssignment
        assert rops.vote_nms is P.ops.vote_nms and rops.cluster_nms is P.ops.cluster_nms
    finally:   # leave the reference as it was for the other tests of this process
        from radet.datasets.pipelines.label_assignment import LabelAssignment as RefLA
        from radet.ops.cluster import cluster_nms as c0
        from radet.ops.vote import global_vote_nms as g0, vote_nms as v0

        R_HEADS.register_module(name="RADetHead", force=True, module=ref_head)
        R_PIPELINES.register_module(name="LabelAssignment", force=True, module=RefLA)
        rops.vote_nms, rops.global_vote_nms, rops.cluster_nms = v0, g0, c0


def test_pack_visible_mask_grid_is_cpu_only_and_collect_compatible():
    """the worker-side step ships the stride-8 sample grid + a seed under the reference's own Collect keys."""
    from radet_b200.plugin.pipelines import is_mask_grid_handoff

    im = syn.make_batch(syn.WORKLOADS["cfg1"], 1)[0]
    step = P.build_from_cfg(dict(type="PackVisibleMaskGrid"), P.PIPELINES)
    np.random.seed(5)
    res = step(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels, distance_maps=im.masks))
    np.random.seed(5)
    want_seed = int(np.random.randint(0, 2 ** 31 - 1))                 # one legacy randint from the global generator
    assert res["points_to_gt_index"].dtype == np.uint8 and res["points_to_gt_index"].shape == (im.gt_bboxes.shape[0], 60, 80)
    assert np.array_equal(res["points_to_gt_index"], im.masks[:, ::8, ::8])
    assert res["points_weight"].dtype == np.int64 and res["points_weight"].tolist() == [want_seed]
    # after DefaultFormatBundle's to_tensor (formating.py:218-223) the head recognises the hand-off
    as_tensors = [torch.from_numpy(res["points_to_gt_index"])]
    assert is_mask_grid_handoff(as_tensors) and not is_mask_grid_handoff([torch.zeros(6400, dtype=torch.int64)])
    assert not is_mask_grid_handoff(None)
    # an explicit per-image seed from the results dict; empty images; graded maps are refused
    res = P.PackVisibleMaskGrid(seed_key="assign_seed")(dict(img_shape=(im.H, im.W, 3), gt_bboxes=np.zeros((0, 4), np.float32),
                                                             distance_maps=np.zeros((0, im.H, im.W), np.uint8), assign_seed=123))
    assert res["points_to_gt_index"].shape == (0, 60, 80) and res["points_weight"].tolist() == [123]
    graded = im.masks.copy()
    graded[0][graded[0] > 0] = 1
    graded[0, :, : im.W // 2][graded[0, :, : im.W // 2] > 0] = 2
    if (graded[0, ::8, ::8] == 1).any() and (graded[0, ::8, ::8] == 2).any():
        with pytest.raises(NotImplementedError, match="graded"):
            step(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels, distance_maps=graded))
    with pytest.raises(NotImplementedError):
        P.build_from_cfg(dict(ASSIGN_CFG, neg_threshold=0.0), P.PIPELINES)      # reference semantics at 0 differ


def test_head_refuses_training_without_an_assignment():
    head = P.build_head(HEAD_CFG)
    assert head.label_assignment_cfg == dict(positive_num=10, neg_threshold=0.2, adapt_positive_num=False, balance_sample=True)
    la = head.assigner()
    assert isinstance(la, P.LabelAssignment) and la.positive_num == 10 and la.balance_sample
    with pytest.raises(Exception, match="points_to_gt_index"):
        head.forward_train([torch.zeros(1, 8, 4, 4)] * 5, [dict(img_shape=(32, 32, 3))], [torch.zeros(0, 4)], [torch.zeros(0)])
    with pytest.raises(NotImplementedError, match="test-time augmentation"):
        head.aug_test([], [])
