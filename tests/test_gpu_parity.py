"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the committed golden vectors.

Bars: bit-exact for indices, labels, weights, targets, NMS keep sets and voted boxes; loss values within 1e-5
relative, gradients within rtol 1e-4 + atol 1e-6*max|grad| of the fp32 oracle (and 2e-5 / 1e-7 of the fp64 oracle
on the small case).
"""
import numpy as np
import pytest
import torch

from oracle import radet_oracle as orc
from radet_b200 import functional as F
from radet_b200 import plugin as P
from radet_b200 import synthetic as syn
from tests import helpers as hp

pytestmark = pytest.mark.gpu
GEOM = F.Geometry()
DEV = "cuda"


def cuda_sigmoid(x):
    return torch.sigmoid(torch.from_numpy(x).to(DEV)).cpu().numpy()


@pytest.fixture(autouse=True)
def _oracle_sigmoid():
    # the reference runs .sigmoid() on CUDA tensors (radet_head.py:106-109); the oracle mirrors that here
    orc.set_sigmoid(cuda_sigmoid)
    yield
    orc.set_sigmoid(None)


# ------------------------------------------------------------------------------------------------ RNG
def test_mt19937_matches_numpy():
    seeds = [0, 1, 777, 12345, 2 ** 31 - 1, 4000000000]
    n = 1500
    out = F.mt19937_uniforms(torch.tensor(seeds, dtype=torch.int64, device=DEV), n).cpu().numpy()
    for i, s in enumerate(seeds):
        assert np.array_equal(out[i], np.random.RandomState(s).random_sample(n)), s


# ------------------------------------------------------------------------------------------------ assignment
def _assign_gpu(cases, mode="seeds"):
    """cases: list of dicts from helpers.image_for sharing (H, W)."""
    H, W = cases[0]["H"], cases[0]["W"]
    shapes = GEOM.level_shapes(H, W)
    counts = [c["boxes"].shape[0] for c in cases]
    gh, gw = cases[0]["grid"].shape[1:] if cases[0]["grid"].ndim == 3 else (H // 8, W // 8)
    boxes = torch.from_numpy(np.concatenate([c["boxes"].reshape(-1, 4) for c in cases]).astype(np.float32)).to(DEV)
    grids = torch.from_numpy(np.concatenate([c["grid"].reshape(-1, gh, gw) for c in cases])).to(DEV)
    bits = F.pack_masks(grids, 1, gh, gw) if grids.shape[0] else torch.zeros((0, gh, (gw + 31) // 32), dtype=torch.int32, device=DEV)
    kw = {}
    if mode == "seeds":
        kw["seeds"] = torch.tensor([c["seed"] for c in cases], dtype=torch.int64, device=DEV)
    elif mode == "uniforms":
        kw["uniforms"] = torch.from_numpy(np.stack([np.random.RandomState(c["seed"]).random_sample(4096) for c in cases])).to(DEV)
    idx, w, used = F.assign(GEOM, shapes, counts, boxes, bits, (gh, gw), **kw)
    return idx.cpu().numpy(), w.cpu().numpy(), used.cpu().numpy()


@pytest.mark.parametrize("mode", ["seeds", "uniforms"])
def test_assignment_golden_bit_exact(mode):
    g, names = hp.assign_cases()
    by_shape = {}
    for name in names:
        c = hp.image_for(g, name)
        c["name"] = name
        by_shape.setdefault((c["H"], c["W"]), []).append(c)
    for cases in by_shape.values():
        idx, w, used = _assign_gpu(cases, mode)
        for i, c in enumerate(cases):
            name = c["name"]
            assert np.array_equal(idx[i], g[f"{name}/idx"].astype(np.int64)), name
            assert np.array_equal(w[i], g[f"{name}/w"]), name
            rs = np.random.RandomState(c["seed"])
            rs.random_sample(int(used[i]))
            assert np.array_equal(rs.random_sample(2), g[f"{name}/tail"]), name   # stream position parity


def test_assignment_vs_oracle_many_images():
    # cfg3 clutter + cfg5 high-res, images beyond the golden set
    for key, first, B in (("cfg3", 8, 24), ("cfg5", 3, 6), ("cfg2", 8, 16)):
        wl = syn.WORKLOADS[key]
        batch = syn.make_batch(wl, B, first)
        cases = [dict(boxes=im.gt_bboxes, grid=syn.sample_grid(im.masks), H=im.H, W=im.W, seed=im.seed) for im in batch]
        idx, w, used = _assign_gpu(cases)
        for i, im in enumerate(batch):
            oi, ow, ou = orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed)
            assert np.array_equal(idx[i], oi), (key, i)
            assert np.array_equal(w[i], ow), (key, i)
            assert used[i] == ou


def test_assignment_many_gt_and_ragged():
    # up to 200 GT per image (multi-word bit sets), plus an empty image in the same batch
    wl = syn.Workload("many", 480, 640, 30, 3, 1, 1, 7)
    rs = np.random.RandomState(5)
    imgs = [syn.make_image(rs, 480, 640, 30, G, occluded_frac=0.15) for G in (200, 0, 70)]
    cases = [dict(boxes=im.gt_bboxes, grid=syn.sample_grid(im.masks), H=480, W=640, seed=100 + i) for i, im in enumerate(imgs)]
    idx, w, used = _assign_gpu(cases)
    for i, im in enumerate(imgs):
        oi, ow, ou = orc.assign_image_seeded(im.gt_bboxes, im.masks, 480, 640, 100 + i)
        assert np.array_equal(idx[i], oi) and np.array_equal(w[i], ow) and used[i] == ou, i


def test_uniform_stream_overflow_is_reported():
    wl = syn.WORKLOADS["cfg1"]
    im = syn.make_batch(wl, 1)[0]
    shapes = GEOM.level_shapes(im.H, im.W)
    grid = syn.sample_grid(im.masks)
    bits = F.pack_masks(torch.from_numpy(grid).to(DEV), 1, grid.shape[1], grid.shape[2])
    u = torch.rand((1, 5), dtype=torch.float64, device=DEV)
    _, _, used = F.assign(GEOM, shapes, [im.gt_bboxes.shape[0]], torch.from_numpy(im.gt_bboxes).to(DEV), bits, grid.shape[1:], uniforms=u)
    assert int(used[0]) == -1


def test_label_assignment_plugin_numpy_global_rng():
    """Drop-in contract: results dict in/out and numpy GLOBAL RNG state identical to the reference afterwards."""
    la = P.LabelAssignment(anchor_generator_cfg=dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                                     strides=[8, 16, 32, 64, 128]),
                           neg_threshold=0.2, positive_num=10, adapt_positive_num=False, balance_sample=True)
    g, names = hp.assign_cases()
    for name in ["cfg1_0", "cfg3_1", "cfg5_0", "edge_occluded_and_tiny", "edge_no_gt"]:
        if name.startswith("edge_"):
            c = hp.image_for(g, name)
            masks = np.zeros((c["boxes"].shape[0], c["H"], c["W"]), np.uint8)
            masks[:, ::8, ::8] = c["grid"]
            boxes, labels, H, W, seed = c["boxes"], c["labels"], c["H"], c["W"], c["seed"]
        else:
            key, i = name.rsplit("_", 1)
            im = syn.make_batch(syn.WORKLOADS[key], 1, int(i))[0]
            masks, boxes, labels, H, W, seed = im.masks, im.gt_bboxes, im.gt_labels, im.H, im.W, im.seed
        np.random.seed(seed)
        res = la(dict(img_shape=(H, W, 3), gt_bboxes=boxes, gt_labels=labels, distance_maps=masks))
        tail = np.random.random_sample(2)
        assert res["points_to_gt_index"].dtype == np.int64 and res["points_weight"].dtype == np.float32
        assert np.array_equal(res["points_to_gt_index"], g[f"{name}/idx"].astype(np.int64)), name
        assert np.array_equal(res["points_weight"], g[f"{name}/w"]), name
        assert np.array_equal(tail, g[f"{name}/tail"]), name


# ------------------------------------------------------------------------------------------------ targets + loss
def _head_inputs(key, B=None):
    wl, batch = hp.head_case(key)
    a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
    idx_l, w_l = [x[0] for x in a], [x[1] for x in a]
    ho = syn.make_head_outputs(wl, batch, idx_l)
    return wl, batch, idx_l, w_l, ho


def _to_dev(ho):
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return [T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou]


def _gt_dev(batch):
    counts = [im.gt_bboxes.shape[0] for im in batch]
    boxes = torch.from_numpy(np.concatenate([im.gt_bboxes for im in batch])).to(DEV)
    labels = torch.from_numpy(np.concatenate([im.gt_labels for im in batch])).to(DEV)
    return counts, boxes, labels


@pytest.mark.parametrize("key", ["small", "cfg1", "cfg3b2"])
def test_get_targets_bit_exact(key):
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    shapes = GEOM.level_shapes(wl.H, wl.W)
    counts, boxes, labels = _gt_dev(batch)
    idx = torch.from_numpy(np.stack(idx_l)).to(DEV)
    w = torch.from_numpy(np.stack(w_l)).to(DEV)
    lab, tg, wt, anc = F.get_targets(GEOM, shapes, wl.C, counts, boxes, labels, idx, w)
    olab, otg, owt, oanc = orc.get_targets([b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
    assert np.array_equal(lab.cpu().numpy(), np.concatenate(olab))
    assert np.array_equal(tg.cpu().numpy(), np.concatenate(otg))
    assert np.array_equal(wt.cpu().numpy(), np.concatenate(owt))
    assert np.array_equal(anc.cpu().numpy(), np.concatenate(oanc))
    g = hp.load("head.npz")
    sizes = [len(batch) * h * w_ for h, w_ in shapes]
    for l, part in enumerate(lab.split(sizes)):
        assert np.array_equal(part.cpu().numpy(), g[f"{key}/labels{l}"].astype(np.int64))


def _check_grads(mine, ref, rtol, atol_rel):
    for a, b in zip(mine, ref):
        a = a.cpu().numpy() if isinstance(a, torch.Tensor) else a
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol_rel * max(1e-30, float(np.abs(b).max())))


@pytest.mark.parametrize("key", ["small", "cfg1", "cfg3b2"])
def test_loss_forward_backward_vs_oracle_and_golden(key):
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    idx = torch.from_numpy(np.stack(idx_l)).to(DEV)
    w = torch.from_numpy(np.stack(w_l)).to(DEV)
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig())
    losses = losses.cpu().numpy()
    gt_b, gt_l = [b.gt_bboxes for b in batch], [b.gt_labels for b in batch]
    o32 = orc.head_loss(ho.cls, ho.bbox, ho.iou, gt_b, gt_l, idx_l, w_l, wl.C, wl.H, wl.W)
    g = hp.load("head.npz")
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o32[k]) <= 1e-5 * abs(o32[k]), (k, losses[i], o32[k])
        assert abs(losses[i] - float(g[f"{key}/{k}"])) <= 1e-5 * abs(float(g[f"{key}/{k}"])), k     # the reference itself
    assert losses[3] == o32["num_pos"]
    _check_grads(grads[0], o32["grad_cls"], 1e-4, 1e-6)
    _check_grads(grads[1], o32["grad_bbox"], 1e-4, 1e-6)
    _check_grads(grads[2], o32["grad_iou"], 1e-4, 1e-6)
    if key == "small":
        o64 = orc.head_loss(ho.cls, ho.bbox, ho.iou, gt_b, gt_l, idx_l, w_l, wl.C, wl.H, wl.W, dtype="float64")
        for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
            assert abs(losses[i] - o64[k]) <= 2e-6 * abs(o64[k]), k
        _check_grads(grads[0], o64["grad_cls"], 2e-5, 1e-7)
        _check_grads(grads[1], o64["grad_bbox"], 2e-5, 1e-6)
        _check_grads(grads[2], o64["grad_iou"], 2e-5, 1e-6)
        for l in range(5):   # golden gradients of the reference (fp32 autograd)
            _check_grads([grads[0][l], grads[1][l], grads[2][l]], [g[f"small/gcls{l}"], g[f"small/gbox{l}"], g[f"small/giou{l}"]], 1e-4, 1e-6)


def test_loss_autograd_through_head_api():
    """RADetHead.loss: dict of autograd-connected scalars; backward gives the fused gradients, scaled by upstream."""
    wl, batch, idx_l, w_l, ho = _head_inputs("small")
    head = P.build_head(dict(type="RADetHead", num_classes=wl.C, in_channels=8, feat_channels=8, stacked_convs=1,
                             norm_cfg=dict(type="GN", num_groups=4, requires_grad=True))).to(DEV)
    cls, bbox, iou = _to_dev(ho)
    for t in cls + bbox + iou:
        t.requires_grad_()
    T = lambda a: torch.from_numpy(a).to(DEV)
    metas = syn.img_metas(batch)
    out = head.loss(cls, bbox, iou, [T(b.gt_bboxes) for b in batch], [T(b.gt_labels) for b in batch], [T(i) for i in idx_l],
                    [T(w) for w in w_l], metas)
    assert set(out) == {"loss_cls", "loss_bbox", "loss_iou"}
    (out["loss_cls"] + 0.5 * out["loss_bbox"] + 2.0 * out["loss_iou"]).backward()
    o32 = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
    _check_grads([t.grad for t in cls], o32["grad_cls"], 1e-4, 1e-6)
    g = hp.load("head.npz")
    # bbox grads scale by 0.5, iou grads by 2 relative to the golden (which used the plain sum)
    _check_grads([t.grad * 2.0 for t in bbox], [g[f"small/gbox{l}"] for l in range(5)], 1e-4, 1e-6)
    _check_grads([t.grad * 0.5 for t in iou], [g[f"small/giou{l}"] for l in range(5)], 1e-4, 1e-6)


def test_loss_no_positive_branch():
    """num_pos == 0 (radet_head.py:279-281): loss_bbox / loss_iou are plain sums over the (ignored) positives."""
    wl, batch, idx_l, w_l, ho = _head_inputs("small")
    idx_l = [np.where(i > 0, 0, i) for i in idx_l]          # positives -> ignored (still members of pos_inds)
    w_l = [np.where(i >= 0, 0.0, w).astype(np.float32) for i, w in zip(idx_l, w_l)]
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, torch.from_numpy(np.stack(idx_l)).to(DEV),
                                   torch.from_numpy(np.stack(w_l)).to(DEV), F.LossConfig())
    o = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
    losses = losses.cpu().numpy()
    assert o["num_pos"] == 0 and losses[3] == 0
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o[k]) <= 1e-5 * abs(o[k]) + 1e-6, k
    _check_grads(grads[1], o["grad_bbox"], 1e-5, 1e-7)
    _check_grads(grads[2], o["grad_iou"], 1e-5, 1e-7)
    _check_grads(grads[0], o["grad_cls"], 1e-4, 1e-6)


def test_loss_odd_plane_sizes_scalar_path():
    """h*w not a multiple of 4 on some levels (100x100 image -> 13x13, 7x7, 4x4, 2x2, 1x1)."""
    wl = syn.Workload("odd", 100, 100, 7, 3, 2, 5, 11)
    batch = syn.make_batch(wl)
    a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
    idx_l, w_l = [x[0] for x in a], [x[1] for x in a]
    ho = syn.make_head_outputs(wl, batch, idx_l)
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, torch.from_numpy(np.stack(idx_l)).to(DEV),
                                   torch.from_numpy(np.stack(w_l)).to(DEV), F.LossConfig())
    o = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
    losses = losses.cpu().numpy()
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o[k]) <= 1e-5 * abs(o[k]), k
    _check_grads(grads[0], o["grad_cls"], 1e-4, 1e-6)
    _check_grads(grads[1], o["grad_bbox"], 1e-4, 1e-6)
    _check_grads(grads[2], o["grad_iou"], 1e-4, 1e-6)
    # and the assignment on the same odd geometry
    cases = [dict(boxes=im.gt_bboxes, grid=syn.sample_grid(im.masks), H=im.H, W=im.W, seed=im.seed) for im in batch]
    idx, w, used = _assign_gpu(cases)
    for i in range(len(batch)):
        assert np.array_equal(idx[i], idx_l[i]) and np.array_equal(w[i], w_l[i])


# ------------------------------------------------------------------------------------------------ inference
NMS_CFG = dict(iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"], iou_enable=False, sima=0.025)


def test_sigmoid_flavour_is_torch_cuda():
    """The kernel's 1/(1+expf(-x)) must be bit-identical to torch's CUDA sigmoid (what the reference runs)."""
    x = torch.randn(1, 1, 64, 64, device=DEV) * 4
    # route through get_bboxes with one class, thr 0: scores come back as score*ctr with ctr = sigmoid(0) = 0.5 -> exact
    geom = F.Geometry(strides=(8,), regress_ranges=((-1, 1e8),))
    cfg = F.DetectConfig(score_thr=0.0, nms_pre=-1, max_per_img=4096, nms_type="nms", iou_threshold=2.0)
    bbox = torch.zeros(1, 4, 64, 64, device=DEV)
    bbox[0, 0] = torch.arange(64 * 64, device=DEV).reshape(64, 64) * 0.001   # distinct boxes
    dets, labels, num = F.get_bboxes(geom, 1, [x], [bbox], [torch.zeros(1, 1, 64, 64, device=DEV)],
                                     torch.tensor([[512, 512]], dtype=torch.int32, device=DEV),
                                     torch.ones(1, 4, device=DEV), cfg)
    assert int(num[0]) == 4096
    got = np.sort(dets[0, :, 4].cpu().numpy())
    want = np.sort((torch.sigmoid(x).reshape(-1) * 0.5).cpu().numpy())
    assert np.array_equal(got, want)


@pytest.mark.parametrize("key", ["small", "cfg1", "cfg3b2"])
@pytest.mark.parametrize("typ", ["vote", "global_vote", "nms"])
@pytest.mark.parametrize("thr", [0.05, 0.1])
def test_get_bboxes_bit_exact_vs_oracle(key, typ, thr):
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    cls, bbox, iou = _to_dev(ho)
    cfg = F.DetectConfig(score_thr=thr, nms_type=typ, **{k: v for k, v in NMS_CFG.items() if k != "sima"})
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.full((len(batch), 4), 1.25, device=DEV)
    dets, labels, num = F.get_bboxes(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    dets, labels, num = dets.cpu().numpy(), labels.cpu().numpy(), num.cpu().numpy()
    for b, im in enumerate(batch):
        od, ol = orc.get_bboxes_image([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou], (im.H, im.W, 3),
                                      np.full(4, 1.25, np.float32), score_thr=thr, nms_cfg=dict(type=typ, **NMS_CFG))
        assert num[b] == od.shape[0], (key, typ, thr, b)
        assert np.array_equal(dets[b, :num[b]].view(np.uint32), od.view(np.uint32)), (key, typ, thr, b)
        assert np.array_equal(labels[b, :num[b]], ol)


@pytest.mark.parametrize("key", ["small", "cfg1"])
def test_get_bboxes_vs_reference_golden(key):
    """Against the reference run on CPU (torch-CPU sigmoid flavour): same detections up to the <=1 ulp score flavour."""
    g = hp.load("head.npz")
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    cls, bbox, iou = _to_dev(ho)
    head = P.build_head(dict(type="RADetHead", num_classes=wl.C, in_channels=8, feat_channels=8, stacked_convs=1,
                             norm_cfg=dict(type="GN", num_groups=4, requires_grad=True),
                             test_cfg=dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type="vote", **NMS_CFG), max_per_img=100)))
    res = head.get_bboxes(cls, bbox, iou, syn.img_metas(batch), rescale=True)
    for b, (d, l) in enumerate(res):
        assert not d.is_cuda and d.shape[1] == 5                      # vote branch hands back CPU tensors like the reference
        ref_d, ref_l = g[f"{key}/det_vote_0.05_{b}"], g[f"{key}/lab_vote_0.05_{b}"].astype(np.int64)
        assert d.shape == ref_d.shape
        assert np.array_equal(l.numpy(), ref_l)
        np.testing.assert_allclose(d.numpy(), ref_d, rtol=2e-6, atol=1e-4)


def test_get_bboxes_empty_and_ragged():
    wl, batch, idx_l, w_l, ho = _head_inputs("small")
    cls, bbox, iou = _to_dev(ho)
    cls = [c.clone() for c in cls]
    for c in cls:
        c[1] = -20.0                                                  # image 1: nothing above threshold
    head = P.build_head(dict(type="RADetHead", num_classes=wl.C, in_channels=8, feat_channels=8, stacked_convs=1,
                             norm_cfg=dict(type="GN", num_groups=4, requires_grad=True),
                             test_cfg=dict(nms_pre=1000, score_thr=0.05, nms=dict(type="vote", **NMS_CFG), max_per_img=100)))
    res = head.get_bboxes(cls, bbox, iou, syn.img_metas(batch), rescale=False)
    assert res[0][0].shape[0] > 0
    assert tuple(res[1][0].shape) == (0, 5) and tuple(res[1][1].shape) == (0, 1) and res[1][1].dtype == torch.int32


# ------------------------------------------------------------------------------------------------ radet.ops
@pytest.mark.parametrize("case", range(5))
@pytest.mark.parametrize("on_gpu", [True, False])
def test_ops_golden_bit_exact(case, on_gpu):
    g = hp.load("ops.npz")
    dev = DEV if on_gpu else "cpu"
    boxes = torch.from_numpy(g[f"c{case}/boxes"]).to(dev)
    labels = torch.from_numpy(g[f"c{case}/labels"].astype(np.int64)).to(dev)
    cls = torch.from_numpy(g[f"c{case}/cls"]).to(dev)
    ctr = torch.from_numpy(g[f"c{case}/ctr"]).to(dev)
    cfg = P.ConfigDict(type="vote", **NMS_CFG)
    for nm, fn in (("vote", P.ops.vote_nms), ("gvote", P.ops.global_vote_nms)):
        d, l = fn(boxes, cls, labels, cfg, score_factor=ctr, max_num=0)
        assert d.is_cuda == on_gpu
        assert np.array_equal(d.cpu().numpy().view(np.uint32), g[f"c{case}/{nm}_dets"].view(np.uint32)), nm
        assert np.array_equal(l.cpu().numpy(), g[f"c{case}/{nm}_labels"].astype(np.int64))
    d, l = P.ops.vote_nms(boxes, cls, labels, cfg, score_factor=ctr, max_num=7)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), g[f"c{case}/vote_dets"][:7].view(np.uint32))
    inst, cnum = P.ops.cluster_nms(g[f"c{case}/boxes"], g[f"c{case}/cls"] * g[f"c{case}/ctr"], g[f"c{case}/labels"].astype(np.int64), 0.65)
    assert np.array_equal(inst.numpy(), g[f"c{case}/inst"]) and np.array_equal(cnum.numpy(), g[f"c{case}/cnum"])


def test_vote_nms_large_list_global_memory_path_and_batch():
    """n > shared-memory capacity (5120) takes the global-memory variant; lists of different length in one launch."""
    rs = np.random.RandomState(99)
    lists = []
    for n, nobj, ncls in ((7000, 300, 4), (0, 1, 1), (513, 40, 30), (1, 1, 1)):
        ctr = rs.uniform(50, 590, (nobj, 2))
        wh = rs.uniform(20, 200, (nobj, 2))
        pick = rs.randint(0, nobj, n)
        c = ctr[pick] + rs.normal(0, 3, (n, 2))
        s = wh[pick] * np.exp(rs.normal(0, 0.06, (n, 2)))
        boxes = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32).reshape(-1, 4)
        labels = rs.randint(0, ncls, n).astype(np.int64)
        sc = rs.permutation(n).astype(np.float32) / max(n, 1) * 0.9 + 0.05     # tie-free
        vs = rs.uniform(0.1, 1, n).astype(np.float32)
        lists.append((boxes, sc, vs, labels))
    cat = lambda i, dt: torch.from_numpy(np.concatenate([l[i] for l in lists]).astype(dt)).to(DEV)
    counts = [l[0].shape[0] for l in lists]
    dets, olab, oidx, num, inst, cnum, off = F.vote_nms_lists(counts, cat(0, np.float32), cat(1, np.float32), cat(2, np.float32),
                                                              cat(3, np.int64), 0.65, want_clusters=True)
    dets, olab, num, inst, cnum = dets.cpu().numpy(), olab.cpu().numpy(), num.cpu().numpy(), inst.cpu().numpy(), cnum.cpu().numpy()
    for i, (boxes, sc, vs, labels) in enumerate(lists):
        ob, ol, os_, oi, oc = orc.vote_nms_c(boxes, sc, vs, labels, 0.65)
        k = ob.shape[0]
        assert num[i] == k, i
        r0 = off[i]
        assert np.array_equal(dets[r0:r0 + k, :4].view(np.uint32), ob.view(np.uint32)), i
        assert np.array_equal(dets[r0:r0 + k, 4], os_) and np.array_equal(olab[r0:r0 + k], ol)
        assert np.array_equal(inst[r0:r0 + counts[i]], oi) and np.array_equal(cnum[r0:r0 + counts[i]], oc)


def test_nms_properties_full_size():
    """Size-independent properties at cfg4 scale (B=64): idempotence of the keep set and score ordering."""
    wl = syn.WORKLOADS["cfg4"]
    B = 8
    batch = syn.make_batch(wl, B)
    a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
    ho = syn.make_head_outputs(wl, batch, [x[0] for x in a])
    cls, bbox, iou = _to_dev(ho)
    cfg = F.DetectConfig(score_thr=wl.score_thr, nms_type="nms", iou_threshold=0.65)
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.ones((B, 4), device=DEV)
    dets, labels, num = F.get_bboxes(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg)
    for b in range(B):
        k = int(num[b])
        d, l = dets[b, :k], labels[b, :k]
        assert bool((d[:-1, 4] >= d[1:, 4]).all())                                  # sorted by score
        d2, keep = P.ops.batched_nms(d[:, :4], d[:, 4], l, dict(type="nms", iou_threshold=0.65))
        assert d2.shape[0] == k and np.array_equal(keep.cpu().numpy(), np.arange(k))  # NMS of a kept set keeps everything


def test_graphed_hot_path_matches_eager_calls():
    """plugin.GraphedHotPath (pinned arena -> H2D -> one CUDA graph -> D2H) gives exactly what the call-by-call API gives."""
    wl, batch, idx_l, w_l, ho = _head_inputs("cfg1")
    la = P.LabelAssignment(anchor_generator_cfg=dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                                     strides=[8, 16, 32, 64, 128]),
                           neg_threshold=0.2, positive_num=10, adapt_positive_num=False, balance_sample=True)
    head = P.build_head(dict(type="RADetHead", num_classes=wl.C, in_channels=8, feat_channels=8, stacked_convs=1,
                             norm_cfg=dict(type="GN", num_groups=4, requires_grad=True),
                             test_cfg=dict(nms_pre=1000, score_thr=0.05, nms=dict(type="vote", **NMS_CFG), max_per_img=100)))
    g = P.GraphedHotPath(head, la, len(batch), (wl.H, wl.W), max_gt_per_image=32).capture()
    buf, views = g.new_host_arena()
    for rep in range(2):      # replay twice: the kernels re-arm their counters
        g.fill(views, [im.gt_bboxes for im in batch], [im.gt_labels for im in batch], [syn.sample_grid(im.masks) for im in batch],
               [im.seed for im in batch], ho.cls, ho.bbox, ho.iou)
        res = g.run(buf)
        o32 = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
        for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
            assert abs(float(res["losses"][i]) - o32[k]) <= 1e-5 * abs(o32[k]), (rep, k)
        _check_grads(g.grads[0], o32["grad_cls"], 1e-4, 1e-6)
        _check_grads(g.grads[1], o32["grad_bbox"], 1e-4, 1e-6)
        for b, im in enumerate(batch):
            od, ol = orc.get_bboxes_image([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou], (im.H, im.W, 3),
                                          np.ones(4, np.float32), score_thr=0.05, nms_cfg=dict(type="vote", **NMS_CFG))
            k = int(res["num"][b])
            assert k == od.shape[0]
            assert np.array_equal(res["dets"][b, :k].numpy().view(np.uint32), od.view(np.uint32))
            assert np.array_equal(res["labels"][b, :k].numpy(), ol)
            assert int(res["consumed"][b]) == orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed)[2]


# ------------------------------------------------------------------------------------------------ standalone LOSSES
from tests.test_oracle_golden import _LOSS_CASES, loss_case_inputs  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("kind,tag,kw", _LOSS_CASES)
def test_standalone_loss_modules(kind, tag, kw):
    """plugin FocalLoss / GIoULoss / CrossEntropyLoss called on their own (models/losses/*.py): value and autograd
    gradient against the reference's golden output (fp32) and the float64 oracle.  Tolerances: loss 1e-5 relative,
    gradient rtol 1e-4 + 1e-6 max|g| (the head's)."""
    g = hp.load("losses.npz")
    pred, target, weight, red, af, lw = loss_case_inputs(g, kind, kw)
    mod = {"focal": lambda: P.FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25), "giou": lambda: P.GIoULoss(eps=1e-6, loss_weight=2.0),
           "bce": lambda: P.CrossEntropyLoss(use_sigmoid=True)}[kind]()
    x = torch.from_numpy(pred).to(DEV).requires_grad_()
    t = torch.from_numpy(target).to(DEV)
    w = None if weight is None else torch.from_numpy(weight).to(DEV)
    out = mod(x, t, weight=w, avg_factor=af, reduction_override=red if tag != "zero_w4" else None)
    out.sum().backward()
    want64, grad64 = orc.standalone_loss(kind, pred, target, weight, red, af, loss_weight=lw, dtype="float64")
    for want, gw in ((g[f"{kind}/{tag}/loss"], g[f"{kind}/{tag}/grad"]), (want64, grad64)):
        np.testing.assert_allclose(out.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(x.grad.cpu().numpy(), gw, rtol=1e-4, atol=1e-6 * max(1e-30, float(np.abs(gw).max())))


@pytest.mark.gpu
def test_standalone_bce_class_index_labels():
    g = hp.load("losses.npz")
    x = torch.from_numpy(g["focal/pred"]).to(DEV).requires_grad_()
    out = P.CrossEntropyLoss(use_sigmoid=True)(x, torch.from_numpy(g["focal/target"].astype(np.int64)).to(DEV),
                                               weight=torch.from_numpy(g["focal/weight"]).to(DEV), avg_factor=11.0)
    out.backward()
    np.testing.assert_allclose(out.item(), g["bce/onehot/loss"], rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["bce/onehot/grad"], rtol=1e-4, atol=1e-8)


@pytest.mark.gpu
def test_standalone_losses_refuse_cpu_tensors():
    with pytest.raises(Exception):
        P.FocalLoss()(torch.zeros(4, 3), torch.zeros(4, dtype=torch.int64))


# ------------------------------------------------------------------------------------------------ with_nms=False
@pytest.mark.gpu
@pytest.mark.parametrize("key", ["small", "cfg1"])
def test_get_bboxes_without_nms_matches_reference(key):
    """RADetHead.get_bboxes(with_nms=False) (radet_head.py:165-169): [box, score*centerness, prior] rows + categories.
    Boxes, priors and categories bit-exact against the reference's golden rows (compared as sets: the reference's order
    is topk(sorted=False)'s); score*centerness within 2 ulp of the torch-CPU golden and bit-exact against the oracle."""
    g = hp.load("head.npz")
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    head = P.build_head(dict(type="RADetHead", num_classes=wl.C, in_channels=8, feat_channels=8, stacked_convs=1,
                             norm_cfg=dict(type="GN", num_groups=4, requires_grad=True),
                             test_cfg=dict(nms_pre=1000, score_thr=0.05, nms=dict(type="vote", iou_threshold=0.65), max_per_img=100))).to(DEV)
    cls, bbox, iou = _to_dev(ho)
    res = head.get_bboxes(cls, bbox, iou, syn.img_metas(batch), rescale=True, with_nms=False)
    assert len(res) == len(batch)
    for b, (rows, cats) in enumerate(res):
        rows, cats = rows.cpu().numpy(), cats.cpu().numpy()
        assert rows.shape[1] == 9 and cats.dtype == np.int64
        # documented order: class-major, descending score*centerness inside a class
        assert np.all(np.diff(cats) >= 0)
        same = np.diff(cats) == 0
        assert np.all(np.diff(rows[:, 4])[same] < 0)
        o = np.lexsort((cats, rows[:, 8], rows[:, 7], rows[:, 6], rows[:, 5]))
        ref = g[f"{key}/cand_{b}"]
        cols = [0, 1, 2, 3, 5, 6, 7, 8]
        assert rows.shape == ref.shape
        assert np.array_equal(rows[o][:, cols].view(np.uint32), ref[:, cols].view(np.uint32))
        # product of two sigmoids: the golden was made by torch-CPU kernels, each factor may sit 1 ulp away from the CUDA
        # flavour the device (and the reference on a GPU) computes; the oracle check below is bit-exact
        np.testing.assert_allclose(rows[o][:, 4], ref[:, 4], rtol=4e-7, atol=0)
        assert np.array_equal(cats[o], g[f"{key}/candlab_{b}"].astype(np.int64))
        im = batch[b]
        bx, sc, ctr, ocats, anc = orc.select_candidates([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou],
                                                        (im.H, im.W, 3), np.ones(4, np.float32), 0.05, 1000)
        full = np.concatenate([bx, (sc * ctr)[:, None], anc], 1).astype(np.float32)
        oo = np.lexsort((ocats, full[:, 8], full[:, 7], full[:, 6], full[:, 5]))
        assert np.array_equal(rows[o].view(np.uint32), full[oo].view(np.uint32))


@pytest.mark.gpu
def test_get_bboxes_without_nms_rescale_and_empty():
    wl, batch, idx_l, w_l, ho = _head_inputs("small")
    cls, bbox, iou = _to_dev(ho)
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.tensor([[2.0, 0.5, 2.0, 0.5]] * len(batch), dtype=torch.float32, device=DEV)
    cfg = F.DetectConfig(score_thr=0.05, nms_pre=7, nms_type="vote", iou_threshold=0.65)
    rows, cats, num = F.get_candidates(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    for b, im in enumerate(batch):
        bx, sc, ctr, ocats, anc = orc.select_candidates([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou],
                                                        (im.H, im.W, 3), np.array([2.0, 0.5, 2.0, 0.5], np.float32), 0.05, 7)
        full = np.concatenate([bx, (sc * ctr)[:, None], anc], 1).astype(np.float32)
        k = int(num[b])
        assert k == full.shape[0] and k <= 7 * 5
        r, c = rows[b, :k].cpu().numpy(), cats[b, :k].cpu().numpy()
        o = np.lexsort((c, r[:, 8], r[:, 7], r[:, 6], r[:, 5], r[:, 4]))
        oo = np.lexsort((ocats, full[:, 8], full[:, 7], full[:, 6], full[:, 5], full[:, 4]))
        assert np.array_equal(r[o].view(np.uint32), full[oo].view(np.uint32))
        assert np.array_equal(c[o], ocats[oo])
    # nothing above the threshold -> the reference's empty return
    cfg2 = F.DetectConfig(score_thr=0.9999999, nms_pre=1000, nms_type="vote", iou_threshold=0.65)
    _, _, num2 = F.get_candidates(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg2, rescale=False)
    assert int(num2.sum()) == 0
    # and a second call on the same workspace still works (counters re-armed)
    rows3, cats3, num3 = F.get_candidates(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    assert torch.equal(num3, num)
    for b in range(len(batch)):
        k = int(num[b])
        assert torch.equal(cats3[b, :k], cats[b, :k]) and torch.equal(rows3[b, :k], rows[b, :k])


# ------------------------------------------------------------------------------------------------ LabelAssignment options
from tests.test_oracle_golden import ASSIGN_OPT_IMAGES, ASSIGN_OPT_VARIANTS  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("variant", sorted(ASSIGN_OPT_VARIANTS))
def test_assignment_options_bit_exact(variant):
    """adapt_positive_num / multiply_samplepro_for_weight / balance_sample=False (label_assignment.py:88-95,104-116,127-128):
    indices, weights and the stream position bit-exact against the reference's golden output, all images in one launch."""
    g = hp.load("assign_opts.npz")
    kw = dict(ASSIGN_OPT_VARIANTS[variant])
    for key in ("cfg2", "cfg3", "cfg5"):
        ids = [i for k, i in ASSIGN_OPT_IMAGES if k == key]
        wl = syn.WORKLOADS[key]
        batch = syn.make_batch(wl, len(ids), ids[0])
        shapes = GEOM.level_shapes(wl.H, wl.W)
        counts = [im.gt_bboxes.shape[0] for im in batch]
        grids = torch.from_numpy(np.concatenate([syn.sample_grid(im.masks) for im in batch])).to(DEV)
        gh, gw = grids.shape[1:]
        boxes = torch.from_numpy(np.concatenate([im.gt_bboxes for im in batch])).to(DEV)
        seeds = torch.tensor([im.seed for im in batch], dtype=torch.int32, device=DEV)
        idx, w, used = F.assign(GEOM, shapes, counts, boxes, F.pack_masks(grids, 1, gh, gw), (gh, gw), seeds=seeds,
                                balance_sample=kw.get("balance_sample", True), adapt_positive_num=kw.get("adapt_positive_num", False),
                                multiply_samplepro_for_weight=kw.get("multiply_samplepro_for_weight", False))
        idx, w, used = idx.cpu().numpy(), w.cpu().numpy(), used.cpu().numpy()
        for b, i in enumerate(ids):
            name = f"{variant}/{key}_{i}"
            assert np.array_equal(idx[b], g[f"{name}/idx"].astype(np.int64)), name
            assert np.array_equal(w[b].view(np.uint32), g[f"{name}/w"].view(np.uint32)), name
            rs = np.random.RandomState(batch[b].seed)
            rs.random_sample(int(used[b]))
            assert np.array_equal(rs.random_sample(2), g[f"{name}/tail"]), name


@pytest.mark.gpu
def test_label_assignment_pipeline_with_options():
    """The PIPELINES entry with the optional switches on: same results dict as the reference, numpy stream advanced alike."""
    g = hp.load("assign_opts.npz")
    la = P.LabelAssignment(anchor_generator_cfg=dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                                     strides=[8, 16, 32, 64, 128]),
                           neg_threshold=0.2, positive_num=10, adapt_positive_num=True, balance_sample=True,
                           multiply_samplepro_for_weight=True)
    im = syn.make_batch(syn.WORKLOADS["cfg3"], 1, 1)[0]
    np.random.seed(im.seed)
    res = la(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels, distance_maps=im.masks))
    assert np.array_equal(res["points_to_gt_index"], g["both/cfg3_1/idx"].astype(np.int64))
    assert np.array_equal(res["points_weight"], g["both/cfg3_1/w"])
    assert np.array_equal(np.random.random_sample(2), g["both/cfg3_1/tail"])


# ------------------------------------------------------------------------------------------------ bbox2result
@pytest.mark.gpu
def test_bbox2result_matches_reference_semantics():
    """core/bbox/transforms.py:99-116: `[bboxes[labels == i, :] for i in range(num_classes)]`; batched + xywh variant."""
    rs = np.random.RandomState(5)
    B, mx, C = 3, 100, 21
    dets = rs.uniform(0, 600, (B, mx, 5)).astype(np.float32)
    labels = rs.randint(0, C, (B, mx)).astype(np.int64)
    labels[0, :5] = C                                   # out-of-range labels appear in no class list
    num = np.array([100, 37, 0], np.int32)
    res = P.ops.bbox2result_batch(torch.from_numpy(dets).to(DEV), torch.from_numpy(labels).to(DEV), torch.from_numpy(num).to(DEV), C)
    resx = P.ops.bbox2result_batch(torch.from_numpy(dets).to(DEV), torch.from_numpy(labels).to(DEV), torch.from_numpy(num).to(DEV), C,
                                   xywh=True)
    for b in range(B):
        d, l = dets[b, :num[b]], labels[b, :num[b]]
        want = [d[l == i, :] for i in range(C)]
        assert len(res[b]) == C
        for i in range(C):
            assert np.array_equal(res[b][i], want[i]), (b, i)
            w = want[i].copy()
            w[:, 2] = want[i][:, 2] - want[i][:, 0]
            w[:, 3] = want[i][:, 3] - want[i][:, 1]
            assert np.array_equal(resx[b][i], w)
    single = P.ops.bbox2result(torch.from_numpy(dets[1, :37]), torch.from_numpy(labels[1, :37]), C)      # CPU tensors are staged
    assert all(np.array_equal(a, d_) for a, d_ in zip(single, [dets[1, :37][labels[1, :37] == i] for i in range(C)]))
    empty = P.ops.bbox2result(torch.zeros((0, 5)), torch.zeros((0,), dtype=torch.int64), C)
    assert len(empty) == C and all(e.shape == (0, 5) for e in empty)


# ------------------------------------------------------------------------------------------------ full BASELINE sizes
def _device_assign(wl, batch, **kw):
    shapes = GEOM.level_shapes(wl.H, wl.W)
    counts = [im.gt_bboxes.shape[0] for im in batch]
    grids = torch.from_numpy(np.concatenate([syn.sample_grid(im.masks) for im in batch])).to(DEV)
    gh, gw = grids.shape[1:]
    boxes = torch.from_numpy(np.concatenate([im.gt_bboxes for im in batch])).to(DEV)
    seeds = torch.tensor([im.seed for im in batch], dtype=torch.int32, device=DEV)
    idx, w, used = F.assign(GEOM, shapes, counts, boxes, F.pack_masks(grids, 1, gh, gw), (gh, gw), seeds=seeds, **kw)
    return shapes, counts, boxes, idx, w, used


@pytest.mark.gpu
def test_full_size_cfg5_assignment_and_loss_properties():
    """cfg5 at its full per-GPU size (16 images of 1280x960, C=30, 5-30 GT): properties that do not need the oracle.
    Assignment: index range, weights of a GT's positives sum to positive_num, ignored weight 0 / negative weight 1,
    one launch of 16 == two launches of 8 (images shard without a collective).  Loss: the batch loss is the
    normaliser-weighted combination of the half-batch losses, gradients scale linearly with the upstream gradient."""
    wl = syn.WORKLOADS["cfg5"]
    batch = syn.make_batch(wl, wl.B)
    shapes, counts, boxes, idx, w, used = _device_assign(wl, batch)
    h = wl.B // 2
    _, c0, b0, idx0, w0, used0 = _device_assign(wl, batch[:h])
    _, c1, b1, idx1, w1, used1 = _device_assign(wl, batch[h:])
    assert torch.equal(idx, torch.cat([idx0, idx1])) and torch.equal(w, torch.cat([w0, w1])) and torch.equal(used, torch.cat([used0, used1]))
    idx_h, w_h = idx.cpu().numpy(), w.cpu().numpy()
    for b, im in enumerate(batch):
        G = im.gt_bboxes.shape[0]
        assert idx_h[b].min() >= -1 and idx_h[b].max() <= G
        assert np.all(w_h[b][idx_h[b] == -1] == 1.0) and np.all(w_h[b][idx_h[b] == 0] == 0.0)
        for g in range(1, G + 1):
            s = w_h[b][idx_h[b] == g].sum()
            assert s in (0.0, 10.0), (b, g, s)                        # a GT either found no candidate or drew 10 (with multiplicity)
        assert int(used[b]) >= 10 * int((np.bincount(idx_h[b][idx_h[b] > 0], minlength=G + 1)[1:] > 0).sum())
    # ---- loss at the 119.5 MB shape (device-generated head outputs: values are irrelevant for the properties)
    g = torch.Generator(device=DEV).manual_seed(7)
    cls = [torch.randn((wl.B, wl.C, hh, ww), device=DEV, generator=g) - 4.6 for hh, ww in shapes]
    bbox = [torch.relu(torch.randn((wl.B, 4, hh, ww), device=DEV, generator=g) + 1) for hh, ww in shapes]
    iou = [torch.randn((wl.B, 1, hh, ww), device=DEV, generator=g) for hh, ww in shapes]
    labels = torch.from_numpy(np.concatenate([im.gt_labels for im in batch])).to(DEV)
    lcfg = F.LossConfig()
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, lcfg)
    losses = losses.cpu().numpy().astype(np.float64)
    gsum = [float(t.double().abs().sum()) for t in grads[0]]
    parts = []
    for sl, cc, bb, ii, ww_ in ((slice(0, h), c0, b0, idx0, w0), (slice(h, None), c1, b1, idx1, w1)):
        lab = torch.from_numpy(np.concatenate([im.gt_labels for im in batch[sl]])).to(DEV)
        l_, _ = F.loss_fwd_bwd(GEOM, wl.C, [t[sl].contiguous() for t in cls], [t[sl].contiguous() for t in bbox],
                               [t[sl].contiguous() for t in iou], cc, bb, lab, ii, ww_, lcfg)
        parts.append(l_.cpu().numpy().astype(np.float64))
    n0, n1 = parts[0][3], parts[1][3]
    assert losses[3] == n0 + n1                                          # num_pos adds up exactly (sums of small integers)
    comb = (parts[0][0] * (n0 + h) + parts[1][0] * (n1 + h)) / (n0 + n1 + wl.B)    # loss_cls: avg_factor = num_pos + B
    assert abs(comb - losses[0]) <= 2e-6 * abs(losses[0])
    comb_iou = (parts[0][2] * n0 + parts[1][2] * n1) / (n0 + n1)          # loss_iou: avg_factor = num_pos
    assert abs(comb_iou - losses[2]) <= 2e-6 * abs(losses[2])
    up = torch.tensor([2.0, 4.0, 0.5], device=DEV)                        # powers of two: scaling is exact
    _, g2 = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, lcfg)
    F.scale_grads(GEOM, wl.C, g2, up)
    for k, f in enumerate((2.0, 4.0, 0.5)):
        for a, b_ in zip(g2[k], grads[k]):
            assert torch.equal(a, b_ * f)
    assert all(s > 0 for s in gsum)


@pytest.mark.gpu
def test_full_size_cfg4_inference_properties():
    """cfg4 at its full size (64 images, score_thr 0.1, nms_pre 1000, vote-NMS, max 100 per image): one launch of 64 ==
    two launches of 32; per image scores descending, labels in range, at most max_per_img, boxes inside the image."""
    wl = syn.WORKLOADS["cfg4"]
    shapes = GEOM.level_shapes(wl.H, wl.W)
    g = torch.Generator(device=DEV).manual_seed(11)
    cls = [torch.randn((wl.B, wl.C, hh, ww), device=DEV, generator=g) * 1.3 - 4.0 for hh, ww in shapes]
    bbox = [torch.relu(torch.randn((wl.B, 4, hh, ww), device=DEV, generator=g) * 2 + 3) for hh, ww in shapes]
    iou = [torch.randn((wl.B, 1, hh, ww), device=DEV, generator=g) for hh, ww in shapes]
    cfg = F.DetectConfig(score_thr=wl.score_thr, nms_pre=1000, max_per_img=100, nms_type="vote", iou_threshold=0.65,
                         cluster_score=["cls", "iou"], vote_score=["iou", "cls"])
    shp = torch.tensor([[wl.H, wl.W]] * wl.B, dtype=torch.int32, device=DEV)
    sf = torch.ones((wl.B, 4), device=DEV)
    dets, labels, num = F.get_bboxes(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    h = wl.B // 2
    for sl in (slice(0, h), slice(h, None)):
        d_, l_, n_ = F.get_bboxes(GEOM, wl.C, [t[sl].contiguous() for t in cls], [t[sl].contiguous() for t in bbox],
                                  [t[sl].contiguous() for t in iou], shp[sl].contiguous(), sf[sl].contiguous(), cfg, rescale=True)
        assert torch.equal(n_, num[sl])
        for b in range(h):
            k = int(n_[b])
            assert torch.equal(d_[b, :k].view(torch.int32), dets[sl][b, :k].view(torch.int32)) and torch.equal(l_[b, :k], labels[sl][b, :k])
    assert int(num.max()) <= 100 and int(num.min()) > 0
    for b in range(wl.B):
        k = int(num[b])
        d, l = dets[b, :k], labels[b, :k]
        assert bool((d[:-1, 4] >= d[1:, 4]).all()) and bool((l >= 0).all()) and bool((l < wl.C).all())
        ok = ~torch.isnan(d[:, :4]).any(1)                                 # a cluster voted from an empty sigma band is NaN, as in the reference
        e = 1e-3                                                           # (s*x)/s of clamped coordinates may round one ulp past the border
        assert bool((d[ok, 0] >= -e).all()) and bool((d[ok, 2] <= wl.W + e).all()) and bool((d[ok, 1] >= -e).all()) and bool((d[ok, 3] <= wl.H + e).all())


@pytest.mark.gpu
def test_get_bboxes_low_threshold_fills_the_select_staging():
    """score_thr so low that nearly every (point, class) passes: detect_select_kernel's staging buffer overflows into
    its direct path, the per-level top-k (nms_pre) has real work.  Candidates as sets + detections bit-exact vs the oracle."""
    wl, batch, idx_l, w_l, ho = _head_inputs("small")
    cls, bbox, iou = _to_dev(ho)
    thr = 1e-4
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.ones((len(batch), 4), device=DEV)
    cfg = F.DetectConfig(score_thr=thr, nms_pre=300, nms_type="vote", **{k: v for k, v in NMS_CFG.items() if k != "sima"})
    rows, cats, num = F.get_candidates(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    dets, labels, nd = F.get_bboxes(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    for b, im in enumerate(batch):
        maps = ([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou])
        assert float((torch.sigmoid(cls[0][b]) > thr).float().mean()) > 0.9          # the staging capacity is exceeded on level 0
        bx, sc, ctr, ocats, anc = orc.select_candidates(*maps, (im.H, im.W, 3), np.ones(4, np.float32), thr, 300)
        full = np.concatenate([bx, (sc * ctr)[:, None], anc], 1).astype(np.float32)
        k = int(num[b])
        assert k == full.shape[0]
        r, c = rows[b, :k].cpu().numpy(), cats[b, :k].cpu().numpy()
        o = np.lexsort((c, r[:, 8], r[:, 7], r[:, 6], r[:, 5], r[:, 4]))
        oo = np.lexsort((ocats, full[:, 8], full[:, 7], full[:, 6], full[:, 5], full[:, 4]))
        assert np.array_equal(r[o].view(np.uint32), full[oo].view(np.uint32)) and np.array_equal(c[o], ocats[oo])
        od, ol = orc.get_bboxes_image(*maps, (im.H, im.W, 3), np.ones(4, np.float32), score_thr=thr, nms_pre=300,
                                      nms_cfg=dict(type="vote", **NMS_CFG))
        kk = int(nd[b])
        assert kk == od.shape[0]
        assert np.array_equal(dets[b, :kk].cpu().numpy().view(np.uint32), od.view(np.uint32)) and np.array_equal(labels[b, :kk].cpu().numpy(), ol)


@pytest.mark.gpu
@pytest.mark.parametrize("nms_pre", [300, 3000, -1])
def test_get_candidates_many_candidates_per_level(nms_pre):
    """~100 k candidates on the finest level of a 640x480 image (score_thr 1e-4): detect_bin_kernel reads the keys beyond
    its register-resident 6 144 from memory, and the three ways it hands out class slots all run -- kept keys compacted in
    shared memory (nms_pre 300), more kept keys than the staging area holds (nms_pre 3 000) and no per-level limit (-1).
    The rows [box, score*centerness, prior] + class are compared with the oracle as sets, bit for bit."""
    wl, batch, idx_l, w_l, ho = _head_inputs("cfg1")
    cls, bbox, iou = _to_dev(ho)
    thr = 1e-4
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.ones((len(batch), 4), device=DEV)
    cfg = F.DetectConfig(score_thr=thr, nms_pre=nms_pre, nms_type="vote", **{k: v for k, v in NMS_CFG.items() if k != "sima"})
    rows, cats, num = F.get_candidates(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    for b, im in enumerate(batch):
        maps = ([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou])
        assert int((torch.sigmoid(cls[0][b]) > thr).sum()) > 6144 * 4
        bx, sc, ctr, ocats, anc = orc.select_candidates(*maps, (im.H, im.W, 3), np.ones(4, np.float32), thr, nms_pre)
        full = np.concatenate([bx, (sc * ctr)[:, None], anc], 1).astype(np.float32)
        k = int(num[b])
        assert k == full.shape[0]
        r, c = rows[b, :k].cpu().numpy(), cats[b, :k].cpu().numpy()
        o = np.lexsort((c, r[:, 8], r[:, 7], r[:, 6], r[:, 5], r[:, 4]))
        oo = np.lexsort((ocats, full[:, 8], full[:, 7], full[:, 6], full[:, 5], full[:, 4]))
        assert np.array_equal(r[o].view(np.uint32), full[oo].view(np.uint32)) and np.array_equal(c[o], ocats[oo])


@pytest.mark.gpu
def test_whole_path_coco_like_shape():
    """80 classes, 40 GT per image (two 32-bit GT words), a 333x500 image (every level size odd, planes not 16-byte
    tileable -> the register-pipelined dense kernel, the scalar select path): assignment bit-exact, loss / gradients
    within the head tolerances, vote-NMS detections bit-exact against the oracle."""
    wl = syn.Workload("coco_like_333x500_B3_C80_G40", 333, 500, 80, 3, 40, 40, 17)
    batch = syn.make_batch(wl)
    a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
    idx_l, w_l = [x[0] for x in a], [x[1] for x in a]
    cases = [dict(boxes=im.gt_bboxes, grid=syn.sample_grid(im.masks), H=im.H, W=im.W, seed=im.seed) for im in batch]
    idx, w, used = _assign_gpu(cases)
    for i in range(len(batch)):
        assert np.array_equal(idx[i], idx_l[i]) and np.array_equal(w[i], w_l[i]) and int(used[i]) == a[i][2]
    ho = syn.make_head_outputs(wl, batch, idx_l)
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, torch.from_numpy(np.stack(idx_l)).to(DEV),
                                   torch.from_numpy(np.stack(w_l)).to(DEV), F.LossConfig())
    o = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
    losses = losses.cpu().numpy()
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o[k]) <= 1e-5 * abs(o[k]), k
    _check_grads(grads[0], o["grad_cls"], 1e-4, 1e-6)
    _check_grads(grads[1], o["grad_bbox"], 1e-4, 1e-6)
    _check_grads(grads[2], o["grad_iou"], 1e-4, 1e-6)
    cfg = F.DetectConfig(score_thr=0.05, nms_type="vote", **{k: v for k, v in NMS_CFG.items() if k != "sima"})
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.full((len(batch), 4), 0.8, device=DEV)
    dets, dl, num = F.get_bboxes(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    for b, im in enumerate(batch):
        od, ol = orc.get_bboxes_image([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou], (im.H, im.W, 3),
                                      np.full(4, 0.8, np.float32), score_thr=0.05, nms_cfg=dict(type="vote", **NMS_CFG))
        k = int(num[b])
        assert k == od.shape[0]
        assert np.array_equal(dets[b, :k].cpu().numpy().view(np.uint32), od.view(np.uint32)) and np.array_equal(dl[b, :k].cpu().numpy(), ol)


# ------------------------------------------------------------------------------------------------ iou_enable=True
@pytest.mark.gpu
@pytest.mark.parametrize("case", range(3))
def test_ops_iou_enable_golden_bit_exact(case):
    """vote_nms / global_vote_nms with the IoU score decay switched on (vote_ext.cpp:164-167: float exp, float multiply)
    against the reference's golden output; the shipped configs never enable it, so it has its own vectors."""
    g = hp.load("ops_iou.npz")
    boxes = torch.from_numpy(g[f"c{case}/boxes"]).to(DEV)
    labels = torch.from_numpy(g[f"c{case}/labels"].astype(np.int64)).to(DEV)
    cls = torch.from_numpy(g[f"c{case}/cls"]).to(DEV)
    ctr = torch.from_numpy(g[f"c{case}/ctr"]).to(DEV)
    cfg = P.ConfigDict(type="vote", iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"], iou_enable=True,
                       sigma=float(g[f"c{case}/sigma"]))
    for nm, fn in (("vote", P.ops.vote_nms), ("gvote", P.ops.global_vote_nms)):
        d, l = fn(boxes, cls, labels, cfg, score_factor=ctr, max_num=0)
        assert np.array_equal(d.cpu().numpy().view(np.uint32), g[f"c{case}/{nm}_dets"].view(np.uint32)), nm
        assert np.array_equal(l.cpu().numpy(), g[f"c{case}/{nm}_labels"].astype(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("typ", ["vote", "global_vote"])
def test_get_bboxes_iou_enable_vs_oracle(typ):
    """The same switch through the detection pipeline (class_nms_kernel's vote) against the oracle."""
    wl, batch, idx_l, w_l, ho = _head_inputs("cfg1")
    cls, bbox, iou = _to_dev(ho)
    nms = dict(iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"], iou_enable=True, sigma=0.3)
    cfg = F.DetectConfig(score_thr=0.05, nms_type=typ, **nms)
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.ones((len(batch), 4), device=DEV)
    dets, labels, num = F.get_bboxes(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    for b, im in enumerate(batch):
        od, ol = orc.get_bboxes_image([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou], (im.H, im.W, 3),
                                      np.ones(4, np.float32), score_thr=0.05, nms_cfg=dict(type=typ, **nms))
        k = int(num[b])
        assert k == od.shape[0]
        assert np.array_equal(dets[b, :k].cpu().numpy().view(np.uint32), od.view(np.uint32)) and np.array_equal(labels[b, :k].cpu().numpy(), ol)
@pytest.mark.gpu
@pytest.mark.parametrize("positive_num", [1, 3, 17, 32])
def test_assignment_other_positive_num(positive_num):
    """positive_num other than the config's 10 (1 and RADET_MAX_POSITIVE_NUM = 32 included), odd image size, against the
    oracle (itself checked against the live reference for 3 and 17 in tests/test_reference_live.py)."""
    wl = syn.Workload("pn_211x333_B4_C5_G0-12", 211, 333, 5, 4, 0, 12, 23)
    batch = syn.make_batch(wl)
    shapes, counts, boxes, idx, w, used = _device_assign(wl, batch, positive_num=positive_num)
    for b, im in enumerate(batch):
        oi, ow, ou = orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed, positive_num=positive_num)
        assert np.array_equal(idx[b].cpu().numpy(), oi) and np.array_equal(w[b].cpu().numpy(), ow) and int(used[b]) == ou


@pytest.mark.gpu
def test_tblr_coder_golden_bit_exact():
    """plugin.TBLRBBoxCoder.encode / decode (radet_tblr_encode / radet_tblr_decode) against the reference's golden output."""
    g = hp.load("coder.npz")
    pri, gts, tblr = (torch.from_numpy(g[k]).to(DEV) for k in ("priors", "gts", "tblr"))
    for nrm in (0.125, 4.0):
        coder = P.TBLRBBoxCoder(normalizer=nrm)
        assert np.array_equal(coder.encode(pri, gts).cpu().numpy().view(np.uint32), g[f"enc_{nrm}"].view(np.uint32))
        assert np.array_equal(coder.decode(pri, tblr).cpu().numpy().view(np.uint32), g[f"dec_{nrm}"].view(np.uint32))
        assert np.array_equal(coder.decode(pri, tblr, max_shape=(480, 640, 3)).cpu().numpy().view(np.uint32),
                              g[f"dec_clip_{nrm}"].view(np.uint32))
    nc = P.TBLRBBoxCoder(normalizer=0.125, clip_border=False)
    assert np.array_equal(nc.decode(pri, tblr, max_shape=(480, 640, 3)).cpu().numpy().view(np.uint32), g["dec_noclip_0.125"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["small", "odd"])
def test_loss_other_hyper_parameters(key):
    """gamma != 2 (the general-gamma instantiation of both dense kernels), other alpha / loss weights / eps, against the
    oracle (fp32 and fp64); `odd` has plane sizes that are not multiples of 4 (register kernel), `small` the TMA one."""
    if key == "odd":
        wl = syn.Workload("odd", 100, 100, 7, 3, 2, 5, 11)
        batch = syn.make_batch(wl)
        a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
        idx_l, w_l = [x[0] for x in a], [x[1] for x in a]
        ho = syn.make_head_outputs(wl, batch, idx_l)
    else:
        wl, batch, idx_l, w_l, ho = _head_inputs("small")
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    hp_ = dict(gamma=1.5, alpha=0.4, w_cls=0.7, w_bbox=1.3, w_iou=0.9, eps=1e-5)
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, torch.from_numpy(np.stack(idx_l)).to(DEV),
                                   torch.from_numpy(np.stack(w_l)).to(DEV), F.LossConfig(**hp_))
    gt_b, gt_l = [b.gt_bboxes for b in batch], [b.gt_labels for b in batch]
    losses = losses.cpu().numpy()
    for dtype, lt, gr, ga in (("float32", 1e-5, 1e-4, 1e-6), ("float64", 5e-6, 5e-5, 1e-6)):
        o = orc.head_loss(ho.cls, ho.bbox, ho.iou, gt_b, gt_l, idx_l, w_l, wl.C, wl.H, wl.W, dtype=dtype, **hp_)
        for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
            assert abs(losses[i] - o[k]) <= lt * abs(o[k]), (k, dtype)
        _check_grads(grads[0], o["grad_cls"], gr, ga)
        _check_grads(grads[1], o["grad_bbox"], gr, ga)
        _check_grads(grads[2], o["grad_iou"], gr, ga)


@pytest.mark.gpu
def test_standalone_focal_other_gamma():
    rs = np.random.RandomState(3)
    pred = rs.normal(-1, 3, (123, 9)).astype(np.float32)
    target = rs.randint(0, 10, 123).astype(np.int64)
    w = rs.uniform(0, 2, 123).astype(np.float32)
    for gamma, alpha in ((1.5, 0.4), (0.5, 0.25), (3.0, 0.75)):
        x = torch.from_numpy(pred).to(DEV).requires_grad_()
        out = P.FocalLoss(use_sigmoid=True, gamma=gamma, alpha=alpha, loss_weight=1.5)(x, torch.from_numpy(target).to(DEV),
                                                                                        torch.from_numpy(w).to(DEV), avg_factor=3.0)
        out.backward()
        l, g = orc.standalone_loss("focal", pred, target, w, "mean", 3.0, loss_weight=1.5, gamma=gamma, alpha=alpha, dtype="float64")
        np.testing.assert_allclose(out.item(), l, rtol=1e-5)
        np.testing.assert_allclose(x.grad.cpu().numpy(), g, rtol=1e-4, atol=1e-6 * float(np.abs(g).max()))


def _fused_loss(wl, cls, bbox, iou, counts, boxes, labels, idx, w, **kw):
    """The experimental single-launch kernel (loss_fused.cu) through the development switch."""
    import os
    os.environ["RADET_LOSS_IMPL"] = "fused"
    try:
        return F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), **kw)
    finally:
        del os.environ["RADET_LOSS_IMPL"]


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["small", "cfg1", "cfg3b2"])
def test_loss_single_launch_variant_vs_oracle(key):
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    idx = torch.from_numpy(np.stack(idx_l)).to(DEV)
    w = torch.from_numpy(np.stack(w_l)).to(DEV)
    losses, grads = _fused_loss(wl, cls, bbox, iou, counts, boxes, labels, idx, w)
    o32 = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
    losses = losses.cpu().numpy()
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o32[k]) <= 1e-5 * abs(o32[k]), (k, losses[i], o32[k])
    assert losses[3] == o32["num_pos"]
    _check_grads(grads[0], o32["grad_cls"], 1e-4, 1e-6)
    _check_grads(grads[1], o32["grad_bbox"], 1e-4, 1e-6)
    _check_grads(grads[2], o32["grad_iou"], 1e-4, 1e-6)
    # forward only, and a second call on the same workspace (the control block re-arms itself)
    l2, g2 = _fused_loss(wl, cls, bbox, iou, counts, boxes, labels, idx, w, want_grads=False)
    assert g2 is None and torch.equal(l2.cpu(), torch.from_numpy(losses))


@pytest.mark.gpu
def test_loss_two_phase_normaliser_sync_path():
    """sync_num_pos: radet_loss_fwd_bwd(phases=1) -> reduce_mean of the two normalisers -> phases=2.  In a single process
    the reduce is the identity, so the split call must give exactly what the one-call form gives (same kernels)."""
    wl, batch, idx_l, w_l, ho = _head_inputs("cfg1")
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    idx = torch.from_numpy(np.stack(idx_l)).to(DEV)
    w = torch.from_numpy(np.stack(w_l)).to(DEV)
    l1, g1 = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig())
    l2, g2 = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), sync_group="world")
    assert torch.equal(l1, l2)
    for a, b in zip(g1[0] + g1[1] + g1[2], g2[0] + g2[1] + g2[2]):
        assert torch.equal(a, b)
    head = P.build_head(dict(type="RADetHead", num_classes=wl.C, in_channels=8, feat_channels=8, stacked_convs=1,
                             norm_cfg=dict(type="GN", num_groups=4, requires_grad=True), sync_num_pos=True)).to(DEV)
    T = lambda a: torch.from_numpy(a).to(DEV)
    out = head.loss(cls, bbox, iou, [T(b.gt_bboxes) for b in batch], [T(b.gt_labels) for b in batch], [T(i) for i in idx_l],
                    [T(x) for x in w_l], syn.img_metas(batch))
    assert abs(float(out["loss_cls"]) - float(l1[0])) <= 1e-7 * abs(float(l1[0]))


@pytest.mark.gpu
def test_forward_train_assigns_from_mask_grids():
    """PackVisibleMaskGrid in the loader + RADetHead.forward_train on the trainer's GPU == the stock flow
    (LabelAssignment in the worker with np.random.seed(seed) before each image, then RADetHead.loss)."""
    wl, batch = hp.head_case("cfg1")
    head = P.build_head(dict(type="RADetHead", num_classes=wl.C, in_channels=8, feat_channels=8, stacked_convs=1,
                             norm_cfg=dict(type="GN", num_groups=4, requires_grad=True))).to(DEV)
    head.init_weights()
    shapes = GEOM.level_shapes(wl.H, wl.W)
    g = torch.Generator(device=DEV).manual_seed(3)
    feats = [torch.randn((len(batch), 8, h, w), device=DEV, generator=g) for h, w in shapes]
    metas = syn.img_metas(batch)
    pack = P.PackVisibleMaskGrid(seed_key="seed")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    grids, seeds = [], []
    for im in batch:                                   # what a DataLoader worker + DefaultFormatBundle hand over (CPU tensors)
        r = pack(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels, distance_maps=im.masks, seed=im.seed))
        grids.append(T(r["points_to_gt_index"]))
        seeds.append(T(r["points_weight"]))
    gtb, gtl = [T(im.gt_bboxes).to(DEV) for im in batch], [T(im.gt_labels).to(DEV) for im in batch]
    losses = head.forward_train(feats, metas, gtb, gtl, grids, seeds)
    # the stock flow: assignment by the oracle (= the reference's LabelAssignment, tests/test_reference_live.py) seeded per image
    a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
    want = head.forward_train(feats, metas, gtb, gtl, [T(x[0]).to(DEV) for x in a], [T(x[1]).to(DEV) for x in a])
    for k in ("loss_cls", "loss_bbox", "loss_iou"):
        assert torch.equal(losses[k], want[k]), k
    (losses["loss_cls"] + losses["loss_bbox"] + losses["loss_iou"]).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in head.atss_cls.parameters())


def _forked_label_assignment(q):
    from radet_b200._lib import RadetError
    la = P.LabelAssignment(anchor_generator_cfg=dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                                     strides=[8, 16, 32, 64, 128]), neg_threshold=0.2, balance_sample=True)
    try:
        la(dict(img_shape=(64, 64, 3), gt_bboxes=np.zeros((0, 4), np.float32), gt_labels=np.zeros(0, np.int64),
                distance_maps=np.zeros((0, 64, 64), np.uint8)))
        q.put("ran")
    except RadetError as e:
        q.put("refused: " + str(e)[:60])
    except Exception as e:  # noqa: BLE001
        q.put("other: " + repr(e)[:80])


@pytest.mark.gpu
def test_label_assignment_refuses_forked_dataloader_workers():
    """mmcv's default DataLoader workers are forked from a process that already uses CUDA: the GPU pipeline step must say so
    instead of dying in `Cannot re-initialize CUDA in forked subprocess`."""
    import multiprocessing as mp
    torch.zeros(1, device=DEV)                          # the parent has initialised CUDA
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_forked_label_assignment, args=(q,))
    p.start()
    msg = q.get(timeout=60)
    p.join(timeout=30)
    assert msg.startswith("refused: LabelAssignment.__call__ runs on the GPU"), msg


@pytest.mark.gpu
def test_head_loss_backward_is_reentrant():
    """a second backward through the fused loss node (retain_graph) must not compound the upstream factor."""
    wl, batch, idx_l, w_l, ho = _head_inputs("small")
    cls, bbox, iou = _to_dev(ho)
    for t in cls + bbox + iou:
        t.requires_grad_()
    counts, boxes, labels = _gt_dev(batch)
    idx = torch.from_numpy(np.stack(idx_l)).to(DEV)
    w = torch.from_numpy(np.stack(w_l)).to(DEV)
    out, _ = F.head_loss(GEOM, wl.C, F.LossConfig(), cls, bbox, iou, counts, boxes, labels, idx, w)
    total = 3.0 * out["loss_cls"] + 0.5 * out["loss_bbox"] + 2.0 * out["loss_iou"]
    g1 = torch.autograd.grad(total, cls + bbox + iou, retain_graph=True)
    g1 = [t.clone() for t in g1]
    g2 = torch.autograd.grad(total, cls + bbox + iou, retain_graph=True)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-6, atol=0), "second backward returned differently scaled gradients"
    g3 = torch.autograd.grad(2.0 * total, cls + bbox + iou)
    for a, b in zip(g1, g3):
        assert torch.allclose(2.0 * a, b, rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_grid_priors_on_the_device():
    """AnchorGenerator.grid_anchors / valid_flags (anchor_generator.py:206-298) through radet_grid_priors == the oracle's priors."""
    ag = P.build_anchor_generator(dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                       strides=[8, 16, 32, 64, 128]))
    for H, W in ((480, 640), (100, 100), (333, 500)):
        sizes = GEOM.level_shapes(H, W)
        anc = ag.grid_anchors(sizes, device=DEV)
        want = orc.grid_anchors(H, W)
        assert len(anc) == 5 and all(a.is_cuda for a in anc)
        assert np.array_equal(torch.cat(anc).cpu().numpy(), want)
        flags = ag.valid_flags(sizes, (H - 40, W - 90, 3), device=DEV)
        ref = ag.valid_flags(sizes, (H - 40, W - 90, 3), device="cpu")
        assert all(torch.equal(f.cpu(), r) for f, r in zip(flags, ref))


def _weight_sums(idx_l, w_l, batch):
    """What radet_assign(weight_sums=...) hands over: per image, the float64 sum of the weights of the points with index >= 0
    (0 for an image without ground truth)."""
    v = [float(w[i >= 0].astype(np.float64).sum()) if im.gt_bboxes.shape[0] > 0 else 0.0 for i, w, im in zip(idx_l, w_l, batch)]
    return torch.tensor(v, dtype=torch.float64, device=DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["small", "cfg1", "cfg3b2"])
@pytest.mark.parametrize("hyper", [dict(), dict(gamma=1.5, alpha=0.4, w_cls=0.7, w_bbox=1.3, w_iou=0.9, eps=1e-5)])
def test_loss_dense_first_order_vs_oracle(key, hyper):
    """With the assignment's per-image weight sums the dense kernel runs FIRST and loss_pos_kernel follows as its programmatic
    dependent (zero-fill / rescale of the regression planes, the three losses): same numbers as the oracle, upstream
    gradient scales applied, forward-only calls and repeated calls on one workspace."""
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    idx = torch.from_numpy(np.stack(idx_l)).to(DEV)
    w = torch.from_numpy(np.stack(w_l)).to(DEV)
    wsum = _weight_sums(idx_l, w_l, batch)
    cfg = F.LossConfig(**hyper)
    o = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W, **hyper)
    for rep in range(2):
        losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, cfg, weight_sums=wsum)
        losses = losses.cpu().numpy()
        for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
            assert abs(losses[i] - o[k]) <= 1e-5 * abs(o[k]), (k, losses[i], o[k], rep)
        assert losses[3] == o["num_pos"]
        _check_grads(grads[0], o["grad_cls"], 1e-4, 1e-6)
        _check_grads(grads[1], o["grad_bbox"], 1e-4, 1e-6)
        _check_grads(grads[2], o["grad_iou"], 1e-4, 1e-6)
    l2, g2 = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, cfg, weight_sums=wsum, want_grads=False)
    assert g2 is None and np.array_equal(l2.cpu().numpy(), losses)
    # the classic order (loss_pos_kernel first) gives the same losses to rounding and the same gradients
    l3, g3 = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, cfg)
    np.testing.assert_allclose(l3.cpu().numpy(), losses, rtol=2e-6)
    for ga, gb in zip(grads, g3):
        for a, b in zip(ga, gb):
            np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-6, atol=1e-12)
    # upstream gradients of the three losses
    gs = torch.tensor([0.5, 2.0, 3.0], device=DEV)
    _, g4 = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, cfg, weight_sums=wsum, grad_scale=gs)
    for grp, base, k in zip(g4, grads, (0.5, 2.0, 3.0)):
        for a, b in zip(grp, base):
            np.testing.assert_allclose(a.cpu().numpy(), k * b.cpu().numpy(), rtol=2e-6, atol=1e-12)


@pytest.mark.gpu
def test_loss_dense_first_order_no_positive_branch_and_empty_image():
    """num_pos == 0 (radet_head.py:279-281) and an image without ground truth, in the dense-first order."""
    wl, batch, idx_l, w_l, ho = _head_inputs("cfg1")   # 640x480: every level h*w % 4 == 0, so the overlapped order runs
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    idx0 = [np.where(i > 0, 0, i) for i in idx_l]          # positives -> ignored (still members of pos_inds)
    w0 = [np.where(i >= 0, 0.0, w).astype(np.float32) for i, w in zip(idx0, w_l)]
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, torch.from_numpy(np.stack(idx0)).to(DEV),
                                   torch.from_numpy(np.stack(w0)).to(DEV), F.LossConfig(), weight_sums=_weight_sums(idx0, w0, batch))
    o = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx0, w0, wl.C, wl.H, wl.W)
    losses = losses.cpu().numpy()
    assert o["num_pos"] == 0 and losses[3] == 0
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o[k]) <= 1e-5 * abs(o[k]) + 1e-6, k
    _check_grads(grads[1], o["grad_bbox"], 1e-5, 1e-7)
    _check_grads(grads[2], o["grad_iou"], 1e-5, 1e-7)
    _check_grads(grads[0], o["grad_cls"], 1e-4, 1e-6)
    # last image without ground truth: its points keep index -1 / the oracle's weights, nothing of it is a positive
    nb = len(batch)
    gt_b = [b.gt_bboxes for b in batch[:-1]] + [np.zeros((0, 4), np.float32)]
    gt_l = [b.gt_labels for b in batch[:-1]] + [np.zeros((0,), np.int64)]
    idx1 = idx_l[:-1] + [np.full_like(idx_l[-1], -1)]
    w1 = w_l[:-1] + [np.ones_like(w_l[-1])]
    counts1 = [g.shape[0] for g in gt_b]
    boxes1 = torch.from_numpy(np.concatenate(gt_b)).to(DEV)
    labels1 = torch.from_numpy(np.concatenate(gt_l)).to(DEV)
    ws1 = torch.tensor([float(w[i >= 0].astype(np.float64).sum()) if c > 0 else 0.0 for i, w, c in zip(idx1, w1, counts1)],
                       dtype=torch.float64, device=DEV)
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts1, boxes1, labels1, torch.from_numpy(np.stack(idx1)).to(DEV),
                                   torch.from_numpy(np.stack(w1)).to(DEV), F.LossConfig(), weight_sums=ws1)
    o = orc.head_loss(ho.cls, ho.bbox, ho.iou, gt_b, gt_l, idx1, w1, wl.C, wl.H, wl.W)
    losses = losses.cpu().numpy()
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o[k]) <= 1e-5 * abs(o[k]), k
    _check_grads(grads[0], o["grad_cls"], 1e-4, 1e-6)
    _check_grads(grads[1], o["grad_bbox"], 1e-4, 1e-6)
    _check_grads(grads[2], o["grad_iou"], 1e-4, 1e-6)
    assert nb >= 2


@pytest.mark.gpu
def test_loss_weight_sums_that_do_not_belong_to_the_weights_are_reported():
    """radet_loss_cfg_t.weight_sums must be the sums of the weights passed: a mismatch must not pass silently (NaN losses)."""
    wl, batch, idx_l, w_l, ho = _head_inputs("cfg1")   # 640x480: the overlapped order applies (h*w % 4 == 0 on every level)
    cls, bbox, iou = _to_dev(ho)
    counts, boxes, labels = _gt_dev(batch)
    idx = torch.from_numpy(np.stack(idx_l)).to(DEV)
    w = torch.from_numpy(np.stack(w_l)).to(DEV)
    good = _weight_sums(idx_l, w_l, batch)
    l_ok, _ = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), weight_sums=good)
    assert torch.isfinite(l_ok).all()
    l_bad, _ = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), weight_sums=good * 1.01)
    assert torch.isnan(l_bad[:3]).all()
    # and the workspace is usable afterwards
    l_again, _ = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), weight_sums=good)
    assert torch.equal(l_again, l_ok)
