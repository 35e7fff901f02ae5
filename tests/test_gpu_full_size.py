"""GPU parity at the FULL sizes of BASELINE.json's configs, against the oracle (not only through properties):

    cfg2  B=8   640x480  C=21   the bench workload: assignment, loss fwd+bwd, get_bboxes -- all three stages
    cfg3  B=8   640x480  C=30   many-GT clutter: assignment + loss
    cfg4  B=64  640x480  C=21   inference, score_thr=0.1: get_bboxes (vote and plain-NMS branches)
    cfg5  B=16  1280x960 C=30   the bandwidth-bound shape: assignment + loss

Same bars as tests/test_gpu_parity.py: indices / weights / RNG position / labels / keep sets / voted boxes bit-exact,
losses 1e-5 relative, gradients rtol 1e-4 + atol 1e-6 * max|grad|.  The oracle's CPU time here is tens of seconds.
"""
import numpy as np
import pytest
import torch

from oracle import radet_oracle as orc
from radet_b200 import functional as F
from radet_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
GEOM = F.Geometry()
DEV = "cuda"
NMS_CFG = dict(iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"], iou_enable=False)


@pytest.fixture(autouse=True)
def _oracle_sigmoid():
    orc.set_sigmoid(lambda x: torch.sigmoid(torch.from_numpy(x).to(DEV)).cpu().numpy())
    yield
    orc.set_sigmoid(None)


_CACHE = {}


def _case(key):
    """Whole batch of the workload: device assignment (checked against the oracle by the caller) + head outputs."""
    if key in _CACHE:
        return _CACHE[key]
    wl = syn.WORKLOADS[key]
    batch = syn.make_batch(wl)
    shapes = GEOM.level_shapes(wl.H, wl.W)
    counts = [im.gt_bboxes.shape[0] for im in batch]
    boxes = torch.from_numpy(np.concatenate([im.gt_bboxes for im in batch])).to(DEV)
    labels = torch.from_numpy(np.concatenate([im.gt_labels for im in batch])).to(DEV)
    grids = np.concatenate([syn.sample_grid(im.masks) for im in batch])
    gh, gw = grids.shape[1:]
    bits = F.pack_masks(torch.from_numpy(grids).to(DEV), 1, gh, gw)
    seeds = torch.tensor([im.seed for im in batch], dtype=torch.int64, device=DEV)
    wsum = torch.full((len(batch),), -1.0, dtype=torch.float64, device=DEV)
    idx, w, used = F.assign(GEOM, shapes, counts, boxes, bits, (gh, gw), seeds=seeds, weight_sums=wsum)
    ref = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
    ho = syn.make_head_outputs(wl, batch, [r[0] for r in ref])
    _CACHE.clear()            # one workload resident at a time (cfg5 holds ~1 GB of masks)
    _CACHE[key] = (wl, batch, counts, boxes, labels, idx, w, used, ref, ho)
    _CACHE[key + "/wsum"] = wsum
    return _CACHE[key]


def _to_dev(ho):
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return [T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou]


@pytest.mark.parametrize("key", ["cfg2", "cfg3", "cfg5"])
def test_full_batch_assignment_bit_exact(key):
    wl, batch, counts, boxes, labels, idx, w, used, ref, ho = _case(key)
    idx, w, used = idx.cpu().numpy(), w.cpu().numpy(), used.cpu().numpy()
    for b in range(len(batch)):
        assert np.array_equal(idx[b], ref[b][0]), (key, b)
        assert np.array_equal(w[b], ref[b][1]), (key, b)
        assert used[b] == ref[b][2], (key, b)          # numpy RNG stream position
    # the per-image weight sums handed to the loss: sum of w over idx >= 0 (radet_head.py:245-254)
    want = np.array([float(ref[b][1][ref[b][0] >= 0].astype(np.float64).sum()) for b in range(len(batch))])
    assert np.array_equal(_CACHE[key + "/wsum"].cpu().numpy(), want), key


@pytest.mark.parametrize("key", ["cfg2", "cfg3", "cfg5"])
def test_full_batch_loss_vs_oracle(key):
    wl, batch, counts, boxes, labels, idx, w, used, ref, ho = _case(key)
    cls, bbox, iou = _to_dev(ho)
    losses, grads = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig())
    losses = losses.cpu().numpy()
    o = orc.head_loss(ho.cls, ho.bbox, ho.iou, [im.gt_bboxes for im in batch], [im.gt_labels for im in batch],
                      [r[0] for r in ref], [r[1] for r in ref], wl.C, wl.H, wl.W)
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(losses[i] - o[k]) <= 1e-5 * abs(o[k]), (key, k, losses[i], o[k])
    assert losses[3] == o["num_pos"]
    for mine, want in ((grads[0], o["grad_cls"]), (grads[1], o["grad_bbox"]), (grads[2], o["grad_iou"])):
        for a, r in zip(mine, want):
            np.testing.assert_allclose(a.cpu().numpy(), r, rtol=1e-4, atol=1e-6 * max(1e-30, float(np.abs(r).max())))
    # dense-first order (per-image weight sums of the assignment handed over): the same bar against the oracle
    l1, g1 = F.loss_fwd_bwd(GEOM, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), weight_sums=_CACHE[key + "/wsum"])
    l1 = l1.cpu().numpy()
    for i, k in enumerate(("loss_cls", "loss_bbox", "loss_iou")):
        assert abs(l1[i] - o[k]) <= 1e-5 * abs(o[k]), (key, k, l1[i], o[k])
    assert l1[3] == o["num_pos"]
    for mine, want in ((g1[0], o["grad_cls"]), (g1[1], o["grad_bbox"]), (g1[2], o["grad_iou"])):
        for a, r in zip(mine, want):
            np.testing.assert_allclose(a.cpu().numpy(), r, rtol=1e-4, atol=1e-6 * max(1e-30, float(np.abs(r).max())))
    # the experimental single-launch kernel agrees, with and without the assignment's per-image weight sums handed over
    from tests.test_gpu_parity import _fused_loss
    for kw in (dict(), dict(weight_sums=_CACHE[key + "/wsum"])):
        l2, g2 = _fused_loss(wl, cls, bbox, iou, counts, boxes, labels, idx, w, **kw)
        np.testing.assert_allclose(l2.cpu().numpy(), losses, rtol=2e-6)
        for grp_a, grp_b in zip(grads, g2):
            for a, b in zip(grp_a, grp_b):
                np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-5, atol=1e-12)


@pytest.mark.parametrize("key,types", [("cfg2", ("vote",)), ("cfg4", ("vote", "nms"))])
def test_full_batch_get_bboxes_bit_exact(key, types):
    wl, batch, counts, boxes, labels, idx, w, used, ref, ho = _case(key)
    cls, bbox, iou = _to_dev(ho)
    shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=DEV)
    sf = torch.ones((len(batch), 4), device=DEV)
    for typ in types:
        cfg = F.DetectConfig(score_thr=wl.score_thr, nms_pre=wl.nms_pre, max_per_img=wl.max_per_img, nms_type=typ, **NMS_CFG)
        dets, dl, num = F.get_bboxes(GEOM, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
        dets, dl, num = dets.cpu().numpy(), dl.cpu().numpy(), num.cpu().numpy()
        for b, im in enumerate(batch):
            od, ol = orc.get_bboxes_image([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou], (im.H, im.W, 3),
                                          np.ones(4, np.float32), score_thr=wl.score_thr, nms_pre=wl.nms_pre,
                                          max_per_img=wl.max_per_img, nms_cfg=dict(type=typ, **NMS_CFG))
            assert num[b] == od.shape[0], (key, typ, b)
            assert np.array_equal(dets[b, :num[b]].view(np.uint32), od.view(np.uint32)), (key, typ, b)
            assert np.array_equal(dl[b, :num[b]], ol), (key, typ, b)
