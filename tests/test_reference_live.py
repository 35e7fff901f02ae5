"""CPU, build container only: the oracle against the UNMODIFIED reference executed live (through oracle/ref_shim.py) on
randomly drawn small problems — wider than the committed goldens (odd image sizes, empty images, many option
combinations).  Skipped wherever /root/reference is absent (the GPU box)."""
import os

import numpy as np
import pytest

from oracle import radet_oracle as orc
from radet_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/radet"), reason="the reference tree only exists in the build container")


@pytest.fixture(scope="module")
def shim():
    from oracle import ref_shim

    ref_shim.install()
    return ref_shim


def _random_workload(rs, i):
    H, W = int(rs.randint(64, 301)), int(rs.randint(64, 401))
    C = int(rs.randint(1, 31))
    g0 = int(rs.randint(0, 4))
    return syn.Workload(f"live{i}_{H}x{W}", H, W, C, int(rs.randint(1, 4)), g0, g0 + int(rs.randint(0, 10)), 1000 + i)


@pytest.mark.parametrize("opts", [dict(), dict(adapt_positive_num=True), dict(multiply_samplepro_for_weight=True),
                                  dict(balance_sample=False), dict(positive_num=3), dict(positive_num=17, neg_threshold=0.5)])
def test_assignment_live(shim, opts):
    la = shim.build_reference_assigner(**opts)
    rs = np.random.RandomState(hash(tuple(sorted(opts))) % 2 ** 31)
    for i in range(10):
        wl = _random_workload(rs, i)
        for im in syn.make_batch(wl):
            np.random.seed(im.seed)
            res = la(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels,
                          distance_maps=shim.BitmapMasksStandIn(im.masks)))
            tail = np.random.random_sample(2)
            idx, w, used = orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed, **opts)
            assert np.array_equal(idx, res["points_to_gt_index"]), (wl.name, opts)
            assert np.array_equal(w, res["points_weight"]), (wl.name, opts)
            r2 = np.random.RandomState(im.seed)
            r2.random_sample(used)
            assert np.array_equal(r2.random_sample(2), tail), (wl.name, opts, "stream position")


def test_targets_loss_and_get_bboxes_live(shim):
    import torch

    rs = np.random.RandomState(4321)
    la = shim.build_reference_assigner()
    for i in range(6):
        wl = _random_workload(rs, 100 + i)
        batch = syn.make_batch(wl)
        assigned = []
        for im in batch:
            np.random.seed(im.seed)
            r = la(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels, distance_maps=shim.BitmapMasksStandIn(im.masks)))
            assigned.append((r["points_to_gt_index"], r["points_weight"]))
        idx_l, w_l = [a[0] for a in assigned], [a[1] for a in assigned]
        ho = syn.make_head_outputs(wl, batch, idx_l)
        thr = float(rs.choice([0.02, 0.05, 0.2]))
        typ = str(rs.choice(["vote", "global_vote"]))
        nms_pre = int(rs.choice([50, 1000]))
        head = shim.build_reference_head(wl.C, nms_type=typ, score_thr=thr, nms_pre=nms_pre, max_per_img=int(rs.choice([5, 100])))
        metas = syn.img_metas(batch, scale=float(rs.choice([1.0, 0.75, 1.6])))
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        cls = [T(m).requires_grad_() for m in ho.cls]
        box = [T(m).requires_grad_() for m in ho.bbox]
        iou = [T(m).requires_grad_() for m in ho.iou]
        gtb, gtl = [T(im.gt_bboxes) for im in batch], [T(im.gt_labels) for im in batch]
        losses = head.loss(cls, box, iou, gtb, gtl, [T(a) for a in idx_l], [T(a) for a in w_l], metas)
        (losses["loss_cls"] + losses["loss_bbox"] + losses["loss_iou"]).backward()
        o = orc.head_loss(ho.cls, ho.bbox, ho.iou, [im.gt_bboxes for im in batch], [im.gt_labels for im in batch], idx_l, w_l,
                          wl.C, wl.H, wl.W)
        for k in ("loss_cls", "loss_bbox", "loss_iou"):
            assert abs(o[k] - float(losses[k].detach())) <= 1e-5 * abs(float(losses[k].detach())) + 1e-7, (wl.name, k)
        for mine, ref in zip(o["grad_cls"] + o["grad_bbox"] + o["grad_iou"], cls + box + iou):
            g = np.zeros(mine.shape, np.float32) if ref.grad is None else ref.grad.numpy()
            np.testing.assert_allclose(mine, g, rtol=1e-4, atol=1e-6 * max(1e-30, float(np.abs(g).max())))
        with torch.no_grad():
            res = head.get_bboxes([T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou], metas, rescale=True)
        sf = np.asarray(metas[0]["scale_factor"], np.float32)
        cfgd = dict(type=typ, iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"], iou_enable=False)
        for b, (im, (dets, labs)) in enumerate(zip(batch, res)):
            od, ol = orc.get_bboxes_image([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou], (im.H, im.W, 3), sf,
                                          score_thr=thr, nms_pre=nms_pre, max_per_img=head.test_cfg.max_per_img, nms_cfg=cfgd)
            dets, labs = dets.numpy(), labs.numpy().reshape(-1)
            assert dets.shape[0] == od.shape[0], (wl.name, typ, thr)
            if od.shape[0]:
                assert np.array_equal(dets[:, :4].view(np.uint32), od[:, :4].view(np.uint32)), (wl.name, typ, thr)
                np.testing.assert_allclose(dets[:, 4], od[:, 4], rtol=4e-7)      # torch-CPU sigmoid flavour
                assert np.array_equal(labs, ol)


def test_candidates_without_nms_live(shim):
    """get_bboxes(with_nms=False) (radet_head.py:165-169) on random problems: rows compared as sets."""
    import torch

    rs = np.random.RandomState(99)
    la = shim.build_reference_assigner()
    for i in range(5):
        wl = _random_workload(rs, 200 + i)
        batch = syn.make_batch(wl)
        idx_l = []
        for im in batch:
            np.random.seed(im.seed)
            idx_l.append(la(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels,
                                 distance_maps=shim.BitmapMasksStandIn(im.masks)))["points_to_gt_index"])
        ho = syn.make_head_outputs(wl, batch, idx_l)
        thr, nms_pre = float(rs.choice([0.03, 0.1])), int(rs.choice([20, 1000]))
        head = shim.build_reference_head(wl.C, score_thr=thr, nms_pre=nms_pre)
        scale = float(rs.choice([1.0, 1.3]))
        metas = syn.img_metas(batch, scale=scale)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        with torch.no_grad():
            res = head.get_bboxes([T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou], metas, rescale=True,
                                  with_nms=False)
        for b, (im, (rows, cats)) in enumerate(zip(batch, res)):
            bx, sc, ctr, ocats, anc = orc.select_candidates([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou],
                                                            (im.H, im.W, 3), np.full(4, scale, np.float32), thr, nms_pre)
            rows, cats = rows.numpy(), cats.numpy().reshape(-1)
            if bx.shape[0] == 0:
                assert rows.shape[0] == 0
                continue
            full = np.concatenate([bx, (sc * ctr)[:, None], anc], 1).astype(np.float32)
            assert rows.shape == full.shape, wl.name
            o = np.lexsort((cats, rows[:, 8], rows[:, 7], rows[:, 6], rows[:, 5], rows[:, 3], rows[:, 2], rows[:, 1], rows[:, 0]))
            oo = np.lexsort((ocats, full[:, 8], full[:, 7], full[:, 6], full[:, 5], full[:, 3], full[:, 2], full[:, 1], full[:, 0]))
            cols = [0, 1, 2, 3, 5, 6, 7, 8]
            assert np.array_equal(rows[o][:, cols].view(np.uint32), full[oo][:, cols].view(np.uint32)), wl.name
            np.testing.assert_allclose(rows[o][:, 4], full[oo][:, 4], rtol=4e-7)
            assert np.array_equal(cats[o], ocats[oo])


def test_standalone_losses_live(shim):
    import torch
    from radet.models.losses import CrossEntropyLoss, FocalLoss, GIoULoss

    rs = np.random.RandomState(7)
    for i in range(8):
        n, C = int(rs.randint(1, 200)), int(rs.randint(1, 40))
        gamma, alpha = float(rs.choice([2.0, 1.5, 0.5])), float(rs.choice([0.25, 0.5]))
        pred = rs.normal(-1, 3, (n, C)).astype(np.float32)
        target = rs.randint(0, C + 1, n).astype(np.int64)
        w = rs.uniform(0, 2, n).astype(np.float32)
        x = torch.from_numpy(pred).requires_grad_()
        out = FocalLoss(use_sigmoid=True, gamma=gamma, alpha=alpha, loss_weight=1.5)(x, torch.from_numpy(target), torch.from_numpy(w), avg_factor=3.0)
        out.backward()
        l, g = orc.standalone_loss("focal", pred, target, w, "mean", 3.0, loss_weight=1.5, gamma=gamma, alpha=alpha, dtype="float32")
        np.testing.assert_allclose(l, out.detach().numpy(), rtol=3e-6)
        np.testing.assert_allclose(g, x.grad.numpy(), rtol=5e-5, atol=1e-8)
        c = rs.uniform(0, 500, (n, 2))
        s = rs.uniform(1, 200, (n, 2))
        tb = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
        pb = (tb + rs.normal(0, 20, (n, 4))).astype(np.float32)
        x = torch.from_numpy(pb).requires_grad_()
        out = GIoULoss(eps=1e-6, loss_weight=2.0)(x, torch.from_numpy(tb), torch.from_numpy(w), reduction_override="sum")
        out.backward()
        l, g = orc.standalone_loss("giou", pb, tb, w, "sum", None, loss_weight=2.0, dtype="float32")
        np.testing.assert_allclose(l, out.detach().numpy(), rtol=3e-6)
        np.testing.assert_allclose(g, x.grad.numpy(), rtol=5e-5, atol=1e-7)
        logit, soft = rs.normal(0, 3, n).astype(np.float32), rs.uniform(0, 1, n).astype(np.float32)
        x = torch.from_numpy(logit).requires_grad_()
        out = CrossEntropyLoss(use_sigmoid=True)(x, torch.from_numpy(soft), torch.from_numpy(w), avg_factor=float(w.sum()) + 1)
        out.backward()
        l, g = orc.standalone_loss("bce", logit, soft, w, "mean", float(w.sum()) + 1, dtype="float32")
        np.testing.assert_allclose(l, out.detach().numpy(), rtol=3e-6)
        np.testing.assert_allclose(g, x.grad.numpy(), rtol=5e-5, atol=1e-8)


def test_get_targets_live(shim):
    """RADetHead.get_targets (radet_head.py:290-392) on random problems: labels (incl. the idx == 0 -> last GT label quirk),
    TBLR targets, weights and priors, level-major / image-minor, bit-exact."""
    import torch

    rs = np.random.RandomState(555)
    la = shim.build_reference_assigner()
    for i in range(6):
        wl = _random_workload(rs, 300 + i)
        batch = syn.make_batch(wl)
        idx_l, w_l = [], []
        for im in batch:
            np.random.seed(im.seed)
            r = la(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels, distance_maps=shim.BitmapMasksStandIn(im.masks)))
            idx_l.append(r["points_to_gt_index"])
            w_l.append(r["points_weight"])
        ho = syn.make_head_outputs(wl, batch, idx_l)
        head = shim.build_reference_head(wl.C)
        metas = syn.img_metas(batch)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        cls, box = [T(m) for m in ho.cls], [T(m) for m in ho.bbox]
        anchors, _ = head.get_anchors([m.shape[-2:] for m in cls], metas, device="cpu")
        lab, tg, wt, anc = head.get_targets(anchors, cls, box, [T(im.gt_bboxes) for im in batch], [T(im.gt_labels) for im in batch],
                                            [T(a) for a in idx_l], [T(a) for a in w_l], metas)
        olab, otg, owt, oanc = orc.get_targets([im.gt_bboxes for im in batch], [im.gt_labels for im in batch], idx_l, w_l, wl.C, wl.H, wl.W)
        for l in range(5):
            assert np.array_equal(lab[l].numpy(), olab[l]), (wl.name, l)
            assert np.array_equal(tg[l].numpy().view(np.uint32), otg[l].view(np.uint32)), (wl.name, l)
            assert np.array_equal(wt[l].numpy(), owt[l]) and np.array_equal(anc[l].numpy(), oanc[l])


def test_mask_grid_handoff_equals_the_stock_pipeline(shim):
    """`PackVisibleMaskGrid` (CPU, in the worker) ships exactly the pixels the reference's LabelAssignment reads,
    in a form the reference's own formatting code accepts; the assignment computed from (grid, seed) equals what the
    UNMODIFIED reference step returns when `np.random.seed(seed)` precedes it."""
    import torch
    from radet.datasets.pipelines.formating import to_tensor

    from radet_b200 import plugin as P
    from radet_b200.plugin.pipelines import is_mask_grid_handoff

    la = shim.build_reference_assigner()
    pack = P.PackVisibleMaskGrid(seed_key="seed")
    rs = np.random.RandomState(99)
    for i in range(6):
        wl = _random_workload(rs, 300 + i)
        for im in syn.make_batch(wl):
            r = pack(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels,
                          distance_maps=shim.BitmapMasksStandIn(im.masks), seed=im.seed))
            grid_t, seed_t = to_tensor(r["points_to_gt_index"]), to_tensor(r["points_weight"])      # DefaultFormatBundle, formating.py:218-223
            assert grid_t.dtype == torch.uint8 and seed_t.dtype == torch.int64 and is_mask_grid_handoff([grid_t])
            np.random.seed(int(seed_t[0]))
            ref = la(dict(img_shape=(im.H, im.W, 3), gt_bboxes=im.gt_bboxes, gt_labels=im.gt_labels,
                          distance_maps=shim.BitmapMasksStandIn(im.masks)))
            idx, w, _ = orc.assign_image_seeded(im.gt_bboxes, grid_t.numpy(), im.H, im.W, int(seed_t[0]), grid_step=8)
            assert np.array_equal(idx, ref["points_to_gt_index"]) and np.array_equal(w, ref["points_weight"]), wl.name
