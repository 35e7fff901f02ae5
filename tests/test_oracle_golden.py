"""CPU: the oracle restatement replayed against fixtures produced by the reference itself."""
import numpy as np
import pytest

from oracle import radet_oracle as orc
from radet_b200 import synthetic as syn
from tests import helpers as hp


def _assign(case):
    return orc.assign_image_seeded(case["boxes"], case["grid"], case["H"], case["W"], case["seed"], grid_step=8)


def test_assignment_matches_reference_bit_exact():
    g, names = hp.assign_cases()
    assert len(names) >= 25
    for name in names:
        case = hp.image_for(g, name)
        idx, w, used = _assign(case)
        assert np.array_equal(idx, g[f"{name}/idx"].astype(np.int64)), name
        assert np.array_equal(w, g[f"{name}/w"]), name
        # stream position after the call: next two doubles of np.random.seed(seed) stream
        rs = np.random.RandomState(case["seed"])
        rs.random_sample(used)
        assert np.array_equal(rs.random_sample(2), g[f"{name}/tail"]), name


def _head_inputs(key):
    wl, batch = hp.head_case(key)
    a = [orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed) for im in batch]
    idx_l, w_l = [x[0] for x in a], [x[1] for x in a]
    ho = syn.make_head_outputs(wl, batch, idx_l)
    return wl, batch, idx_l, w_l, ho


@pytest.mark.parametrize("key", ["small", "cfg1", "cfg3b2"])
def test_targets_and_loss_match_reference(key):
    g = hp.load("head.npz")
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    assert (hp.sha(*ho.cls, *ho.bbox, *ho.iou) == g[f"{key}/sha"]).all()
    lab, tg, wt, anc = orc.get_targets([b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l, wl.C, wl.H, wl.W)
    for l in range(5):
        assert np.array_equal(lab[l], g[f"{key}/labels{l}"].astype(np.int64))
        if key == "small":
            assert np.array_equal(tg[l], g[f"{key}/tg{l}"])          # bit-exact: power-of-two scalings
            assert np.array_equal(wt[l], g[f"{key}/wt{l}"])
            assert np.array_equal(anc[l], g[f"{key}/anc{l}"])
        else:
            nz = g[f"{key}/tg_nz{l}"]
            assert np.array_equal(np.nonzero(tg[l].any(1))[0], nz)
            assert np.array_equal(tg[l][nz], g[f"{key}/tg_val{l}"])
    out = orc.head_loss(ho.cls, ho.bbox, ho.iou, [b.gt_bboxes for b in batch], [b.gt_labels for b in batch], idx_l, w_l,
                        wl.C, wl.H, wl.W)
    for k in ("loss_cls", "loss_bbox", "loss_iou"):
        assert abs(out[k] - float(g[f"{key}/{k}"])) <= 1e-5 * abs(float(g[f"{key}/{k}"])), k
    for l in range(5):
        if key == "small":
            refs = (g[f"{key}/gcls{l}"], g[f"{key}/gbox{l}"], g[f"{key}/giou{l}"])
            mine = (out["grad_cls"][l], out["grad_bbox"][l], out["grad_iou"][l])
        else:
            refs = [g[f"{key}/gcls_s{l}"], g[f"{key}/gbox{l}"], g[f"{key}/giou{l}"]]
            mine = [out["grad_cls"][l].reshape(-1)[::97], out["grad_bbox"][l], out["grad_iou"][l]]
            if l < 2:
                mine[1] = mine[1].reshape(-1)[::13]
                mine[2] = mine[2].reshape(-1)[::13]
        for a, b in zip(mine, refs):
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6 * max(1e-12, float(np.abs(b).max())))


@pytest.mark.parametrize("key", ["small", "cfg1", "cfg3b2"])
@pytest.mark.parametrize("typ", ["vote", "global_vote"])
@pytest.mark.parametrize("thr", [0.05, 0.1])
def test_get_bboxes_matches_reference_bit_exact(key, typ, thr):
    g = hp.load("head.npz")
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    sf = np.ones(4, np.float32)
    for b, im in enumerate(batch):
        dets, labs = orc.get_bboxes_image([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou],
                                          (im.H, im.W, 3), sf, score_thr=thr,
                                          nms_cfg=dict(type=typ, iou_threshold=0.65, cluster_score=["cls", "iou"],
                                                       vote_score=["iou", "cls"], iou_enable=False, sima=0.025))
        assert np.array_equal(dets.view(np.uint32), g[f"{key}/det_{typ}_{thr}_{b}"].view(np.uint32)), (key, typ, thr, b)
        assert np.array_equal(labs, g[f"{key}/lab_{typ}_{thr}_{b}"].astype(np.int64))


@pytest.mark.parametrize("key", ["small", "cfg1"])
def test_candidates_match_reference(key):
    g = hp.load("head.npz")
    wl, batch, idx_l, w_l, ho = _head_inputs(key)
    for b, im in enumerate(batch):
        bx, sc, ctr, cats, anc = orc.select_candidates([m[b] for m in ho.cls], [m[b] for m in ho.bbox], [m[b] for m in ho.iou],
                                                       (im.H, im.W, 3), np.ones(4, np.float32), 0.05, 1000)
        full = np.concatenate([bx, (sc * ctr)[:, None], anc], 1).astype(np.float32)
        o = np.lexsort((cats, full[:, 8], full[:, 7], full[:, 6], full[:, 5]))
        ref = g[f"{key}/cand_{b}"]
        cols = [0, 1, 2, 3, 5, 6, 7, 8]
        assert np.array_equal(full[o][:, cols].view(np.uint32), ref[:, cols].view(np.uint32))
        # score*centerness: torch-CPU sigmoid is alignment dependent (SIMD vs scalar path), <= 2 ulp apart
        np.testing.assert_allclose(full[o][:, 4], ref[:, 4], rtol=2.5e-7, atol=0)
        assert np.array_equal(cats[o], g[f"{key}/candlab_{b}"].astype(np.int64))


@pytest.mark.parametrize("case", range(5))
def test_ops_match_reference_bit_exact(case):
    g = hp.load("ops.npz")
    boxes, labels = g[f"c{case}/boxes"], g[f"c{case}/labels"].astype(np.int64)
    cls, ctr = g[f"c{case}/cls"], g[f"c{case}/ctr"]
    cfg = dict(iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"], iou_enable=False, sima=0.025)
    for nm, gm in (("vote", False), ("gvote", True)):
        d, l = orc.vote_nms_wrapper(boxes, cls, labels, cfg, score_factor=ctr, global_mode=gm)
        assert np.array_equal(d.view(np.uint32), g[f"c{case}/{nm}_dets"].view(np.uint32)), nm
        assert np.array_equal(l, g[f"c{case}/{nm}_labels"].astype(np.int64))
    _, _, _, inst, cnum = orc.vote_nms_c(boxes, cls * ctr, cls * ctr, labels, 0.65)
    assert np.array_equal(inst, g[f"c{case}/inst"])
    assert np.array_equal(cnum, g[f"c{case}/cnum"])
    if boxes.shape[0] <= 64:    # python restatement agrees with the C one
        b2, l2, s2 = orc.vote_nms_py(boxes, cls * ctr, cls * ctr, labels, 0.65)
        b1, l1, s1, _, _ = orc.vote_nms_c(boxes, cls * ctr, cls * ctr, labels, 0.65)
        assert np.array_equal(b1.view(np.uint32), b2.view(np.uint32)) and np.array_equal(l1, l2) and np.array_equal(s1, s2)


def test_anchor_docstring_kat():
    # anchor_generator.py:40-55 known answers, with this config's scale (8) instead of 9 checked structurally
    a = orc.grid_anchors(32, 32, strides=(16,), octave_base_scale=9)
    assert np.array_equal(a, np.array([[-72, -72, 72, 72], [-56, -72, 88, 72], [-72, -56, 72, 88], [-56, -56, 88, 88]], np.float32))


_LOSS_CASES = [("focal", "mean_avg", dict(reduction="mean", avg_factor=37.5, use_w=True)), ("focal", "none", dict(reduction="none")),
               ("focal", "sum_w", dict(reduction="sum", use_w=True)), ("focal", "mean", dict(reduction="mean")),
               ("giou", "mean_avg", dict(reduction="mean", avg_factor="sum_w", use_w=True, loss_weight=2.0)),
               ("giou", "none", dict(reduction="none", loss_weight=2.0)), ("giou", "zero_w4", dict(zero_w4=True, loss_weight=2.0)),
               ("giou", "w4", dict(reduction="sum", w4=True, loss_weight=2.0)),
               ("bce", "mean_avg", dict(reduction="mean", avg_factor="sum_w", use_w=True)), ("bce", "none", dict(reduction="none"))]


def loss_case_inputs(g, kind, kw):
    """(pred, target, weight, reduction, avg_factor, loss_weight) of one losses.npz record."""
    pred = g[f"{kind}/pred"]
    target = g["focal/target"].astype(np.int64) if kind == "focal" else g["giou/target"] if kind == "giou" else g["bce/label"]
    w = g[f"{kind}/weight"]
    weight = w if kw.get("use_w") else None
    if kw.get("zero_w4"):
        weight = np.zeros((pred.shape[0], 4), np.float32)
    if kw.get("w4"):
        weight = np.repeat(w[:, None], 4, 1)
    af = kw.get("avg_factor")
    if af == "sum_w":
        af = float(w.sum())
    return pred, target, weight, kw.get("reduction", "mean"), af, kw.get("loss_weight", 1.0)


@pytest.mark.parametrize("kind,tag,kw", _LOSS_CASES)
def test_standalone_losses_match_reference(kind, tag, kw):
    """oracle restatement of FocalLoss / GIoULoss / CrossEntropyLoss vs the reference modules (fp32 autograd golden)."""
    g = hp.load("losses.npz")
    pred, target, weight, red, af, lw = loss_case_inputs(g, kind, kw)
    loss, grad = orc.standalone_loss(kind, pred, target, weight, red, af, loss_weight=lw, dtype="float32")
    np.testing.assert_allclose(loss, g[f"{kind}/{tag}/loss"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(grad, g[f"{kind}/{tag}/grad"], rtol=2e-5, atol=1e-7)


def test_standalone_bce_onehot_matches_reference():
    g = hp.load("losses.npz")
    loss, grad = orc.standalone_loss("bce", g["focal/pred"], g["focal/target"].astype(np.int64), g["focal/weight"], "mean", 11.0,
                                     dtype="float32")
    np.testing.assert_allclose(loss, g["bce/onehot/loss"], rtol=2e-6)
    np.testing.assert_allclose(grad, g["bce/onehot/grad"], rtol=2e-5, atol=1e-8)


ASSIGN_OPT_VARIANTS = {"adapt": dict(adapt_positive_num=True), "mult": dict(multiply_samplepro_for_weight=True),
                       "both": dict(adapt_positive_num=True, multiply_samplepro_for_weight=True),
                       "adapt_nobal": dict(adapt_positive_num=True, balance_sample=False)}
ASSIGN_OPT_IMAGES = [("cfg2", i) for i in range(4)] + [("cfg3", i) for i in range(3)] + [("cfg5", i) for i in range(2)]


@pytest.mark.parametrize("variant", sorted(ASSIGN_OPT_VARIANTS))
def test_assignment_options_match_reference_bit_exact(variant):
    """adapt_positive_num (label_assignment.py:88-95) / multiply_samplepro_for_weight (:127-128) / balance_sample=False."""
    g = hp.load("assign_opts.npz")
    for key, i in ASSIGN_OPT_IMAGES:
        im = syn.make_batch(syn.WORKLOADS[key], 1, i)[0]
        idx, w, used = orc.assign_image_seeded(im.gt_bboxes, im.masks, im.H, im.W, im.seed, **ASSIGN_OPT_VARIANTS[variant])
        name = f"{variant}/{key}_{i}"
        assert np.array_equal(idx, g[f"{name}/idx"].astype(np.int64)), name
        assert np.array_equal(w, g[f"{name}/w"]), name
        rs = np.random.RandomState(im.seed)
        rs.random_sample(used)
        assert np.array_equal(rs.random_sample(2), g[f"{name}/tail"]), name


def test_legacy_choice_restatement_equals_numpy():
    """The third-party arithmetic on the assignment path is numpy's legacy RandomState.choice (label_assignment.py:112,119;
    numpy is unpinned in requirements.txt).  The oracle's restatement must reproduce numpy's own outputs AND leave the
    stream at the same position, for uniform p (binary masks) and for arbitrary p, with and without replacement."""
    rs = np.random.RandomState(2024)
    for case in range(400):
        n = int(rs.randint(1, 60))
        k = int(rs.randint(1, 33))
        replace = bool(rs.randint(0, 2)) or k > n
        if case % 2:
            p32 = np.full(n, np.float32(1.0)) / np.float32(n)                # what binary masks give (:103)
        else:
            w = rs.uniform(0.05, 1.0, n).astype(np.float32)
            p32 = w / np.sum(w)
        seed = int(rs.randint(0, 2 ** 31 - 1))
        ref_rs = np.random.RandomState(seed)
        want = ref_rs.choice(a=n, size=k, p=p32, replace=replace)
        tail_want = ref_rs.random_sample(3)
        mine_rs = np.random.RandomState(seed)

        class S:
            pos = 0

            def random_sample(self, m):
                self.pos += m
                return mine_rs.random_sample(m)

        got = orc.legacy_choice(S(), p32, k, replace)
        assert np.array_equal(got, want), (case, n, k, replace)
        assert np.array_equal(mine_rs.random_sample(3), tail_want), (case, "stream position")


def _live_reference_ext(name):
    """The reference's own C++ op, compiled unmodified into oracle/_ref/ by oracle/build_ref.py (build container only)."""
    import importlib.util
    import os

    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", name, name + ".so")
    if not os.path.exists(so):
        pytest.skip(f"{so} not built (the reference tree only exists in the build container)")
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(name, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("iou_enable,sigma", [(False, 0.025), (True, 0.025), (True, 0.5)])
def test_c_oracle_equals_the_compiled_reference_ops(iou_enable, sigma):
    """Fuzz: oracle/vote_oracle.c against the reference's compiled vote_ext / cluster_ext on clustered random boxes,
    including the iou_enable score decay that no shipped config (and hence no golden) switches on."""
    import torch

    vote_ext = _live_reference_ext("vote_ext")
    cluster_ext = _live_reference_ext("cluster_ext")
    rs = np.random.RandomState(77)
    for case in range(12):
        n, ncls, nobj = int(rs.randint(1, 400)), int(rs.randint(1, 6)), int(rs.randint(1, 40))
        ctr = rs.uniform(50, 590, (nobj, 2))
        wh = rs.uniform(20, 200, (nobj, 2))
        pick = rs.randint(0, nobj, n)
        c = ctr[pick] + rs.normal(0, 3, (n, 2))
        s = wh[pick] * np.exp(rs.normal(0, 0.06, (n, 2)))
        boxes = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
        labels = rs.randint(0, ncls, n).astype(np.int64)
        cs = rs.uniform(0.05, 1, n).astype(np.float32)
        vs = rs.uniform(0.05, 1, n).astype(np.float32)
        if np.unique(cs).size != n:
            continue
        T = torch.from_numpy
        for gmode, fn in ((False, vote_ext.vote_nms), (True, vote_ext.global_vote_nms)):
            rb, rl, rsc = fn(T(boxes), T(cs), T(vs.copy()), T(labels), 0.65, iou_enable, sigma)      # vote_wrapper.py:32
            ob, ol, os_, inst, cnum = orc.vote_nms_c(boxes, cs, vs, labels, 0.65, global_mode=gmode, iou_enable=iou_enable, sigma=sigma)
            assert rb.shape[0] == ob.shape[0], (case, gmode)
            assert np.array_equal(rb.numpy().view(np.uint32), ob.view(np.uint32)), (case, gmode)
            assert np.array_equal(rl.numpy(), ol) and np.array_equal(rsc.numpy(), os_)
        ri, rn = cluster_ext.cluster_nms(T(boxes), T(cs), T(labels), 0.65)
        _, _, _, inst, cnum = orc.vote_nms_c(boxes, cs, cs, labels, 0.65)
        assert np.array_equal(ri.numpy(), inst) and np.array_equal(rn.numpy(), cnum)


@pytest.mark.parametrize("case", range(3))
def test_ops_iou_enable_match_reference_bit_exact(case):
    """iou_enable=True (vote_ext.cpp:164-167) through the reference's wrappers: golden ops_iou.npz vs the C oracle."""
    g = hp.load("ops_iou.npz")
    cfg = dict(type="vote", iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"], iou_enable=True,
               sigma=float(g[f"c{case}/sigma"]))
    for nm, gm in (("vote", False), ("gvote", True)):
        d, l = orc.vote_nms_wrapper(g[f"c{case}/boxes"], g[f"c{case}/cls"], g[f"c{case}/labels"].astype(np.int64), cfg,
                                    score_factor=g[f"c{case}/ctr"], global_mode=gm)
        assert np.array_equal(d.view(np.uint32), g[f"c{case}/{nm}_dets"].view(np.uint32)), nm
        assert np.array_equal(l, g[f"c{case}/{nm}_labels"].astype(np.int64))


def test_greedy_nms_restatement_equals_torchvision():
    """mmcv.ops.batched_nms is not in the reference tree (mmcv==1.3.18, requirements.txt:8) and its branch of
    _get_bboxes_single (radet_head.py:159-163) has no reference golden; torchvision.ops.batched_nms implements the same
    greedy `iou > thr` rule (SURVEY §8c names it as the second oracle): keep sets and order must agree."""
    import torch
    import torchvision

    rs = np.random.RandomState(5)
    for case in range(30):
        n, ncls, nobj = int(rs.randint(1, 500)), int(rs.randint(1, 8)), int(rs.randint(1, 40))
        ctr = rs.uniform(50, 590, (nobj, 2))
        wh = rs.uniform(20, 200, (nobj, 2))
        pick = rs.randint(0, nobj, n)
        c = ctr[pick] + rs.normal(0, 3, (n, 2))
        s = wh[pick] * np.exp(rs.normal(0, 0.06, (n, 2)))
        boxes = np.concatenate([c - s / 2, c + s / 2], 1).astype(np.float32)
        labels = rs.randint(0, ncls, n).astype(np.int64)
        sc = rs.uniform(0.05, 1, n).astype(np.float32)
        if np.unique(sc).size != n:
            continue
        keep = orc.greedy_nms_keep(boxes, sc, labels, 0.65)
        tv = torchvision.ops.boxes._batched_nms_vanilla(torch.from_numpy(boxes), torch.from_numpy(sc), torch.from_numpy(labels), 0.65)
        assert np.array_equal(keep, tv.numpy()), case


def test_tblr_coder_matches_reference_bit_exact():
    """TBLRBBoxCoder.encode / decode (tblr_bbox_coder.py:29-68, 71-172) on explicit lists: square and non-square priors,
    normalizer 1/8 (the config's) and 4.0 (the class default), with / without clipping."""
    g = hp.load("coder.npz")
    for nrm in (0.125, 4.0):
        assert np.array_equal(orc.tblr_encode(g["priors"], g["gts"], nrm).view(np.uint32), g[f"enc_{nrm}"].view(np.uint32))
        assert np.array_equal(orc.tblr_decode(g["priors"], g["tblr"], nrm).view(np.uint32), g[f"dec_{nrm}"].view(np.uint32))
        assert np.array_equal(orc.tblr_decode(g["priors"], g["tblr"], nrm, max_shape=(480, 640, 3)).view(np.uint32),
                              g[f"dec_clip_{nrm}"].view(np.uint32))
    assert np.array_equal(orc.tblr_decode(g["priors"], g["tblr"], 0.125, max_shape=(480, 640, 3), clip_border=False).view(np.uint32),
                          g["dec_noclip_0.125"].view(np.uint32))
