#!/usr/bin/env python
"""One pass over every kernel of libradet_b200.so at small sizes, for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck python scripts/sanitize_driver.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
from radet_b200 import functional as F, plugin as P, synthetic as syn

entry.smoke()                                   # assign (seeded) -> loss fwd+bwd (TMA dense kernel) -> get_bboxes (vote), checked vs the oracle
dev = "cuda"
geom = F.Geometry()
# many GT (two-word bit sets), uniforms mode, options; odd plane sizes (register dense kernel); single-launch loss variant
for wl, fused in ((syn.Workload("s1", 480, 640, 30, 2, 40, 40, 3), False), (syn.Workload("s2", 100, 100, 7, 2, 3, 5, 11), False),
                  (syn.Workload("s3", 256, 320, 21, 2, 5, 9, 5), True)):
    batch = syn.make_batch(wl)
    shapes = geom.level_shapes(wl.H, wl.W)
    counts = [im.gt_bboxes.shape[0] for im in batch]
    boxes = torch.from_numpy(np.concatenate([im.gt_bboxes for im in batch])).to(dev)
    labels = torch.from_numpy(np.concatenate([im.gt_labels for im in batch])).to(dev)
    full = torch.from_numpy(np.concatenate([im.masks for im in batch])).to(dev)
    gh, gw = -(-wl.H // 8), -(-wl.W // 8)
    bits = F.pack_masks(full, 8, gh, gw)                                  # general (strided, warp-per-word) packing kernel
    u = torch.from_numpy(np.stack([np.random.RandomState(im.seed).random_sample(4096) for im in batch])).to(dev)
    wsum = torch.zeros(len(batch), dtype=torch.float64, device=dev)
    idx, w, used = F.assign(geom, shapes, counts, boxes, bits, (gh, gw), uniforms=u, adapt_positive_num=True,
                            multiply_samplepro_for_weight=True, weight_sums=wsum)
    st = F.seed_states(torch.tensor([im.seed for im in batch], dtype=torch.int32, device=dev))
    idx, w, used = F.assign(geom, shapes, counts, boxes, bits, (gh, gw), mt_states=st, weight_sums=wsum)
    ho = syn.make_head_outputs(wl, batch, list(idx.cpu().numpy()))
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cls, bbox, iou = [T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou]
    if fused:
        os.environ["RADET_LOSS_IMPL"] = "fused"
    losses, grads = F.loss_fwd_bwd(geom, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), weight_sums=wsum if fused else None)
    losses, grads = F.loss_fwd_bwd(geom, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(gamma=1.5))
    os.environ.pop("RADET_LOSS_IMPL", None)
    # the overlapped order (dense kernel as the programmatic dependent of loss_pos), general gamma, 256- and 128-point items
    for pts in ("256", "128"):
        os.environ["RADET_DENSE_PTS"] = pts
        F.loss_fwd_bwd(geom, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(gamma=1.5), weight_sums=wsum)
        F.loss_fwd_bwd(geom, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig(), weight_sums=wsum)
        F.loss_fwd_bwd(geom, wl.C, cls, bbox, iou, counts, boxes, labels, idx, w, F.LossConfig())
    os.environ.pop("RADET_DENSE_PTS", None)
    F.scale_grads(geom, wl.C, grads, torch.tensor([2.0, 0.5, 1.0], device=dev))
    F.get_targets(geom, shapes, wl.C, counts, boxes, labels, idx, w)
    shp = torch.tensor([[wl.H, wl.W]] * len(batch), dtype=torch.int32, device=dev)
    sf = torch.full((len(batch), 4), 1.25, device=dev)
    for typ in ("vote", "global_vote", "nms"):
        cfg = F.DetectConfig(score_thr=0.02, nms_pre=200, max_per_img=50, nms_type=typ, iou_threshold=0.65, cluster_score=["cls", "iou"],
                             vote_score=["iou", "cls"], iou_enable=(typ == "vote"))
        dets, dl, num = F.get_bboxes(geom, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    rows, cats, n = F.get_candidates(geom, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    F.bbox2result_batch(dets, dl, num, wl.C, xywh=True)
# detect_bin_kernel beyond its register-resident keys: ~100 k candidates on the finest level (score_thr 1e-4), kept keys
# compacted in shared memory (nms_pre 300), more than the staging area holds (3 000), no per-level limit (-1); many seeds
# per image through detect_rank_kernel (max_per_img 1 000 -> bitonic path)
wl = syn.Workload("s4", 480, 640, 21, 1, 4, 4, 9)
batch = syn.make_batch(wl)
shapes = geom.level_shapes(wl.H, wl.W)
Pn = geom.num_points(shapes)
ho = syn.make_head_outputs(wl, batch, [np.zeros(Pn, np.int64) for _ in batch])
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, bbox, iou = [T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou]
shp = torch.tensor([[wl.H, wl.W]] * len(batch), dtype=torch.int32, device=dev)
sf = torch.ones((len(batch), 4), device=dev)
for nms_pre in (300, 3000, -1):
    cfg = F.DetectConfig(score_thr=1e-4, nms_pre=nms_pre, max_per_img=1000, nms_type="vote", iou_threshold=0.65,
                         cluster_score=["cls", "iou"], vote_score=["iou", "cls"])
    F.get_candidates(geom, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
    if nms_pre == 300:
        F.get_bboxes(geom, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
# radet.ops lists (shared-memory and global-memory variants of nms_list_kernel), standalone coder / losses
rs = np.random.RandomState(0)
for n in (300, 6000):
    c = rs.uniform(0, 500, (n // 10, 2)).repeat(10, 0) + rs.normal(0, 3, (n // 10 * 10, 2))
    wh = rs.uniform(20, 60, (n // 10 * 10, 2))
    b = torch.from_numpy(np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)).to(dev)
    sc = torch.from_numpy(rs.uniform(0.05, 1, b.shape[0]).astype(np.float32)).to(dev)
    lb = torch.from_numpy(rs.randint(0, 5, b.shape[0])).to(dev)
    P.ops.vote_nms(b, sc, lb, dict(iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"]), score_factor=sc.clone(), max_num=100)
    P.ops.cluster_nms(b, sc, lb, 0.65)
pri = torch.rand(100, 4, device=dev) * 100
pri[:, 2:] += pri[:, :2] + 1
F.tblr_decode(pri, F.tblr_encode(pri, pri + 1, 0.125), 0.125, max_shape=(100, 100))
x = torch.randn(64, 7, device=dev, requires_grad=True)
F.sigmoid_focal_loss_elementwise(x, torch.randint(0, 8, (64,), device=dev), 2.0, 0.25).sum().backward()
F.giou_loss_elementwise(pri.clone().requires_grad_(), pri + 2, 1e-6).sum().backward()
F.bce_with_logits_elementwise(torch.randn(50, device=dev, requires_grad=True), torch.rand(50, device=dev)).sum().backward()
# head-tower epilogues (csrc/tower.cu): aligned and odd plane sizes, with and without affine parameters
for shape, groups in (((2, 64, 12, 20), 32), ((2, 24, 7, 5), 4)):
    xt = torch.randn(shape, device=dev, requires_grad=True)
    wt = torch.randn(shape[1], device=dev, requires_grad=True)
    bt = torch.randn(shape[1], device=dev, requires_grad=True)
    F.gn_relu(xt, wt, bt, groups).sum().backward()
    F.gn_relu(xt.detach().requires_grad_(), None, None, groups).sum().backward()
    F.scale_relu(xt, torch.tensor(1.3, device=dev, requires_grad=True)).sum().backward()
torch.cuda.synchronize()
print("sanitize driver done")
