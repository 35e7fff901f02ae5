#!/usr/bin/env python
"""Development aid: per-phase clock64() stamps of class_nms_kernel for one get_bboxes call on the bench workload."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radet_b200 import functional as F, synthetic as syn, _lib

wl = syn.WORKLOADS["cfg2"]; dev = "cuda"; geom = F.Geometry(); B = wl.B
shapes = geom.level_shapes(wl.H, wl.W)
batch = syn.make_batch(wl, B)
counts = [im.gt_bboxes.shape[0] for im in batch]
grids = torch.from_numpy(np.concatenate([syn.sample_grid(im.masks) for im in batch])).to(dev)
gh, gw = grids.shape[1:]
idx, w, _ = F.assign(geom, shapes, counts, torch.from_numpy(np.concatenate([im.gt_bboxes for im in batch])).to(dev),
                     F.pack_masks(grids, 1, gh, gw), (gh, gw), seeds=torch.tensor([im.seed for im in batch], dtype=torch.int32, device=dev))
ho = syn.make_head_outputs(wl, batch, list(idx.cpu().numpy()))
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cls, bbox, iou = [T(m) for m in ho.cls], [T(m) for m in ho.bbox], [T(m) for m in ho.iou]
cfg = F.DetectConfig(score_thr=wl.score_thr, nms_type="vote", iou_threshold=0.65, cluster_score=["cls", "iou"], vote_score=["iou", "cls"])
shp = torch.tensor([[im.H, im.W] for im in batch], dtype=torch.int32, device=dev); sf = torch.ones((B, 4), device=dev)
lib = _lib.load()
lib.radet_debug_set_buffer.argtypes = [ctypes.c_void_p]
for _ in range(3):
    F.get_bboxes(geom, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
dbg = torch.zeros((B * wl.C, 16), dtype=torch.int64, device=dev)
lib.radet_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
F.get_bboxes(geom, wl.C, cls, bbox, iou, shp, sf, cfg, rescale=True)
torch.cuda.synchronize()
lib.radet_debug_set_buffer(None)
d = dbg.cpu().numpy()
names = ["load", "sort", "decode", "mask", "scan", "vote"]
ph = np.diff(d[:, :7], axis=1)
print("m: mean %.0f max %d   seeds: mean %.0f max %d" % (d[:, 8].mean(), d[:, 8].max(), d[:, 9].mean(), d[:, 9].max()))
print("cycles per phase (mean / max over the %d class CTAs):" % len(d))
for k, n in enumerate(names):
    print(f"  {n:7s} {ph[:, k].mean():10.0f} {ph[:, k].max():10.0f}")
print("  total   %10.0f %10.0f" % ((d[:, 6] - d[:, 0]).mean(), (d[:, 6] - d[:, 0]).max()))
print("  span of all CTAs (first start .. last end): %d cycles" % (d[:, 6].max() - d[:, 0].min()))

L = len(shapes)
print("detect_bin per (image 0, level): cycles load | select | slots | emit")
for l in range(L):
    r = d[0 * L + l, 10:15]
    print("  level %d: %s" % (l, " ".join("%6d" % v for v in np.diff(r))))
print("detect_rank per image: cycles load | select")
for b_ in range(min(B, 4)):
    r = d[64 + b_, 10:13]
    print("  image %d: %s" % (b_, " ".join("%6d" % v for v in np.diff(r))))

# ---- assign_resolve_kernel phases
dbg2 = torch.zeros((B, 16), dtype=torch.int64, device=dev)
seeds = torch.tensor([im.seed for im in batch], dtype=torch.int32, device=dev)
boxes = torch.from_numpy(np.concatenate([im.gt_bboxes for im in batch])).to(dev)
bits = F.pack_masks(grids, 1, gh, gw)
for _ in range(3):
    F.assign(geom, shapes, counts, boxes, bits, (gh, gw), seeds=seeds)
lib.radet_debug_set_buffer(ctypes.c_void_p(dbg2.data_ptr()))
F.assign(geom, shapes, counts, boxes, bits, (gh, gw), seeds=seeds)
torch.cuda.synchronize()
lib.radet_debug_set_buffer(None)
d = dbg2.cpu().numpy()
print("assign_resolve per image: G, uniforms used, candidate points M, then cycles: seed+twist | workers phases 1-3 | join | sampling | tail(phase 5)")
for r in d:
    print(f"  G={r[8]:3d} used={r[9]:4d} M={r[10]:5d}  state={r[1]-r[0]:6d} | compaction={r[11]-r[0]:6d} claiming={r[12]-r[11]:6d} owners={r[2]-r[12]:6d} | join={r[3]-r[0]:6d} sampling={r[4]-r[3]:6d} tail={r[5]-r[4]:6d} total={r[5]-r[0]:6d}")
