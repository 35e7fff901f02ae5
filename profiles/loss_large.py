#!/usr/bin/env python
"""Runs the fused loss a few times at the cfg5 per-GPU shape (1280x960, B=16, C=30; ~120 MB per launch) — for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radet_b200 import functional as F, synthetic as syn

wl = syn.WORKLOADS["cfg5"]; B, C = wl.B, wl.C; dev = "cuda"
geom = F.Geometry(); shapes = geom.level_shapes(wl.H, wl.W); P = geom.num_points(shapes)
imgs = [syn.make_image(np.random.RandomState(50 + i), wl.H, wl.W, C, 12 + 5 * i) for i in range(2)]
counts = [imgs[i % 2].gt_bboxes.shape[0] for i in range(B)]
off = F.offsets_of(counts, dev)
boxes = torch.from_numpy(np.concatenate([imgs[i % 2].gt_bboxes for i in range(B)])).to(dev)
labels = torch.from_numpy(np.concatenate([imgs[i % 2].gt_labels for i in range(B)])).to(dev)
grids = torch.from_numpy(np.concatenate([syn.sample_grid(imgs[i % 2].masks) for i in range(B)])).to(dev)
gh, gw = grids.shape[1:]
idx, w, _ = F.assign(geom, shapes, counts, boxes, F.pack_masks(grids, 1, gh, gw), (gh, gw),
                     seeds=torch.arange(B, dtype=torch.int32, device=dev), gt_offsets=off)
g = torch.Generator(device=dev).manual_seed(0)
sets = []
for r in range(3):
    sets.append(([torch.randn((B, C, h, w_), device=dev, generator=g) - 4.6 for h, w_ in shapes],
                 [torch.relu(torch.randn((B, 4, h, w_), device=dev, generator=g) + 1) for h, w_ in shapes],
                 [torch.randn((B, 1, h, w_), device=dev, generator=g) for h, w_ in shapes]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for i in range(n):
    c, b, o = sets[i % 3]
    F.loss_fwd_bwd(geom, C, c, b, o, counts, boxes, labels, idx, w, F.LossConfig(), gt_offsets=off)
torch.cuda.synchronize()
print("done", P * B)
