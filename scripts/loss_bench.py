#!/usr/bin/env python
"""Times radet_loss_fwd_bwd alone (CUDA events around graph replays over rotating input sets > L2).

    python scripts/loss_bench.py [cfg5|cfg2|cfg3] [iters]      env: RADET_FUSED_CG=2|4, RADET_LOSS_IMPL=fused|reg
    python scripts/loss_bench.py cfg5 ncu                       a few eager launches, for ncu
"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radet_b200 import functional as F, synthetic as syn

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
mode = sys.argv[2] if len(sys.argv) > 2 else "100"
wl = syn.WORKLOADS[name]; B, C = wl.B, wl.C; dev = "cuda"
geom = F.Geometry(); shapes = geom.level_shapes(wl.H, wl.W); P = geom.num_points(shapes)
imgs = [syn.make_image(np.random.RandomState(50 + i), wl.H, wl.W, C, wl.g_lo + (wl.g_hi - wl.g_lo) * i) for i in range(2)]
counts = [imgs[i % 2].gt_bboxes.shape[0] for i in range(B)]
off = F.offsets_of(counts, dev)
boxes = torch.from_numpy(np.concatenate([imgs[i % 2].gt_bboxes for i in range(B)])).to(dev)
labels = torch.from_numpy(np.concatenate([imgs[i % 2].gt_labels for i in range(B)])).to(dev)
grids = torch.from_numpy(np.concatenate([syn.sample_grid(imgs[i % 2].masks) for i in range(B)])).to(dev)
gh, gw = grids.shape[1:]
wsum = torch.zeros(B, dtype=torch.float64, device=dev) if os.environ.get("LOSS_HINT") else None
idx, w, _ = F.assign(geom, shapes, counts, boxes, F.pack_masks(grids, 1, gh, gw), (gh, gw), weight_sums=wsum,
                     seeds=torch.arange(B, dtype=torch.int32, device=dev), gt_offsets=off)
g = torch.Generator(device=dev).manual_seed(0)
per_set = B * P * (C + 5) * 4 * 2
R = max(3, int(np.ceil(2.2 * 126e6 / per_set)))
sets = []
for r in range(R):
    sets.append(([torch.randn((B, C, h, w_), device=dev, generator=g) - 4.6 for h, w_ in shapes],
                 [torch.relu(torch.randn((B, 4, h, w_), device=dev, generator=g) + 1) for h, w_ in shapes],
                 [torch.randn((B, 1, h, w_), device=dev, generator=g) for h, w_ in shapes]))
run = lambda s: F.loss_fwd_bwd(geom, C, s[0], s[1], s[2], counts, boxes, labels, idx, w, F.LossConfig(), gt_offsets=off, weight_sums=wsum)
if mode == "ncu":
    for i in range(6):
        run(sets[i % R])
    torch.cuda.synchronize()
    print("done")
    sys.exit(0)
iters = int(mode)
for i in range(3):
    out = run(sets[i % R])
torch.cuda.synchronize()
keep = []
pool = torch.cuda.graph_pool_handle()
gr = torch.cuda.CUDAGraph()      # one graph = the call on every rotating set, back to back (one graph-launch gap per R calls)
with torch.cuda.graph(gr, pool=pool):
    for s in sets:
        keep.append(run(s))
gr.replay()
torch.cuda.synchronize()
ts = []
nrep = max(1, iters // R)
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(nrep):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3 / (nrep * R))
us = float(np.median(ts))
nbytes = B * P * (8 * C + 52)
print(json.dumps({"workload": name, "us": us, "all_us": [round(t, 2) for t in ts], "bytes": nbytes, "GBps": nbytes / us / 1e3, "frac_of_6553.9": nbytes / us / 1e3 / 6553.9,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("RADET_") or k == "LOSS_HINT"}, "losses": keep[0][0].tolist()}))
