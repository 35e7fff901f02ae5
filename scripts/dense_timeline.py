#!/usr/bin/env python
"""Development aid: globaltimer stamps of loss_dense_w_kernel (per item) for one launch at cfg5 / cfg2."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radet_b200 import functional as F, synthetic as syn, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
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
idx, w, _ = F.assign(geom, shapes, counts, boxes, F.pack_masks(grids, 1, gh, gw), (gh, gw), weight_sums=wsum, seeds=torch.arange(B, dtype=torch.int32, device=dev), gt_offsets=off)
g = torch.Generator(device=dev).manual_seed(0)
sets = [([torch.randn((B, C, h, w_), device=dev, generator=g) - 4.6 for h, w_ in shapes],
         [torch.relu(torch.randn((B, 4, h, w_), device=dev, generator=g) + 1) for h, w_ in shapes],
         [torch.randn((B, 1, h, w_), device=dev, generator=g) for h, w_ in shapes]) for _ in range(3)]
run = lambda s: F.loss_fwd_bwd(geom, C, s[0], s[1], s[2], counts, boxes, labels, idx, w, F.LossConfig(), gt_offsets=off, weight_sums=wsum)
for i in range(4):
    run(sets[i % 3])
torch.cuda.synchronize()
lib = _lib.load(); lib.radet_debug_set_buffer.argtypes = [ctypes.c_void_p]
dbg = torch.zeros((100000, 8), dtype=torch.int64, device=dev)
lib.radet_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
if os.environ.get("TIMELINE_EAGER"):
    run(sets[1]); torch.cuda.synchronize()
else:
    # the stamps of the LAST of many graph replays (clocks ramped, launches paced by the device, not by the host)
    gs, keep = [], []
    pool = torch.cuda.graph_pool_handle()
    for s_ in sets:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, pool=pool):
            keep.append(run(s_))
        gs.append(gr)
    for i in range(301):
        gs[i % 3].replay()
    torch.cuda.synchronize()
lib.radet_debug_set_buffer(None)
d = dbg.cpu().numpy().astype(np.float64)
dp = d[50000:]
d = d[:50000]
d = d[d[:, 0] > 0]
dp = dp[dp[:, 0] > 0]
t0 = min(d[:, 0].min(), dp[:, 0].min()) if len(dp) else d[:, 0].min()
d = np.where(d > 0, (d - t0) / 1e3, np.nan)
dp = np.where(dp > 0, (dp - t0) / 1e3, np.nan)
q = lambda a: "n=%d min %.1f p10 %.1f med %.1f p90 %.1f max %.1f" % (np.isfinite(a).sum(), np.nanmin(a), np.nanpercentile(a, 10), np.nanmedian(a), np.nanpercentile(a, 90), np.nanmax(a))
print(f"{name}: items={len(d)} (us since the first CTA started)  env={ {k: v for k, v in os.environ.items() if k.startswith('RADET_') or k == 'LOSS_HINT'} }")
for k, nm in enumerate(["pos CTA start", "pos scan done", "pos terms done", "pos atomic done", "pos last block end"]):
    if len(dp) and np.isfinite(dp[:, k]).any():
        print(f" {nm:18s}", q(dp[:, k]))
for k, nm in enumerate(["CTA start", "setup done", "stage 0 arrived", "stage 7 arrived", "planes done", "box planes done", "exit"]):
    print(f" {nm:18s}", q(d[:, k]))
print(" per item: start->setup", q(d[:, 1] - d[:, 0])); print(" setup->stage0", q(d[:, 2] - d[:, 1])); print(" stage0->stage7", q(d[:, 3] - d[:, 2]))
print(" after wait", q(d[:, 7]))
print(" stage7->planes done", q(d[:, 4] - d[:, 3])); print(" box planes", q(d[:, 5] - d[:, 4])); print(" exit", q(d[:, 6] - d[:, 5]))
