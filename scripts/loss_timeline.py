#!/usr/bin/env python
"""Development aid: globaltimer stamps of the fused loss kernel (per ticket) for one launch."""
import ctypes, os, sys
os.environ.setdefault("RADET_LOSS_IMPL", "fused")
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radet_b200 import functional as F, synthetic as syn, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
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
n1 = (B * P + 511) // 512
dbg = torch.zeros((200000, 8), dtype=torch.int64, device=dev)
lib.radet_debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
run(sets[1]); torch.cuda.synchronize()
lib.radet_debug_set_buffer(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(sets[2]); e1.record(); torch.cuda.synchronize()
print("event time of one more launch (us):", e0.elapsed_time(e1) * 1e3)
d = dbg.cpu().numpy().astype(np.float64)
cta = d[190000:199000]; fin = d[199990]
d = d[:190000]
used = d[:, 0] > 0
t0 = min(d[used, 0].min(), cta[cta[:, 0] > 0, 0].min())
cta = np.where(cta > 0, (cta - t0) / 1e3, np.nan); fin = (fin - t0) / 1e3
d = np.where(d > 0, (d - t0) / 1e3, np.nan)
p1 = d[:n1]; rest = d[n1:][used[n1:]]
it = rest[np.isfinite(rest[:, 2])]; bx = rest[~np.isfinite(rest[:, 2])]
q = lambda a: "n=%d min %.1f med %.1f max %.1f" % (np.isfinite(a).sum(), np.nanmin(a), np.nanmedian(a), np.nanmax(a))
print(f"{name}: n1={n1} items={len(it)}   (us since the first stamp)")
print(" CTA start          ", q(cta[:, 0])); print(" CTA got tickets    ", q(cta[:, 1])); print(" warp exits         ", q(cta[:, 2:6]))
print(" finalise begin/end ", fin[0], fin[1])
print(" phase1 start       ", q(p1[:, 0])); print(" phase1 a summed    ", q(p1[:, 1])); print(" phase1 a reported  ", q(p1[:, 2]))
print(" flag_a published   ", q(p1[:, 3])); print(" phase1 b computed  ", q(p1[:, 4])); print(" flag_b published   ", q(p1[:, 6])); print(" phase1 end         ", q(p1[:, 5]))
print(" item start         ", q(it[:, 0])); print(" item has num_pos   ", q(it[:, 1])); print(" item planes done   ", q(it[:, 2])); print(" item end           ", q(it[:, 3]))
