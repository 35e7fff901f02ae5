#!/usr/bin/env python
"""Times the head-tower epilogue kernels (csrc/tower.cu) against the torch ops they replace, per FPN level of one tower layer at
the bench batch (B=8, 256 channels, 640x480): CUDA events around graph replays over rotating inputs > L2, kernels called
directly (no autograd bookkeeping on either side)."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radet_b200 import functional as F, _lib

dev = "cuda"
B, C, G = int(os.environ.get("TOWER_B", 8)), 256, 32
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(0)
w = torch.randn(C, device=dev, generator=g); b = torch.randn(C, device=dev, generator=g)

def timeit(fn, sets, iters=20):
    for s in sets[:2]:
        fn(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for s in sets:
            fn(s)
    gr.replay(); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        gr.replay()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) * 1e3 / (iters * len(sets))

P = F._ptr
for h, w_ in [(60, 80), (30, 40), (15, 20)]:
    hw = h * w_
    el = B * C * hw
    R = max(3, int(np.ceil(2.2 * 126e6 / (el * 4 * 5))))
    sets = []
    for _ in range(R):
        x = torch.randn((B, C, h, w_), device=dev, generator=g)
        sets.append(dict(x=x, dy=torch.randn_like(x), y=torch.empty_like(x), dx=torch.empty_like(x), mean=torch.empty((B, G), device=dev),
                         rstd=torch.empty((B, G), device=dev), dg=torch.empty((B, C), device=dev), db=torch.empty((B, C), device=dev)))
    def f_fwd(s):
        assert lib.radet_gn_relu_forward(P(s["x"]), P(w), P(b), B, C, hw, G, 1e-5, P(s["y"]), P(s["mean"]), P(s["rstd"]), F._stream()) == 0
    def f_bwd(s):
        assert lib.radet_gn_relu_backward(P(s["dy"]), P(s["x"]), P(w), P(b), P(s["mean"]), P(s["rstd"]), B, C, hw, G, P(s["dx"]), P(s["dg"]),
                                          P(s["db"]), F._stream()) == 0
    def t_fwd(s):
        s["ty"], s["tm"], s["tr"] = torch.ops.aten.native_group_norm(s["x"], w, b, B, C, hw, G, 1e-5)
        s["tz"] = torch.relu(s["ty"])
    def t_bwd(s):
        d = torch.ops.aten.threshold_backward(s["dy"], s["tz"], 0)
        s["tdx"], s["tdg"], s["tdb"] = torch.ops.aten.native_group_norm_backward(d, s["x"], s["tm"], s["tr"], w, B, C, hw, G, [True, True, True])
    for s in sets:
        f_fwd(s); t_fwd(s)
    res = {"level": f"{h}x{w_}", "B": B, "C": C, "elements": el}
    res["fwd_us"] = timeit(f_fwd, sets)
    res["bwd_us"] = timeit(f_bwd, sets)
    res["fwd_torch_us"] = timeit(t_fwd, sets)
    res["bwd_torch_us"] = timeit(t_bwd, sets)
    res["fwd_GBps_at_8B"] = el * 8 / res["fwd_us"] / 1e3
    res["bwd_GBps_at_12B"] = el * 12 / res["bwd_us"] / 1e3
    print(json.dumps(res))
