"""Shared helpers for the parity tests: golden loading and synthetic-case reconstruction."""
import hashlib
import os

import numpy as np

from radet_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = syn.Workload("small_160x128_B2_C21_G4", 128, 160, 21, 2, 4, 4, 9)
CFG3B2 = syn.Workload("cfg3_B2", 480, 640, 30, 2, 10, 30, 3)
HEAD_CASES = {"small": SMALL, "cfg1": syn.WORKLOADS["cfg1"], "cfg3b2": CFG3B2}


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def sha(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


def assign_cases():
    """(name, ImageGT-or-edge-dict) for every record of assign.npz."""
    g = load("assign.npz")
    names = sorted({k.split("/")[0] for k in g.files})
    return g, names


def image_for(g, name):
    """Re-create the inputs of an assignment golden record; returns dict(boxes, labels, grid(u8 [G,h8,w8]), H, W, seed)."""
    if name.startswith("edge_"):
        H, W = 480, 640
        grid = np.unpackbits(g[f"{name}/grid"], axis=-1)[..., : (W + 7) // 8]
        return dict(boxes=g[f"{name}/gt_bboxes"], labels=g[f"{name}/gt_labels"], grid=grid.astype(np.uint8),
                    H=H, W=W, seed=int(g[f"{name}/seed"]))
    key, i = name.rsplit("_", 1)
    wl = syn.WORKLOADS[key]
    img = syn.make_batch(wl, 1, int(i))[0]
    grid = syn.sample_grid(img.masks)
    assert (sha(img.gt_bboxes, img.gt_labels, grid) == g[f"{name}/sha"]).all(), "synthetic generator drifted"
    return dict(boxes=img.gt_bboxes, labels=img.gt_labels, grid=grid, H=img.H, W=img.W, seed=img.seed)


def head_case(key):
    """batch, idx_list, w_list, head outputs for a head.npz case (assignment from assign oracle is NOT used:
    idx/w come from the oracle run in the test, pinned separately by assign.npz)."""
    wl = HEAD_CASES[key]
    batch = syn.make_batch(wl)
    return wl, batch
