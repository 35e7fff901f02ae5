"""Seeded synthetic YCB-V / LM-O / ITODD-shaped inputs for the dense-head hot path.

Follows the input protocol of SURVEY.md §8(d): the same tensors feed the
oracle, the golden-vector generator, the parity tests and ``bench.py``.
numpy only — no dependency on the CUDA library or on ``oracle/``.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

STRIDES = (8, 16, 32, 64, 128)
# label_assignment.py:32 — inclusive on both ends (label_assignment.py:73-74)
REGRESS_RANGES = ((-1.0, 64.0), (64.0, 128.0), (128.0, 256.0), (256.0, 512.0), (512.0, 1e8))


def level_shapes(H: int, W: int, strides=STRIDES) -> List[Tuple[int, int]]:
    """ceil(H/s) x ceil(W/s) per level (label_assignment.py:138)."""
    return [(math.ceil(H / s), math.ceil(W / s)) for s in strides]


def num_points(H: int, W: int, strides=STRIDES) -> int:
    return sum(h * w for h, w in level_shapes(H, W, strides))


@dataclass
class Workload:
    """One of BASELINE.json's configs."""
    name: str
    H: int
    W: int
    C: int
    B: int
    g_lo: int
    g_hi: int
    cfg_id: int
    score_thr: float = 0.05
    nms_pre: int = 1000
    max_per_img: int = 100
    iou_threshold: float = 0.65


WORKLOADS = {
    # configs[0]: CPU-runnable parity case
    "cfg1": Workload("cfg1_640x480_B2_C21_G8", 480, 640, 21, 2, 8, 8, 1),
    # configs[1]: the bench workload (metric is quoted on this one)
    "cfg2": Workload("cfg2_ycbv_640x480_B8_C21_G3-21", 480, 640, 21, 8, 3, 21, 2),
    "cfg3": Workload("cfg3_lmo_tless_640x480_B8_C30_G10-30", 480, 640, 30, 8, 10, 30, 3),
    "cfg4": Workload("cfg4_infer_640x480_B64_C21", 480, 640, 21, 64, 3, 21, 4, score_thr=0.1),
    "cfg5": Workload("cfg5_itodd_hb_1280x960_B16_C30_G5-30", 960, 1280, 30, 16, 5, 30, 5),
}


@dataclass
class ImageGT:
    gt_bboxes: np.ndarray   # [G,4] f32 x1,y1,x2,y2
    gt_labels: np.ndarray   # [G] i64
    masks: np.ndarray       # [G,H,W] u8 (visible masks, 0/1)
    H: int
    W: int
    seed: int = 0           # np.random.seed(seed) right before the assignment call


def make_image(rs: np.random.RandomState, H: int, W: int, C: int, G: int, occluded_frac: float = 0.05) -> ImageGT:
    cx = rs.uniform(0, W, G)
    cy = rs.uniform(0, H, G)
    lo, hi = math.log(24.0), math.log(0.6 * min(H, W))
    bw = np.exp(rs.uniform(lo, hi, G))
    bh = np.exp(rs.uniform(lo, hi, G))
    x1 = np.clip(cx - bw / 2, 0, W - 1)
    x2 = np.clip(cx + bw / 2, 0, W - 1)
    y1 = np.clip(cy - bh / 2, 0, H - 1)
    y2 = np.clip(cy + bh / 2, 0, H - 1)
    boxes = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
    labels = rs.randint(0, C, G).astype(np.int64)
    masks = np.zeros((G, H, W), np.uint8)
    for g in range(G):
        bx1, by1, bx2, by2 = boxes[g]
        ix1, iy1 = int(math.floor(bx1)), int(math.floor(by1))
        ix2, iy2 = min(W, int(math.ceil(bx2)) + 1), min(H, int(math.ceil(by2)) + 1)
        if ix2 <= ix1 or iy2 <= iy1:
            continue
        ys = np.arange(iy1, iy2, dtype=np.float32)[:, None]
        xs = np.arange(ix1, ix2, dtype=np.float32)[None, :]
        ex, ey = (bx1 + bx2) / 2, (by1 + by2) / 2
        rx, ry = max((bx2 - bx1) / 2, 1e-3), max((by2 - by1) / 2, 1e-3)
        m = (((xs - ex) / rx) ** 2 + ((ys - ey) / ry) ** 2) <= 1.0
        for _ in range(rs.randint(1, 4)):
            ow = rs.uniform(0.1, 0.6) * (bx2 - bx1)
            oh = rs.uniform(0.1, 0.6) * (by2 - by1)
            ox = rs.uniform(bx1, bx2)
            oy = rs.uniform(by1, by2)
            m &= ~((np.abs(xs - ox) <= ow / 2) & (np.abs(ys - oy) <= oh / 2))
        if rs.uniform() < occluded_frac:
            m[:] = False   # fully occluded: exercises the "everyone is non-neg" fallback
        masks[g, iy1:iy2, ix1:ix2] = m
    return ImageGT(boxes, labels, masks, H, W)


def make_batch(wl: Workload, B: Optional[int] = None, first_image: int = 0) -> List[ImageGT]:
    """Images first_image .. first_image+B-1 of the workload (RNG 1000*cfg + image_index)."""
    B = wl.B if B is None else B
    out = []
    for i in range(first_image, first_image + B):
        rs = np.random.RandomState(1000 * wl.cfg_id + i)
        G = int(rs.randint(wl.g_lo, wl.g_hi + 1))
        img = make_image(rs, wl.H, wl.W, wl.C, G)
        img.seed = 777 + i
        out.append(img)
    return out


def sample_grid(masks: np.ndarray, step: int = 8) -> np.ndarray:
    """The only mask pixels the assignment ever reads: (y*step, x*step) (label_assignment.py:80-85)."""
    return np.ascontiguousarray(masks[:, ::step, ::step])


@dataclass
class HeadOutputs:
    """NCHW maps per level, f32: cls [B,C,h,w], bbox [B,4,h,w] (post-ReLU, TBLR), iou [B,1,h,w]."""
    cls: List[np.ndarray] = field(default_factory=list)
    bbox: List[np.ndarray] = field(default_factory=list)
    iou: List[np.ndarray] = field(default_factory=list)


def tblr_targets(img: ImageGT, idx: np.ndarray, strides=STRIDES) -> np.ndarray:
    """(T,B,L,R)/stride for points with idx>0, zero elsewhere (radet_head.py:391, tblr_bbox_coder.py:71-114)."""
    shapes = level_shapes(img.H, img.W, strides)
    P = idx.shape[0]
    out = np.zeros((P, 4), np.float32)
    off = 0
    for (h, w), s in zip(shapes, strides):
        n = h * w
        ys, xs = np.divmod(np.arange(n), w)
        cx = (xs * s).astype(np.float32)
        cy = (ys * s).astype(np.float32)
        sel = idx[off:off + n] > 0
        g = idx[off:off + n][sel] - 1
        b = img.gt_bboxes[g]
        t = np.stack([cy[sel] - b[:, 1], b[:, 3] - cy[sel], cx[sel] - b[:, 0], b[:, 2] - cx[sel]], 1) / np.float32(s)
        out[off:off + n][sel] = t
        off += n
    return out


def make_head_outputs(wl: Workload, batch: List[ImageGT], idx_list: List[np.ndarray], seed_base: Optional[int] = None,
                      strides=STRIDES, boost: float = 6.0) -> HeadOutputs:
    """cls ~ N(-4.6,1) (+boost at assigned positives w.p. 0.7); bbox = relu(target + N(0,.5)) at positives,
    relu(N(1,1)) elsewhere; iou ~ N(0,1).  Generator seed cfg*100+level."""
    B = len(batch)
    H, W, C = wl.H, wl.W, wl.C
    shapes = level_shapes(H, W, strides)
    seed_base = wl.cfg_id * 100 if seed_base is None else seed_base
    tg = [tblr_targets(img, idx, strides) for img, idx in zip(batch, idx_list)]
    out = HeadOutputs()
    off = 0
    for lvl, (h, w) in enumerate(shapes):
        n = h * w
        rs = np.random.RandomState(seed_base + lvl)
        cls = rs.normal(-4.6, 1.0, (B, C, n)).astype(np.float32)
        bbox = rs.normal(1.0, 1.0, (B, 4, n)).astype(np.float32)
        iou = rs.normal(0.0, 1.0, (B, 1, n)).astype(np.float32)
        for b, (img, idx) in enumerate(zip(batch, idx_list)):
            li = idx[off:off + n]
            pos = np.nonzero(li > 0)[0]
            if pos.size:
                lab = img.gt_labels[li[pos] - 1]
                hit = rs.uniform(size=pos.size) < 0.7
                cls[b, lab[hit], pos[hit]] += np.float32(boost)
                noise = rs.normal(0.0, 0.5, (pos.size, 4)).astype(np.float32)
                bbox[b][:, pos] = (tg[b][off:off + n][pos] + noise).T
        np.maximum(bbox, 0, out=bbox)
        out.cls.append(cls.reshape(B, C, h, w))
        out.bbox.append(bbox.reshape(B, 4, h, w))
        out.iou.append(iou.reshape(B, 1, h, w))
        off += n
    return out


def img_metas(batch: List[ImageGT], scale: float = 1.0):
    sf = np.array([scale, scale, scale, scale], np.float32)
    return [dict(img_shape=(im.H, im.W, 3), pad_shape=(im.H, im.W, 3), scale_factor=sf) for im in batch]
