"""GraphedHotPath — the whole hot path of one batch as ONE CUDA-graph replay fed from a pinned host arena.

The eager plugin calls (`LabelAssignment.assign_batch`, `RADetHead.loss`, `RADetHead.get_bboxes`) cost a few hundred
microseconds of Python / launch overhead per batch, several times the GPU time of the kernels themselves.  For fixed
shapes (batch size, image size, an upper bound on GT per image) this class lays every host input of a batch out in one
pinned arena, mirrors it with one device arena, and captures

    [1 H2D copy is issued by `run`]  ->  seed MT19937 states || pack masks -> assign -> loss fwd+bwd (-> gradients)
                                     ||  select -> top-k/bin -> per-class NMS + vote -> rank
                                     ->  D2H of (losses, num_pos, dets, labels, num_dets) into a pinned result block

into a single graph (three forked branches, see bench.py).  Per batch the host then does: fill the arena views (or
hand over an arena that already holds the data), one async copy, one graph launch, one event wait.

It is a thin composition of the same C-ABI calls the eager classes make; numerics are identical.
"""
import math
from typing import Dict, Optional

import numpy as np
import torch

from .. import functional as F
from .._lib import MT_STATE_WORDS, RadetError


def _align(n, a=256):
    return (n + a - 1) // a * a


class GraphedHotPath:
    def __init__(self, head, assigner, batch_size: int, img_shape, max_gt_per_image: int = 32, device=None,
                 rescale: bool = True, with_inference: bool = True):
        if not torch.cuda.is_available():
            raise RadetError("GraphedHotPath needs a CUDA device")
        self.head, self.assigner = head, assigner
        self.B = int(batch_size)
        self.H, self.W = int(img_shape[0]), int(img_shape[1])
        self.cap = int(max_gt_per_image)
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.geom = head.geom
        self.shapes = self.geom.level_shapes(self.H, self.W)
        self.P = self.geom.num_points(self.shapes)
        self.C = head.num_classes
        self.rescale = rescale
        self.with_inference = with_inference
        step = self.geom.mask_step
        self.gh, self.gw = math.ceil(self.H / step), math.ceil(self.W / step)
        self.dcfg = F.DetectConfig.from_test_cfg(head.test_cfg) if with_inference else None
        B, cap, C = self.B, self.cap, self.C
        # ---- arena layout (bytes); every field 256-byte aligned
        fields = [("gt_bboxes", (B * cap, 4), np.float32), ("gt_labels", (B * cap,), np.int64), ("gt_offsets", (B + 1,), np.int32),
                  ("seeds", (B,), np.int32), ("img_shapes", (B, 2), np.int32), ("scale_factors", (B, 4), np.float32)]
        for l, (h, w) in enumerate(self.shapes):
            fields += [(f"cls{l}", (B, C, h, w), np.float32), (f"bbox{l}", (B, 4, h, w), np.float32), (f"iou{l}", (B, 1, h, w), np.float32)]
        # the only field whose used size varies from batch to batch goes last: `launch` copies the arena only up to
        # the last GT's mask grid (the padding up to max_gt_per_image never crosses the bus)
        fields += [("mask_grids", (B * cap, self.gh, self.gw), np.uint8)]
        self._layout, off = {}, 0
        for name, shape, dt in fields:
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            self._layout[name] = (off, shape, dt)
            off += _align(nbytes)
        self.arena_bytes = off
        self._dev_arena = torch.empty(self.arena_bytes, dtype=torch.uint8, device=self.dev)
        self._dviews = self._views(self._dev_arena)
        # worst-case host offsets: only used by the library to size shared memory / bit-set words
        self._off_host_cap = (np.arange(B + 1, dtype=np.int32) * cap)
        mp = self.dcfg.max_per_img if with_inference else 1
        self._res_fields = [("losses", (4,), torch.float32), ("dets", (B, mp, 5), torch.float32), ("labels", (B, mp), torch.int64),
                            ("num", (B,), torch.int32), ("consumed", (B,), torch.int32)]
        self._graph = None
        self._stream = torch.cuda.Stream(device=self.dev)
        self._side = [torch.cuda.Stream(device=self.dev) for _ in range(2)]
        self._wsum = torch.zeros((self.B,), dtype=torch.float64, device=self.dev)   # assignment -> loss: per-image weight sums
        with torch.cuda.device(self.dev):
            self._done = torch.cuda.Event()
        self.grads = None

    # ------------------------------------------------------------------ arenas
    def _views(self, buf: torch.Tensor) -> Dict[str, torch.Tensor]:
        tmap = {np.float32: torch.float32, np.int64: torch.int64, np.int32: torch.int32, np.uint8: torch.uint8}
        out = {}
        for name, (off, shape, dt) in self._layout.items():
            n = int(np.prod(shape)) * np.dtype(dt).itemsize
            out[name] = buf[off:off + n].view(tmap[dt]).view(shape)
        return out

    def new_host_arena(self):
        """(pinned uint8 buffer, dict of numpy views into it).  Fill the views, then call run(arena)."""
        buf = torch.empty(self.arena_bytes, dtype=torch.uint8).pin_memory()
        views = {k: v.numpy() for k, v in self._views(buf).items()}
        views["scale_factors"][:] = 1.0
        views["img_shapes"][:] = (self.H, self.W)
        views["gt_offsets"][:] = 0
        return buf, views

    def fill(self, views, gt_bboxes_list, gt_labels_list, mask_grids, seeds, cls, bbox, iou, scale_factors=None):
        """Copy one batch into arena views (GT rows packed contiguously, reference formats otherwise)."""
        counts = [int(b.shape[0]) for b in gt_bboxes_list]
        if max(counts, default=0) > self.cap or len(counts) != self.B:
            raise RadetError(f"batch of {len(counts)} images with up to {max(counts, default=0)} GT does not fit "
                             f"GraphedHotPath(batch_size={self.B}, max_gt_per_image={self.cap})")
        off = np.zeros(self.B + 1, np.int32)
        np.cumsum(counts, out=off[1:])
        views["gt_offsets"][:] = off
        n = int(off[-1])
        if n:
            views["gt_bboxes"][:n] = np.concatenate([np.asarray(b, np.float32).reshape(-1, 4) for b in gt_bboxes_list])
            views["gt_labels"][:n] = np.concatenate([np.asarray(l, np.int64).reshape(-1) for l in gt_labels_list])
            views["mask_grids"][:n] = np.concatenate([np.asarray(g, np.uint8).reshape(-1, self.gh, self.gw) for g in mask_grids])
        views["seeds"][:] = np.asarray(seeds, np.int64).astype(np.int32)
        if scale_factors is not None:
            views["scale_factors"][:] = np.asarray(scale_factors, np.float32).reshape(self.B, 4)
        for l in range(len(self.shapes)):
            views[f"cls{l}"][...] = cls[l]
            views[f"bbox{l}"][...] = bbox[l]
            views[f"iou{l}"][...] = iou[l]

    # ------------------------------------------------------------------ capture
    def _body(self):
        d = self._dviews
        L = len(self.shapes)
        cls, bbox, iou = [d[f"cls{l}"] for l in range(L)], [d[f"bbox{l}"] for l in range(L)], [d[f"iou{l}"] for l in range(L)]
        main = torch.cuda.current_stream()
        s_seed, s_det = self._side
        s_seed.wait_stream(main)
        with torch.cuda.stream(s_seed):
            states = F.seed_states(d["seeds"])
        dets = labels = num = None
        if self.with_inference:
            s_det.wait_stream(main)
            with torch.cuda.stream(s_det):
                dets, labels, num = F.get_bboxes(self.geom, self.C, cls, bbox, iou, d["img_shapes"], d["scale_factors"], self.dcfg,
                                                 rescale=self.rescale)
        off = (self._off_host_cap, d["gt_offsets"])
        bits = F.pack_masks(d["mask_grids"], 1, self.gh, self.gw)
        main.wait_stream(s_seed)
        idx, w, consumed = F.assign(self.geom, self.shapes, None, d["gt_bboxes"], bits, (self.gh, self.gw), mt_states=states,
                                    positive_num=self.assigner.positive_num, balance_sample=self.assigner.balance_sample,
                                    adapt_positive_num=self.assigner.adapt_positive_num,
                                    multiply_samplepro_for_weight=self.assigner.multiply_sample_pro_for_weight,
                                    gt_offsets=off, weight_sums=self._wsum)
        losses, grads = F.loss_fwd_bwd(self.geom, self.C, cls, bbox, iou, None, d["gt_bboxes"], d["gt_labels"], idx, w,
                                       self.head.loss_cfg, gt_offsets=off, weight_sums=self._wsum)
        if self.with_inference:
            main.wait_stream(s_det)
        return dict(losses=losses, grads=grads, dets=dets, labels=labels, num=num, idx=idx, w=w, consumed=consumed)

    def capture(self):
        """Warm up eagerly (workspaces, lazy module loading), then capture the step."""
        self._stream.wait_stream(torch.cuda.current_stream(self.dev))    # the arena may have been allocated on the caller's stream
        with torch.cuda.stream(self._stream):
            # initialise the arena on the stream the warm-up runs on (side streams do not order against the caller's)
            self._dev_arena.zero_()
            self._dviews["img_shapes"][:] = torch.tensor([self.H, self.W], dtype=torch.int32, device=self.dev)
            self._dviews["scale_factors"].fill_(1.0)
            for _ in range(2):
                self._body()
            self._stream.synchronize()
            # pinned result buffers (allocated before the capture)
            mp = self.dcfg.max_per_img if self.with_inference else 1
            self._res_host = {"losses": torch.empty(4, dtype=torch.float32).pin_memory(),
                              "consumed": torch.empty(self.B, dtype=torch.int32).pin_memory()}
            if self.with_inference:
                self._res_host["dets"] = torch.empty((self.B, mp, 5), dtype=torch.float32).pin_memory()
                self._res_host["labels"] = torch.empty((self.B, mp), dtype=torch.int64).pin_memory()
                self._res_host["num"] = torch.empty(self.B, dtype=torch.int32).pin_memory()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, stream=self._stream):
                self._out = self._body()
                for k, t in self._res_host.items():
                    t.copy_(self._out[k], non_blocking=True)
        self.grads = self._out["grads"]
        self.h2d_bytes = self.arena_bytes
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self._res_host.values())
        return self

    # ------------------------------------------------------------------ run
    def used_bytes(self, host_arena: torch.Tensor) -> int:
        """Bytes of the arena that carry data for the batch it holds (everything up to the last GT's mask grid)."""
        off, shape, _ = self._layout["gt_offsets"]
        n_gt = int(host_arena[off:off + 4 * (self.B + 1)].view(torch.int32)[self.B])
        moff = self._layout["mask_grids"][0]
        return min(self.arena_bytes, _align(moff + n_gt * self.gh * self.gw))

    def launch(self, host_arena: torch.Tensor, maps_resident: bool = False):
        """Enqueue: H2D of the used part of the arena, the graph, D2H of the results.  Returns immediately.

        maps_resident=True is the deployment shape: the head outputs (cls / bbox / iou maps) are already in the device
        arena (written there by the conv towers, `device_views()`), so only the ground truth, the seeds and the mask
        grids cross the bus (two copies: the small fields in front of the maps, the used mask grids behind them)."""
        if self._graph is None:
            self.capture()
        n = self.used_bytes(host_arena)
        with torch.cuda.stream(self._stream):
            if maps_resident:
                head = self._layout["cls0"][0]
                moff = self._layout["mask_grids"][0]
                self._dev_arena[:head].copy_(host_arena[:head], non_blocking=True)
                if n > moff:
                    self._dev_arena[moff:n].copy_(host_arena[moff:n], non_blocking=True)
                n = head + max(0, n - moff)
            else:
                self._dev_arena[:n].copy_(host_arena[:n], non_blocking=True)
            self._graph.replay()
            self._done.record(self._stream)
        self.last_h2d_bytes = n

    def device_views(self) -> Dict[str, torch.Tensor]:
        """Views into the device arena (e.g. `cls0`, `bbox0`, `iou0`, ...: where a model writes its head outputs for
        `launch(..., maps_resident=True)`)."""
        return self._dviews

    def wait(self):
        """Block until the last launch has finished; returns the pinned result block (valid until the next launch)
        — losses f32[4] = (loss_cls, loss_bbox, loss_iou, num_pos), dets [B,max,5], labels, num, consumed — while the
        gradients of the head outputs stay on the device in `self.grads` (cls, bbox, iou lists)."""
        self._done.synchronize()
        return self._res_host

    def run(self, host_arena: torch.Tensor):
        self.launch(host_arena)
        return self.wait()
