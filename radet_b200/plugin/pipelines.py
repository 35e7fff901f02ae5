"""PIPELINES mirror: LabelAssignment (radet/datasets/pipelines/label_assignment.py:14-201) on the GPU.

`LabelAssignment.__call__(results)` keeps the reference contract: it reads results['img_shape'], ['gt_bboxes'],
['gt_labels'], ['distance_maps'] (a BitmapMasks-like object with .to_ndarray(), or a uint8 ndarray [G,H,W]) and writes
results['points_to_gt_index'] (np.int64 [P]) and results['points_weight'] (np.float32 [P]), consuming numpy's GLOBAL
legacy RNG exactly like np.random.choice does in the reference: the MT19937 state is shipped to the device, advanced
there and written back, so interleaved np.random users (RandomFlip, ...) see the same stream as with the reference.

`assign_batch(...)` is the batched entry point (new capability): many images per launch, masks shipped as
the stride-8 sample grid only.

`PackVisibleMaskGrid` is the CPU-only pipeline step for DataLoader workers: it replaces `LabelAssignment`
in `train_pipeline`, emits the sample grid and a per-image seed under the reference's own `Collect` keys
(`points_to_gt_index`, `points_weight`; configs/base/datasets/bop_detection.py:36), and `RADetHead.forward_train` then
runs the assignment for the whole batch on the training process's GPU.
"""
import math

import numpy as np
import torch

from .. import functional as F
from .._lib import RadetError
from .registry import PIPELINES, build_anchor_generator

INF = 1e8


def check_binary_grid(grid):
    """A visible mask has one non-zero value; a graded (real-valued, quantised) distance map would be binarised silently
    by the bit packing, so it is refused here (label_assignment.py:98 thresholds the continuous value)."""
    if grid.size == 0:
        return
    g = grid.reshape(grid.shape[0], -1)
    hi = g.max(axis=1)
    lo = np.where(g > 0, g, 255).min(axis=1)
    if np.any((hi > 0) & (lo != hi)):
        raise NotImplementedError("distance_maps with more than one non-zero value per ground truth (graded distance maps) are "
                                  "outside the implemented surface: binary visible masks only")


def refuse_forked_cuda(what):
    """CUDA cannot be re-initialised in a forked child of a process that already uses it -- which is exactly where mmcv's
    default DataLoader workers (fork, workers_per_gpu >= 1) would run a pipeline step."""
    if torch.cuda._is_in_bad_fork():
        raise RadetError(f"{what} runs on the GPU, but this is a forked DataLoader worker of a process that already "
                         "initialised CUDA. Use the CPU-only `PackVisibleMaskGrid` step in the pipeline instead (the assignment "
                         "then runs batched inside RADetHead.forward_train), or set workers_per_gpu=0 / "
                         "multiprocessing_context='spawn'.")


@PIPELINES.register_module()
class LabelAssignment:
    def __init__(self, strides=(8, 16, 32, 64, 128),
                 regress_ranges=((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF)),
                 anchor_generator_cfg=None, positive_num=10, neg_threshold=0.2, adapt_positive_num=False,
                 balance_sample=False, multiply_samplepro_for_weight=False, ambiguous_sample='min_area',
                 random_sample_by_distance=True, device=None):
        assert len(strides) == len(regress_ranges)
        # configuration surface implemented on the device (include/radet_b200.h, radet_assign)
        if not random_sample_by_distance:
            raise NotImplementedError("radet_b200.LabelAssignment implements random_sample_by_distance=True (every shipped "
                                      "config); the unweighted np.random.choice / permutation streams are not on the device")
        if ambiguous_sample != 'min_area':
            # 'max_dis' references an undefined variable in the reference (label_assignment.py:158-161) and crashes there
            raise NotImplementedError("ambiguous_sample must be 'min_area'")
        if not (0.0 < neg_threshold < 1.0):
            # binary masks: pro is 1 (visible) or the 1e-8 clip, so every threshold in (0, 1) selects the visible candidates
            # (label_assignment.py:98-100); at 0 the reference keeps EVERY candidate with p ~ {1, 1e-8}, which the device
            # path does not implement
            raise NotImplementedError("neg_threshold must be in (0,1) (binary visible masks)")
        self.num_levels = len(strides)
        self.strides = tuple(strides)
        self.regress_ranges = tuple(tuple(r) for r in regress_ranges)
        self.positive_num = positive_num
        self.ambiguous_sample = ambiguous_sample
        self.neg_threshold = neg_threshold
        self.adapt_positive_num = adapt_positive_num
        self.balance_sample = balance_sample
        self.random_sample_by_distance = random_sample_by_distance
        self.multiply_sample_pro_for_weight = multiply_samplepro_for_weight
        self.anchor_generator = build_anchor_generator(anchor_generator_cfg)
        if tuple(s[0] for s in self.anchor_generator.strides) != tuple(self.strides):
            raise RadetError("anchor_generator strides and LabelAssignment strides differ")
        self.geom = F.Geometry(self.strides, self.regress_ranges, self.anchor_generator.anchor_scale)
        self.device = device

    # ------------------------------------------------------------------ helpers
    def _dev(self):
        if not torch.cuda.is_available():
            raise RadetError("LabelAssignment needs a CUDA device: radet_b200 has no CPU path")
        return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())

    def _mask_grid(self, distance_maps, G, H, W):
        """The only pixels the assignment reads are (y*step, x*step) (label_assignment.py:80-85): ship just those."""
        m = distance_maps.to_ndarray() if hasattr(distance_maps, "to_ndarray") else np.asarray(distance_maps)
        step = self.geom.mask_step
        gh, gw = math.ceil(H / step), math.ceil(W / step)
        if G == 0:
            return np.zeros((0, gh, gw), np.uint8), gh, gw
        if m.dtype != np.uint8 and m.dtype != np.bool_:
            raise NotImplementedError("real-valued distance maps are outside the implemented surface (binary visible masks only)")
        if m.ndim != 3 or m.shape[0] != G or m.shape[1] < H or m.shape[2] < W:
            raise RadetError(f"distance_maps must be [G,>=H,>=W] = [{G},{H},{W}], got {m.shape}")
        grid = np.ascontiguousarray(m[:, :H:step, :W:step]).view(np.uint8)
        check_binary_grid(grid)
        return grid, gh, gw

    def assign_batch(self, img_shapes, gt_bboxes_list, mask_grids, *, seeds=None, mt_states=None, uniforms=None, weight_sums=None):
        """Batched device entry point.  All images must share (H, W).
        gt_bboxes_list: list of np/torch [G_i,4]; mask_grids: list of uint8 [G_i, ceil(H/step), ceil(W/step)]
        (the stride-`step` sample grid, see `_mask_grid`).  Exactly one RNG source: seeds (np.random.seed per image),
        mt_states [B,625] (full legacy states, advanced in place) or uniforms [B,n].
        weight_sums: optional CUDA float64 [B], written: per-image sum of the weights of the assigned points -- handed to the
        loss (functional.loss_fwd_bwd(weight_sums=...)) it spares the dense pass the wait for num_pos.
        Returns device tensors points_to_gt_index [B,P] int64, points_weight [B,P] f32, consumed [B] int32."""
        dev = self._dev()
        H, W = int(img_shapes[0][0]), int(img_shapes[0][1])
        if any((int(s[0]), int(s[1])) != (H, W) for s in img_shapes):
            raise RadetError("assign_batch: all images of a batch must share img_shape")
        shapes = self.geom.level_shapes(H, W)
        counts = [int(b.shape[0]) for b in gt_bboxes_list]
        step = self.geom.mask_step
        gh, gw = math.ceil(H / step), math.ceil(W / step)
        tot = sum(counts)
        if tot:
            boxes = torch.cat([torch.as_tensor(b, dtype=torch.float32).reshape(-1, 4) for b in gt_bboxes_list]).to(dev, non_blocking=True)
            grids = torch.cat([torch.as_tensor(g).reshape(-1, gh, gw) for g in mask_grids]).to(dev, non_blocking=True)
            bits = F.pack_masks(grids, 1, gh, gw)
        else:
            boxes = torch.zeros((0, 4), dtype=torch.float32, device=dev)
            bits = torch.zeros((0, gh, (gw + 31) // 32), dtype=torch.int32, device=dev)
        return F.assign(self.geom, shapes, counts, boxes, bits, (gh, gw), seeds=seeds, mt_states=mt_states, uniforms=uniforms,
                        positive_num=self.positive_num, balance_sample=self.balance_sample,
                        adapt_positive_num=self.adapt_positive_num,
                        multiply_samplepro_for_weight=self.multiply_sample_pro_for_weight, weight_sums=weight_sums)

    # ------------------------------------------------------------------ reference contract
    def __call__(self, results):
        refuse_forked_cuda("LabelAssignment.__call__")
        image_h, image_w, _ = results['img_shape']
        gt_bboxes = np.asarray(results['gt_bboxes'], np.float32).reshape(-1, 4)
        G = gt_bboxes.shape[0]
        grid, gh, gw = self._mask_grid(results['distance_maps'], G, image_h, image_w)
        dev = self._dev()
        # numpy's global legacy MT19937 state -> device -> back (np.random.choice parity, label_assignment.py:112,119)
        kind, key, pos, has_gauss, cached = np.random.get_state()
        if kind != 'MT19937':
            raise RadetError("numpy global RNG is not the legacy MT19937")
        st = np.empty(625, np.uint32)
        st[:624] = key
        st[624] = pos
        st_d = torch.from_numpy(st.view(np.int32)).to(dev).reshape(1, 625)
        idx, w, consumed = self.assign_batch([(image_h, image_w)], [gt_bboxes], [grid], mt_states=st_d)
        if int(consumed[0]) < 0:
            raise RadetError("LabelAssignment: adaptive positive_num above 32 for a ground truth of this image "
                             "(radet_assign consumed = %d)" % int(consumed[0]))
        st_back = st_d.cpu().numpy().view(np.uint32).reshape(625)
        np.random.set_state((kind, st_back[:624].copy(), int(st_back[624]), has_gauss, cached))
        results['points_to_gt_index'] = idx[0].cpu().numpy()
        results['points_weight'] = w[0].cpu().numpy()
        return results

    def __repr__(self):
        return (f"{self.__class__.__name__}(strides={self.strides}, positive_num={self.positive_num}, "
                f"balance_sample={self.balance_sample}) [radet_b200/sm_100a]")


@PIPELINES.register_module()
class PackVisibleMaskGrid:
    """CPU-only replacement of `LabelAssignment` inside DataLoader workers (numpy; never touches CUDA).

    The assignment reads the visible masks only at (y*step, x*step), step = gcd(strides) (label_assignment.py:80-85), so
    the worker ships that uint8 sample grid -- 4.8 KB per ground truth at 640x480 instead of a 51 KB index/weight pair
    per image computed by ~25 ms of numpy -- and a seed.  To stay a drop-in for `Collect(keys=[..., 'points_to_gt_index',
    'points_weight'])` and for `RADet.forward_train`'s signature (radet/models/detectors/radet.py:19-32), the two values
    travel under those two keys:

        results['points_to_gt_index']  uint8 [G, ceil(H/step), ceil(W/step)]   the sample grid of every GT's visible mask
        results['points_weight']       int64 [1]                               seed of the image's assignment RNG stream

    `RADetHead.forward_train` recognises the pair (uint8 rank-3 tensors) and runs `LabelAssignment.assign_batch` for the
    whole batch on its own GPU; the result equals what the reference's `LabelAssignment` returns when `np.random.seed(seed)`
    is called right before it.  The seed is drawn from numpy's global generator (one `randint`), so runs stay reproducible
    under mmdet's worker seeding; pass `seed_key` to take it from the results dict instead."""

    def __init__(self, strides=(8, 16, 32, 64, 128), seed_key=None):
        self.strides = tuple(int(s) for s in strides)
        self.step = int(np.gcd.reduce(np.asarray(self.strides)))
        self.seed_key = seed_key

    def __call__(self, results):
        H, W = int(results['img_shape'][0]), int(results['img_shape'][1])
        G = int(np.asarray(results['gt_bboxes']).reshape(-1, 4).shape[0])
        dm = results['distance_maps']
        m = dm.to_ndarray() if hasattr(dm, "to_ndarray") else np.asarray(dm)
        gh, gw = math.ceil(H / self.step), math.ceil(W / self.step)
        if G == 0:
            grid = np.zeros((0, gh, gw), np.uint8)
        else:
            if m.dtype != np.uint8 and m.dtype != np.bool_:
                raise NotImplementedError("real-valued distance maps are outside the implemented surface (binary visible masks only)")
            if m.ndim != 3 or m.shape[0] != G or m.shape[1] < H or m.shape[2] < W:
                raise RadetError(f"distance_maps must be [G,>=H,>=W] = [{G},{H},{W}], got {m.shape}")
            grid = np.ascontiguousarray(m[:, :H:self.step, :W:self.step]).view(np.uint8)
            check_binary_grid(grid)
        seed = int(results[self.seed_key]) if self.seed_key is not None else int(np.random.randint(0, 2 ** 31 - 1))
        results['points_to_gt_index'] = grid
        results['points_weight'] = np.asarray([seed], np.int64)
        return results

    def __repr__(self):
        return f"{self.__class__.__name__}(strides={self.strides}) [radet_b200, CPU]"


def is_mask_grid_handoff(points_to_gt_index):
    """True when `points_to_gt_index` carries PackVisibleMaskGrid's sample grids instead of assigned indices."""
    if points_to_gt_index is None:
        return False
    first = points_to_gt_index[0] if isinstance(points_to_gt_index, (list, tuple)) and len(points_to_gt_index) else points_to_gt_index
    return isinstance(first, torch.Tensor) and first.dtype in (torch.uint8, torch.bool) and first.dim() == 3
