"""BBOX_CODERS / ANCHOR_GENERATORS mirrors the configs name (configs/bop/r50_ycbv_pbr.py:37-45)."""
import numpy as np
import torch

from .. import functional as F
from .registry import ANCHOR_GENERATORS, BBOX_CODERS


@BBOX_CODERS.register_module()
class TBLRBBoxCoder:
    """core/bbox/coder/tblr_bbox_coder.py:8-68.  Inside the fused head the coder is folded into the kernels; the
    standalone encode/decode run the radet_tblr_* kernels (CUDA tensors only)."""

    def __init__(self, normalizer=4.0, clip_border=True):
        if not isinstance(normalizer, float):
            raise NotImplementedError("per-dimension normalizer lists are outside the implemented surface")
        self.normalizer = normalizer
        self.clip_border = clip_border

    def encode(self, bboxes, gt_bboxes):
        assert bboxes.size(0) == gt_bboxes.size(0)
        assert bboxes.size(-1) == gt_bboxes.size(-1) == 4
        return F.tblr_encode(bboxes, gt_bboxes, self.normalizer)

    def decode(self, bboxes, pred_bboxes, max_shape=None):
        assert pred_bboxes.size(0) == bboxes.size(0)
        return F.tblr_decode(bboxes, pred_bboxes, self.normalizer, max_shape=max_shape, clip_border=self.clip_border)


@ANCHOR_GENERATORS.register_module()
class AnchorGenerator:
    """core/anchor/anchor_generator.py:10-347 restricted to what RADet uses: ONE square prior per cell
    (ratios=[1], scales_per_octave=1, center_offset=0).  The kernels compute the priors in closed form; this class
    carries the geometry and can still materialise them (`grid_anchors`) for callers that want tensors."""

    def __init__(self, strides, ratios, scales=None, base_sizes=None, scale_major=True, octave_base_scale=None,
                 scales_per_octave=None, centers=None, center_offset=0.):
        assert ((octave_base_scale is not None and scales_per_octave is not None) ^ (scales is not None)), \
            'scales and octave_base_scale with scales_per_octave cannot be set at the same time'
        if scales is None:
            scales = [octave_base_scale * 2 ** (i / scales_per_octave) for i in range(scales_per_octave)]
        if len(ratios) != 1 or float(ratios[0]) != 1.0 or len(scales) != 1 or center_offset != 0 or centers is not None \
                or base_sizes is not None:
            raise NotImplementedError("radet_b200 supports the RADet prior layout only: one square prior per cell "
                                      "(ratios=[1], one scale, center_offset=0)")
        self.strides = [(int(s), int(s)) if not isinstance(s, (tuple, list)) else tuple(s) for s in strides]
        if any(s[0] != s[1] for s in self.strides):
            raise NotImplementedError("anisotropic strides are outside the implemented surface")
        self.base_sizes = [min(s) for s in self.strides]
        self.scales = torch.Tensor([float(scales[0])])
        self.ratios = torch.Tensor([1.0])
        self.octave_base_scale = octave_base_scale
        self.scales_per_octave = scales_per_octave
        self.scale_major = scale_major
        self.centers = None
        self.center_offset = 0.

    @property
    def anchor_scale(self):
        return float(self.scales[0])

    @property
    def num_base_anchors(self):
        return [1 for _ in self.strides]

    @property
    def num_levels(self):
        return len(self.strides)

    def _geometry(self):
        from .. import functional as F
        return F.Geometry(strides=[sw for sw, _ in self.strides], regress_ranges=[(0.0, 0.0)] * len(self.strides),
                          anchor_scale=self.anchor_scale)

    def grid_anchors(self, featmap_sizes, device='cuda'):
        """anchor_generator.py:206-271.  On a CUDA device: one launch of `radet_grid_priors`; a CPU device is only served for
        the docstring KAT / host-side shape checks (small-integer arithmetic, exact either way)."""
        assert self.num_levels == len(featmap_sizes)
        sizes = [(int(h), int(w)) for h, w in featmap_sizes]
        if torch.device(device).type == 'cuda':
            from .. import functional as F
            flat = F.grid_priors(self._geometry(), sizes, torch.device(device))[0]
            return list(flat.split([h * w for h, w in sizes]))
        out = []
        for (h, w), (s, _) in zip(sizes, self.strides):
            xs = np.tile(np.arange(w, dtype=np.float32) * s, h)
            ys = np.repeat(np.arange(h, dtype=np.float32) * s, w)
            half = np.float32(0.5 * self.anchor_scale * s)
            a = np.stack([xs - half, ys - half, xs + half, ys + half], 1)
            out.append(torch.from_numpy(a).to(device))
        return out

    def valid_flags(self, featmap_sizes, pad_shape, device='cuda'):
        """anchor_generator.py:273-298."""
        sizes = [(int(h), int(w)) for h, w in featmap_sizes]
        if torch.device(device).type == 'cuda':
            from .. import functional as F
            flags = F.grid_priors(self._geometry(), sizes, torch.device(device), pad_shape=pad_shape[:2], want_anchors=False)[1]
            return list(flags.bool().split([h * w for h, w in sizes]))
        flags = []
        for (fh, fw), (sw, sh) in zip(sizes, self.strides):
            h, w = pad_shape[:2]
            vh, vw = min(int(np.ceil(h / sh)), int(fh)), min(int(np.ceil(w / sw)), int(fw))
            f = np.zeros((int(fh), int(fw)), bool)
            f[:vh, :vw] = True
            flags.append(torch.from_numpy(f.reshape(-1)).to(device))
        return flags
