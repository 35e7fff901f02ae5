"""ctypes binding of libradet_b200.so (the C ABI declared in include/radet_b200.h).

The product path has NO fallback: if the library is missing or a call fails, this module raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libradet_b200.so")
MAX_LEVELS = 8
MAX_GT_PER_IMAGE = 256
MT_STATE_WORDS = 625

NMS_VOTE, NMS_GLOBAL_VOTE, NMS_PLAIN = 0, 1, 2

_ERR = {-1: "RADET_E_BADARG", -2: "RADET_E_TOO_MANY_GT", -3: "RADET_E_WORKSPACE", -4: "RADET_E_UNSUPPORTED"}


class RadetError(RuntimeError):
    pass


class Grid(Structure):
    _fields_ = [("num_levels", c_int32), ("level_h", c_int32 * MAX_LEVELS), ("level_w", c_int32 * MAX_LEVELS),
                ("stride", c_int32 * MAX_LEVELS), ("range_lo", c_float * MAX_LEVELS), ("range_hi", c_float * MAX_LEVELS),
                ("anchor_scale", c_float), ("tblr_normalizer", c_float)]


class Maps(Structure):
    _fields_ = [("cls", c_void_p * MAX_LEVELS), ("bbox", c_void_p * MAX_LEVELS), ("iou", c_void_p * MAX_LEVELS)]


class LossCfg(Structure):
    _fields_ = [("gamma", c_float), ("alpha", c_float), ("w_cls", c_float), ("w_bbox", c_float), ("w_iou", c_float),
                ("eps", c_float), ("avg_extra", c_float), ("weight_sums", c_void_p)]


class DetectCfg(Structure):
    _fields_ = [("score_thr", c_float), ("nms_pre", c_int32), ("max_per_img", c_int32), ("nms_mode", c_int32),
                ("iou_threshold", c_float), ("cluster_score_mode", c_int32), ("vote_score_mode", c_int32),
                ("iou_enable", c_int32), ("sigma", c_float), ("rescale", c_int32)]


_SIGS = {
    "radet_version": (c_char_p, []),
    "radet_launch_count": (c_uint64, []),
    "radet_num_points": (c_int64, [POINTER(Grid)]),
    "radet_stream_gate": (c_int32, [c_void_p, c_int64, c_void_p]),
    "radet_pack_masks": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "radet_mt19937_uniforms": (c_int32, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "radet_mt19937_seed": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p]),
    "radet_assign_workspace_bytes": (c_size_t, [POINTER(Grid), c_int32]),
    "radet_assign": (c_int32, [POINTER(Grid), c_int32, c_void_p, POINTER(c_int32), c_void_p, c_void_p, c_int32, c_int32, c_int32,
                               c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p]),
    "radet_grid_priors": (c_int32, [POINTER(Grid), c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "radet_get_targets": (c_int32, [POINTER(Grid), c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "radet_loss_workspace_bytes": (c_size_t, [POINTER(Grid), c_int32, c_int32]),
    "radet_loss_fwd_bwd": (c_int32, [POINTER(Grid), c_int32, c_int32, POINTER(Maps), c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, POINTER(LossCfg), c_void_p, POINTER(Maps), c_void_p, c_int32, c_void_p, c_size_t, c_void_p]),
    "radet_scale_grads": (c_int32, [POINTER(Grid), c_int32, c_int32, POINTER(Maps), c_void_p, c_void_p]),
    "radet_tblr_encode": (c_int32, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p]),
    "radet_tblr_decode": (c_int32, [c_void_p, c_void_p, c_int64, c_float, c_int32, c_float, c_float, c_void_p, c_void_p]),
    "radet_sigmoid_focal_loss": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "radet_giou_loss": (c_int32, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "radet_bce_with_logits": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "radet_vote_nms_workspace_bytes": (c_size_t, [c_int32, c_int64, c_int64]),
    "radet_vote_nms": (c_int32, [c_int32, POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int32, c_float,
                                 c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "radet_get_bboxes_workspace_bytes": (c_size_t, [POINTER(Grid), c_int32, c_int32, POINTER(DetectCfg)]),
    "radet_bbox2result": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "radet_candidates_capacity": (c_int64, [POINTER(Grid), c_int32, c_int32]),
    "radet_get_candidates": (c_int32, [POINTER(Grid), c_int32, c_int32, POINTER(Maps), c_void_p, c_void_p, POINTER(DetectCfg),
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "radet_gn_relu_forward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "radet_gn_relu_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                         c_void_p, c_void_p, c_void_p, c_void_p]),
    "radet_scale_relu_partials": (c_int32, [c_int64]),
    "radet_scale_relu_forward": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "radet_scale_relu_backward": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "radet_get_bboxes": (c_int32, [POINTER(Grid), c_int32, c_int32, POINTER(Maps), c_void_p, c_void_p, POINTER(DetectCfg),
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def load():
    """Load the shared library (raises RadetError with build instructions when it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RadetError(f"{LIB_PATH} not found: build it with `python -m radet_b200.build` "
                         "(there is no CPU or PyTorch fallback for this path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise RadetError(f"{what}: {_ERR.get(rc, rc)}")
    raise RadetError(f"{what}: cudaError {rc}")


def make_grid(level_shapes, strides, regress_ranges, anchor_scale=8.0, tblr_normalizer=0.125):
    g = Grid()
    n = len(level_shapes)
    if n > MAX_LEVELS or n != len(strides) or n != len(regress_ranges):
        raise RadetError("grid: inconsistent level description")
    g.num_levels = n
    for i, ((h, w), s, (lo, hi)) in enumerate(zip(level_shapes, strides, regress_ranges)):
        g.level_h[i], g.level_w[i], g.stride[i] = int(h), int(w), int(s)
        g.range_lo[i], g.range_hi[i] = float(lo), float(hi)
    g.anchor_scale = float(anchor_scale)
    g.tblr_normalizer = float(tblr_normalizer)
    return g


def make_maps(cls, bbox, iou):
    m = Maps()
    for i, (c, b, o) in enumerate(zip(cls, bbox, iou)):
        m.cls[i], m.bbox[i], m.iou[i] = c, b, o
    return m


def launch_count():
    return int(load().radet_launch_count())
