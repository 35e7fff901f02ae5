"""radet_b200 — B200-native (sm_100a) dense-head hot path of RADet: visibility-guided sample assignment,
fused head loss (forward+backward), decode + class-aware vote-NMS.  CUDA kernels behind a C ABI
(include/radet_b200.h, radet_b200/csrc), Python host side mirroring the reference's plugin API (radet_b200.plugin)."""
__version__ = "0.1.0"
