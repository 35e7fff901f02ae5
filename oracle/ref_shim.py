"""TEST INFRASTRUCTURE — container-only import shim for the UNMODIFIED reference.

Imports ``/root/reference/radet`` (read-only, never copied) so that its own
``LabelAssignment``, ``RADetHead`` and ``radet.ops`` can be executed here to
(1) validate the restatement in ``oracle/radet_oracle.py`` and (2) generate the
golden vectors committed under ``tests/golden/`` (see ``tests/golden/make_golden.py``).

``/root/reference`` does not exist on the GPU box, so nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` imports this module.

The reference depends on ``mmcv==1.3.18`` (requirements.txt:8), which is neither
vendored nor installed.  The shim fabricates auto-mocking modules for ``mmcv``,
``pycocotools``, ``terminaltables`` and ``matplotlib`` and supplies real
stand-ins only for the handful of symbols the hot path executes:
``Registry``/``build_from_cfg``, ``force_fp32``/``auto_fp16`` (identity),
``ConvModule``/``Scale`` and ``mmcv.ops.sigmoid_focal_loss`` (restated from the
reference's own ``py_sigmoid_focal_loss``, models/losses/focal_loss.py:10-41,
because the mmcv op is CUDA-only and absent).
"""
import importlib.abc
import importlib.machinery
import inspect
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RADET_REFERENCE_ROOT", "/root/reference")
_REF_BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_MOCK_ROOTS = ("mmcv", "pycocotools", "terminaltables", "matplotlib")


class _Mock:
    """Callable, subscriptable, attribute-bearing placeholder."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        # used as decorator factory or decorator: hand back the function itself
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Mock()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Mock()

    def __getitem__(self, k):
        return _Mock()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _MockModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Mock()
        setattr(self, name, m)
        return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _MOCK_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _MockModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


class Registry:
    """Stand-in for mmcv.utils.Registry (dict registry)."""

    def __init__(self, name, *a, **k):
        self.name = name
        self.module_dict = {}

    def get(self, key):
        return self.module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self.module_dict[name or module.__name__] = module
            return module

        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls

        return _reg


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    return cls(**args)


def _identity_decorator_factory(*a, **k):
    def deco(fn):
        return fn

    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return deco


class ConfigDict(dict):
    """dict with attribute access and .copy() preserving the type (mmcv.Config stand-in)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v

    def __setattr__(self, k, v):
        self[k] = v

    def copy(self):
        return ConfigDict(dict.copy(self))


def _build_reference_ext(name, rel_src):
    """JIT-build one reference C++ op from where it lies, output under oracle/_ref/."""
    from torch.utils.cpp_extension import load

    os.makedirs(os.path.join(_REF_BUILD, name), exist_ok=True)
    return load(
        name=name,
        sources=[os.path.join(REFERENCE_ROOT, rel_src)],
        build_directory=os.path.join(_REF_BUILD, name),
        extra_cflags=["-O2"],
        verbose=False,
    )


_installed = False


def install():
    """Make ``import radet`` resolve to the unmodified reference."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} (container-only shim)")
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    sys.dont_write_bytecode = True
    sys.meta_path.insert(0, _Finder())
    import mmcv  # the mock

    mmcv.__version__ = "1.3.18"  # version gate radet/__init__.py:19-27
    import pycocotools

    pycocotools.__version__ = "12.0.2"  # gate datasets/coco.py:21
    import mmcv.utils as mu

    mu.Registry = Registry
    mu.build_from_cfg = build_from_cfg
    mmcv.is_tuple_of = lambda seq, typ: isinstance(seq, tuple) and all(isinstance(s, typ) for s in seq)
    import mmcv.runner as mr

    mr.force_fp32 = _identity_decorator_factory
    mr.auto_fp16 = _identity_decorator_factory
    mr.Hook = type("Hook", (), {})
    mr.OptimizerHook = type("OptimizerHook", (), {})
    import mmcv.cnn as mc

    class Scale(nn.Module):
        def __init__(self, scale=1.0):
            super().__init__()
            self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

        def forward(self, x):
            return x * self.scale

    class ConvModule(nn.Module):
        def __init__(self, cin, cout, k, stride=1, padding=0, conv_cfg=None, norm_cfg=None, **kw):
            super().__init__()
            self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=norm_cfg is None)
            self.gn = nn.GroupNorm(norm_cfg["num_groups"], cout) if norm_cfg else None

        def forward(self, x):
            x = self.conv(x)
            if self.gn is not None:
                x = self.gn(x)
            return F.relu(x)

    mc.Scale = Scale
    mc.ConvModule = ConvModule
    mc.normal_init = lambda *a, **k: None
    mc.bias_init_with_prob = lambda p: 0.0

    # the three native ops, compiled from the reference's own sources
    for mod, name, src in (
        ("radet.ops.vote.vote_ext", "vote_ext", "radet/ops/vote/vote_ext.cpp"),
        ("radet.ops.cluster.cluster_ext", "cluster_ext", "radet/ops/cluster/cluster_ext.cpp"),
    ):
        sys.modules[mod] = _build_reference_ext(name, src)
    # bbox2distance is out of scope: mock it instead of building
    sys.modules["radet.ops.bbox2distance.bbox2distance_ext"] = _MockModule("bbox2distance_ext")

    sys.path.insert(0, REFERENCE_ROOT)
    import radet  # noqa: F401
    import radet.models.losses.focal_loss as fl

    def _focal_none(pred, target, gamma, alpha, weight, reduction):
        # mmcv.ops.sigmoid_focal_loss stand-in: one-hot targets (label C = all-zero row)
        C = pred.size(1)
        onehot = F.one_hot(target, C + 1)[:, :C]
        return fl.py_sigmoid_focal_loss(pred, onehot, weight=None, gamma=gamma, alpha=alpha, reduction="none")

    fl._sigmoid_focal_loss = _focal_none
    _installed = True


# kwargs of configs/bop/r50_ycbv_pbr.py:30-56 (head) and :70-80 (test_cfg)
def head_kwargs(num_classes=21):
    return dict(
        num_classes=num_classes,
        in_channels=256,
        stacked_convs=4,
        feat_channels=256,
        strides=[8, 16, 32, 64, 128],
        anchor_generator=dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                              strides=[8, 16, 32, 64, 128]),
        bbox_coder=dict(type="TBLRBBoxCoder", normalizer=1 / 8),
        loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
        loss_bbox=dict(type="GIoULoss", loss_weight=2.0),
        loss_centerness=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0),
    )


def test_cfg(score_thr=0.05, nms_type="vote", nms_pre=1000, max_per_img=100, iou_threshold=0.65):
    return ConfigDict(
        nms_pre=nms_pre, min_bbox_size=0, score_thr=score_thr,
        nms=ConfigDict(type=nms_type, iou_threshold=iou_threshold, cluster_score=["cls", "iou"],
                       vote_score=["iou", "cls"], iou_enable=False, sima=0.025),
        max_per_img=max_per_img)


def assignment_kwargs():
    # configs/base/datasets/bop_detection.py:19-32
    return dict(
        anchor_generator_cfg=dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                                  strides=[8, 16, 32, 64, 128]),
        neg_threshold=0.2, positive_num=10, adapt_positive_num=False, balance_sample=True)


def build_reference_head(num_classes=21, **tc):
    install()
    from radet.models.dense_heads.radet_head import RADetHead

    return RADetHead(train_cfg=None, test_cfg=test_cfg(**tc), **head_kwargs(num_classes))


def build_reference_assigner(**overrides):
    install()
    from radet.datasets.pipelines.label_assignment import LabelAssignment

    kw = assignment_kwargs()
    kw.update(overrides)
    return LabelAssignment(**kw)


class BitmapMasksStandIn:
    """What LabelAssignment needs from radet.core.mask.BitmapMasks (structures.py:473-475)."""

    def __init__(self, masks):
        self.masks = masks

    def to_ndarray(self):
        return self.masks
