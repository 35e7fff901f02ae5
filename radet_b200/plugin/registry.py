"""Minimal registries mirroring the ones RADet's configs are built through.

Reference: radet/models/builder.py:6-57 (HEADS, LOSSES), radet/core/bbox/builder.py (BBOX_CODERS),
radet/core/anchor/builder.py:3-7 (ANCHOR_GENERATORS), radet/datasets/builder.py:22-23 (PIPELINES), all instances of
mmcv.utils.Registry + build_from_cfg.  `install_into_reference()` (plugin/__init__.py) registers the same classes
into the reference's own registries when RADet is importable, which is how the drop-in is wired in production.
"""


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def get(self, key):
        return self.module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self.module_dict[key] = cls
            return cls

        return _reg(module) if module is not None else _reg


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict) or "type" not in cfg:
        raise TypeError(f"cfg must be a dict with a `type` key, got {cfg!r}")
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop("type")
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f"{typ} is not in the {registry.name} registry")
    return cls(**args)


HEADS = Registry("head")
LOSSES = Registry("loss")
BBOX_CODERS = Registry("bbox_coder")
ANCHOR_GENERATORS = Registry("Anchor generator")
PIPELINES = Registry("pipeline")


def build_head(cfg):
    return build_from_cfg(cfg, HEADS)


def build_loss(cfg):
    return build_from_cfg(cfg, LOSSES)


def build_bbox_coder(cfg, **default_args):
    return build_from_cfg(cfg, BBOX_CODERS, default_args)


def build_anchor_generator(cfg, default_args=None):
    return build_from_cfg(cfg, ANCHOR_GENERATORS, default_args)


class ConfigDict(dict):
    """dict with attribute access (stand-in for mmcv.Config nodes such as test_cfg)."""

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
