"""Minimal mmcv-style registries so ``stereoscene.py`` builds by type name.

The reference builds its model from string type names through the mmdet / mmdet3d registries
(tools/test.py:130-210 -> build_model -> DETECTORS['BEVDepthOccupancy'] -> builder.build_neck /
build_backbone, detectors/bevdepth.py:16-34).  mmcv, mmdet and mmdet3d are not installed in this
image, so the same surface is provided here: same registry names, ``register_module`` decorator,
``build(cfg)`` with the mmcv semantics (``type`` popped, remaining keys are ctor kwargs,
``default_args`` filled in when absent).  If the real mmdet3d is importable, ``register_into_mmdet3d``
adds our classes to ITS registries instead (force=True), which is what a maintainer switching an
existing checkout over would use.
"""
from __future__ import annotations

import inspect


class Registry:
    def __init__(self, name: str):
        self.name = name
        self._module_dict = {}

    @property
    def module_dict(self):
        return self._module_dict

    def __contains__(self, key):
        return key in self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._module_dict and not force and self._module_dict[key] is not cls:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._module_dict[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, default_args=None):
        if cfg is None:
            return None
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"{self.name}: cfg must be a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        t = args.pop("type")
        cls = t if inspect.isclass(t) else self._module_dict.get(t)
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry")
        return cls(**args)


BACKBONES = Registry("backbone")
NECKS = Registry("neck")
HEADS = Registry("head")
DETECTORS = Registry("detector")


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_neck(cfg):
    return NECKS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return DETECTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_model(cfg, train_cfg=None, test_cfg=None):
    """mmdet3d.models.build_model for the detector on this path."""
    return build_detector(cfg, train_cfg=train_cfg, test_cfg=test_cfg)


def register_into_mmdet3d():
    """Register the B200 modules into a real mmdet3d / mmdet installation (if present), replacing
    the reference's Python modules of the same names."""
    from mmdet.models import DETECTORS as D, HEADS as H, NECKS as N1          # type: ignore
    from mmdet3d.models.builder import BACKBONES as B3, NECKS as N3           # type: ignore
    for name, cls in NECKS.module_dict.items():
        N3.register_module(name=name, force=True, module=cls)
        N1.register_module(name=name, force=True, module=cls)
    for name, cls in BACKBONES.module_dict.items():
        B3.register_module(name=name, force=True, module=cls)
    for name, cls in HEADS.module_dict.items():
        H.register_module(name=name, force=True, module=cls)
    for name, cls in DETECTORS.module_dict.items():
        D.register_module(name=name, force=True, module=cls)
