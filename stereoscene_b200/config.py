"""Minimal ``mmcv.Config.fromfile`` for python config files with ``_base_`` inheritance.

Semantics kept from mmcv 1.4 (the version the reference pins, README.md:62): a config file is
executed as Python; ``_base_`` (str or list, paths relative to the file) are loaded first and
merged -- dicts merge recursively, every other value (lists, scalars) is replaced by the child;
``_delete_=True`` inside a child dict replaces the base dict instead of merging; keys starting
with ``__`` and modules/functions are dropped.  ``merge_from_dict`` takes dotted keys like
``--cfg-options`` (tools/test.py:131-132).  ``projects/configs/occupancy/semantickitti/
stereoscene.py`` loads unchanged through this.
"""
from __future__ import annotations

import copy
import os
import types

BASE_KEY = "_base_"
DELETE_KEY = "_delete_"


class ConfigDict(dict):
    """dict with attribute access (cfg.model.img_view_transformer.numC_Trans)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def _merge(child: dict, base: dict) -> dict:
    out = copy.deepcopy(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get(DELETE_KEY, False):
            out[k] = _merge(v, out[k])
        else:
            if isinstance(v, dict):
                v = {a: b for a, b in v.items() if a != DELETE_KEY}
            out[k] = copy.deepcopy(v)
    return out


def _exec_file(path: str) -> dict:
    with open(path, "r") as f:
        src = f.read()
    ns = {"__file__": path, "__name__": "__config__"}
    exec(compile(src, path, "exec"), ns)
    return {k: v for k, v in ns.items()
            if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType, type))}


def _load(path: str) -> dict:
    path = os.path.abspath(os.path.expanduser(path))
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    cfg = _exec_file(path)
    bases = cfg.pop(BASE_KEY, None)
    if bases is None:
        return cfg
    if isinstance(bases, str):
        bases = [bases]
    merged: dict = {}
    for b in bases:
        bcfg = _load(os.path.join(os.path.dirname(path), b))
        dup = set(merged) & set(bcfg)
        if dup:
            raise KeyError(f"duplicate keys in _base_ files of {path}: {sorted(dup)}")
        merged.update(bcfg)
    return _merge(cfg, merged)


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, "_cfg_dict", _wrap(cfg_dict or {}))
        object.__setattr__(self, "filename", filename)

    @staticmethod
    def fromfile(filename: str) -> "Config":
        return Config(_load(filename), filename=filename)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __contains__(self, name):
        return name in self._cfg_dict

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def keys(self):
        return self._cfg_dict.keys()

    def to_dict(self) -> dict:
        def unwrap(v):
            if isinstance(v, dict):
                return {k: unwrap(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return [unwrap(x) for x in v]
            return v
        return unwrap(self._cfg_dict)

    def merge_from_dict(self, options: dict):
        """Dotted-key overrides, e.g. {'model.img_view_transformer.grid_config.xbound': [0, 51.2, 0.8]}."""
        for full, val in options.items():
            d = self._cfg_dict
            keys = full.split(".")
            for k in keys[:-1]:
                if k not in d or not isinstance(d[k], dict):
                    d[k] = ConfigDict()
                d = d[k]
            d[keys[-1]] = _wrap(val)
