"""Minimal OmegaConf stand-in for the runner config API (omegaconf is not installed in this image).

Covers exactly what run_group_ddim_inversion.py / run_group_pnp_edit.py use: ``load`` a YAML template,
``create`` from a dict, ``merge`` (later wins, recursive), attribute + item access, ``${a.b}`` interpolation resolved
against the ROOT at access time (template.yaml:13,22 rely on the JSON entry overriding ``video_name`` before
``output_dir`` is read), assignment of new keys, ``to_yaml``, and ``items()``.
If the real omegaconf is importable it is NOT used — behaviour stays identical on every box.
"""
from __future__ import annotations

import copy
import re

import yaml

_INTERP = re.compile(r"\$\{([^}]+)\}")


class Config:
    def __init__(self, data=None, root=None):
        object.__setattr__(self, "_data", {} if data is None else data)
        object.__setattr__(self, "_root", root if root is not None else self)

    # -- construction -------------------------------------------------------------------------------------------
    @staticmethod
    def _wrap(value, root):
        if isinstance(value, dict):
            return Config(value, root)
        return value

    def _resolve(self, value, depth=0):
        if isinstance(value, str) and "${" in value:
            if depth > 20:
                raise ValueError(f"interpolation too deep / cyclic: {value}")
            whole = _INTERP.fullmatch(value)
            if whole:  # "${image_size}" keeps the referenced node's type (a list)
                return self._resolve(self._root._lookup(whole.group(1)), depth + 1)
            return _INTERP.sub(lambda m: str(self._resolve(self._root._lookup(m.group(1)), depth + 1)), value)
        if isinstance(value, list):
            return [self._resolve(v, depth) for v in value]
        return value

    def _lookup(self, dotted: str):
        node = self._data
        for part in dotted.strip().split("."):
            if not isinstance(node, dict) or part not in node:
                raise KeyError(f"interpolation key '{dotted}' not found")
            node = node[part]
        return node

    # -- access -------------------------------------------------------------------------------------------------
    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"Missing key {name}") from None

    def __getitem__(self, name):
        value = self._data[name]
        if isinstance(value, dict):
            return Config(value, self._root)
        return self._resolve(value)

    def __setattr__(self, name, value):
        self._data[name] = value._data if isinstance(value, Config) else value

    __setitem__ = __setattr__

    def __contains__(self, name):
        return name in self._data

    def get(self, name, default=None):
        return self[name] if name in self._data else default

    def keys(self):
        return self._data.keys()

    def items(self):
        return [(k, self[k]) for k in self._data]

    def to_container(self, resolve=True):
        def conv(node):
            if isinstance(node, dict):
                return {k: conv(v) for k, v in node.items()}
            return self._resolve(node) if resolve else node
        return conv(self._data)

    def __repr__(self):
        return f"Config({self._data!r})"


class OmegaConf:
    """The subset of the omegaconf.OmegaConf static API the runners call."""

    @staticmethod
    def load(path) -> Config:
        with open(path, "r") as fh:
            return Config(yaml.safe_load(fh) or {})

    @staticmethod
    def create(obj=None) -> Config:
        if isinstance(obj, Config):
            return Config(copy.deepcopy(obj._data))
        return Config(copy.deepcopy(obj) if obj is not None else {})

    @staticmethod
    def merge(*configs) -> Config:
        def rec(dst, src):
            for k, v in src.items():
                if isinstance(v, dict) and isinstance(dst.get(k), dict):
                    rec(dst[k], v)
                else:
                    dst[k] = copy.deepcopy(v)
        out: dict = {}
        for c in configs:
            rec(out, c._data if isinstance(c, Config) else c)
        return Config(out)

    @staticmethod
    def to_yaml(cfg: Config, resolve: bool = False) -> str:
        return yaml.safe_dump(cfg.to_container(resolve=resolve), sort_keys=False)
