"""Recipe handling: ``DotDict`` and ``instantiate``.

The reference gets these from ``anemoi.utils.config.DotDict`` and ``hydra.utils.instantiate``
(/root/reference/src/anemoi/graphs/create.py:35-40,79,85,138; edges/builder.py:133).  Neither
package is in this image and neither can be installed, so the two behaviours the hot path
relies on are restated here:

* ``DotDict`` - a ``dict`` whose keys are also attributes, recursively, loadable from YAML;
* ``instantiate(cfg, **kwargs)`` - import ``cfg["_target_"]`` and call it with the remaining
  keys (nested ``_target_`` dictionaries are instantiated first) merged with ``kwargs``.

``_target_`` strings written for the reference (``anemoi.graphs.edges.KNNEdges`` ...) are
accepted unchanged: when ``anemoi.graphs`` itself is not importable the prefix is mapped to
this package, which mirrors the reference's module layout for the hot path.
"""

from __future__ import annotations

import importlib
from pathlib import Path
from typing import Any

REFERENCE_PREFIX = "anemoi.graphs."
LOCAL_PREFIX = "anemoi_graphs_b200."


class DotDict(dict):
    """Dictionary with attribute access, applied recursively to nested dicts and lists."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        for k, v in list(self.items()):
            super().__setitem__(k, self._wrap(v))

    @classmethod
    def _wrap(cls, v: Any) -> Any:
        if isinstance(v, dict) and not isinstance(v, DotDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(i) for i in v)
        return v

    @classmethod
    def from_file(cls, path: str | Path) -> "DotDict":
        path = Path(path)
        suffix = path.suffix.lower()
        if suffix in (".yaml", ".yml"):
            import yaml

            with open(path) as f:
                return cls(yaml.safe_load(f) or {})
        if suffix == ".json":
            import json

            with open(path) as f:
                return cls(json.load(f))
        raise ValueError(f"Unknown file extension {suffix!r} for recipe {path}")

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __setitem__(self, key, value) -> None:
        super().__setitem__(key, self._wrap(value))

    def __delattr__(self, name: str) -> None:
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name) from None


def resolve_target(target: str):
    """Import the object named by a ``_target_`` string."""
    candidates = [target]
    if target.startswith(REFERENCE_PREFIX):
        # Prefer this package for the hot-path classes; it is the drop-in for that path.
        candidates.insert(0, LOCAL_PREFIX + target[len(REFERENCE_PREFIX):])
    last_err: Exception | None = None
    for cand in candidates:
        module_name, _, attr = cand.rpartition(".")
        try:
            module = importlib.import_module(module_name)
            return getattr(module, attr)
        except (ImportError, AttributeError) as err:
            last_err = err
    raise ImportError(f"Cannot resolve _target_ {target!r}: {last_err}")


def instantiate(config: Any, *args, **kwargs) -> Any:
    """Minimal ``hydra.utils.instantiate``: honour ``_target_`` recursively and merge kwargs."""
    if config is None:
        return None
    if not isinstance(config, dict) or "_target_" not in config:
        raise ValueError(f"instantiate() needs a mapping with a _target_ key, got {config!r}")
    params = {}
    for k, v in config.items():
        if k in ("_target_", "_convert_", "_recursive_", "_partial_"):
            continue
        params[k] = _instantiate_nested(v)
    params.update(kwargs)
    return resolve_target(config["_target_"])(*args, **params)


def _instantiate_nested(v: Any) -> Any:
    if isinstance(v, dict):
        if "_target_" in v:
            return instantiate(v)
        return {k: _instantiate_nested(i) for k, i in v.items()}
    if isinstance(v, (list, tuple)):
        return [_instantiate_nested(i) for i in v]
    return v
