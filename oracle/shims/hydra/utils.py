import importlib


def _nested(v):
    if isinstance(v, dict):
        return instantiate(v) if "_target_" in v else {k: _nested(i) for k, i in v.items()}
    if isinstance(v, (list, tuple)):
        return [_nested(i) for i in v]
    return v


def instantiate(config, *args, **kwargs):
    module_name, _, attr = config["_target_"].rpartition(".")
    cls = getattr(importlib.import_module(module_name), attr)
    params = {k: _nested(v) for k, v in config.items() if not (k.startswith("_") and k.endswith("_"))}
    params.update(kwargs)
    return cls(*args, **params)
