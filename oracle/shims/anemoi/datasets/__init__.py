def open_dataset(*a, **kw):
    raise RuntimeError("anemoi.datasets is not available in this image (shim)")
