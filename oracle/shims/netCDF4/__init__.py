"""Empty shim: the real package is not in this image."""
