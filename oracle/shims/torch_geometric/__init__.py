"""Shim: see oracle/shims/README.md."""
