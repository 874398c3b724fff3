"""Shim for trimesh: see oracle/trimesh_icosphere.py for the restated algorithm."""
import importlib.util
import pathlib

import numpy as np

_spec = importlib.util.spec_from_file_location(
    "_oracle_trimesh_icosphere", pathlib.Path(__file__).resolve().parents[2] / "trimesh_icosphere.py"
)
_impl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_impl)


class Trimesh:
    def __init__(self, vertices, faces):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces, dtype=np.int64)

    @property
    def edges_unique(self):
        return _impl.edges_unique(self.faces)


from . import creation  # noqa: E402,F401
