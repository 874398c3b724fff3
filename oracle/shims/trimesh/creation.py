from . import Trimesh, _impl


def icosphere(subdivisions=3, radius=1.0, **kwargs):
    v, f = _impl.icosphere(subdivisions, radius)
    return Trimesh(v, f)
