"""sph_raytracer_b200 -- forward / adjoint line integration through spherical voxel grids
on NVIDIA B200 (sm_100a).

Public surface mirrors the reference package (``sph_raytracer/__init__.py:1-2``):
``Operator`` plus the geometry classes.
"""
from .geometry import *          # noqa: F401,F403
from .geometry import __all__ as _geometry_all
from .raytracer import Operator, find_starts   # noqa: F401
from . import model, loss, retrieval, parallel   # noqa: F401  (same sub-modules as the reference)

__all__ = ["Operator", "find_starts", *_geometry_all]
