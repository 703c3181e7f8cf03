"""Parametric volume models for retrievals (mirror of reference ``sph_raytracer/model.py``).

A model maps a coefficient tensor to a density volume on a ``SphericalGrid``; ``gd`` optimises
the coefficients.  Pure host-side tensor code, no kernels.
"""
import torch

__all__ = ["Model", "FullyDenseModel", "CubesModel", "AxisAlignmentModel"]


class Model:
    """Base class: subclasses provide ``__call__(coeffs) -> density`` and ``coeffs_shape``."""

    def __init__(self, grid):
        raise NotImplementedError

    def __call__(self, coeffs):
        raise NotImplementedError

    @property
    def coeffs_shape(self):
        raise NotImplementedError

    def __repr__(self):
        return f"{type(self).__name__}({tuple(self.grid.shape)})"


class FullyDenseModel(Model):
    """One coefficient per voxel: the coefficients are the density (reference model.py:41-52)."""

    def __init__(self, grid):
        self.grid = grid

    def __call__(self, coeffs):
        return coeffs

    @property
    def coeffs_shape(self):
        return self.grid.shape


class _FixedVolume(Model):
    """Models without free parameters: they always return ``self.volume``."""

    def __call__(self, coeffs):
        return self.volume

    @property
    def coeffs_shape(self):
        return ()


class CubesModel(_FixedVolume):
    """Two boxes in spherical coordinates (reference model.py:55-83)."""

    def __init__(self, grid):
        self.grid = grid
        nr, ne, na = grid.shape.r, grid.shape.e, grid.shape.a
        # the reference multiplies the axis length by a float32 tensor and truncates
        # (model.py:58-66); Python float64 products truncate differently for some sizes
        # (ne = 180: .7 -> 126 in float32, 125 in float64)
        edges = lambda n, fracs: tuple(int(v) for v in n * torch.tensor(fracs))
        self.r0, self.r1 = edges(nr, (.333, .666))
        self.e00, self.e01 = edges(ne, (.2, .3))
        self.e10, self.e11 = edges(ne, (.7, .9))
        self.a0, self.a1 = edges(na, (.4, .6))
        self.volume = torch.zeros(grid.shape)
        self.volume[self.r0:self.r1, self.e00:self.e01, self.a0:self.a1] = 1
        self.volume[self.r0:self.r1, self.e10:self.e11, self.a0:self.a1] = 1


class AxisAlignmentModel(_FixedVolume):
    """Marks the X, Y and Z axes to check that projections are not mirrored
    (reference model.py:86-113)."""

    def __init__(self, grid):
        self.grid = grid
        nr, ne, na = grid.shape.r, grid.shape.e, grid.shape.a
        self.volume = torch.zeros(grid.shape)
        self.volume[:nr // 3, ne // 2, 0] = 1                 # X axis
        self.volume[:nr // 2, ne // 2, (na * 3) // 4] = 1     # Y axis
        self.volume[:, 0, :] = 1                              # Z axis
