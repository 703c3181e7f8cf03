"""Host-side geometry definitions consumed by the line-integration operator.

These classes are the *inputs* of the hot path: a
spherical voxel grid described by its shell / cone / half-plane boundaries, and
detector descriptions that expand to one (origin, direction) pair per pixel.
They mirror the public surface of the reference's ``sph_raytracer/geometry.py``
(class names, constructor arguments, attribute names, shapes, dtypes) so code
written against the reference keeps working, but they are written fresh and
carry no plotting code (visualisation is out of scope).

Everything here is tiny host work in float64.  The arithmetic *order* of the
pixel-direction formulas is kept identical to the reference
(``geometry.py:496-508`` rectangular, ``:573-582`` circular, ``:632-655``
parallel, note the float32 ``linspace`` pixel coordinates) because the tracing
kernels are compared bit-for-bit on voxel indices against the reference and a
one-ulp different direction is a different ray.
"""

from collections import namedtuple
import math

import torch

__all__ = [
    "SphericalGrid", "ViewGeom", "ViewGeomCollection",
    "ConeRectGeom", "ConeCircGeom", "ParallelGeom",
]

F64 = torch.float64

# same field names as the reference's namedtuples (geometry.py:20-23)
_Shape3 = namedtuple("Shape", ["r", "e", "a"])
_Shape4 = namedtuple("Shape", ["t", "r", "e", "a"])
_Size3 = namedtuple("Size", ["r", "e", "a"])
_Size4 = namedtuple("Size", ["t", "r", "e", "a"])


def _f64(x):
    return torch.asarray(x, dtype=F64)


def _midpoints(b):
    return (b[1:] + b[:-1]) / 2


def _extent(v):
    return float(min(v)), float(max(v))


class SphericalGrid:
    """Voxel grid in spherical coordinates (radius, elevation from +Z, azimuth from +X).

    Either give ``shape`` (3 entries static, 4 entries dynamic with a leading
    time axis) together with the ``size_*`` extents, or give the boundary
    vectors ``r_b, e_b, a_b`` (and ``t`` for a dynamic grid) explicitly.
    Mirrors reference ``geometry.py:107-183``.

    Attributes used by the operator: ``r_b, e_b, a_b`` (float64 boundary
    vectors, one longer than the voxel count), ``shape`` (namedtuple) and
    ``dynamic``.  ``r, e, a, t`` are the bin centres / time samples.
    """

    def __init__(self, shape=(50, 50, 50),
                 size_t=(0, 1), size_r=(0, 1), size_e=(0, math.pi),
                 size_a=(-math.pi, math.pi), spacing="lin",
                 t=None, r_b=None, e_b=None, a_b=None, timeunit="s",
                 rs_b=None, phis_b=None, thetas_b=None):
        ndim = len(shape)
        if ndim not in (3, 4):
            raise ValueError("shape must be 3D or 4D")
        self.dynamic = ndim == 4

        # legacy keyword spelling accepted by the reference (geometry.py:129-130)
        if rs_b is not None and phis_b is not None and thetas_b is not None:
            r_b, e_b, a_b = rs_b, phis_b, thetas_b

        explicit = r_b is not None and e_b is not None and a_b is not None
        if explicit:
            size_r, size_e, size_a = _extent(r_b), _extent(e_b), _extent(a_b)
            nvox = (len(r_b) - 1, len(e_b) - 1, len(a_b) - 1)
            if t is None:
                shape, size = _Shape3(*nvox), _Size3(size_r, size_e, size_a)
                self.dynamic = False
            else:
                size_t = _extent(t)
                t = _f64(t)
                shape = _Shape4(len(t), *nvox)
                size = _Size4(size_t, size_r, size_e, size_a)
                self.dynamic = True
            r_b, e_b, a_b = _f64(r_b), _f64(e_b), _f64(a_b)
            r = _midpoints(r_b)
        else:
            if self.dynamic:
                shape = _Shape4(*shape)
                size = _Size4(size_t, size_r, size_e, size_a)
                t = torch.linspace(size_t[0], size_t[1], shape.t, dtype=F64)
            else:
                shape = _Shape3(*shape)
                size = _Size3(size_r, size_e, size_a)
            if spacing == "lin":
                r_b = torch.linspace(size_r[0], size_r[1], shape.r + 1, dtype=F64)
                r = _midpoints(r_b)
            elif spacing == "log":
                r_b = torch.logspace(math.log10(size_r[0]), math.log10(size_r[1]),
                                     shape.r + 1, dtype=F64)
                r = torch.sqrt(r_b[1:] * r_b[:-1])
            else:
                raise ValueError("Invalid value for spacing")
            e_b = torch.linspace(size_e[0], size_e[1], shape.e + 1, dtype=F64)
            a_b = torch.linspace(size_a[0], size_a[1], shape.a + 1, dtype=F64)

        self.shape, self.size, self.spacing = shape, size, spacing
        self.r_b, self.e_b, self.a_b = r_b, e_b, a_b
        self.t, self.r = t, r
        self.e, self.a = _midpoints(e_b), _midpoints(a_b)
        self.timeunit = timeunit
        # deprecated aliases kept by the reference (geometry.py:182-183)
        self.rs_b, self.phis_b, self.thetas_b = r_b, e_b, a_b
        self.rs, self.phis, self.thetas = self.r, self.e, self.a

    def __repr__(self):
        lines = [f"{type(self).__name__}(", f"    shape={tuple(self.shape)},"]
        for name, (lo, hi) in self.size._asdict().items():
            lines.append(f"    size_{name}=({lo:.2f}, {hi:.2f}),")
        lines.append(")")
        return "\n".join(lines)

    @property
    def coords(self):
        keys = ("t", "r", "e", "a") if self.dynamic else ("r", "e", "a")
        return {k: getattr(self, k) for k in keys}

    @property
    def mesh(self):
        """Dense coordinate mesh, (N_r,N_e,N_a,3) or (N_t,N_r,N_e,N_a,4)."""
        return torch.stack(
            torch.meshgrid(list(self.coords.values()), indexing="ij"), dim=-1)

    @property
    def nptime(self):
        return self.t.numpy().astype(f"datetime64[{self.timeunit}]")


def _normalise_rows_(v):
    """In-place ``v /= ||v||`` over the last axis (same op the reference applies)."""
    v /= torch.linalg.norm(v, axis=-1)[..., None]
    return v


class ViewGeom:
    """Arbitrary set of lines of sight: ``ray_starts (..., 3)`` and ``rays (..., 3)``.

    ``shape`` (everything but the trailing 3) is the shape of the images the
    operator returns.  Mirrors reference ``geometry.py:284-288``.
    """

    def __init__(self, ray_starts, rays):
        self.ray_starts = _f64(ray_starts)
        self.rays = _normalise_rows_(_f64(rays))
        self.shape = self.rays.shape[:-1]

    def __add__(self, other):
        if other == 0 or other is None:
            return ViewGeomCollection(self)
        if isinstance(other, ViewGeomCollection):
            other.geoms.append(self)
            return other
        return ViewGeomCollection(self, other)

    __radd__ = __add__

    def __repr__(self):
        return f"{type(self).__name__}(\n    shape={tuple(self.shape)}\n)"


class ViewGeomCollection(ViewGeom):
    """Ordered set of same-shaped view geometries (one per observation).

    ``rays`` / ``ray_starts`` stack the members on a new leading axis
    (reference ``geometry.py:385-395``).
    """

    def __init__(self, *geoms):
        first = geoms[0].shape
        if any(g.shape != first for g in geoms):
            raise ValueError("ViewGeoms must all have same shape")
        if len(geoms) == 1 and hasattr(geoms[0], "geoms"):
            self.geoms = geoms[0].geoms
        else:
            self.geoms = list(geoms)

    def __add__(self, other):
        if isinstance(other, ViewGeomCollection):
            self.geoms += other.geoms
            other.geoms += self.geoms
        else:
            self.geoms.append(other)
        return self

    __radd__ = __add__

    def __getitem__(self, i):
        return self.geoms[i]

    def __len__(self):
        return len(self.geoms)

    @property
    def shape(self):
        return (len(self.geoms), *self.geoms[0].shape)

    def _stacked(self, name):
        return torch.concat([getattr(g, name)[None, ...] for g in self.geoms])

    @property
    def rays(self):
        return self._stacked("rays")

    @property
    def ray_starts(self):
        return self._stacked("ray_starts")

    @property
    def pos(self):
        if all(hasattr(g, "pos") for g in self.geoms):
            return self._stacked("pos")
        return None


def _detector_frame(pos, lookdir, updir):
    """Shared constructor logic of the cone / parallel detectors
    (reference ``geometry.py:474-485`` and ``:619-629``)."""
    pos = _f64(pos)
    lookdir = -pos if lookdir is None else _f64(lookdir)
    if updir is None:
        updir = torch.cross(lookdir, _f64((0, 0, 1)), dim=-1)
    else:
        updir = _f64(updir)
    lookdir /= torch.linalg.norm(lookdir, axis=-1)
    updir /= torch.linalg.norm(updir, axis=-1)
    return pos, lookdir, updir


class ConeRectGeom(ViewGeom):
    """Rectangular pinhole detector: ``shape=(npix_x, npix_y)``, ``fov`` in degrees.

    Pixel (0, 0) is the top-left corner (image convention of the reference).
    """

    def __init__(self, shape, pos, lookdir=None, updir=None, fov=(45, 45)):
        self.pos, self.lookdir, self.updir = _detector_frame(pos, lookdir, updir)
        self.shape = shape
        self.fov = _f64(fov)

    def _half_extent(self, axis):
        # a 1-pixel axis degenerates to the optical axis (geometry.py:500-501)
        if self.shape[axis] > 1:
            return torch.tan(torch.deg2rad(self.fov[axis] / 2))
        return 0

    @property
    def rays(self):
        """Unit line-of-sight vectors, ``(*shape, 3)`` float64."""
        u = torch.cross(self.lookdir, self.updir, dim=-1)
        v = self.updir
        ulim, vlim = self._half_extent(0), self._half_extent(1)
        # pixel coordinates are float32 linspaces promoted on multiply
        cu = torch.linspace(-ulim, ulim, self.shape[0])[:, None, None]
        cv = torch.linspace(-vlim, vlim, self.shape[1])[None, :, None]
        d = (self.lookdir[None, None, :] + u[None, None, :] * cu + v[None, None, :] * cv)
        return _normalise_rows_(d.reshape((*self.shape, 3)))

    @property
    def ray_starts(self):
        """All pixels share the pinhole position; broadcastable ``(1, 1, 3)``."""
        return self.pos[None, None, :]

    def __repr__(self):
        return (f"{type(self).__name__}(\n    shape={self.shape}\n"
                f"    pos={self.pos.tolist()},\n    lookdir={self.lookdir.tolist()},\n"
                f"    fov={self.fov.tolist()}\n)")


class ConeCircGeom(ConeRectGeom):
    """Polar-sampled pinhole detector: ``shape=(npix_r, npix_theta)``,
    ``fov=(inner, outer)`` full cone angles in degrees."""

    def __init__(self, *args, fov=(0, 45), spacing="lin", **kwargs):
        super().__init__(*args, fov=fov, **kwargs)
        lims = [torch.tan(torch.deg2rad(self.fov[i] / 2)) for i in (0, 1)]
        if spacing == "lin":
            self.r = torch.linspace(*lims, self.shape[0])
        elif spacing == "log":
            self.r = torch.logspace(*lims, self.shape[0])
        else:
            raise ValueError(f"Invalid spacing {spacing}")
        self.theta = torch.linspace(0, 2 * torch.pi, self.shape[1]) + torch.pi / 2

    @property
    def rays(self):
        u = torch.cross(self.lookdir, self.updir, dim=-1)
        v = self.updir
        rad = self.r[:, None, None]
        ang = self.theta[None, :, None]
        d = (self.lookdir[None, None, :]
             + rad * torch.cos(ang) * u[None, None, :]
             + rad * torch.sin(ang) * v[None, None, :])
        return _normalise_rows_(d)


class ParallelGeom(ViewGeom):
    """Rectangular parallel-beam detector of physical ``size=(width, height)``.

    All pixels share the direction ``lookdir``; each has its own start point.
    (The reference cannot trace this class, SURVEY.md §8a quirk 5 — the
    operator here expands the shared direction instead of failing.)
    """

    def __init__(self, shape, pos, lookdir=None, updir=None, size=(1, 1)):
        self.pos, self.lookdir, self.updir = _detector_frame(pos, lookdir, updir)
        u = torch.cross(self.lookdir, self.updir, dim=-1)
        v = self.updir
        ulim = size[0] / 2 if shape[0] > 1 else 0
        vlim = size[1] / 2 if shape[1] > 1 else 0
        self._u_arr = u[None, None, :] * torch.linspace(ulim, -ulim, shape[0])[:, None, None]
        self._v_arr = v[None, None, :] * torch.linspace(-vlim, vlim, shape[1])[None, :, None]
        self.shape = shape
        self.size = size

    @property
    def rays(self):
        return self.lookdir[None, None, :]

    @property
    def ray_starts(self):
        return (self.pos[None, None, :] + self._u_arr + self._v_arr).reshape((*self.shape, 3))

    def __repr__(self):
        return (f"ParallelGeom(\n    shape={self.shape}\n    pos={self.pos.tolist()},\n"
                f"    lookdir={self.lookdir.tolist()},\n)")
