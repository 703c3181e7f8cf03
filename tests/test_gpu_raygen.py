"""GPU tests of the analytic in-kernel ray generation for pinhole detectors.

The kernels form every direction of a ConeRectGeom / ConeCircGeom from the observation's frame and
the float32 pixel-coordinate tables with the arithmetic of the reference's detector classes
(geometry.py:494-508, 573-582).  The exported directions must equal ``geom.rays`` BIT FOR BIT -- a
direction that differs in the last bit is a different ray -- for the detectors of configs 1, 2 / 4
and 5, and an operator built from the generator must give exactly the results of one built from
uploaded directions.
"""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sph_raytracer_b200 as S
    return S


def _directions(S, grid, geom):
    op = S.Operator(grid, geom)
    assert op._rays._gen is not None, "pinhole detectors must take the analytic generator"
    return op, op._rays.directions(op._lib, op._stream()).cpu()


def test_generated_directions_are_bit_identical_to_geom_rays(S):
    grid = S.SphericalGrid(shape=(50, 50, 50))
    # config 1: one rectangular detector on the equator cone and the plane a = 0
    g1 = S.ConeRectGeom((256, 256), (5, 0, 0), fov=(45, 45))
    _, d = _directions(S, grid, g1)
    assert torch.equal(d, g1.rays)
    # configs 2 / 4: polar detectors on an orbit (the reference's float32 radius x cos / sin products)
    geoms = [S.ConeCircGeom((100, 50), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(0, 45))
             for t in torch.linspace(0, 2 * torch.pi, 7)]
    coll = S.ViewGeomCollection(*geoms)
    _, d = _directions(S, grid, coll)
    assert tuple(d.shape) == (7, 100, 50, 3) and torch.equal(d, coll.rays)
    # config 5: 512 x 512 rectangular detectors at seeded positions
    import bench
    grid5, geoms5, _ = bench.workload("c5", 3)
    coll5 = S.ViewGeomCollection(*geoms5)
    _, d = _directions(S, grid5, coll5)
    assert torch.equal(d, coll5.rays)
    # a one-pixel axis, log-spaced radii, explicit look / up directions, different fields of view
    odd = [S.ConeRectGeom((1, 9), (3.0, -2.0, 1.5), lookdir=(-1.0, 0.5, -0.2), updir=(0.1, 0.2, 1.0), fov=(10, 25)),
           S.ConeRectGeom((1, 9), (-3.0, 2.0, 0.5), fov=(12, 20))]
    _, d = _directions(S, grid, S.ViewGeomCollection(*odd))
    assert torch.equal(d, S.ViewGeomCollection(*odd).rays)
    circ = S.ConeCircGeom((12, 9), (0.0, 6.0, 2.0), fov=(5, 40), spacing="log")
    _, d = _directions(S, grid, circ)
    assert torch.equal(d, circ.rays)


def test_operator_from_the_generator_equals_operator_from_uploaded_directions(S):
    grid = S.SphericalGrid(shape=(40, 36, 48), size_r=(0.5, 3))
    geoms = [S.ConeCircGeom((40, 30), (7 * math.cos(t), 7 * math.sin(t), 1.0), fov=(0, 40)) for t in (0.3, 2.2, 4.0)]
    coll = S.ViewGeomCollection(*geoms)
    x = torch.rand(tuple(grid.shape), device="cuda", dtype=torch.float64)
    w = torch.rand(3, 40, 30, device="cuda", dtype=torch.float64)
    a = S.Operator(grid, coll)
    os.environ["SPHRT_RAYGEN"] = "0"
    try:
        b = S.Operator(grid, coll)
    finally:
        os.environ.pop("SPHRT_RAYGEN")
    assert a._rays._gen is not None and b._rays._gen is None and b._rays.dirs is not None
    assert torch.equal(a._apply_forward(x), b._apply_forward(x))
    ga, gb = a._apply_adjoint(w, x.shape), b._apply_adjoint(w, x.shape)
    assert torch.allclose(ga, gb, rtol=1e-12, atol=1e-13)          # atomics order only
    ca, ta = a.segment_counts(); cb, tb = b.segment_counts()
    assert ta == tb and torch.equal(ca, cb)
    # the generator keeps 96 bytes per observation instead of 24 bytes per ray
    assert a._rays.dirs is None
