"""GPU tests of the stored operator (Operator(..., store_segments=True)): the
canonical segments are traced once and kept as an 8-byte-per-segment list; applying it must
give what the on-the-fly kernels give, and what the reference gives (golden fixtures)."""
import math
import types

import numpy as np
import pytest
import torch

from helpers import load_case, case_density, case_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sph_raytracer_b200 as S
    return S


def _case_operator(S, name, **kw):
    c = load_case(name)
    dens = case_density(c)
    dyn = name == "c4_dynamic"
    g = types.SimpleNamespace(r_b=c.grid.r_b, e_b=c.grid.e_b, a_b=c.grid.a_b, dynamic=dyn)
    g.shape = ((dens.shape[0],) if dyn else ()) + tuple(c.nvox)
    shape = tuple(int(v) for v in c.ray_shape)
    geom = types.SimpleNamespace(ray_starts=torch.as_tensor(c.xs).reshape(*shape, 3),
                                 rays=torch.as_tensor(c.dirs).reshape(*shape, 3), shape=shape)
    return c, dens, shape, S.Operator(g, geom, **kw)


@pytest.mark.parametrize("name", ["c1_rand", "c2_static_rand", "c2_static_rand_f64", "c2_multichannel",
                                  "c4_dynamic", "rand_full_inside", "rand_partial", "rand_log",
                                  "parallel_hollow"])
def test_stored_operator_matches_reference_golden(S, name):
    c, dens, shape, op = _case_operator(S, name, store_segments=torch.float64 if "f64" in name else True)
    assert op.n_segments == int(c.count.sum()) or not c.stable.all()
    assert op.stored_bytes >= op.n_segments * (12 if "f64" in name else 8)
    x = torch.from_numpy(dens).cuda().requires_grad_(True)
    y = op(x)
    want = c.integral
    tol = 1e-5 if dens.dtype == np.float32 else 1e-10
    st = np.broadcast_to(c.stable.reshape(shape), want.shape[-len(shape):] if shape else ())
    st = np.broadcast_to(st, want.shape)
    got = y.detach().cpu().numpy()
    assert got.shape == want.shape
    assert np.allclose(got[st], want[st], rtol=tol, atol=tol), np.abs(got[st] - want[st]).max()
    if hasattr(c, "grad"):
        w = case_weights(c)
        (y * torch.from_numpy(w).cuda()).sum().backward()
        g = x.grad.cpu().numpy()
        atol = tol * max(1.0, float(np.abs(c.grad).max()))
        assert np.allclose(g, c.grad, rtol=10 * tol, atol=atol), np.abs(g - c.grad).max()


def test_stored_operator_equals_on_the_fly(S):
    """Same geometry, both modes, float64 volume and float64 stored lengths: equal to rounding
    of the sums; float32 stored lengths: equal to 1e-6."""
    grid = S.SphericalGrid((40, 36, 48), size_r=(0.5, 3))
    geoms = sum(S.ConeRectGeom((48, 40), (7 * math.cos(t), 7 * math.sin(t), 1.5 * math.sin(3 * t)), fov=(50, 50))
                for t in (0.2, 1.1, 2.7, 4.0))
    fly = S.Operator(grid, geoms)
    st64 = S.Operator(grid, geoms, store_segments=torch.float64)
    st32 = S.Operator(grid, geoms, store_segments=True)
    _, total = fly.segment_counts()
    assert st64.n_segments == total and st32.stored_bytes == 8 * total + 8 * (fly._rays.n + 1)
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(tuple(grid.shape), generator=gen, device="cuda", dtype=torch.float64)
    y = torch.rand(fly._rays.shape, generator=gen, device="cuda", dtype=torch.float64)
    a, b, c = fly(x), st64(x), st32(x)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-13)
    assert torch.allclose(a, c, rtol=1e-6, atol=1e-7)
    ga, gb, gc = fly.T(y), st64.T(y), st32.T(y)
    assert torch.allclose(ga, gb, rtol=1e-11, atol=1e-12)
    assert torch.allclose(ga, gc, rtol=1e-6, atol=1e-7)
    # channels and autograd go through the same list
    xs = torch.rand((3,) + tuple(grid.shape), generator=gen, device="cuda").requires_grad_(True)
    ys = st32(xs)
    assert ys.shape == (3,) + fly._rays.shape
    ys.sum().backward()
    ref = fly.T(torch.ones(fly._rays.shape, device="cuda"))
    assert torch.allclose(xs.grad[1], ref, rtol=1e-5, atol=1e-5)
