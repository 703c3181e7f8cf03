"""GPU tests of the fused normal-equations step (``Operator.normal`` / ``sphrt_normal``,
``loss.FusedSquareLoss``): one walk of the rays must give what the forward pass followed by the
adjoint pass gives -- the integrals bit for bit (same kernels, same order of the float32 sums), the
gradient to the rounding of its atomic sums."""
import math
import types

import numpy as np
import pytest
import torch

from helpers import load_case, case_density

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


@pytest.mark.parametrize("walker", ["coop", "auto"])
@pytest.mark.parametrize("name", ["c1_rand", "c2_static_rand", "c2_static_rand_f64", "c4_dynamic",
                                  "rand_full_inside", "rand_partial", "rand_log", "onplane_16"])
def test_normal_equals_forward_then_adjoint(S, monkeypatch, name, walker):
    """Golden cases incl. the ones whose rays mostly go to the per-thread fallback passes (starts
    inside the grid, partial azimuth range): y bit-identical, gradient to 1e-5 (f32) / 1e-11 (f64)."""
    if walker == "coop":
        monkeypatch.setenv("SPHRT_WALKER", "coop")
    c, dens, shape, op = _case_operator(S, name)
    x = torch.from_numpy(dens).cuda()
    gen = torch.Generator(device="cuda").manual_seed(3)
    y_shape = tuple(op._plan(tuple(x.shape))[2])
    b = torch.rand(y_shape, generator=gen, device="cuda", dtype=x.dtype)
    mask = (torch.rand(y_shape, generator=gen, device="cuda") > 0.2).to(x.dtype)
    # rays that are chaotic in the reference (through the origin) may differ
    # between the two walkers: they carry no weight and are left out of the comparison of y
    st = torch.from_numpy(np.broadcast_to(c.stable.reshape(shape), y_shape).copy()).cuda()
    mask = mask * st.to(x.dtype)
    for bb, mm, scale in ((b, mask, 0.37), (None, mask, 1.0), (b, mask, 2.0 / b.numel())):
        y, g = op.normal(x, b=bb, mask=mm, scale=scale)
        if walker == "coop" and bb is None:            # also without a mask
            y1, g1 = op.normal(x)
            assert torch.equal(y1, op(x)) and torch.allclose(g1, op._apply_adjoint(y1, tuple(x.shape)), rtol=1e-5,
                                                             atol=1e-5 * float(g1.abs().max()) + 1e-30)
        y0 = op(x)
        w = y0 if bb is None else y0 - bb
        if mm is not None:
            w = w * mm
        g0 = op.T(w * scale) if not op.grid.dynamic else op._apply_adjoint(w * scale, tuple(x.shape))
        tol = 1e-5 if x.dtype == torch.float32 else 1e-11
        if walker == "coop":        # the same kernel code walked the rays: the same bits
            assert torch.equal(y, y0)
        else:                       # (the operator may have chosen the per-thread walker for its own passes)
            assert torch.allclose(y[st], y0[st], rtol=tol, atol=tol)
        assert torch.allclose(g, g0, rtol=tol, atol=tol * float(g0.abs().max()) + 1e-30), float((g - g0).abs().max())


def test_normal_on_stored_and_tiny_operators_falls_back_to_two_passes(S):
    grid = S.SphericalGrid((6, 5, 8), size_r=(0.5, 2))
    geom = S.ConeRectGeom((8, 8), pos=(6.0, 1.5, 2.0), fov=(40, 40))
    x = torch.rand(tuple(grid.shape), device="cuda")
    b = torch.rand((8, 8), device="cuda")
    ref = None
    for kw in ({}, {"store_segments": True}, {"store_segments": "auto"}):
        op = S.Operator(grid, geom, **kw)
        y, g = op.normal(x, b=b, scale=0.5)
        if ref is None:
            ref = (y, g)
            assert torch.allclose(g, op.T((op(x) - b) * 0.5), rtol=1e-5, atol=1e-6)
        assert torch.allclose(y, ref[0], rtol=1e-5, atol=1e-6) and torch.allclose(g, ref[1], rtol=1e-5, atol=1e-6)


def test_fused_square_loss_is_square_loss(S):
    """Same value and gradient as SquareLoss (with a projection and a volume mask), and the same gd()
    trajectory."""
    from sph_raytracer_b200 import loss as L, model as M, retrieval as RT
    grid = S.SphericalGrid((24, 20, 28), size_r=(0.5, 3))
    geoms = sum(S.ConeRectGeom((24, 20), (7 * math.cos(t), 7 * math.sin(t), 1.5 * math.sin(3 * t)), fov=(50, 50))
                for t in (0.2, 1.1, 2.7, 4.0, 5.1))
    op = S.Operator(grid, geoms)
    gen = torch.Generator(device="cuda").manual_seed(1)
    truth = torch.rand(tuple(grid.shape), generator=gen, device="cuda")
    y = op(truth)
    pm = (torch.rand(y.shape, generator=gen, device="cuda") > 0.1).float()
    vm = (torch.rand(truth.shape, generator=gen, device="cuda") > 0.05).float()
    d = torch.rand(truth.shape, generator=gen, device="cuda")
    vals, grads = [], []
    for cls in (L.SquareLoss, L.FusedSquareLoss):
        dd = d.clone().requires_grad_(True)
        v = (3 * cls(projection_mask=pm, volume_mask=vm))(op, y, dd, None)
        v.backward()
        vals.append(float(v)); grads.append(dd.grad)
    assert math.isclose(vals[0], vals[1], rel_tol=1e-5)
    assert torch.allclose(grads[0], grads[1], rtol=1e-4, atol=1e-6 * float(grads[0].abs().max()))
    hist = []
    for cls in (L.SquareLoss, L.FusedSquareLoss):
        fns = [1 * cls(), 1 * L.NegRegularizer()]
        _, _, losses = RT.gd(op, y, M.FullyDenseModel(grid), lr=1e-1, loss_fns=fns, num_iterations=25,
                             progress_bar=False)
        hist.append(np.array([losses[k] for k in fns]))
    assert np.allclose(hist[0], hist[1], rtol=2e-3, atol=1e-6 * hist[0].max())
