"""GPU parity tests of the launch configuration the full config-5 benchmark runs.

From 4 Mi rays per launch on the library walks the ordinary rays and the all-cones rays
(snapped cones, starts on a grid cone / half-plane) in two separate launches of the
cooperative walker; below that one launch takes both.  Every headline number comes from the
two-launch path, so it is forced here on the small parity cases
(`SPHRT_COOP_SPLIT_RAYS=0`) and compared with the reference's golden values, the CPU oracle
and the one-launch path, and it is run once at its natural size (16 whole config-5
observations, 4.19 M rays) against size-independent properties.
"""
import os
import types

import numpy as np
import pytest
import torch

from helpers import OPERATOR_CASES, load_case, case_density, case_weights

pytestmark = pytest.mark.gpu

CASES = [n for n in OPERATOR_CASES if not (n.startswith("chords_") or n == "hemisphere")]


@pytest.fixture(scope="module")
def S():
    import sph_raytracer_b200 as S
    return S


@pytest.fixture
def env():
    """Set launch-selection variables of the library (read at every launch); restored afterwards."""
    names = ("SPHRT_WALKER", "SPHRT_COOP_SPLIT_RAYS")
    old = {k: os.environ.get(k) for k in names}

    def setenv(**kw):
        for k, v in kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
    yield setenv
    setenv(**old)


def _geom(xs, dirs, shape):
    xs, dirs = torch.as_tensor(xs).reshape(*shape, 3), torch.as_tensor(dirs).reshape(*shape, 3)
    return types.SimpleNamespace(ray_starts=xs, rays=dirs, shape=tuple(shape))


def _grid(c, dynamic=False, nt=0):
    g = types.SimpleNamespace(r_b=c.grid.r_b, e_b=c.grid.e_b, a_b=c.grid.a_b, dynamic=dynamic)
    g.shape = ((nt,) if dynamic else ()) + tuple(c.nvox)
    return g


@pytest.mark.parametrize("name", CASES)
def test_split_launch_matches_reference_golden(S, env, name):
    """Integrals and gradients through the forced two-launch path against the reference's own
    values (rtol 1e-5 for float32 volumes), and bit-for-bit forward agreement with the
    one-launch path in float64."""
    c = load_case(name)
    dens = case_density(c)
    if dens is None:
        pytest.skip("no integral in this fixture")
    dyn = name == "c4_dynamic"
    shape = tuple(int(v) for v in c.ray_shape)
    op = S.Operator(_grid(c, dyn, dens.shape[0] if dyn else 0), _geom(c.xs, c.dirs, shape))
    env(SPHRT_WALKER="coop", SPHRT_COOP_SPLIT_RAYS=0)
    x = torch.from_numpy(dens).cuda().requires_grad_(True)
    y = op(x)
    want = c.integral
    tol = 1e-5 if dens.dtype == np.float32 else 1e-10
    st = np.broadcast_to(c.stable.reshape(shape), want.shape[-len(shape):] if shape else ())
    st = np.broadcast_to(st, want.shape)
    got = y.detach().cpu().numpy()
    assert np.allclose(got[st], want[st], rtol=tol, atol=tol), np.abs(got[st] - want[st]).max()
    if hasattr(c, "grad"):
        w = case_weights(c)
        (y * torch.from_numpy(w).cuda()).sum().backward()
        g = x.grad.cpu().numpy()
        atol = tol * max(1.0, float(np.abs(c.grad).max()))
        assert np.allclose(g, c.grad, rtol=10 * tol, atol=atol), np.abs(g - c.grad).max()
    # the same rays through one launch: identical sums in float64 (same events, same order)
    x64 = torch.from_numpy(dens.astype(np.float64)).cuda()
    y_split = op._apply_forward(x64)
    env(SPHRT_COOP_SPLIT_RAYS=1 << 40)
    y_one = op._apply_forward(x64)
    assert torch.equal(y_split, y_one)


@pytest.mark.parametrize("seed", [11, 12])
def test_split_launch_matches_oracle_on_fuzzed_rays(S, env, seed):
    """Random rays (from outside, aimed at the grid) on full-azimuth grids through the forced
    two-launch path against the CPU oracle: integrals and gradients."""
    from oracle import oracle as O
    gen = torch.Generator().manual_seed(seed)
    n = 4096
    xs = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    xs = xs / xs.norm(dim=-1, keepdim=True) * (3 + 4 * torch.rand(n, 1, generator=gen, dtype=torch.float64))
    d = -xs + 1.2 * torch.randn(n, 3, generator=gen, dtype=torch.float64)
    rng = np.random.default_rng(seed)
    for grid in (S.SphericalGrid((20, 16, 24), size_r=(0, 2.5)),
                 S.SphericalGrid((30, 31, 33), size_r=(0.5, 2)),
                 S.SphericalGrid((12, 40, 17), spacing="log", size_r=(0.1, 2.9))):
        geom = S.ViewGeom(xs, d)
        og = O.OracleGrid(grid)
        ox, od, _ = O.broadcast_rays(geom.ray_starts, geom.rays)
        regs, lens = O.trace(og, ox, od)
        dens = rng.random(tuple(grid.shape), dtype=np.float32)
        w = rng.random((1, n), dtype=np.float32)
        want = O.forward(og, regs, lens, dens)[0]
        want_g = O.adjoint(og, regs, lens, w)[0]
        op = S.Operator(grid, geom)
        env(SPHRT_WALKER="coop", SPHRT_COOP_SPLIT_RAYS=0)
        got = op._apply_forward(torch.from_numpy(dens).cuda()).cpu().numpy()
        got_g = op._apply_adjoint(torch.from_numpy(w[0]).cuda(), dens.shape).cpu().numpy()
        assert np.allclose(got, want, rtol=1e-5, atol=1e-5), np.abs(got - want).max()
        assert np.allclose(got_g, want_g, rtol=1e-4, atol=1e-5 * float(np.abs(want_g).max()))


def test_split_launch_at_its_natural_size(S, env):
    """16 whole config-5 observations = 4 194 304 rays: the smallest launch that takes the
    two-launch path without any override (what `python bench.py` runs).  Checked through
    size-independent properties: chord lengths through a volume of ones (analytic), the
    adjoint dot test in float64, and one observation launched alone (one-launch path) giving
    bit-identical integrals."""
    import bench
    env(SPHRT_WALKER=None, SPHRT_COOP_SPLIT_RAYS=None)
    grid, geoms, _ = bench.workload("c5", 16)
    op = S.Operator(grid, S.ViewGeomCollection(*geoms))
    assert op._rays.n >= 4 << 20 and op.walker() == "cooperative"
    ones = torch.ones(tuple(grid.shape), device="cuda", dtype=torch.float64)
    y = op._apply_forward(ones).reshape(16, 512, 512)
    # chord through the hollow sphere r in (3, 25) of a ray with impact parameter b
    for k in (0, 7, 15):
        g = geoms[k]
        pos = torch.as_tensor(g.ray_starts, dtype=torch.float64).reshape(3)
        d = torch.as_tensor(g.rays, dtype=torch.float64)
        d = d / d.norm(dim=-1, keepdim=True)
        tc = -(d * pos).sum(-1)
        b2 = (pos * pos).sum() - tc * tc
        chord = 2 * (torch.sqrt((625.0 - b2).clamp(min=0)) - torch.sqrt((9.0 - b2).clamp(min=0)))
        got = y[k].cpu()
        # rays grazing a shell within rounding may gain or lose a sliver
        assert torch.allclose(got, chord, rtol=1e-9, atol=1e-6), (got - chord).abs().max()
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(tuple(grid.shape), generator=gen, device="cuda", dtype=torch.float64)
    w = torch.rand(op._rays.shape, generator=gen, device="cuda", dtype=torch.float64)
    ax = op._apply_forward(x)
    lhs = (ax * w).sum().item()
    rhs = (x * op._apply_adjoint(w, x.shape)).sum().item()
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
    alone = S.Operator(grid, geoms[7])
    assert alone._rays.n < 4 << 20
    assert torch.equal(alone._apply_forward(x).reshape(512, 512), ax.reshape(16, 512, 512)[7])
