"""GPU tests of the plan cache (``sphrt_rays.plan_cache``): the planning kernel's records and masks
are kept after the first call; later calls must give the same bits as a call that plans."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sph_raytracer_b200 as S
    return S


def _problem(S, dynamic=False):
    shape = (40, 36, 48)
    grid = S.SphericalGrid(((5,) + shape) if dynamic else shape, size_r=(0.5, 3))
    geoms = sum(S.ConeRectGeom((48, 40), (7 * math.cos(t), 7 * math.sin(t), 1.5 * math.sin(3 * t)), fov=(50, 50))
                for t in (0.2, 1.1, 2.7, 4.0, 5.3))
    return grid, geoms


@pytest.mark.parametrize("chunk", [None, "2048"])
@pytest.mark.parametrize("dynamic", [False, True])
def test_cached_plan_gives_the_same_results(S, monkeypatch, chunk, dynamic):
    """Forward, adjoint and the fused step, first (planning) call against the cached ones and
    against an operator that never caches; also with several chunks per launch, whose record
    areas and counts are kept apart."""
    monkeypatch.setenv("SPHRT_WALKER", "coop")
    if chunk:
        monkeypatch.setenv("SPHRT_WARP_CHUNK", chunk)
    grid, geoms = _problem(S, dynamic)
    op = S.Operator(grid, geoms)
    monkeypatch.setenv("SPHRT_PLAN_CACHE", "0")
    ref = S.Operator(grid, geoms)
    gen = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand(tuple(grid.shape), generator=gen, device="cuda")
    y = torch.rand(tuple(geoms.shape), generator=gen, device="cuda")
    y0, g0 = ref(x), ref._apply_adjoint(y, tuple(x.shape))
    assert ref.plan_cache_bytes == 0
    monkeypatch.delenv("SPHRT_PLAN_CACHE")
    for k in range(3):                                   # call 0 fills the cache, 1 and 2 use it
        assert torch.equal(op(x), y0), k
        g = op._apply_adjoint(y, tuple(x.shape))
        assert torch.allclose(g, g0, rtol=1e-5, atol=1e-6 * float(g0.abs().max())), k
        yn, gn = op.normal(x, b=y, scale=0.5)
        assert torch.equal(yn, y0), k
        assert op.plan_cache_bytes > 0
    assert op._rays.c.plan_cache_state == 2
    # multi-channel calls walk the same cached records once per group of four volumes
    if not dynamic:
        xs = torch.rand((5,) + tuple(grid.shape), generator=gen, device="cuda")
        assert torch.equal(op(xs), ref(xs))
    # another kernel path chosen through the environment: the cache is refilled, not trusted
    monkeypatch.setenv("SPHRT_WARP_CHUNK", "4096" if chunk is None else "1024")   # (smaller: the scratch still fits)
    assert torch.equal(op(x), y0)
    assert op._rays.c.plan_cache_state == 1
    assert torch.equal(op(x), y0) and op._rays.c.plan_cache_state == 2
