"""GPU parity tests of the warp-cooperative walker (csrc/sphrt_coop.cuh), the kernel that
serves large grids (BASELINE config 5).  It is checked three ways:

* its raw event export (every event of a ray in merged order, `sphrt_trace_events`) is put
  in canonical form and compared with the golden fixtures of the reference and with the CPU
  oracle: voxel indices and crossing counts bit-exact, lengths to 1e-9;
* line integrals and gradients computed through it (forced with SPHRT_WALKER=coop, or
  selected automatically on the config-5 grid) against golden / oracle values, rtol 1e-5;
* against the per-thread walker on the same rays (two independent implementations of the
  same reference algorithm on the GPU).
Rays that walker does not handle itself (starts inside the grid / on a boundary, snapped
cones, failed verification) take the per-thread kernels and are covered by
test_gpu_parity.py; here their number is bounded.
"""
import math
import os
import types

import numpy as np
import pytest
import torch

from helpers import (OPERATOR_CASES, load_case, case_density, case_weights, ragged_equal,
                     canonical_from_events)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import sph_raytracer_b200 as S
    return S


@pytest.fixture
def walker():
    """Select the walker for the ordinary rays through the environment (read at every launch)."""
    old = os.environ.get("SPHRT_WALKER")

    def choose(mode):
        if mode is None:
            os.environ.pop("SPHRT_WALKER", None)
        else:
            os.environ["SPHRT_WALKER"] = mode
    yield choose
    choose(old)


def _geom(xs, dirs, shape):
    xs, dirs = torch.as_tensor(xs).reshape(*shape, 3), torch.as_tensor(dirs).reshape(*shape, 3)
    return types.SimpleNamespace(ray_starts=xs, rays=dirs, shape=tuple(shape))


def _grid(c, dynamic=False, nt=0):
    g = types.SimpleNamespace(r_b=c.grid.r_b, e_b=c.grid.e_b, a_b=c.grid.a_b, dynamic=dynamic)
    g.shape = ((nt,) if dynamic else ()) + tuple(c.nvox)
    return g


def _events_canonical(op, max_seg):
    vox, lens, count, stats = op.events()
    cc, cv, cl = canonical_from_events(vox.cpu().numpy(), lens.cpu().numpy(), count.cpu().numpy(), max_seg)
    return cc, cv, cl, stats


# ---- sequences -------------------------------------------------------------------------------

@pytest.mark.parametrize("name", OPERATOR_CASES)
def test_coop_events_match_reference_golden(S, name):
    c = load_case(name)
    dens = case_density(c)
    dyn = name == "c4_dynamic"
    shape = tuple(int(v) for v in c.ray_shape)
    op = S.Operator(_grid(c, dyn, dens.shape[0] if dyn else 0), _geom(c.xs, c.dirs, shape))
    max_seg = max(int(c.count.max()), 1) + 8
    cc, cv, cl, stats = _events_canonical(op, max_seg)
    handled = cc >= 0
    assert stats[0] + stats[1] + stats[2] <= len(cc)
    degenerate = name.startswith("chords_") or name == "hemisphere"
    n, bad = ragged_equal(c, cc, cv, cl, mask=c.stable & handled & (not degenerate))
    assert len(bad) == 0, f"{name}: {len(bad)}/{n} rays differ, first {bad[:8]}"
    if name in ("c5_tile", "parallel_hollow"):
        # ordinary rays of a full-azimuth grid that start outside are walked cooperatively,
        # except the ~1 % whose inclination matches a cone's within the |aa| < tau snap
        assert stats[2] <= 0.03 * len(cc) and stats[3] == 0 and stats[0] > 0.5 * len(cc)


def _oracle_canonical(S, grid, geom, ms):
    from oracle import oracle as O
    xs, d, _ = O.broadcast_rays(geom.ray_starts, geom.rays)
    og = O.OracleGrid(grid)
    regs, lens = O.trace(og, xs, d)
    return og, regs, lens, O.canonical(og, regs, lens, max_seg=ms)


def _compare_handled(cc, cv, cl, v0, s0, c0):
    ms = v0.shape[1]
    handled = cc >= 0
    idx = np.arange(ms)[None, :] < c0[:, None]
    ok = (c0 == cc) & ((v0 == cv).all(-1) | ~idx).all(-1) & \
        (np.isclose(s0, cl, rtol=1e-9, atol=1e-12) | ~idx).all(-1)
    return handled, np.where(handled & ~ok)[0]


def test_coop_events_match_oracle_c5_tile(S):
    """BASELINE config 5 grid (200,180,360), one 512x512 detector sub-sampled 8x: the event
    sequences of the cooperative walker against the oracle, every ray."""
    grid = S.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
    g = S.ConeRectGeom((512, 512), (60.0, -70.0, -40.0), fov=(30, 30))
    geom = S.ViewGeom(g.ray_starts.expand(64, 64, 3), g.rays[3::8, 1::8])
    ms = 900
    og, regs, lens, (v0, s0, c0) = _oracle_canonical(S, grid, geom, ms)
    op = S.Operator(grid, geom)
    cc, cv, cl, stats = _events_canonical(op, ms)
    handled, bad = _compare_handled(cc, cv, cl, v0, s0, c0)
    assert len(bad) == 0, f"{len(bad)} rays differ from the oracle: {bad[:10]}"
    hits = c0 > 0
    assert (handled & hits).sum() >= 0.99 * hits.sum(), (stats, hits.sum())
    assert (cc[~hits & handled] == 0).all()


@pytest.mark.parametrize("seed", [0, 1])
def test_coop_events_match_oracle_random(S, seed):
    """Random rays from outside and inside on full-azimuth grids (linear, log, hollow, odd
    sizes): whatever the cooperative walker accepts must equal the oracle exactly."""
    gen = torch.Generator().manual_seed(seed)
    n = 6000
    xs = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    xs = xs / xs.norm(dim=-1, keepdim=True) * (3 + 4 * torch.rand(n, 1, generator=gen, dtype=torch.float64))
    d = -xs + 1.2 * torch.randn(n, 3, generator=gen, dtype=torch.float64)
    grids = [
        S.SphericalGrid((20, 16, 24), size_r=(0, 2.5)),
        S.SphericalGrid((30, 31, 33), size_r=(0.5, 2)),
        S.SphericalGrid((12, 40, 17), spacing="log", size_r=(0.1, 2.9)),
        S.SphericalGrid((9, 12, 64), size_r=(1, 2), size_e=(0.4, 2.6)),
    ]
    total = 0
    for grid in grids:
        geom = S.ViewGeom(xs, d)
        ms = 2 * (grid.shape[0] + grid.shape[1] + grid.shape[2]) + 8
        og, regs, lens, (v0, s0, c0) = _oracle_canonical(S, grid, geom, ms)
        op = S.Operator(grid, geom)
        cc, cv, cl, stats = _events_canonical(op, ms)
        handled, bad = _compare_handled(cc, cv, cl, v0, s0, c0)
        assert len(bad) == 0, f"grid {tuple(grid.shape)}: {len(bad)} rays differ: {bad[:10]}"
        total += int((handled & (c0 > 0)).sum())
        # the per-thread walker (which also takes the rays handed over) on ALL these rays; they
        # include nearly horizontal ones whose equator crossing the reference misplaces
        # (ray_has_snapped_aa in sphrt_walker.cuh)
        vox, sl, cnt = op.segments(ms)
        tc = cnt.reshape(-1).cpu().numpy()
        _, bad_t = _compare_handled(tc, vox.reshape(-1, ms, 3).cpu().numpy(),
                                    sl.reshape(-1, ms).cpu().numpy(), v0, s0, c0)
        assert len(bad_t) == 0, f"grid {tuple(grid.shape)}: per-thread walker differs: {bad_t[:10]}"
    assert total > 0.5 * n          # the test must exercise the walker, not its hand-over


# ---- integrals and gradients --------------------------------------------------------------------

@pytest.mark.parametrize("name", [n for n in OPERATOR_CASES
                                  if not (n.startswith("chords_") or n == "hemisphere")])
def test_coop_integrals_match_reference_golden(S, walker, name):
    """Forward and gradient through the forced cooperative path against the reference's values."""
    c = load_case(name)
    dens = case_density(c)
    if dens is None:
        pytest.skip("no integral in this fixture")
    dyn = name == "c4_dynamic"
    shape = tuple(int(v) for v in c.ray_shape)
    op = S.Operator(_grid(c, dyn, dens.shape[0] if dyn else 0), _geom(c.xs, c.dirs, shape))
    walker("coop")
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


def test_config5_grid_takes_cooperative_walker_and_matches_oracle(S, walker):
    """On the config-5 grid the cooperative walker is selected automatically; integrals and
    gradients against the oracle, and against the per-thread walker in float64."""
    from oracle import oracle as O
    walker(None)
    grid = S.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
    g = S.ConeRectGeom((512, 512), (-20.0, 95.0, 25.0), fov=(30, 30))
    geom = S.ViewGeom(g.ray_starts.expand(64, 48, 3), g.rays[2::8, 5::10][:, :48])
    op = S.Operator(grid, geom)
    _, _, _, stats = op.events(max_ev=8)
    assert stats[0] > 0.6 * 64 * 48 and stats[3] == 0
    xs, d, _ = O.broadcast_rays(geom.ray_starts, geom.rays)
    og = O.OracleGrid(grid)
    regs, lens = O.trace(og, xs, d)
    rng = np.random.default_rng(3)
    dens = rng.random(tuple(grid.shape), dtype=np.float32)
    w = rng.random((64, 48), dtype=np.float32)
    x = torch.from_numpy(dens).cuda().requires_grad_(True)
    y = op(x)                                               # auto -> cooperative
    (y * torch.from_numpy(w).cuda()).sum().backward()
    want = O.forward(og, regs, lens, dens)[0].reshape(64, 48)
    assert np.allclose(y.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    want_g = O.adjoint(og, regs, lens, w.reshape(1, -1))[0]
    got_g = x.grad.cpu().numpy()
    assert np.allclose(got_g, want_g, rtol=1e-4, atol=1e-5 * float(np.abs(want_g).max()))
    # float64 volume: both GPU walkers agree to rounding of the sum
    x64 = torch.from_numpy(dens.astype(np.float64)).cuda()
    w64 = torch.from_numpy(w.astype(np.float64)).cuda()
    walker("tracked"); y_t = op._apply_forward(x64); g_t = op._apply_adjoint(w64, x64.shape)
    walker("coop"); y_c = op._apply_forward(x64); g_c = op._apply_adjoint(w64, x64.shape)
    assert torch.allclose(y_t, y_c, rtol=1e-12, atol=1e-12)
    assert torch.allclose(g_t, g_c, rtol=1e-11, atol=1e-12)


def test_coop_adjoint_is_transpose_at_scale(S, walker):
    """Dot test <A x, y> = <x, A^T y> on two whole config-5 observations (524 288 rays), float64."""
    import bench
    walker(None)
    grid, geoms, _ = bench.workload("c5", 2)
    op = S.Operator(grid, S.ViewGeomCollection(*geoms))
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(tuple(grid.shape), generator=gen, device="cuda", dtype=torch.float64)
    y = torch.rand(op._rays.shape, generator=gen, device="cuda", dtype=torch.float64)
    lhs = (op._apply_forward(x) * y).sum().item()
    rhs = (x * op._apply_adjoint(y, x.shape)).sum().item()
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)
    # linearity of the forward pass
    a = op._apply_forward(2.5 * x)
    assert torch.allclose(a, 2.5 * op._apply_forward(x), rtol=1e-12)


def test_coop_multichannel_and_dynamic(S, walker):
    """Channel batches (4 volumes per walk) and time-indexed volumes through the cooperative path."""
    walker("coop")
    grid = S.SphericalGrid((24, 20, 36))
    geoms = sum(S.ConeRectGeom((24, 20), (4 * math.cos(t), 4 * math.sin(t), 0.7), fov=(35, 35))
                for t in (0.3, 1.7, 2.9))
    op = S.Operator(grid, geoms)
    vols = torch.rand((6,) + tuple(grid.shape), device="cuda")
    y = op(vols)
    for ch in range(6):
        assert torch.allclose(y[ch], op(vols[ch]), rtol=1e-6, atol=1e-6)
    walker("tracked")
    y_t = op(vols)
    assert torch.allclose(y, y_t, rtol=1e-6, atol=1e-6)
    gd = S.SphericalGrid((3, 24, 20, 36))
    opd = S.Operator(gd, geoms)
    vol = torch.rand(tuple(gd.shape), device="cuda", dtype=torch.float64)
    w = torch.rand(3, 24, 20, device="cuda", dtype=torch.float64)
    walker("coop"); yc = opd(vol); gc = opd.T(w)
    walker("tracked"); yt = opd(vol); gt = opd.T(w)
    assert torch.allclose(yc, yt, rtol=1e-12, atol=1e-13)
    assert torch.allclose(gc, gt, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("n", [150, 200, 230])
def test_grid_sizes_around_the_48k_shared_memory_boundary(S, walker, n):
    """Boundary tables of 8-13 KB put static + dynamic shared memory of the fallback kernels next
    to the 48 KB opt-in limit (a (200,200,200) grid failed to launch once)."""
    grid = S.SphericalGrid((n, n, n), size_r=(0.5, 5))
    g = S.ConeRectGeom((64, 64), (80.0, 10.0, 20.0), fov=(7, 7))
    op = S.Operator(grid, g)
    x = torch.rand(tuple(grid.shape), device="cuda", dtype=torch.float64)
    walker("tracked"); a = op._apply_forward(x)
    walker("coop"); b = op._apply_forward(x)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-12)
    cnt, total = op.segment_counts()
    assert total > 0


def test_rays_with_a_snapped_leading_coefficient(S):
    """A ray whose inclination matches a cone's angle within the |aa| < tau snap (raytracer.py:376)
    gets that cone's root displaced, and the reference's elevation sequence stops being adjacent
    (71 -> 73 -> 72 ...).  Both walkers must take such rays from the sorted list of all cone roots
    (found by tools/xcheck.py: the per-thread bin tracker used to walk them)."""
    rays = [([41.87165449473164, -66.74590637381635, 26.85256007936145],
             [-44.167720339167914, 68.42400838356247, -27.08770691345806]),
            ([14.996988686600575, -80.19249035690713, -6.798118500014652],
             [-12.336793973804266, 80.54483212839102, 6.381406568204781]),
            ([-22.193157227915137, -58.82050434403097, -29.834515536095495],
             [23.964697249196313, 64.51301336421729, 34.094206626620924]),
            ([-53.171844506825096, -14.80506756622339, 45.923453476975766],
             [62.09546843462786, 17.85875431446437, -52.85088629478713])]
    grid = S.SphericalGrid((64, 181, 91), size_r=(0, 20))
    geom = S.ViewGeom(torch.tensor([r[0] for r in rays], dtype=torch.float64),
                      torch.tensor([r[1] for r in rays], dtype=torch.float64))
    ms = 800
    og, regs, lens, (v0, s0, c0) = _oracle_canonical(S, grid, geom, ms)
    op = S.Operator(grid, geom)
    cc, cv, cl, stats = _events_canonical(op, ms)
    handled, bad = _compare_handled(cc, cv, cl, v0, s0, c0)
    assert handled.all() and len(bad) == 0, (stats, bad)
    vox, sl, cnt = op.segments(ms)
    _, bad_t = _compare_handled(cnt.reshape(-1).cpu().numpy(), vox.reshape(-1, ms, 3).cpu().numpy(),
                                sl.reshape(-1, ms).cpu().numpy(), v0, s0, c0)
    assert len(bad_t) == 0, bad_t


def test_masked_rays_are_skipped_in_the_adjoint(S, walker):
    """Rays with zero weight (loss.py projection_mask) contribute nothing and are not walked."""
    grid = S.SphericalGrid((40, 36, 48), size_r=(0.5, 3))
    geom = S.ConeRectGeom((64, 48), (7.0, 1.0, 2.0), fov=(50, 50))
    op = S.Operator(grid, geom)
    gen = torch.Generator(device="cuda").manual_seed(2)
    y = torch.rand(64, 48, generator=gen, device="cuda", dtype=torch.float64)
    mask = (torch.rand(64, 48, generator=gen, device="cuda") > 0.5).double()
    for mode in ("coop", "tracked"):
        walker(mode)
        full = op.T(y)
        part = op.T(y * mask) + op.T(y * (1 - mask))
        assert torch.allclose(full, part, rtol=1e-11, atol=1e-12)
    ych = torch.stack([y * mask, torch.zeros_like(y), y])          # channels: zero only in some
    walker("coop")
    g3 = op.adjoint(ych, density_shape=(3,) + tuple(grid.shape))
    assert torch.allclose(g3[0], op.T(y * mask), rtol=1e-11, atol=1e-12) and g3[1].abs().max() == 0
    assert torch.allclose(g3[2], op.T(y), rtol=1e-11, atol=1e-12)
