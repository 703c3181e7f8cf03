"""GPU parity tests: the CUDA path (through the C ABI / Operator) against the golden
fixtures produced by the reference and against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): voxel indices and per-ray crossing counts bit-exact on
oracle-stable rays; segment lengths to 1e-9; line integrals and gradients to rtol 1e-5 in
fp32 (1e-10 in fp64).
"""
import math
import types

import numpy as np
import pytest
import torch

from helpers import OPERATOR_CASES, load_case, case_density, case_weights, ragged_equal, GOLDEN

pytestmark = pytest.mark.gpu


def _geom(xs, dirs, shape=None):
    """Geometry stand-in that hands the operator exactly these origins / directions."""
    xs, dirs = torch.as_tensor(xs), torch.as_tensor(dirs)
    if shape is not None:
        xs, dirs = xs.reshape(*shape, 3), dirs.reshape(*shape, 3)
    return types.SimpleNamespace(ray_starts=xs, rays=dirs, shape=tuple(dirs.shape[:-1]))


def _grid(c, dynamic=False, nt=0):
    g = types.SimpleNamespace(r_b=c.grid.r_b, e_b=c.grid.e_b, a_b=c.grid.a_b, dynamic=dynamic)
    g.shape = ((nt,) if dynamic else ()) + tuple(c.nvox)
    return g


@pytest.fixture(scope="module")
def S():
    import sph_raytracer_b200 as S
    return S


@pytest.mark.parametrize("name", OPERATOR_CASES)
def test_cuda_matches_reference_golden(S, name):
    c = load_case(name)
    dyn = name == "c4_dynamic"
    dens = case_density(c)
    grid = _grid(c, dynamic=dyn, nt=dens.shape[0] if dyn else 0)
    shape = tuple(int(v) for v in c.ray_shape)
    op = S.Operator(grid, _geom(c.xs, c.dirs, shape))
    max_seg = max(int(c.count.max()), 1) + 8
    vox, lens, count = op.segments(max_seg)
    vox = vox.reshape(-1, max_seg, 3).cpu().numpy()
    lens = lens.reshape(-1, max_seg).cpu().numpy()
    count = count.reshape(-1).cpu().numpy()
    degenerate = name.startswith("chords_") or name == "hemisphere"
    n, bad = ragged_equal(c, count, vox, lens, mask=c.stable & (not degenerate))
    assert len(bad) == 0, f"{name}: {len(bad)}/{n} stable rays differ, first {bad[:8]}"
    cnt2, total = op.segment_counts()
    assert np.array_equal(cnt2.reshape(-1).cpu().numpy(), count)
    assert total == int(count.sum())
    if dens is None:
        return

    x = torch.from_numpy(dens).cuda().requires_grad_(True)
    y = op(x)
    assert y.dtype == x.dtype and y.is_cuda
    want = c.integral
    assert tuple(y.shape) == want.shape, (y.shape, want.shape)
    tol = 1e-5 if dens.dtype == np.float32 else 1e-10
    st = np.broadcast_to(c.stable.reshape(shape), want.shape[-len(shape):] if shape else ())
    st = np.broadcast_to(st, want.shape)
    got = y.detach().cpu().numpy()
    assert np.allclose(got[st], want[st], rtol=tol, atol=tol), \
        f"{name}: max err {np.abs(got[st] - want[st]).max()}"
    if name.startswith("chords_"):
        assert np.allclose(got, 2 * (c.r_b[-1] - c.r_b[0]), atol=1e-2)   # test_raytracer.py:49-52
    if name == "hemisphere":
        assert np.allclose(got, 50, atol=1e-2)                          # test_raytracer.py:85-116
    if hasattr(c, "grad"):
        w = case_weights(c)                       # zero on the oracle-unstable rays
        (y * torch.from_numpy(w).cuda()).sum().backward()
        g = x.grad.cpu().numpy()
        want_g = c.grad
        atol = tol * max(1.0, float(np.abs(want_g).max()))
        assert np.allclose(g, want_g, rtol=10 * tol, atol=atol), \
            f"{name}: grad max err {np.abs(g - want_g).max()}"


def test_known_answer_tables_r_e_a_torch(S):
    """r_torch / e_torch / a_torch of the mirror (CUDA) on the known-answer inputs of the
    reference's own tests (test_all.py:18-173): regions exact wherever t is finite, t equal up to
    the reference build's 1-ulp sqrt, same +-inf pattern."""
    from sph_raytracer_b200.raytracer import r_torch, e_torch, a_torch
    z = np.load(f"{GOLDEN}/known_answers.npz")
    fn = {"r": r_torch, "e": e_torch, "a": a_torch}
    for i in range(int(z["n_cases"])):
        kind = str(z[f"c{i:02d}_kind"])
        t, reg, inds, neg = fn[kind](z[f"c{i:02d}_b"], z[f"c{i:02d}_xs"], z[f"c{i:02d}_dirs"])
        t, reg = t.numpy(), reg.numpy()
        T, R = z[f"c{i:02d}_t"], z[f"c{i:02d}_reg"]
        assert t.shape == T.shape and tuple(inds.shape) == T.shape and tuple(neg.shape) == T.shape
        fin = np.isfinite(T)
        assert np.array_equal(np.isfinite(t), fin), (i, kind)
        assert np.array_equal(t[~fin], T[~fin]), (i, kind)
        assert np.allclose(t[fin], T[fin], rtol=4e-16, atol=0), (i, kind)
        assert np.array_equal(reg[fin], R[fin]), (i, kind)
        assert set(np.unique(neg.numpy())) <= {0, 1}


def test_find_starts_known_answers(S):
    z = np.load(f"{GOLDEN}/known_answers.npz")
    for row in z["find_starts"]:
        shape, x, want = row[:3].astype(int), row[3:6], row[6:9].astype(int)
        got = S.find_starts(S.SphericalGrid(shape=tuple(shape)), x)
        assert got.tolist() == want.tolist()
        from sph_raytracer_b200.raytracer import find_starts_cuda
        assert find_starts_cuda(S.SphericalGrid(shape=tuple(shape)), x).tolist() == want.tolist()


def _oracle_compare(S, grid, geom, max_seg=None, allow_unstable=0):
    from oracle import oracle as O
    op = S.Operator(grid, geom)
    xs, d, shape = O.broadcast_rays(geom.ray_starts, geom.rays)
    og = O.OracleGrid(grid)
    regs, lens = O.trace(og, xs, d)
    ms = max_seg or og.K
    v0, s0, c0 = O.canonical(og, regs, lens, max_seg=ms)
    vox, sl, cnt = op.segments(ms)
    v1 = vox.reshape(-1, ms, 3).cpu().numpy(); s1 = sl.reshape(-1, ms).cpu().numpy()
    c1 = cnt.reshape(-1).cpu().numpy()
    idx = np.arange(ms)[None, :] < c0[:, None]
    ok = (c0 == c1) & ((v0 == v1).all(-1) | ~idx).all(-1) & \
        (np.isclose(s0, s1, rtol=1e-9, atol=1e-12) | ~idx).all(-1)
    bad = np.where(~ok)[0]
    assert len(bad) <= allow_unstable, f"{len(bad)} rays differ from the oracle: {bad[:10]}"
    return op, og, regs, lens, ok


def test_cuda_matches_oracle_c1_full(S):
    """BASELINE config 1 (examples/single_vantage.py) at full size, every ray."""
    grid = S.SphericalGrid(shape=(50, 50, 50))
    geom = S.ConeRectGeom((256, 256), (5, 0, 0), fov=(45, 45))
    op, og, regs, lens, ok = _oracle_compare(S, grid, geom)
    from oracle import oracle as O
    x = torch.zeros(grid.shape); x[-1] += 1; x[-10] += 1
    y = op(x.cuda()).cpu().numpy().reshape(-1)
    want = O.forward(og, regs, lens, x.numpy())[0]
    assert np.allclose(y, want, rtol=1e-5, atol=1e-6)


def test_cuda_matches_oracle_c2_observations(S):
    """BASELINE config 2/3 geometry (static_retrieval.py), 6 of the 50 observations whole.
    Row 0 of every ConeCircGeom passes exactly through the origin (50 coincident rays per
    observation) -- the reference is chaotic there."""
    grid = S.SphericalGrid(shape=(50, 50, 50))
    thetas = torch.linspace(0, 2 * torch.pi, 50)[[0, 3, 11, 24, 37, 49]]
    geoms = sum(S.ConeCircGeom((100, 50), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(0, 45))
                for t in thetas)
    op, og, regs, lens, ok = _oracle_compare(S, grid, geoms, allow_unstable=6 * 50)
    bad_rows = np.where(~ok)[0] % 5000 // 50
    assert (bad_rows == 0).all(), "mismatches outside the through-origin detector row"


def test_cuda_matches_oracle_c5_tile(S):
    """BASELINE config 5 grid (200,180,360), one 512x512 detector sub-sampled 8x."""
    grid = S.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
    g = S.ConeRectGeom((512, 512), (60.0, -70.0, -40.0), fov=(30, 30))
    geom = S.ViewGeom(g.ray_starts.expand(64, 64, 3), g.rays[3::8, 1::8])
    _oracle_compare(S, grid, geom, max_seg=900)


@pytest.mark.parametrize("seed", [0, 1])
def test_cuda_matches_oracle_random(S, seed):
    """Random rays, starts inside and outside, partial / log / hollow / odd grids."""
    gen = torch.Generator().manual_seed(seed)
    n = 20000
    xs = torch.randn(n, 3, generator=gen, dtype=torch.float64) * 2
    d = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    grids = [
        (S.SphericalGrid((13, 9, 11), size_r=(0.5, 3), size_e=(0.3, 2.9), size_a=(-2.5, 2.9)), xs, d),
        (S.SphericalGrid((7, 6, 5), spacing="log", size_r=(0.1, 4)), xs, d),
        (S.SphericalGrid((20, 16, 24)), xs * 0.3, d),
        (S.SphericalGrid((30, 31, 33), size_r=(0, 2)), xs * 3, -xs + 0.8 * d),
        (S.SphericalGrid((5, 7, 9), size_r=(1, 2)), xs * 2, -xs + 0.5 * d),
        (S.SphericalGrid((1, 1, 1)), xs, d),
    ]
    for grid, starts, dirs in grids:
        _oracle_compare(S, grid, S.ViewGeom(starts, dirs))


def test_operator_shapes_and_devices(S):
    """Output shape / dtype / device rules (reference test_raytracer.py:55-82)."""
    geom = S.ConeRectGeom((64, 64), (1, 0, 0))
    for gshape, dshape in [((2, 3, 4), (2, 3, 4)), ((2, 3, 4), (10, 2, 3, 4)),
                           ((10, 2, 3, 4), (10, 2, 3, 4))]:
        op = S.Operator(S.SphericalGrid(gshape), geom)
        d = torch.rand(dshape)
        y = op(d)                                  # host tensor in -> host tensor out
        assert tuple(y.shape) == tuple(dshape[:-3]) + (64, 64) and y.device.type == "cpu"
        y2 = op(d.cuda().double())
        assert y2.is_cuda and y2.dtype == torch.float64
        assert torch.allclose(y2.float().cpu(), y, rtol=1e-5, atol=1e-6)
    # multi-channel volume through a single ray (test_raytracer.py:55-60)
    grid = S.SphericalGrid(shape=(25, 25, 25), size_r=(5, 10))
    op = S.Operator(grid, S.ViewGeom([-100, 0, 0], [1, 0, 0]))
    assert op(torch.rand((5,) + tuple(grid.shape))).shape == (5,)
    assert repr(op) == f"Operator({grid.shape} → {op.geom.shape})"
    # dynamic volume against a collection: T must equal the number of observations
    geoms = sum(S.ConeRectGeom((8, 8), (3 * math.cos(t), 3 * math.sin(t), 0.5)) for t in (0.1, 1.0, 2.5))
    gd = S.SphericalGrid((3, 4, 5, 6))
    opd = S.Operator(gd, geoms)
    vol = torch.rand(tuple(gd.shape)).cuda()
    yd = opd(vol)
    assert yd.shape == (3, 8, 8)
    for o in range(3):      # observation o reads time-slice o
        single = S.Operator(S.SphericalGrid((4, 5, 6)), geoms[o])(vol[o])
        assert torch.allclose(yd[o], single, rtol=1e-6, atol=1e-7)
    with pytest.raises(IndexError):
        opd(torch.rand((4, 4, 5, 6)).cuda())
    # adjoint of the dynamic operator (the reference raises NotImplementedError here)
    w = torch.rand(3, 8, 8, dtype=torch.float64).cuda()
    vol64 = vol.double()
    lhs = (opd(vol64) * w).sum().item()
    rhs = (vol64 * opd.T(w)).sum().item()
    assert abs(lhs - rhs) <= 1e-11 * abs(lhs)


def test_adjoint_dot_test_and_T(S):
    """<A x, y> == <x, A^T y> for the static operator; T equals the autograd gradient."""
    grid = S.SphericalGrid(shape=(50, 50, 50))
    geoms = sum(S.ConeCircGeom((100, 50), (5 * math.cos(t), 5 * math.sin(t), 1), fov=(0, 45))
                for t in (0.3, 2.0, 4.4))
    op = S.Operator(grid, geoms)
    gen = torch.Generator().manual_seed(0)
    for dt, tol in ((torch.float64, 1e-11), (torch.float32, 2e-5)):
        x = torch.rand(tuple(grid.shape), generator=gen, dtype=dt).cuda().requires_grad_(True)
        y = torch.rand(op.geom.shape, generator=gen, dtype=dt).cuda()
        Ax = op(x)
        lhs = (Ax.double() * y.double()).sum()
        ATy = op.T(y)
        rhs = (x.detach().double() * ATy.double()).sum()
        assert abs(lhs.item() - rhs.item()) <= tol * abs(lhs.item())
        (Ax * y).sum().backward()
        assert torch.allclose(x.grad, ATy, rtol=10 * tol, atol=10 * tol * ATy.abs().max().item())
    # linearity of the forward map
    a, b = (torch.rand(tuple(grid.shape), generator=gen, dtype=torch.float64).cuda() for _ in range(2))
    assert torch.allclose(op(2 * a - 3 * b), 2 * op(a) - 3 * op(b), rtol=1e-10, atol=1e-10)


def test_parallel_geom_traces(S):
    """ParallelGeom cannot be traced by the reference; here it equals the
    same rays given explicitly."""
    grid = S.SphericalGrid((12, 10, 16), size_r=(3, 25))
    pg = S.ParallelGeom((24, 24), (30, 4, 7), size=(50, 50))
    x = torch.rand(tuple(grid.shape), dtype=torch.float64).cuda()
    y1 = S.Operator(grid, pg)(x)
    y2 = S.Operator(grid, S.ViewGeom(pg.ray_starts, pg.rays.expand(24, 24, 3).clone()))(x)
    assert y1.shape == (24, 24) and torch.equal(y1, y2)


def test_large_config_properties(S):
    """Size-independent properties at BASELINE config 5 scale (a slab of whole detectors):
    adjoint dot-test, linearity and chord lengths through a unit volume."""
    grid = S.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
    gen = torch.Generator().manual_seed(0)
    th = torch.rand(2, generator=gen) * 2 * math.pi
    ph = (0.2 + 0.6 * torch.rand(2, generator=gen)) * math.pi
    geoms = sum(S.ConeRectGeom((512, 512), (100 * math.cos(t) * math.sin(p), 100 * math.sin(t) * math.sin(p),
                                             100 * math.cos(p)), fov=(30, 30)) for t, p in zip(th, ph))
    op = S.Operator(grid, geoms)
    ones = torch.ones(tuple(grid.shape), device="cuda")
    chord = op(ones)
    # a ray at impact parameter b crosses the hollow shell 3 <= r <= 25 over
    # 2 (sqrt(25^2-b^2) - sqrt(max(3^2-b^2, 0)))
    rays = geoms.rays.cuda(); pos = geoms.ray_starts.cuda().expand_as(rays)
    tc = -(pos * rays).sum(-1)
    b2 = ((pos * pos).sum(-1) - tc * tc).clamp_min(0)
    want = 2 * ((625 - b2).clamp_min(0).sqrt() - (9 - b2).clamp_min(0).sqrt())
    assert torch.allclose(chord.double(), want, rtol=1e-5, atol=1e-4)
    x = torch.rand(tuple(grid.shape), generator=gen).cuda()
    y = torch.rand(op.geom.shape, generator=gen).cuda()
    lhs = (op(x).double() * y.double()).sum().item()
    rhs = (x.double() * op.T(y).double()).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * abs(lhs)
    count, total = op.segment_counts()
    assert total == int(count.sum().item()) and total > 0
