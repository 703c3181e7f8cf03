"""CPU tests of the host side: geometry mirror, C-ABI surface, batching plan, sharding.
No kernel is launched here."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

import sph_raytracer_b200 as S
from sph_raytracer_b200 import _cabi, raytracer as R, parallel as P

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_cabi_exports_every_declared_symbol():
    """libsphrt.so loads and exports exactly the functions include/sphrt.h declares."""
    from sph_raytracer_b200 import build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "sphrt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sphrt_[a-z_]+)\s*\(", hdr))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = ctypes.CDLL(build.LIB)
    for name in declared:
        assert hasattr(lib, name), name
    lib = _cabi.load()
    assert lib.sphrt_version() == 7
    assert lib.sphrt_error_string(-3).decode().startswith("scratch")
    # argument validation happens before any CUDA call
    g = _cabi.Grid(0, 1, 1, 0, *([0] * 7))
    r = _cabi.Rays(0, 0, 1, 0, 1, None, 0)
    assert lib.sphrt_forward(g, r, 0, 0, 1, 0, 0, 0, 0, 0) == -1
    with pytest.raises(_cabi.SphrtError):
        _cabi.check(-1)


def test_operator_requires_cuda_and_rejects_unsupported_modes():
    grid = S.SphericalGrid((4, 4, 4))
    geom = S.ViewGeom([-100, 0, 0], [1, 0, 0])
    with pytest.raises(NotImplementedError):
        S.Operator(grid, geom, ftype=torch.float32)
    with pytest.raises(NotImplementedError):
        S.Operator(grid, geom, invalid=True)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):          # no CPU fallback
            S.Operator(grid, geom)
    op = S.Operator(grid, geom, _compute=False)
    assert repr(op) == f"Operator({grid.shape} → {geom.shape})"
    assert op.grid is grid and op.geom is geom and op.dynamic is False


def test_compact_ray_layout():
    """Shared pinholes / directions are stored once per group of rays."""
    g = S.ConeRectGeom((6, 5), (5, 0, 0))
    s, per = R._compact(g.ray_starts, (6, 5))
    assert s.shape == (1, 3) and per == 30
    coll = sum(S.ConeRectGeom((6, 5), (5 * math.cos(t), 5 * math.sin(t), 1)) for t in (0.1, 0.2, 0.3))
    s, per = R._compact(coll.ray_starts, (3, 6, 5))
    assert s.shape == (3, 3) and per == 30
    d, per = R._compact(coll.rays, (3, 6, 5))
    assert d.shape == (90, 3) and per == 1
    pg = S.ParallelGeom((4, 4), (30, 4, 7), size=(2, 2))
    d, per = R._compact(pg.rays, (4, 4))
    assert d.shape == (1, 3) and per == 16
    # a start that varies along a middle axis only cannot be compacted: materialised
    odd = torch.zeros(1, 6, 1, 3, dtype=torch.float64)
    s, per = R._compact(odd, (3, 6, 5))
    assert s.shape == (90, 3) and per == 1


class _FakeOp(R.Operator):
    """Operator with the launch plumbing stubbed out, to exercise the batching plan."""

    def __init__(self, grid, ray_shape):
        super().__init__(grid, None, _compute=False)
        self._tables = type("T", (), dict(nr=grid.shape[-3], ne=grid.shape[-2], na=grid.shape[-1]))()
        self._rays = type("Rr", (), dict(shape=tuple(ray_shape), n=math.prod(ray_shape)))()


def test_batching_plan_matches_reference_rules():
    """Output shapes / time-slice mapping of Operator.__call__ (reference raytracer.py:703-713,
    SURVEY 3.2)."""
    st = S.SphericalGrid((2, 3, 4))
    op = _FakeOp(st, (64, 64))
    assert op._plan((2, 3, 4)) == (1, 0, (64, 64), None)
    assert op._plan((10, 2, 3, 4)) == (10, 0, (10, 64, 64), None)
    assert op._plan((5, 7, 2, 3, 4)) == (35, 0, (5, 7, 64, 64), None)
    with pytest.raises(IndexError):
        op._plan((2, 3, 5))
    dyn = S.SphericalGrid((10, 2, 3, 4))
    assert _FakeOp(dyn, (64, 64))._plan((10, 2, 3, 4)) == (10, 0, (10, 64, 64), None)
    assert _FakeOp(dyn, (10, 8, 8))._plan((10, 2, 3, 4)) == (10, 64, (10, 8, 8), None)
    assert _FakeOp(dyn, (1, 8, 8))._plan((10, 2, 3, 4)) == (10, 0, (10, 8, 8), 0)
    assert _FakeOp(dyn, (7,))._plan((10, 2, 3, 4)) == (10, 0, (10, 1, 7), None)
    with pytest.raises(IndexError):
        _FakeOp(dyn, (9, 8, 8))._plan((10, 2, 3, 4))


def test_shard_bounds_cover_everything_once():
    for n, w in [(1000, 8), (50, 4), (7, 3), (3, 3), (20, 1)]:
        spans = [P.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_bounds(10, 4, 4)


# ---- world_size-2 gloo test of the sharded operator (host logic of the multi-GPU path) --------
class _DenseOp:
    """CPU stand-in with the Operator interface: y = A x for a fixed random matrix per ray."""

    def __init__(self, grid, geom, **kw):
        self.grid, self.geom, self.device = grid, geom, "cpu"
        g = torch.Generator().manual_seed(1234)
        n_all = 4 * 6                                  # 4 observations of 6 rays in the test
        A = torch.rand(n_all, math.prod(grid.shape), generator=g, dtype=torch.float64)
        self.A = torch.cat([A[o.tag * 6:(o.tag + 1) * 6] for o in geom.geoms])

    def __call__(self, density):
        return (self.A @ density.reshape(-1)).reshape(len(self.geom.geoms), 6)

    def T(self, y):
        return (self.A.T @ y.reshape(-1)).reshape(tuple(self.grid.shape))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = S.SphericalGrid((2, 3, 4))
    geoms = []
    for k in range(4):
        g = S.ViewGeom(torch.zeros(6, 3), torch.ones(6, 3)); g.tag = k
        geoms.append(g)
    coll = S.ViewGeomCollection(*geoms)
    sop = P.ShardedOperator(grid, coll, _DenseOp)
    x = torch.arange(24, dtype=torch.float64).reshape(2, 3, 4).requires_grad_(True)
    y = sop(x)
    w = torch.ones_like(y) * (rank + 1)
    (y * w).sum().backward()
    # fused normal step through the sharded operator (dense stand-in: the two-pass form) and the
    # drop-in loss that uses it: same value and GLOBAL gradient as SquareLoss
    from sph_raytracer_b200 import loss as L
    b = torch.full_like(y.detach(), 0.25)
    yn, gn = sop.normal(x.detach(), b=b, scale=0.5)
    vals = []
    for cls in (L.SquareLoss, L.FusedSquareLoss):
        xx = x.detach().clone().requires_grad_(True)
        v = cls()(sop, b, xx, None)
        v.backward()
        vals.append((float(v), xx.grad.numpy()))
    q.put((rank, sop.lo, sop.hi, y.detach().numpy(), x.grad.numpy(), sop.T(w).numpy(),
           yn.numpy(), gn.numpy(), vals))
    dist.destroy_process_group()


def test_sharded_operator_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(1234)
    A = torch.rand(24, 24, generator=g, dtype=torch.float64)
    x = torch.arange(24, dtype=torch.float64)
    full = (A @ x).reshape(4, 6).numpy()
    w_full = torch.cat([torch.ones(12) * 1, torch.ones(12) * 2]).double()
    grad = (A.T @ w_full).reshape(2, 3, 4).numpy()
    assert [(r[1], r[2]) for r in res] == [(0, 2), (2, 4)]
    gnorm = (A.T @ (0.5 * (A @ x - 0.25))).reshape(2, 3, 4).numpy()
    for rank, lo, hi, y, gx, ty, yn, gn, vals in res:
        assert np.allclose(y, full[lo:hi])
        assert np.allclose(gx, grad)       # gradient is global (all-reduced) on every rank
        assert np.allclose(ty, grad)
        assert np.allclose(yn, full[lo:hi]) and np.allclose(gn, gnorm)
        assert np.isclose(vals[0][0], vals[1][0]) and np.allclose(vals[0][1], vals[1][1])


class _DynOp:
    """CPU stand-in for a dynamic operator: observation o integrates time-slice o with a fixed
    random matrix per ray."""

    def __init__(self, grid, geom, **kw):
        self.grid, self.geom, self.device = grid, geom, "cpu"
        assert grid.shape[0] == len(geom.geoms)            # local time axis = local observations
        g = torch.Generator().manual_seed(77)
        A = torch.rand(4, 6, math.prod(grid.shape[1:]), generator=g, dtype=torch.float64)
        self.A = torch.stack([A[o.tag] for o in geom.geoms])

    def __call__(self, density):
        return torch.einsum("orv,ov->or", self.A, density.reshape(len(self.geom.geoms), -1))

    def T(self, y):
        return torch.einsum("orv,or->ov", self.A, y).reshape(tuple(self.grid.shape))


def _dyn_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = S.SphericalGrid((4, 2, 3, 4))                   # four time-slices
    geoms = []
    for k in range(4):
        g = S.ViewGeom(torch.zeros(6, 3), torch.ones(6, 3)); g.tag = k
        geoms.append(g)
    sop = P.ShardedOperator(grid, S.ViewGeomCollection(*geoms), _DynOp)
    x = (torch.arange(96, dtype=torch.float64) / 7).reshape(4, 2, 3, 4).requires_grad_(True)
    y = sop(x)
    w = torch.ones_like(y) * (rank + 1)
    (y * w).sum().backward()
    q.put((rank, sop.lo, sop.hi, y.detach().numpy(), x.grad.numpy(), sop.T(w).numpy()))
    dist.destroy_process_group()


def test_sharded_dynamic_operator_gloo_world2():
    """Dynamic volume: the time axis is sharded with the observations; the all-reduce assembles the
    ranks' disjoint gradient slices."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dyn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(77)
    A = torch.rand(4, 6, 24, generator=g, dtype=torch.float64)
    x = (torch.arange(96, dtype=torch.float64) / 7).reshape(4, 24)
    full = torch.einsum("orv,ov->or", A, x).numpy()
    w_full = torch.cat([torch.ones(2, 6), 2 * torch.ones(2, 6)]).double()
    grad = torch.einsum("orv,or->ov", A, w_full).reshape(4, 2, 3, 4).numpy()
    for rank, lo, hi, y, gx, ty in res:
        assert np.allclose(y, full[lo:hi])
        assert np.allclose(gx, grad) and np.allclose(ty, grad)


def test_sharded_dynamic_needs_one_slice_per_observation():
    """A dynamic volume is sharded along time with the observations: the counts must agree, and the
    local grid keeps the spatial boundaries and this rank's time samples."""
    grid = S.SphericalGrid((4, 2, 3, 4))
    geoms = S.ViewGeomCollection(*[S.ViewGeom(torch.zeros(2, 3), torch.ones(2, 3)) for _ in range(3)])
    with pytest.raises(IndexError):
        P.ShardedOperator(grid, geoms, _DynOp, rank=0, world=1)
    local = P._time_slices(grid, 1, 3)
    assert tuple(local.shape) == (2, 2, 3, 4) and local.dynamic
    assert local.shape.t == 2 and torch.equal(local.t, grid.t[1:3])
    assert local.r_b is grid.r_b and tuple(grid.shape) == (4, 2, 3, 4)     # the original is untouched


def test_reference_memory_estimate_known_answers():
    """bench.py's figure for what the reference stores: 2.048 GB of regs / lens for configs[1]
    (250 000 rays x 256 slots x 32 B, measured in BASELINE.md) and 3.06 GB from the reference's own
    estimator (examples/memory_usage.py:23-44 at the README's parameters)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    m = bench.reference_memory_gb(S.SphericalGrid((50, 50, 50)), 50 * 50 * 100)
    assert m["slots_per_ray"] == 256
    assert m["stored_operator"] == pytest.approx(2.048)
    assert m["estimator"] == pytest.approx(3.06)


def test_start_cells_match_find_starts():
    """The host-side binning of start points (what sphrt_rays.start_cells carries) is the
    reference's find_starts: known answers of the reference's own tests, the C oracle on random
    points, and the on-boundary flag."""
    import numpy as np
    from oracle import oracle as O
    z = np.load(os.path.join(ROOT, "tests", "golden", "known_answers.npz"))
    for row in z["find_starts"]:
        shape, x, want = row[:3].astype(int), row[3:6], row[6:9].astype(int)
        grid = S.SphericalGrid(shape=tuple(shape))
        cells = R._start_cells(grid.r_b.double(), grid.e_b.double(), grid.a_b.double(),
                               torch.tensor(x, dtype=torch.float64)[None])
        assert cells[0, 3:6].tolist() == want.tolist()
    grid = S.SphericalGrid((7, 9, 11), size_r=(0.5, 3), size_e=(0.3, 2.9), size_a=(-2.5, 2.9))
    pts = torch.randn(5000, 3, generator=torch.Generator().manual_seed(0), dtype=torch.float64) * 2
    cells = R._start_cells(grid.r_b.double(), grid.e_b.double(), grid.a_b.double(), pts).numpy()
    og = O.OracleGrid(grid)
    want = np.asarray(O.find_starts(og, pts.numpy())).reshape(3, -1)
    assert (cells[:, 3:6].T == want).all()
    assert not cells[:, 6].any()                       # random points are on no boundary
    # geometric bins keep the "beyond the last boundary" bin that the states map to -1
    assert ((cells[:, 0:3] == cells[:, 3:6]) | (cells[:, 3:6] == -1)).all()
    full = S.SphericalGrid((4, 4, 8))
    on = R._start_cells(full.r_b.double(), full.e_b.double(), full.a_b.double(),
                        torch.tensor([[5.0, 0.0, 0.0], [3.0, 3.0, 1.0], [1.0, 0.2, 0.3]], dtype=torch.float64))
    assert on[:, 6].tolist() == [1, 1, 0]              # (5,0,0): equator cone and plane a = 0; (3,3,1): a = pi/4


def test_canonical_from_events():
    import numpy as np
    from helpers import canonical_from_events
    vox = np.array([[[1, 1, 1], [1, 1, 1], [1, 2, 1], [-1, 2, 1], [0, 0, 0]]], dtype=np.int32)
    lens = np.array([[0.5, 0.25, 0.0, 0.0, 9.0]])
    c, v, l = canonical_from_events(vox, lens, np.array([4]), 4)
    assert c.tolist() == [1] and v[0, 0].tolist() == [1, 1, 1] and l[0, 0] == 0.75
    c, v, l = canonical_from_events(vox, lens, np.array([-1]), 4)
    assert c.tolist() == [-1]


def test_find_starts_runs_on_the_host_and_matches_the_reference_known_answers():
    """find_starts (reference raytracer.py:605-644) takes the host torch path of _start_cells (the
    last bit of atan2 decides the bin of a point on a cone / half-plane, and CUDA's differs)."""
    import sph_raytracer_b200 as S
    z = np.load(os.path.join(ROOT, "tests", "golden", "known_answers.npz"))
    for row in z["find_starts"]:
        shape, x, want = row[:3].astype(int), row[3:6], row[6:9].astype(int)
        got = S.find_starts(S.SphericalGrid(shape=tuple(shape)), torch.from_numpy(x), device="cpu")
        assert got.dtype == torch.int64 and got.tolist() == want.tolist()
    # batched points keep their leading shape, like the reference
    pts = torch.randn(4, 5, 3, dtype=torch.float64)
    out = S.find_starts(S.SphericalGrid(shape=(6, 7, 8)), pts, device="cpu")
    assert tuple(out.shape) == (3, 4, 5)


def test_cubes_model_boxes_equal_the_reference_float32_arithmetic():
    """CubesModel box edges: the reference truncates n * float32 fraction (model.py:58-66); on the
    config-5 grid (ne = 180) float64 arithmetic would give 125 instead of 126."""
    import sph_raytracer_b200 as S
    from sph_raytracer_b200.model import CubesModel
    m = CubesModel(S.SphericalGrid(shape=(200, 180, 360)))
    assert (m.r0, m.r1, m.e00, m.e01, m.e10, m.e11, m.a0, m.a1) == (66, 133, 36, 54, 126, 162, 144, 216)
    m = CubesModel(S.SphericalGrid(shape=(90, 170, 330)))
    assert (m.e10, m.e11, m.a0, m.a1) == (119, 153, 132, 198)


# ---- fewer observations than ranks: detector rows are sharded ------------
def test_shard_detector_rows_covers_every_row_once():
    g1 = S.ConeRectGeom((5, 4), (5.0, 0.0, 0.0), fov=(45, 45))
    g2 = S.ConeRectGeom((3, 4), (0.0, 5.0, 1.0), fov=(30, 30))
    coll = S.ViewGeomCollection.__new__(S.ViewGeomCollection)       # ragged heights: a bare list
    coll.geoms = [g1, g2]
    want_d = torch.cat([g1.rays, g2.rays])
    want_x = torch.cat([g1.ray_starts.broadcast_to(g1.rays.shape), g2.ray_starts.broadcast_to(g2.rays.shape)])
    for world in (1, 2, 3, 8):
        got_d, got_x, blocks = [], [], []
        for rank in range(world):
            geom, (lo, hi) = P.shard_detector_rows(coll, rank, world)
            assert tuple(geom.rays.shape) == (hi - lo, 4, 3) == tuple(geom.ray_starts.shape)
            got_d.append(geom.rays); got_x.append(geom.ray_starts); blocks.append((lo, hi))
        assert blocks[0][0] == 0 and blocks[-1][1] == 8
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert torch.equal(torch.cat(got_d), want_d) and torch.equal(torch.cat(got_x), want_x)
    with pytest.raises(ValueError):
        P.shard_detector_rows(g2, 3, 4)                              # 3 rows, 4 ranks: rank 3 owns nothing


class _RowOp:
    """CPU stand-in: y[row, pixel] = <direction of the pixel, fixed vector> * sum(x) -- enough to see
    that every rank traces its own rows and that the gradient is all-reduced."""

    def __init__(self, grid, geom, **kw):
        self.grid, self.geom, self.device = grid, geom, "cpu"
        self.w = (torch.asarray(geom.rays) * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)).sum(-1)

    def __call__(self, density):
        return self.w * density.sum()

    def T(self, y):
        return torch.full(tuple(self.grid.shape), float((self.w * y).sum()), dtype=torch.float64)


def _row_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = S.SphericalGrid((2, 3, 4))
    geom = S.ConeRectGeom((6, 5), (5.0, 0.0, 0.0), fov=(45, 45))      # ONE observation, two ranks
    sop = P.ShardedOperator(grid, geom, _RowOp)
    x = torch.ones(2, 3, 4, dtype=torch.float64, requires_grad=True)
    y = sop(x)
    y.sum().backward()
    q.put((rank, sop.by_rows, sop.lo, sop.hi, y.detach().numpy(), x.grad.numpy()))
    dist.destroy_process_group()


def test_sharded_operator_shards_detector_rows_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_row_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    geom = S.ConeRectGeom((6, 5), (5.0, 0.0, 0.0), fov=(45, 45))
    w = (geom.rays * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)).sum(-1)
    assert [(r[1], r[2], r[3]) for r in res] == [(True, 0, 3), (True, 3, 6)]
    for rank, by_rows, lo, hi, y, gx in res:
        assert np.allclose(y, (w[lo:hi] * 24).numpy())
        assert np.allclose(gx, float(w.sum()))                      # global gradient on every rank
