"""The REFERENCE's own, unchanged callers of the operator running on the CUDA ``Operator``
(north_star: "model.py / loss.py / retrieval.py run unchanged"):

* ``retrieval.gd`` with the reference's ``loss`` and ``model`` classes at the full size of
  examples/static_retrieval.py (BASELINE configs[2]), loss curve against the reference's own run
* ``plotting.preview3d`` (plotting.py:210-297) with its ``Operator`` symbol pointing at the CUDA one
* examples/dynamic_measurements.py (BASELINE configs[3]) at full size against the reference's output

The reference package is the copy the bench uses (``baseline/_ref``, installed from
/root/reference, git-ignored, shipped with the tree); the golden values were produced by running
that same package on the CPU (tests/golden/make_golden_callers.py).  Nothing here reads
/root/reference.
"""
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def S():
    import sph_raytracer_b200 as S
    return S


@pytest.fixture(scope="module")
def R():
    """The unmodified reference package (matplotlib, which only its drawing code uses, stubbed)."""
    if not os.path.isdir(os.path.join(REF, "sph_raytracer")):
        pytest.skip("baseline/_ref (the installed reference) is not in this tree")
    for name in ("matplotlib", "matplotlib.animation", "matplotlib.pyplot", "matplotlib.ticker"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["matplotlib.ticker"], "EngFormatter"):
        sys.modules["matplotlib.ticker"].EngFormatter = object
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import sph_raytracer
    return sph_raytracer


def _orbit(mod, n):
    return sum(mod.ConeCircGeom(shape=(100, 50), pos=(5 * torch.cos(th), 5 * torch.sin(th), 1), fov=(0, 45))
               for th in torch.linspace(0, 2 * torch.pi, n))


@pytest.mark.parametrize("store", [False, "auto", "fused"])
def test_reference_gd_loss_and_model_run_unchanged_on_the_cuda_operator(S, R, store):
    """static_retrieval.py:18-57 at full size: 100 Adam iterations of the reference's gd() with its
    SquareLoss + NegRegularizer and FullyDenseModel; only ``Operator`` is this package's."""
    from sph_raytracer.model import FullyDenseModel
    from sph_raytracer.retrieval import gd
    from sph_raytracer.loss import SquareLoss, NegRegularizer
    z = np.load(os.path.join(GOLDEN, "retrieval_c3.npz"))
    grid = R.SphericalGrid(shape=(50, 50, 50))                # the reference's grid and detectors too
    geoms = _orbit(R, 50)
    x = torch.zeros(grid.shape, device="cuda")
    x[:, 25:, :25] = 1
    x[:, :25, 25:] = 1
    fused = store == "fused"      # this package's drop-in FusedSquareLoss inside the reference's gd()
    if fused:
        from sph_raytracer_b200.loss import FusedSquareLoss as SquareLoss
        store = False
    op = S.Operator(grid, geoms, device=x.device, store_segments=store)
    meas = op(x)
    stable = torch.ones(meas.shape, dtype=torch.bool)
    stable[:, 0, :] = False        # ring r = 0: the ray through the origin (chaotic)
    stable = stable.reshape(-1)[::211].numpy()
    got = meas.reshape(-1)[::211].cpu().numpy()
    assert np.allclose(got[stable], z["meas"][stable], rtol=1e-5, atol=1e-6), np.abs(got - z["meas"])[stable].max()
    mask = torch.ones(meas.shape, device="cuda")
    mask[:, 0, :] = 0              # (as in the golden run: the chaotic ring is masked out of the fidelity term)
    fns = [1 * SquareLoss(projection_mask=mask), 1 * NegRegularizer()]
    coeffs, y_fit, losses = gd(op, meas, FullyDenseModel(grid), lr=1e-1, loss_fns=fns,
                               num_iterations=100, progress_bar=False)
    hist = np.array([losses[k] for k in fns])
    assert hist.shape == z["hist"].shape == (2, 100)
    # float32 sums in another order, 100 Adam steps deep: the curve agrees to a few 1e-4 relative
    # (of the starting loss for the regulariser, which hovers around zero)
    scale = np.maximum(np.abs(z["hist"]), 1e-3 * np.abs(z["hist"]).max(axis=1, keepdims=True))
    assert (np.abs(hist - z["hist"]) / scale).max() < 5e-3, (np.abs(hist - z["hist"]) / scale).max()
    assert hist[0, -1] < 2e-3 * hist[0, 0]                   # and it has converged like the reference's
    c = coeffs.detach().reshape(-1)[::97].cpu().numpy()
    assert np.abs(c - z["coeffs"]).max() < 2e-3 * max(1.0, np.abs(z["coeffs"]).max())
    yf = y_fit.detach().reshape(-1)[::211].cpu().numpy()
    assert np.allclose(yf[stable], z["y_fit"][stable], rtol=2e-3, atol=2e-3 * np.abs(z["y_fit"]).max())


def test_reference_preview3d_runs_unchanged_on_the_cuda_operator(S, R, monkeypatch):
    """plotting.preview3d of the reference, its module-level ``Operator`` replaced by the CUDA one;
    and this package's own preview3d: both against the reference's images."""
    import sph_raytracer.plotting as RP
    z = np.load(os.path.join(GOLDEN, "preview3d.npz"))
    monkeypatch.setattr(RP, "Operator", lambda grid, geom: S.Operator(grid, geom, device="cuda"))
    from sph_raytracer_b200.plotting import preview3d as mine
    grid, dgrid = R.SphericalGrid(shape=(10, 8, 12)), R.SphericalGrid(shape=(3, 10, 8, 12))
    vol, dvol = torch.from_numpy(z["vol"]), torch.from_numpy(z["dvol"])
    rgb = torch.stack([vol, 2 * vol, vol ** 2], dim=-1)
    cases = [("static", vol, grid, dict(shape=(32, 32))),
             ("static_noorbit", vol, grid, dict(shape=(16, 16), orbit=False, elev=35, azim=20)),
             ("rgb", rgb, grid, dict(shape=(16, 16))),
             ("dynamic", dvol, dgrid, dict(shape=(16, 16)))]
    for name, v, g, kw in cases:
        want = z[name]
        got = RP.preview3d(v, g, **kw)                       # the reference's code, CUDA operator
        assert tuple(got.shape) == want.shape, name
        assert np.allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6), (name, "reference caller")
        sg = S.SphericalGrid(shape=tuple(g.shape))
        got = mine(v.cuda(), sg, **kw)                       # this package's preview3d
        assert tuple(got.shape) == want.shape, name
        assert np.allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-6), (name, "mirror")


@pytest.mark.parametrize("store", [False, "auto"])
def test_dynamic_measurements_at_full_size_match_the_reference(S, store):
    """dynamic_measurements.py:18-45 (BASELINE configs[3]) whole: 20 time-slices of 50^3, 20
    observations of 100 x 50 rays, observation o integrates slice o; forward and autograd adjoint."""
    z = np.load(os.path.join(GOLDEN, "c4_full.npz"))
    grid = S.SphericalGrid(shape=(20, 50, 50, 50))
    op = S.Operator(grid, _orbit(S, 20), store_segments=store)
    x = torch.from_numpy(np.random.default_rng(4).random(tuple(grid.shape), dtype=np.float32)).cuda().requires_grad_(True)
    y = op(x)
    assert tuple(y.shape) == (20, 100, 50)
    w = torch.from_numpy(np.random.default_rng(104).random(tuple(y.shape), dtype=np.float32)).cuda()
    # ring r = 0 of these polar detectors is the ray through the origin, 50 times: chaotic in the
    # reference itself.  Left out of the comparison (zero weight in the adjoint,
    # as in the golden run); every other ray and voxel to the float32 tolerance of BASELINE.json
    w[:, 0, :] = 0
    (y * w).sum().backward()
    stable = torch.ones(y.shape, dtype=torch.bool)
    stable[:, 0, :] = False
    stable = stable.reshape(-1)[::7].numpy()
    got = y.detach().reshape(-1)[::7].cpu().numpy()
    assert np.allclose(got[stable], z["integral"][stable], rtol=1e-5, atol=1e-5), \
        np.abs(got - z["integral"])[stable].max()
    g = x.grad.reshape(-1)[::13].cpu().numpy()
    assert np.allclose(g, z["grad"], rtol=1e-4, atol=1e-5 * np.abs(z["grad"]).max()), np.abs(g - z["grad"]).max()


def test_debug_printer_prints_the_reference_slot_table(S, capsys):
    """Operator(..., debug=True, debug_los=...) (raytracer.py:175-225): same header and, for every
    slot that carries a segment, the same row (typ, filled region, length, distance, boundary
    index, negative-crossing flag) as the reference printed for this line of sight."""
    want = open(os.path.join(GOLDEN, "debug_los.txt")).read().splitlines()
    grid = S.SphericalGrid(shape=(6, 5, 8), size_r=(1, 3))
    geom = S.ConeRectGeom((8, 8), pos=(6.0, 1.5, 2.0), fov=(40, 40))
    S.Operator(grid, geom, debug=True, debug_los=(3, 4))
    got = capsys.readouterr().out.splitlines()
    assert got[2:4] == want[2:4]                                   # column header and rule
    assert got[0].startswith("ray_start:") and got[1].startswith("ray:")
    assert len(got) == len(want)                                   # one row per slot (K = candidates + start)
    rows = lambda lines: [" ".join(l.split()) for l in lines[4:] if "l:0.00" not in l]
    assert rows(got) == rows(want) and len(rows(want)) > 8
    # the first slot is the start point, of no family
    assert got[4].split()[0] == "?" and want[4].split()[0] == "?"
