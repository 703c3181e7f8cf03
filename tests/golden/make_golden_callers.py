#!/usr/bin/env python3
"""Golden outputs of the reference's own CALLERS of the operator, generated from the UNMODIFIED
reference on the CPU (build container only; the fixtures are committed):

* ``retrieval_c3.npz`` -- examples/static_retrieval.py at full size (BASELINE configs[2]: 50^3 grid,
  50 x ConeCircGeom (100, 50), two-shell volume, gd() with SquareLoss + NegRegularizer, Adam
  lr 0.1, 100 iterations): loss history, and every 97th retrieved coefficient / 211th fitted
  measurement
* ``preview3d.npz`` -- plotting.preview3d (plotting.py:210-297) for a static, a static multi-channel
  and a dynamic volume (matplotlib is stubbed: only the compute path runs)
* ``c4_full.npz`` -- examples/dynamic_measurements.py at full size (BASELINE configs[3]: 20 x 50^3
  grid, 20 x ConeCircGeom (100, 50)): every 7th ray's integral of a seeded volume and the autograd
  gradient of every 13th voxel

* ``debug_los.txt`` -- what ``Operator(..., debug=True, debug_los=(3, 4))`` prints (raytracer.py:175-225)

    python tests/golden/make_golden_callers.py [retrieval] [preview] [c4] [debug]
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

# the reference's plotting module imports matplotlib at the top; none of it is used by preview3d
for name in ("matplotlib", "matplotlib.animation", "matplotlib.pyplot", "matplotlib.ticker"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib.ticker"].EngFormatter = object
sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]

from sph_raytracer import SphericalGrid, ConeCircGeom, Operator   # noqa: E402  (the reference)


def orbit(n):
    return sum(ConeCircGeom(shape=(100, 50), pos=(5 * torch.cos(th), 5 * torch.sin(th), 1), fov=(0, 45))
               for th in torch.linspace(0, 2 * torch.pi, n))


def retrieval():
    from sph_raytracer.model import FullyDenseModel
    from sph_raytracer.retrieval import gd
    from sph_raytracer.loss import SquareLoss, NegRegularizer
    grid = SphericalGrid(shape=(50, 50, 50))
    x = torch.zeros(grid.shape)
    x[:, 25:, :25] = 1
    x[:, :25, 25:] = 1
    op = Operator(grid, orbit(50), device="cpu")
    meas = op(x)
    # ring r = 0 of the polar detectors is the ray through the origin (50 copies per observation):
    # chaotic in the reference itself, so it is masked out of the fidelity
    # term with the reference's own projection_mask -- the only departure from the example script
    mask = torch.ones(meas.shape)
    mask[:, 0, :] = 0
    fns = [1 * SquareLoss(projection_mask=mask), 1 * NegRegularizer()]
    coeffs, y_fit, losses = gd(op, meas, FullyDenseModel(grid), lr=1e-1, loss_fns=fns,
                               num_iterations=100, progress_bar=False)
    np.savez_compressed(os.path.join(HERE, "retrieval_c3.npz"),
                        hist=np.array([losses[k] for k in fns]),
                        meas=meas.detach().reshape(-1)[::211].numpy(),
                        coeffs=coeffs.detach().reshape(-1)[::97].numpy(),
                        y_fit=y_fit.detach().reshape(-1)[::211].numpy())
    print("retrieval_c3: final losses", [losses[k][-1] for k in fns])


def preview():
    from sph_raytracer.plotting import preview3d
    gen = torch.Generator().manual_seed(0)
    out = {}
    grid = SphericalGrid(shape=(10, 8, 12))
    vol = torch.rand(tuple(grid.shape), generator=gen)
    out["vol"] = vol.numpy()
    out["static"] = preview3d(vol, grid, shape=(32, 32)).numpy()
    out["static_noorbit"] = preview3d(vol, grid, shape=(16, 16), orbit=False, elev=35, azim=20).numpy()
    rgb = torch.stack([vol, 2 * vol, vol ** 2], dim=-1)
    # the reference's multi-channel branch rolls axis -2 of (r, e, a, C): the azimuth
    out["rgb"] = preview3d(rgb, grid, shape=(16, 16)).numpy()
    dgrid = SphericalGrid(shape=(3, 10, 8, 12))
    dvol = torch.rand(tuple(dgrid.shape), generator=gen)
    out["dvol"] = dvol.numpy()
    out["dynamic"] = preview3d(dvol, dgrid, shape=(16, 16)).numpy()
    np.savez_compressed(os.path.join(HERE, "preview3d.npz"), **out)
    print("preview3d:", {k: v.shape for k, v in out.items()})


def c4():
    grid = SphericalGrid(shape=(20, 50, 50, 50))
    op = Operator(grid, orbit(20), device="cpu")
    x = torch.from_numpy(np.random.default_rng(4).random(tuple(grid.shape), dtype=np.float32)).requires_grad_(True)
    y = op(x)
    w = torch.from_numpy(np.random.default_rng(104).random(tuple(y.shape), dtype=np.float32))
    w[:, 0, :] = 0        # ring r = 0 of the polar detector: the ray through the origin, chaotic (quirk 2)
    (y * w).sum().backward()
    np.savez_compressed(os.path.join(HERE, "c4_full.npz"),
                        integral=y.detach().reshape(-1)[::7].numpy(),
                        grad=x.grad.reshape(-1)[::13].numpy())
    print("c4_full:", tuple(y.shape), float(y.sum()))


def debug():
    """The reference's debug printer (raytracer.py:175-225) for one line of sight."""
    import contextlib
    import io
    from sph_raytracer import ConeRectGeom
    grid = SphericalGrid(shape=(6, 5, 8), size_r=(1, 3))
    geom = ConeRectGeom((8, 8), pos=(6.0, 1.5, 2.0), fov=(40, 40))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Operator(grid, geom, debug=True, debug_los=(3, 4))
    with open(os.path.join(HERE, "debug_los.txt"), "w") as f:
        f.write(buf.getvalue())
    print(buf.getvalue()[:600])


if __name__ == "__main__":
    todo = sys.argv[1:] or ["retrieval", "preview", "c4", "debug"]
    for name in todo:
        {"retrieval": retrieval, "preview": preview, "c4": c4, "debug": debug}[name]()
