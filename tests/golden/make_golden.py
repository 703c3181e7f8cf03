#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed because the reference cannot travel to the GPU box.

    python tests/golden/make_golden.py

For every case it stores the inputs (grid boundary vectors, broadcast ray starts and
directions as given to Operator) and the reference's outputs:

* ``cand_t / cand_reg``  -- r_torch / e_torch / a_torch tables (known-answer cases)
* ``count / vox / seg_len`` -- Operator.regs / Operator.lens in canonical form
  (non-zero segments, consecutive identical voxels merged), ragged
* ``stable`` -- rays whose canonical sequence survives a 4e-16 relative perturbation
  of the direction (the oracle-stable set; bit-exact grading is restricted to it)
* ``integral`` -- op(density) for a seeded density, ``grad`` -- autograd adjoint
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from sph_raytracer import raytracer as RR, geometry as RG   # noqa: E402  (the reference)
from oracle import oracle as O                              # noqa: E402  (canonical() formatter only)


def density_from_seed(shape, seed, dtype=np.float32):
    """Seeded test volume; tests regenerate it with the same call (PCG64 is stable)."""
    return torch.from_numpy(np.random.default_rng(seed).random(tuple(shape), dtype=np.float32).astype(dtype))


def grid_arrays(grid):
    return dict(r_b=grid.r_b.numpy(), e_b=grid.e_b.numpy(), a_b=grid.a_b.numpy())


def canonical_ragged(grid, regs, lens):
    n = lens.reshape(-1, lens.shape[-1]).shape[0]
    K = lens.shape[-1]
    vox, seg, cnt = O.canonical(grid, regs.reshape(3, n, K).numpy(), lens.reshape(n, K).numpy())
    keep = np.arange(vox.shape[1])[None, :] < cnt[:, None]
    return cnt, vox[keep].astype(np.int16), seg[keep]


def operator_case(name, grid, starts, rays, density=None, want_grad=False, per_obs=False, seed=0):
    """Trace with the reference and store canonical sequences, stability mask, integrals."""
    xs, d, shape = O.broadcast_rays(starts, rays)
    geom = RG.ViewGeom(torch.from_numpy(xs.copy()).reshape(*shape, 3),
                       torch.from_numpy(d.copy()).reshape(*shape, 3))
    d_in = geom.rays.clone().reshape(-1, 3).numpy()     # what Operator receives (ViewGeom renormalised)
    op = RR.Operator(grid, RG.ViewGeom(geom.ray_starts.clone(), geom.rays.clone()))
    cnt, vox, seg = canonical_ragged(grid, op.regs, op.lens)

    # oracle-stable set: perturb the directions by 4e-16 (relative) and trace again
    g = torch.Generator().manual_seed(1234)
    pert = geom.rays * (1 + 4e-16 * torch.randn(geom.rays.shape, generator=g, dtype=torch.float64)) \
        + 4e-16 * torch.randn(geom.rays.shape, generator=g, dtype=torch.float64)
    op2 = RR.Operator(grid, RG.ViewGeom(geom.ray_starts.clone(), pert))
    cnt2, vox2, seg2 = canonical_ragged(grid, op2.regs, op2.lens)
    off = np.concatenate(([0], np.cumsum(cnt))); off2 = np.concatenate(([0], np.cumsum(cnt2)))
    stable = np.zeros(len(cnt), dtype=bool)
    for i in range(len(cnt)):
        if cnt[i] == cnt2[i]:
            a = vox[off[i]:off[i + 1]]; b = vox2[off2[i]:off2[i + 1]]
            la = seg[off[i]:off[i + 1]]; lb = seg2[off2[i]:off2[i + 1]]
            stable[i] = np.array_equal(a, b) and np.allclose(la, lb, rtol=1e-6, atol=1e-9)

    out = dict(xs=xs, dirs=d_in, ray_shape=np.array(shape), count=cnt, vox=vox, seg_len=seg,
               stable=stable, **grid_arrays(grid))
    if density is not None:
        density = density.clone().requires_grad_(want_grad)
        # dynamic grid: a ViewGeom of shape (T, ...) reads time-slice = leading index
        y = op(density)
        # densities are formulas or seeded (see density_from_seed); not stored
        out["integral"] = y.detach().numpy()
        if want_grad:
            # adjoint input: seeded weights, zero on the oracle-unstable rays so that the
            # stored gradient only contains contributions that are graded
            w = density_from_seed(y.shape, 100 + seed, np.dtype(str(y.dtype).replace("torch.", "")))
            w = w * torch.from_numpy(stable.reshape(shape)).to(w.dtype)
            (y * w).sum().backward()
            out["grad"] = density.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name:24s} rays={len(cnt):6d} segs={int(cnt.sum()):8d} unstable={int((~stable).sum()):4d} "
          f"max_seg={int(cnt.max())}")


def known_answer_cases():
    """Inputs of the reference's own unit tests (test_all.py:18-173, 225-234) with the
    outputs the reference produces for them here."""
    pi = math.pi
    cases = []

    def run(kind, b, xs, rays):
        fn = dict(r=RR.r_torch, e=RR.e_torch, a=RR.a_torch)[kind]
        b = torch.as_tensor(b, dtype=torch.float64)
        xs_t = torch.tensor(xs, dtype=torch.float64).reshape(-1, 3)
        r_t = torch.tensor(rays, dtype=torch.float64).reshape(-1, 3)
        d_in = r_t.clone().numpy()
        t, reg = fn(b, xs_t, r_t)[:2]
        cases.append(dict(kind=kind, b=b.numpy(), xs=xs_t.numpy(), dirs=d_in,
                          t=t.numpy(), reg=reg.numpy()))

    # test_r
    for rays in ([(1, 0, 0)], [(-1, 0, 0)], [(0, 0, 1)]):
        run("r", (0.1, 1, 2), [(-3, 0, 0)], rays)
    run("r", [2], [(-3, 2, 0), (-3, -2, 0), (-3, -2, 0)], [(1, 0, 0), (1, 0, 0), (-1, 0, 0)])
    run("r", [0], [(-3, 0, 0)], [(1, 0, 0)])
    # test_e
    e = [pi / 6, pi / 4]
    run("e", e, [(-1, 0, 0)], [(0, 0, 1)])
    run("e", e, [(-100, 0, 1)], [(1, 0, 0)])
    run("e", [pi - pi / 6, pi - pi / 4], [(-100, 0, -1)], [(1, 0, 0)])
    run("e", e, [(-1, 0, -1)], [(1, 0, 0)])
    run("e", [pi / 4], [(0, 0, 1)], [(1, 0, 1)])
    run("e", [pi / 4], [(-1, 0, 1)], [(1, 0, -1)])
    run("e", [pi / 4], [(1, 1, 1)], [(0, -1, 0)])
    run("e", e, [(-1, 0, 0)], [(1, 0, 0)])
    run("e", [0, pi], [(-1, 0, 0)], [(1, 0, 0)])
    # test_a
    a_b = [pi / 4, pi / 2]
    run("a", a_b, [(-1, 1, 0)], [(1, 0, 0)])
    run("a", a_b, [(-1, 1, 0)], [(-1, 0, 0)])
    run("a", a_b, [(-1, -1, 0)], [(1, 0, 0)])
    run("a", [0], [(0, 1, 0)], [(1, 0, 0)])
    run("a", a_b, [(-1, 0, 0)], [(1, 0, 0)])
    run("a", [pi / 2], [(0, 1, 0)], [(0, -1, 0)])

    out = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            out[f"c{i:02d}_{k}"] = np.asarray(v)
    out["n_cases"] = np.array(len(cases))
    # test_find_starts
    fs = []
    for shape, x in (((5, 5, 1), [0, 0, 100]), ((5, 5, 1), [0, 0, -100]), ((5, 5, 5), [100, 0, 0])):
        g = RG.SphericalGrid(shape=shape)
        fs.append(np.concatenate((np.array(shape), np.array(x, dtype=float),
                                  RR.find_starts(g, x).numpy().astype(float))))
    out["find_starts"] = np.array(fs)
    np.savez_compressed(os.path.join(HERE, "known_answers.npz"), **out)
    print(f"known_answers            cases={len(cases)} + {len(fs)} find_starts")


def main():
    torch.manual_seed(0)
    known_answer_cases()

    # --- reference integration tests (test_raytracer.py:8-60, 85-116) ---
    u = 0.001
    starts = [[-100, u, u], [u, -100, u], [u, u, -100], [-100, 0, u], [0, -100, u], [0, u, -100],
              [-100, u, 0], [u, -100, 0], [u, 0, -100], [5, 0, 0]]
    rays = [[1, 0, 0], [0, 1, 0], [0, 0, 1]] * 3 + \
        [[-0.99998629093170166016, 0.00413372274488210678, 0.00321511807851493359]]
    grids = [RG.SphericalGrid(shape=(50, 50, 50), size_r=(3, 25)), RG.SphericalGrid(shape=(4, 4, 4)),
             RG.SphericalGrid(shape=(1, 4, 4)), RG.SphericalGrid(shape=(4, 1, 4)),
             RG.SphericalGrid(shape=(4, 4, 1))]
    for i, grid in enumerate(grids):
        operator_case(f"chords_{i}", grid, starts, rays, density=torch.ones(grid.shape))
    grid = RG.SphericalGrid(shape=(1, 2, 1), size_r=(0, 25))
    operator_case("hemisphere", grid, [[-200, u, u]], [[1, 0, 0]],
                  density=torch.tensor([[[1.0], [0]]]))

    # --- C1 single_vantage.py: every 4th pixel of the 256x256 detector ---
    grid = RG.SphericalGrid(shape=(50, 50, 50))
    geom = RG.ConeRectGeom((256, 256), (5, 0, 0), fov=(45, 45))
    x = torch.zeros(grid.shape); x[-1] += 1; x[-10] += 1
    operator_case("c1_single_vantage", grid, geom.ray_starts, geom.rays[::4, ::4], density=x)
    dens = density_from_seed(grid.shape, 0)
    operator_case("c1_rand", grid, geom.ray_starts, geom.rays[1::8, 2::8], density=dens, want_grad=True)

    # --- C2 static_retrieval.py geometry: 3 of the 50 observations, every 2nd radial pixel ---
    thetas = torch.linspace(0, 2 * torch.pi, 50)[[0, 7, 31]]
    geoms = [RG.ConeCircGeom((100, 50), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(0, 45)) for t in thetas]
    st = torch.stack([g.ray_starts.expand(50, 50, 3) for g in geoms])
    ry = torch.stack([g.rays[::2] for g in geoms])
    x = torch.zeros(grid.shape); x[:, 25:, :25] = 1; x[:, :25, 25:] = 1
    operator_case("c2_static_pattern", grid, st, ry, density=x)
    operator_case("c2_static_rand", grid, st, ry, density=dens, want_grad=True)
    # float64 density, as retrieval.gd uses (retrieval.py:64-70)
    operator_case("c2_static_rand_f64", grid, st[:1, ::5], ry[:1, ::5], density=density_from_seed(grid.shape, 0, np.float64), want_grad=True)
    # multi-channel static volume (plotting.preview3d call pattern)
    operator_case("c2_multichannel", grid, st[:1, ::5, ::5], ry[:1, ::5, ::5],
                  density=density_from_seed((3,) + tuple(grid.shape), 3),
                  want_grad=True)

    # --- starts that lie exactly ON a half-plane: pos (5, 0, 1) with an even azimuth count has
    #     a_b[N/2] == 0.0, so the crossing of that plane ties with the start slot at t = 0 ---
    g16 = RG.SphericalGrid(shape=(16, 16, 16))
    geoms16 = [RG.ConeCircGeom((24, 16), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(5, 45))
               for t in torch.linspace(0, 2 * torch.pi, 7)[:6]]
    operator_case("onplane_16", g16, torch.stack([g.ray_starts.expand(24, 16, 3) for g in geoms16]),
                  torch.stack([g.rays for g in geoms16]), density=density_from_seed(g16.shape, 9),
                  want_grad=True)

    # --- C4 dynamic_measurements.py: 4 time steps, coarse detector ---
    gdyn = RG.SphericalGrid(shape=(4, 50, 50, 50))
    thetas = torch.linspace(0, 2 * torch.pi, 20)[[0, 5, 11, 17]]
    geoms = [RG.ConeCircGeom((100, 50), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(0, 45)) for t in thetas]
    st = torch.stack([g.ray_starts.expand(20, 10, 3) for g in geoms])
    ry = torch.stack([g.rays[::5, ::5] for g in geoms])
    x = density_from_seed(gdyn.shape, 4)
    operator_case("c4_dynamic", gdyn, st, ry, density=x, want_grad=True, per_obs=True)

    # --- C5 large sweep tile: (200,180,360) grid, one 512x512 detector, every 32nd pixel ---
    g5 = RG.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
    geom = RG.ConeRectGeom((512, 512), (95.5, 29.6, 20.0), fov=(30, 30))
    operator_case("c5_tile", g5, geom.ray_starts, geom.rays[7::32, 5::32])

    # --- random rays: starts inside and outside, partial / log / hollow grids ---
    gen = torch.Generator().manual_seed(7)
    n = 1500
    xs = torch.randn(n, 3, generator=gen, dtype=torch.float64) * 2
    d = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    gp = RG.SphericalGrid((13, 9, 11), size_r=(0.5, 3), size_e=(0.3, 2.9), size_a=(-2.5, 2.9))
    operator_case("rand_partial", gp, xs, d,
                  density=density_from_seed(gp.shape, 5), want_grad=True)
    gl = RG.SphericalGrid((7, 6, 5), spacing="log", size_r=(0.1, 4))
    operator_case("rand_log", gl, xs, d, density=density_from_seed(gl.shape, 6), want_grad=True)
    gf = RG.SphericalGrid((20, 16, 24))
    operator_case("rand_full_inside", gf, xs * 0.3, d,
                  density=density_from_seed(gf.shape, 7), want_grad=True)
    # parallel beam through a hollow grid (per-pixel starts, shared direction)
    pg = RG.ParallelGeom((24, 24), (30, 4, 7), size=(50, 50))
    gh = RG.SphericalGrid((12, 10, 16), size_r=(3, 25))
    operator_case("parallel_hollow", gh, pg.ray_starts, pg.rays.expand(24, 24, 3),
                  density=density_from_seed(gh.shape, 8), want_grad=True)




# ---- the retrieval loop that calls forward + adjoint (reference retrieval.gd) ----
class DenseOp:
    """Deterministic dense stand-in for an operator (CPU tests of the retrieval loop)."""

    def __init__(self, grid, n_meas=40, seed=11):
        self.grid, self.device = grid, "cpu"
        n = int(np.prod(tuple(grid.shape)))
        self.A = torch.from_numpy(np.random.default_rng(seed).random((n_meas, n)))

    def __call__(self, d):
        return self.A @ d.reshape(-1).to(self.A.dtype)


def retrieval_cases():
    from sph_raytracer import retrieval as RT, loss as RL, model as RM
    out = {}
    # (1) loop logic on a dense operator: loss history must be reproduced exactly on the CPU
    grid = RG.SphericalGrid(shape=(4, 3, 5))
    f = DenseOp(grid)
    truth = density_from_seed(grid.shape, 21, np.float64)
    y = f(truth)
    fns = [1 * RL.SquareLoss(), 0.5 * RL.NegRegularizer(), RL.CheaterLoss(truth, use_grad=False)]
    coeffs, y_fit, losses = RT.gd(f, y, RM.FullyDenseModel(grid), lr=1e-1, loss_fns=fns,
                                  num_iterations=40, progress_bar=False)
    out["dense_hist"] = np.array([losses[k] for k in fns])
    out["dense_coeffs"] = coeffs.detach().numpy()
    # fixed test volumes of model.py
    gm = RG.SphericalGrid(shape=(12, 10, 8))
    out["cubes"] = RM.CubesModel(gm).volume.numpy()
    out["axes"] = RM.AxisAlignmentModel(gm).volume.numpy()

    # (2) static_retrieval.py scaled down, with the reference operator: 6 observations of a
    #     (24,16) circular detector, grid 16^3, 30 Adam steps, float64 coefficients
    grid = RG.SphericalGrid(shape=(16, 16, 16))
    x = torch.zeros(grid.shape)
    x[:, 8:, :8] = 1
    x[:, :8, 8:] = 1
    geoms = sum(RG.ConeCircGeom((24, 16), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(5, 45))
                for t in torch.linspace(0, 2 * torch.pi, 7)[:6])
    op = RR.Operator(grid, geoms)
    meas = op(x)
    fns = [1 * RL.SquareLoss(), 1 * RL.NegRegularizer()]
    coeffs, y_fit, losses = RT.gd(op, meas.clone(), RM.FullyDenseModel(grid), lr=1e-1, loss_fns=fns,
                                  num_iterations=30, progress_bar=False)
    out["ret_meas"] = meas.detach().numpy()
    out["ret_hist"] = np.array([losses[k] for k in fns])
    out["ret_coeffs"] = coeffs.detach().numpy()
    out["ret_yfit"] = y_fit.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "retrieval.npz"), **out)
    print("retrieval                dense loss", out["dense_hist"][0][[0, -1]],
          " operator loss", out["ret_hist"][0][[0, -1]])


if __name__ == "__main__":
    if "--retrieval-only" not in sys.argv:
        main()
    retrieval_cases()
