"""the retrieval loop (gd + losses + models), the caller of forward + adjoint.
Golden loss histories come from the reference's own gd (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import sph_raytracer_b200 as S
from sph_raytracer_b200 import loss as L, model as M, retrieval as RT
from helpers import GOLDEN, density_from_seed


class DenseOp:
    def __init__(self, grid, n_meas=40, seed=11):
        self.grid, self.device = grid, "cpu"
        n = int(np.prod(tuple(grid.shape)))
        self.A = torch.from_numpy(np.random.default_rng(seed).random((n_meas, n)))

    def __call__(self, d):
        return self.A @ d.reshape(-1).to(self.A.dtype)


def test_models_match_reference_volumes():
    z = np.load(f"{GOLDEN}/retrieval.npz")
    g = S.SphericalGrid(shape=(12, 10, 8))
    assert np.array_equal(M.CubesModel(g).volume.numpy(), z["cubes"])
    assert np.array_equal(M.AxisAlignmentModel(g).volume.numpy(), z["axes"])
    assert M.FullyDenseModel(g).coeffs_shape == g.shape and M.CubesModel(g).coeffs_shape == ()
    assert repr(M.FullyDenseModel(g)) == "FullyDenseModel((12, 10, 8))"
    assert repr(3 * L.SquareLoss()) == "3e+00 * SquareLoss"


def test_gd_reproduces_reference_history_on_dense_operator():
    """Same optimiser steps as the reference's gd: 40 Adam iterations, three loss terms."""
    z = np.load(f"{GOLDEN}/retrieval.npz")
    grid = S.SphericalGrid(shape=(4, 3, 5))
    f = DenseOp(grid)
    truth = torch.from_numpy(density_from_seed(grid.shape, 21, np.float64))
    y = f(truth)
    fns = [1 * L.SquareLoss(), 0.5 * L.NegRegularizer(), L.CheaterLoss(truth, use_grad=False)]
    coeffs, y_fit, losses = RT.gd(f, y, M.FullyDenseModel(grid), lr=1e-1, loss_fns=fns,
                                  num_iterations=40, progress_bar=False)
    hist = np.array([losses[k] for k in fns])
    assert np.allclose(hist, z["dense_hist"], rtol=1e-12, atol=0)
    assert np.allclose(coeffs.detach().numpy(), z["dense_coeffs"], rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        RT.gd(f, y, M.FullyDenseModel(S.SphericalGrid(shape=(4, 3, 5))), progress_bar=False)


@pytest.mark.gpu
def test_gd_on_cuda_operator_matches_reference_retrieval():
    """static_retrieval.py scaled down: the loss history of 30 Adam steps on the CUDA operator
    follows the reference's (which used its own stored operator) to 1e-6."""
    z = np.load(f"{GOLDEN}/retrieval.npz")
    grid = S.SphericalGrid(shape=(16, 16, 16))
    x = torch.zeros(grid.shape); x[:, 8:, :8] = 1; x[:, :8, 8:] = 1
    geoms = sum(S.ConeCircGeom((24, 16), (5 * torch.cos(t), 5 * torch.sin(t), 1), fov=(5, 45))
                for t in torch.linspace(0, 2 * torch.pi, 7)[:6])
    op = S.Operator(grid, geoms, device="cuda")
    meas = op(x.cuda())
    assert np.allclose(meas.cpu().numpy(), z["ret_meas"], rtol=1e-5, atol=1e-6)
    fns = [1 * L.SquareLoss(), 1 * L.NegRegularizer()]
    y = torch.from_numpy(z["ret_meas"]).cuda()
    coeffs, y_fit, losses = RT.gd(op, y, M.FullyDenseModel(grid), lr=1e-1, loss_fns=fns,
                                  num_iterations=30, progress_bar=False)
    assert coeffs.dtype == torch.float64 and coeffs.is_cuda
    hist = np.array([losses[k] for k in fns])
    assert np.allclose(hist, z["ret_hist"], rtol=1e-6, atol=1e-9), np.abs(hist - z["ret_hist"]).max()
    assert np.allclose(coeffs.detach().cpu().numpy(), z["ret_coeffs"], rtol=1e-5, atol=1e-6)
    assert np.allclose(y_fit.detach().cpu().numpy(), z["ret_yfit"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_preview3d_is_a_multichannel_client():
    """plotting.preview3d: N_a rolled copies of a volume through one detector."""
    from sph_raytracer_b200.plotting import preview3d
    grid = S.SphericalGrid(shape=(10, 8, 12))
    vol = torch.rand(tuple(grid.shape), generator=torch.Generator().manual_seed(0)).cuda()
    img = preview3d(vol, grid, shape=(32, 32))
    assert img.shape == (12, 32, 32)
    rgb = preview3d(torch.stack([vol, 2 * vol, 3 * vol], dim=-1), grid, shape=(32, 32))
    assert rgb.shape == (12, 32, 32, 3)
    assert torch.allclose(rgb[..., 0], img) and torch.allclose(rgb[..., 2], 3 * img, rtol=1e-5)
    # frame k looks at the volume rolled by k azimuth bins
    import math
    r, e, a = 4 * grid.size.r[1], math.radians(60), 0.125 * 2 * math.pi / 12
    pos = (r * math.sin(e) * math.cos(a), r * math.sin(e) * math.sin(a), r * math.cos(e))
    op = S.Operator(grid, S.ConeRectGeom((32, 32), pos=pos, fov=(30, 30)))
    assert torch.allclose(img[5], op(torch.roll(vol, 5, dims=-1)), rtol=1e-5, atol=1e-6)
    gd_ = S.SphericalGrid(shape=(3, 10, 8, 12))
    dyn = preview3d(torch.rand(tuple(gd_.shape)).cuda(), gd_, shape=(16, 16))
    assert dyn.shape == (3, 16, 16)
