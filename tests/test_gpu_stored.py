"""GPU tests of the stored operator (Operator(..., store_segments=True)): the
canonical segments are traced once and kept as an 8-byte-per-segment list; applying it must
give what the on-the-fly kernels give, and what the reference gives (golden fixtures)."""
import math
import types

import numpy as np
import pytest
import torch

from helpers import load_case, case_density, case_weights

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


@pytest.mark.parametrize("name", ["c1_rand", "c2_static_rand", "c2_static_rand_f64", "c2_multichannel",
                                  "c4_dynamic", "rand_full_inside", "rand_partial", "rand_log",
                                  "parallel_hollow"])
def test_stored_operator_matches_reference_golden(S, name):
    c, dens, shape, op = _case_operator(S, name, store_segments=torch.float64 if "f64" in name else True)
    assert op.n_segments == int(c.count.sum()) or not c.stable.all()
    assert op.stored_bytes >= op.n_segments * (12 if "f64" in name else 8)
    x = torch.from_numpy(dens).cuda().requires_grad_(True)
    y = op(x)
    want = c.integral
    tol = 1e-5 if dens.dtype == np.float32 else 1e-10
    st = np.broadcast_to(c.stable.reshape(shape), want.shape[-len(shape):] if shape else ())
    st = np.broadcast_to(st, want.shape)
    got = y.detach().cpu().numpy()
    assert got.shape == want.shape
    assert np.allclose(got[st], want[st], rtol=tol, atol=tol), np.abs(got[st] - want[st]).max()
    if hasattr(c, "grad"):
        w = case_weights(c)
        (y * torch.from_numpy(w).cuda()).sum().backward()
        g = x.grad.cpu().numpy()
        atol = tol * max(1.0, float(np.abs(c.grad).max()))
        assert np.allclose(g, c.grad, rtol=10 * tol, atol=atol), np.abs(g - c.grad).max()


def test_stored_operator_equals_on_the_fly(S):
    """Same geometry, both modes, float64 volume and float64 stored lengths: equal to rounding
    of the sums; float32 stored lengths: equal to 1e-6."""
    grid = S.SphericalGrid((40, 36, 48), size_r=(0.5, 3))
    geoms = sum(S.ConeRectGeom((48, 40), (7 * math.cos(t), 7 * math.sin(t), 1.5 * math.sin(3 * t)), fov=(50, 50))
                for t in (0.2, 1.1, 2.7, 4.0))
    fly = S.Operator(grid, geoms)
    st64 = S.Operator(grid, geoms, store_segments=torch.float64)
    st32 = S.Operator(grid, geoms, store_segments=True)
    _, total = fly.segment_counts()
    assert st64.n_segments == total and st32.stored_bytes == 8 * total + 8 * (fly._rays.n + 1)
    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.rand(tuple(grid.shape), generator=gen, device="cuda", dtype=torch.float64)
    y = torch.rand(fly._rays.shape, generator=gen, device="cuda", dtype=torch.float64)
    a, b, c = fly(x), st64(x), st32(x)
    assert torch.allclose(a, b, rtol=1e-12, atol=1e-13)
    assert torch.allclose(a, c, rtol=1e-6, atol=1e-7)
    ga, gb, gc = fly.T(y), st64.T(y), st32.T(y)
    assert torch.allclose(ga, gb, rtol=1e-11, atol=1e-12)
    assert torch.allclose(ga, gc, rtol=1e-6, atol=1e-7)
    # channels and autograd go through the same list
    xs = torch.rand((3,) + tuple(grid.shape), generator=gen, device="cuda").requires_grad_(True)
    ys = st32(xs)
    assert ys.shape == (3,) + fly._rays.shape
    ys.sum().backward()
    ref = fly.T(torch.ones(fly._rays.shape, device="cuda"))
    assert torch.allclose(xs.grad[1], ref, rtol=1e-5, atol=1e-5)


# ---- packed list (4 bytes per segment) and the hybrid operator -------------------------------------

def _decode_packed(offsets, words, ne, na):
    """Host restatement of the packed format of include/sphrt.h: per ray (flat voxel offsets, lengths)."""
    off = offsets.cpu().numpy()
    w = words.cpu().numpy().view(np.uint32)
    delta = np.array([1, -1, na, -na, ne * na, -ne * na, 0, 0], dtype=np.int64)
    rays = []
    for i in range(len(off) - 1):
        ww = w[off[i]:off[i + 1]]
        code, payload = ww >> 29, ww & 0x1FFFFFFF
        vox, lens, idx = [], [], None
        for c, p in zip(code.tolist(), payload.tolist()):
            if c == 7:
                idx = p
                continue
            idx = idx + int(delta[c])
            vox.append(idx)
            lens.append(np.array([p << 2], dtype=np.uint32).view(np.float32)[0])
        rays.append((np.array(vox, dtype=np.int64), np.array(lens, dtype=np.float64)))
    return rays


@pytest.mark.parametrize("name", ["c2_static_rand", "rand_full_inside", "rand_partial", "parallel_hollow"])
def test_packed_words_decode_to_the_canonical_segments(S, name):
    c, dens, shape, op = _case_operator(S, name, store_segments="packed")
    offsets, words, n_s = op._packed
    assert n_s == op._rays.n and op.n_stored_rays == n_s
    nr, ne, na = c.nvox
    vox, lens, count = op.segments()
    vox = vox.reshape(-1, vox.shape[-2], 3).cpu().numpy().astype(np.int64)
    lens = lens.reshape(-1, lens.shape[-1]).cpu().numpy()
    count = count.reshape(-1).cpu().numpy()
    n_esc = 0
    for i, (pv, pl) in enumerate(_decode_packed(offsets, words, ne, na)):
        k = count[i]
        flat = (vox[i, :k, 0] * ne + vox[i, :k, 1]) * na + vox[i, :k, 2]
        assert len(pv) == k and np.array_equal(pv, flat), i
        assert np.allclose(pl, lens[i, :k], rtol=3e-7, atol=1e-30), i
        n_esc += int(offsets[i + 1] - offsets[i]) - k
    # one escape per ray that has segments, a few more at seams and holes: the list stays near 4 B/segment
    assert n_esc < (count > 0).sum() + 0.05 * count.sum() + 8


@pytest.mark.parametrize("name", ["c1_rand", "c2_static_rand", "c2_multichannel", "c4_dynamic",
                                  "rand_full_inside", "rand_partial", "rand_log", "parallel_hollow"])
def test_packed_operator_matches_reference_golden(S, name):
    c, dens, shape, op = _case_operator(S, name, store_segments="packed")
    x = torch.from_numpy(dens).cuda().requires_grad_(True)
    y = op(x)
    want = c.integral
    tol = 1e-5            # BASELINE.json: integrals and gradients within rtol 1e-5 in float32
    st = np.broadcast_to(c.stable.reshape(shape), want.shape[-len(shape):] if shape else ())
    st = np.broadcast_to(st, want.shape)
    got = y.detach().cpu().numpy()
    assert got.shape == want.shape
    assert np.allclose(got[st], want[st], rtol=tol, atol=tol), np.abs(got[st] - want[st]).max()
    if hasattr(c, "grad"):
        w = case_weights(c)
        (y * torch.from_numpy(w).cuda()).sum().backward()
        g = x.grad.cpu().numpy()
        atol = tol * max(1.0, float(np.abs(c.grad).max()))
        assert np.allclose(g, c.grad, rtol=10 * tol, atol=atol), np.abs(g - c.grad).max()


def _orbit(S, n_obs, shape=(40, 32)):
    return sum(S.ConeRectGeom(shape, (7 * math.cos(t), 7 * math.sin(t), 1.5 * math.sin(3 * t)), fov=(50, 50))
               for t in np.linspace(0.2, 5.9, n_obs))


@pytest.mark.parametrize("explicit_rays", [False, True])
def test_hybrid_operator_equals_on_the_fly(S, monkeypatch, explicit_rays):
    """store_segments='auto' with a budget that holds only some of the observations: the packed
    prefix and the regenerated rest together give what the regenerated operator gives (static,
    multi-channel and time-indexed volumes; analytic and uploaded ray directions)."""
    if explicit_rays:
        monkeypatch.setenv("SPHRT_RAYGEN", "0")
    grid = S.SphericalGrid((40, 36, 48), size_r=(0.5, 3))
    geoms = _orbit(S, 6)
    fly = S.Operator(grid, geoms)
    full = S.Operator(grid, geoms, store_segments="packed")
    assert full.n_stored_rays == fly._rays.n
    monkeypatch.setenv("SPHRT_STORE_BUDGET_GB", str(0.55 * full.stored_bytes / 1e9))
    hyb = S.Operator(grid, geoms, store_segments="auto")
    per_obs = 40 * 32
    assert 0 < hyb.n_stored_rays < fly._rays.n and hyb.n_stored_rays % per_obs == 0
    assert hyb.stored_bytes <= 0.55 * full.stored_bytes
    gen = torch.Generator(device="cuda").manual_seed(5)
    for dt, tol in ((torch.float32, 2e-6), (torch.float64, 2e-6)):
        x = torch.rand(tuple(grid.shape), generator=gen, device="cuda", dtype=dt)
        y = torch.rand(fly._rays.shape, generator=gen, device="cuda", dtype=dt)
        for op in (full, hyb):
            assert torch.allclose(op(x), fly(x), rtol=tol, atol=tol)
            assert torch.allclose(op.T(y), fly.T(y), rtol=tol, atol=tol * float(fly.T(y).abs().max()))
    # a budget that holds nothing: everything regenerated, same numbers
    monkeypatch.setenv("SPHRT_STORE_BUDGET_GB", "1e-9")
    none = S.Operator(grid, geoms, store_segments="auto")
    assert none.n_stored_rays == 0
    x = torch.rand(tuple(grid.shape), generator=gen, device="cuda")
    assert torch.equal(none(x), fly(x))
    monkeypatch.setenv("SPHRT_STORE_BUDGET_GB", str(0.55 * full.stored_bytes / 1e9))
    xs = torch.rand((3,) + tuple(grid.shape), generator=gen, device="cuda").requires_grad_(True)
    ys = hyb(xs)
    assert ys.shape == (3,) + fly._rays.shape
    assert torch.allclose(ys, fly(xs.detach()), rtol=2e-6, atol=2e-6)
    w = torch.rand(ys.shape, generator=gen, device="cuda")
    (ys * w).sum().backward()
    ref = torch.stack([fly.T(w[k]) for k in range(3)])
    assert torch.allclose(xs.grad, ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()))
    # time-indexed volume: observation o reads time-slice o, in both parts
    dgrid = S.SphericalGrid((6, 40, 36, 48), size_r=(0.5, 3))
    dfly = S.Operator(dgrid, geoms)
    dhyb = S.Operator(dgrid, geoms, store_segments="auto")
    assert 0 < dhyb.n_stored_rays < dfly._rays.n
    xd = torch.rand(tuple(dgrid.shape), generator=gen, device="cuda")
    yd = torch.rand(dfly._rays.shape, generator=gen, device="cuda")
    assert torch.allclose(dhyb(xd), dfly(xd), rtol=2e-6, atol=2e-6)
    gd, gr = dhyb.T(yd), dfly.T(yd)
    assert torch.allclose(gd, gr, rtol=2e-6, atol=2e-6 * float(gr.abs().max()))
