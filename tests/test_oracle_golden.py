"""Pins the CPU oracle (oracle/sph_oracle.c) to the reference's own outputs.

The fixtures in tests/golden/ were produced by importing the unmodified reference
(tests/golden/make_golden.py).  Everything here runs on the CPU.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from helpers import OPERATOR_CASES, load_case, case_density, case_weights, ragged_equal, GOLDEN


def _same(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


def test_known_answer_tables():
    """r_torch / e_torch / a_torch known-answer inputs of reference test_all.py:18-173:
    regions bit-exact wherever t is finite, t equal up to the reference build's 1-ulp sqrt."""
    z = np.load(f"{GOLDEN}/known_answers.npz")
    big = np.array([0.0, 1.0])
    for i in range(int(z["n_cases"])):
        kind = str(z[f"c{i:02d}_kind"])
        b = z[f"c{i:02d}_b"]
        xs, dirs = z[f"c{i:02d}_xs"], z[f"c{i:02d}_dirs"]
        # build a grid in which only the family under test carries the case's boundaries
        grid = dict(r_b=big, e_b=big, a_b=big)
        grid[{"r": "r_b", "e": "e_b", "a": "a_b"}[kind]] = b
        from types import SimpleNamespace
        og = O.OracleGrid(SimpleNamespace(**{k: torch.from_numpy(np.asarray(v, dtype=np.float64))
                                             for k, v in grid.items()}))
        t, reg = O.candidates(og, np.ascontiguousarray(xs), np.ascontiguousarray(dirs))
        nr1, ne1 = og.nr + 1, og.ne + 1
        sl = {"r": slice(0, 2 * nr1), "e": slice(2 * nr1, 2 * nr1 + 2 * ne1),
              "a": slice(2 * nr1 + 2 * ne1, None)}[kind]
        T, R = z[f"c{i:02d}_t"], z[f"c{i:02d}_reg"]
        assert t[:, sl].shape == T.shape
        fin = np.isfinite(T)
        assert np.array_equal(np.isfinite(t[:, sl]), fin), (i, kind)
        assert np.array_equal(t[:, sl][~fin], T[~fin]), (i, kind)          # same +-inf pattern
        assert np.allclose(t[:, sl][fin], T[fin], rtol=4e-16, atol=0), (i, kind)
        assert np.array_equal(reg[:, sl][fin], R[fin]), (i, kind)


def test_known_answer_find_starts():
    """reference test_all.py:225-234"""
    from sph_raytracer_b200.geometry import SphericalGrid
    z = np.load(f"{GOLDEN}/known_answers.npz")
    for row in z["find_starts"]:
        shape, x, want = row[:3].astype(int), row[3:6], row[6:9].astype(int)
        got = O.find_starts(SphericalGrid(shape=tuple(shape)), x[None, :])[:, 0]
        assert np.array_equal(got, want)
    assert np.array_equal(z["find_starts"][:, 6:9], [[-1, 0, 0], [-1, 4, 0], [-1, 2, 2]])


@pytest.mark.parametrize("name", OPERATOR_CASES)
def test_oracle_matches_reference_operator(name):
    """Canonical voxel sequences / crossing counts bit-exact on the oracle-stable rays,
    lengths to 1e-9, integrals and autograd adjoint to the fp32 tolerance."""
    c = load_case(name)
    og = O.OracleGrid(c.grid)
    regs, lens = O.trace(og, c.xs, c.dirs)
    vox, seg, cnt = O.canonical(og, regs, lens, max_seg=max(int(c.count.max()), 1) + 8)
    # the reference's own integration rays (chords_*, hemisphere) run exactly along grid
    # planes / through the axis: exact ties whose order the reference leaves unspecified
    # (SURVEY.md 8a quirks 2-3), so only their integrals are graded, like upstream does.
    degenerate = name.startswith("chords_") or name == "hemisphere"
    n, bad = ragged_equal(c, cnt, vox, seg, mask=c.stable & (not degenerate))
    assert len(bad) == 0, f"{name}: {len(bad)}/{n} stable rays differ, first {bad[:5]}"
    # unstable rays: only require that they are few
    assert (~c.stable).sum() <= 0.06 * len(c.stable) + 1

    dens = case_density(c)
    if dens is None:
        return
    dyn = name == "c4_dynamic"
    nray = len(c.count)
    prv = None
    if dyn:
        prv = np.repeat(np.arange(dens.shape[0]), nray // dens.shape[0]).astype(np.int32)
    y = O.forward(og, regs, lens, dens, per_ray_vol=prv)
    want = c.integral.reshape(y.shape)
    tol = 1e-5 if dens.dtype == np.float32 else 1e-12
    st = np.broadcast_to(c.stable.reshape((1, -1) if y.ndim == 2 else (-1,)), y.shape)
    assert np.allclose(y[st], want[st], rtol=tol, atol=tol), name
    if name.startswith("chords_"):
        # reference test_raytracer.py:49-52: chord through the solid sphere = 2 (r_max - r_min)
        assert np.allclose(y, 2 * (c.r_b[-1] - c.r_b[0]), atol=1e-2)
    if name == "hemisphere":
        assert np.allclose(y, 50, atol=1e-2)          # test_raytracer.py:85-116
    if hasattr(c, "grad"):
        w = case_weights(c).reshape(y.shape)
        g = O.adjoint(og, regs, lens, w, n_vol=dens.reshape(-1, *c.nvox).shape[0], per_ray_vol=prv)
        want_g = c.grad.reshape(g.shape)
        assert np.allclose(g, want_g, rtol=10 * tol, atol=tol * max(1.0, float(np.abs(want_g).max()))), name
