"""Shared helpers of the test-suite: golden fixture access and canonical comparisons."""
import glob
import os
import types

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

OPERATOR_CASES = sorted(
    os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
    if os.path.basename(p) not in ("known_answers.npz", "retrieval.npz",
                                   # outputs of the reference's callers (make_golden_callers.py)
                                   "retrieval_c3.npz", "preview3d.npz", "c4_full.npz"))

# seeds / formulas of the volumes used by tests/golden/make_golden.py
DENSITY = {
    "c1_rand": (0, np.float32), "c2_static_rand": (0, np.float32),
    "c2_static_rand_f64": (0, np.float64), "c2_multichannel": (3, np.float32),
    "c4_dynamic": (4, np.float32), "rand_partial": (5, np.float32), "rand_log": (6, np.float32),
    "rand_full_inside": (7, np.float32), "parallel_hollow": (8, np.float32),
    "onplane_16": (9, np.float32),
}


def density_from_seed(shape, seed, dtype=np.float32):
    return np.random.default_rng(seed).random(tuple(shape), dtype=np.float32).astype(dtype)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    c = types.SimpleNamespace(**{k: z[k] for k in z.files})
    c.name = name
    c.grid = types.SimpleNamespace(r_b=torch.from_numpy(c.r_b), e_b=torch.from_numpy(c.e_b),
                                   a_b=torch.from_numpy(c.a_b))
    c.nvox = (len(c.r_b) - 1, len(c.e_b) - 1, len(c.a_b) - 1)
    c.offsets = np.concatenate(([0], np.cumsum(c.count)))
    return c


def case_density(c):
    """The volume make_golden.py traced for this case (None if the case has no integral)."""
    if not hasattr(c, "integral"):
        return None
    nr, ne, na = c.nvox
    if c.name.startswith("chords_"):
        return np.ones(c.nvox, dtype=np.float32)
    if c.name == "hemisphere":
        return np.array([[[1.0], [0.0]]], dtype=np.float32)
    if c.name == "c1_single_vantage":
        x = np.zeros(c.nvox, dtype=np.float32); x[-1] += 1; x[-10] += 1
        return x
    if c.name == "c2_static_pattern":
        x = np.zeros(c.nvox, dtype=np.float32); x[:, 25:, :25] = 1; x[:, :25, 25:] = 1
        return x
    seed, dt = DENSITY[c.name]
    lead = tuple(c.integral.shape[:1]) if c.name in ("c2_multichannel", "c4_dynamic") else ()
    return density_from_seed(lead + tuple(c.nvox), seed, dt)


def case_weights(c):
    """Adjoint input used by make_golden.py: seeded weights, zero on oracle-unstable rays."""
    w = density_from_seed(c.integral.shape, 100, c.integral.dtype)
    shape = tuple(int(v) for v in c.ray_shape)
    return w * c.stable.reshape(shape).astype(w.dtype)


def ragged_equal(c, count, vox, seg_len, mask=None, len_rtol=1e-9, len_atol=1e-12):
    """Compare (count, vox[n,max,3], seg_len[n,max]) padded arrays with the ragged golden
    sequences of case ``c``.  Returns (n_checked, index array of mismatching rays)."""
    bad = []
    idx = np.arange(len(c.count)) if mask is None else np.where(mask)[0]
    for i in idx:
        k = c.count[i]
        ok = count[i] == k
        if ok and k:
            gv = c.vox[c.offsets[i]:c.offsets[i + 1]]
            gl = c.seg_len[c.offsets[i]:c.offsets[i + 1]]
            ok = np.array_equal(vox[i, :k], gv) and \
                np.allclose(seg_len[i, :k], gl, rtol=len_rtol, atol=len_atol)
        if not ok:
            bad.append(i)
    return len(idx), np.array(bad, dtype=np.int64)


def canonical_from_events(vox, lens, count, max_seg):
    """Canonical segments (non-zero length, consecutive identical voxels merged) from the raw
    event export of the warp-cooperative walker.  Rays with count < 0 (not handled by that
    walker) get count -1.  Returns (count, vox[n,max_seg,3], seg_len[n,max_seg])."""
    n = len(count)
    out_c = np.full(n, -1, dtype=np.int64)
    out_v = np.full((n, max_seg, 3), -9, dtype=np.int32)
    out_l = np.zeros((n, max_seg), dtype=np.float64)
    for i in range(n):
        k = int(count[i])
        if k < 0:
            continue
        v, l = vox[i, :k], lens[i, :k]
        keep = l > 0
        v, l = v[keep], l[keep]
        m = 0
        for j in range(len(l)):
            if m > 0 and np.array_equal(out_v[i, m - 1], v[j]):
                out_l[i, m - 1] += l[j]
                continue
            if m < max_seg:
                out_v[i, m] = v[j]; out_l[i, m] = l[j]
            m += 1
        out_c[i] = m
    return out_c, out_v, out_l
