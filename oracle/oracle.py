"""ctypes front-end of the CPU oracle (oracle/sph_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of sph_oracle.c.  The product package
never imports this module; tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs do.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_i32p = ctypes.POINTER(ctypes.c_int32)


class _Grid(ctypes.Structure):
    _fields_ = [("nr", ctypes.c_int32), ("ne", ctypes.c_int32), ("na", ctypes.c_int32),
                ("r_b", c_dp), ("e_b", c_dp), ("cos_e", c_dp),
                ("a_b", c_dp), ("cos_a", c_dp), ("sin_a", c_dp)]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.sphor_slots.restype = ctypes.c_int64
        _lib.sphor_slots.argtypes = [ctypes.c_int32] * 3
    return _lib


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


class OracleGrid:
    """Boundary vectors + host-computed trig tables of a SphericalGrid-like object."""

    def __init__(self, grid):
        self.grid = grid
        f = lambda x: np.ascontiguousarray(torch.asarray(x, dtype=torch.float64).numpy())
        e_b = torch.asarray(grid.e_b, dtype=torch.float64)
        a_b = torch.asarray(grid.a_b, dtype=torch.float64)
        self.r_b, self.e_b, self.a_b = f(grid.r_b), f(e_b), f(a_b)
        self.cos_e = f(torch.cos(e_b))
        self.cos_a, self.sin_a = f(torch.cos(a_b)), f(torch.sin(a_b))
        self.nr, self.ne, self.na = len(self.r_b) - 1, len(self.e_b) - 1, len(self.a_b) - 1
        self.K = int(lib().sphor_slots(self.nr, self.ne, self.na))
        self.c = _Grid(self.nr, self.ne, self.na,
                       _ptr(self.r_b, c_dp), _ptr(self.e_b, c_dp), _ptr(self.cos_e, c_dp),
                       _ptr(self.a_b, c_dp), _ptr(self.cos_a, c_dp), _ptr(self.sin_a, c_dp))

    @property
    def ref(self):
        return ctypes.byref(self.c)


def broadcast_rays(ray_starts, rays):
    """(n,3) contiguous float64 starts / directions, broadcast like trace_indices does
    (raytracer.py:77-80).  Returns (xs, dirs, ray_shape)."""
    xs = torch.asarray(ray_starts, dtype=torch.float64)
    d = torch.asarray(rays, dtype=torch.float64)
    shape = torch.broadcast_shapes(xs.shape, d.shape)
    xs = xs.broadcast_to(shape).reshape(-1, 3).contiguous().numpy()
    d = d.broadcast_to(shape).reshape(-1, 3).contiguous().numpy()
    return xs, d, tuple(shape[:-1])


def find_starts(grid, xs):
    og = grid if isinstance(grid, OracleGrid) else OracleGrid(grid)
    xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64).reshape(-1, 3))
    out = np.empty((3, len(xs)), dtype=np.int64)
    lib().sphor_find_starts(og.ref, ctypes.c_int64(len(xs)), _ptr(xs, c_dp), _ptr(out, c_i64p))
    return out


def candidates(grid, xs, dirs):
    """Unsorted candidate table (n, K-1): [r first | r second | e first | e second | a]."""
    og = grid if isinstance(grid, OracleGrid) else OracleGrid(grid)
    n = len(xs)
    t = np.empty((n, og.K - 1), dtype=np.float64)
    reg = np.empty((n, og.K - 1), dtype=np.int64)
    lib().sphor_candidates(og.ref, ctypes.c_int64(n), _ptr(xs, c_dp), _ptr(dirs, c_dp),
                           _ptr(t, c_dp), _ptr(reg, c_i64p))
    return t, reg


def trace(grid, xs, dirs, nthreads=0):
    """regs (3,n,K) int64 and lens (n,K) float64, the reference's stored operator."""
    og = grid if isinstance(grid, OracleGrid) else OracleGrid(grid)
    n = len(xs)
    regs = np.empty((3, n, og.K), dtype=np.int64)
    lens = np.empty((n, og.K), dtype=np.float64)
    lib().sphor_trace(og.ref, ctypes.c_int64(n), _ptr(xs, c_dp), _ptr(dirs, c_dp),
                      _ptr(regs, c_i64p), _ptr(lens, c_dp), ctypes.c_int32(nthreads))
    return regs, lens


def canonical(grid, regs, lens, max_seg=None):
    """Canonical comparison form: per-ray merged non-zero segments.

    Returns (vox (n,max_seg,3) int32, seg_len (n,max_seg) f64, count (n) int32)."""
    og = grid if isinstance(grid, OracleGrid) else OracleGrid(grid)
    n = lens.shape[0]
    if max_seg is None:
        max_seg = og.K
    regs = np.ascontiguousarray(regs, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.float64)
    vox = np.full((n, max_seg, 3), -9, dtype=np.int32)
    seg = np.zeros((n, max_seg), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    lib().sphor_canonical(og.ref, ctypes.c_int64(n), _ptr(regs, c_i64p), _ptr(lens, c_dp),
                          ctypes.c_int64(max_seg), _ptr(vox, c_i32p), _ptr(seg, c_dp),
                          _ptr(cnt, c_i32p))
    return vox, seg, cnt


def _apply_args(og, regs, lens, arr, per_ray_vol):
    regs = np.ascontiguousarray(regs, dtype=np.int64)
    lens = np.ascontiguousarray(lens, dtype=np.float64)
    if arr.dtype == np.float32:
        suffix, ctype = "f32", ctypes.POINTER(ctypes.c_float)
    elif arr.dtype == np.float64:
        suffix, ctype = "f64", c_dp
    else:
        raise TypeError(arr.dtype)
    prv = None if per_ray_vol is None else np.ascontiguousarray(per_ray_vol, dtype=np.int32)
    return regs, lens, suffix, ctype, prv


def forward(grid, regs, lens, density, per_ray_vol=None, nthreads=0):
    """Operator.__call__ on stored regs/lens.  density (n_vol,nr,ne,na) -> (n_vol,n),
    or with per_ray_vol (n,) int -> (n,) (dynamic volume, one time-slice per ray)."""
    og = grid if isinstance(grid, OracleGrid) else OracleGrid(grid)
    density = np.ascontiguousarray(density).reshape(-1, og.nr, og.ne, og.na)
    regs, lens, suffix, ctype, prv = _apply_args(og, regs, lens, density, per_ray_vol)
    n = lens.shape[0]
    out = np.zeros((n,) if prv is not None else (density.shape[0], n), dtype=density.dtype)
    getattr(lib(), f"sphor_forward_{suffix}")(
        og.ref, ctypes.c_int64(n), _ptr(regs, c_i64p), _ptr(lens, c_dp), _ptr(density, ctype),
        ctypes.c_int64(density.shape[0]), None if prv is None else _ptr(prv, c_i32p),
        _ptr(out, ctype), ctypes.c_int32(nthreads))
    return out


def adjoint(grid, regs, lens, y, n_vol=1, per_ray_vol=None):
    """Scatter-add adjoint.  y (n_vol,n) -> grad (n_vol,nr,ne,na); with per_ray_vol y is (n,)."""
    og = grid if isinstance(grid, OracleGrid) else OracleGrid(grid)
    y = np.ascontiguousarray(y)
    regs, lens, suffix, ctype, prv = _apply_args(og, regs, lens, y, per_ray_vol)
    n = lens.shape[0]
    grad = np.zeros((n_vol, og.nr, og.ne, og.na), dtype=y.dtype)
    getattr(lib(), f"sphor_adjoint_{suffix}")(
        og.ref, ctypes.c_int64(n), _ptr(regs, c_i64p), _ptr(lens, c_dp), _ptr(y, ctype),
        ctypes.c_int64(n_vol), None if prv is None else _ptr(prv, c_i32p), _ptr(grad, ctype))
    return grad
