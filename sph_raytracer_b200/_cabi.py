"""ctypes binding of libsphrt.so (the C ABI declared in include/sphrt.h).

There is deliberately no fallback: if the CUDA extension is missing or a call fails the
caller gets an exception, never a silently slower path.
"""
import ctypes
import os

from . import build as _build

c_void_p, c_int, c_int32, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64

F32, F64 = 0, 1

EXPORTS = ("sphrt_version", "sphrt_error_string", "sphrt_last_cuda_error", "sphrt_scratch_bytes",
           "sphrt_forward", "sphrt_adjoint", "sphrt_count_segments", "sphrt_trace_segments",
           "sphrt_find_starts", "sphrt_trace_events", "sphrt_walker_kind",
           "sphrt_store_segments", "sphrt_forward_stored", "sphrt_adjoint_stored", "sphrt_candidates",
           "sphrt_launch_count", "sphrt_export_rays",
           "sphrt_pack_count", "sphrt_pack_segments", "sphrt_forward_packed", "sphrt_adjoint_packed",
           "sphrt_normal", "sphrt_plan_cache_bytes")


class Grid(ctypes.Structure):
    """struct sphrt_grid"""
    _fields_ = [("nr", c_int32), ("ne", c_int32), ("na", c_int32), ("n_c2", c_int32),
                ("r_b", c_void_p), ("e_b", c_void_p), ("cos_e", c_void_p), ("a_b", c_void_p),
                ("cos_a", c_void_p), ("sin_a", c_void_p), ("c2_sorted", c_void_p)]


class Rays(ctypes.Structure):
    """struct sphrt_rays"""
    _fields_ = [("n_rays", c_int64), ("starts", c_void_p), ("rays_per_start", c_int64),
                ("dirs", c_void_p), ("rays_per_dir", c_int64), ("start_cells", c_void_p),
                ("walker", c_int64),
                ("gen_kind", c_int32), ("gen_h", c_int32), ("gen_w", c_int32), ("gen_per_obs_tables", c_int32),
                ("gen_frames", c_void_p), ("gen_ta", c_void_p), ("gen_tb", c_void_p), ("gen_tc", c_void_p),
                ("plan_cache", c_void_p), ("plan_cache_bytes", c_int64), ("plan_cache_state", c_int64)]


class SphrtError(RuntimeError):
    pass


_lib = None


def load(path=None):
    """Load libsphrt.so (building it is the job of __graft_entry__.build / build.build)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # SPHRT_LIB selects another build of the same extension (kernel tuning experiments)
    path = path or os.environ.get("SPHRT_LIB") or _build.LIB
    if not os.path.exists(path):
        raise SphrtError(
            f"CUDA extension {path} is missing. Build it with "
            "`python -m sph_raytracer_b200.build` (needs nvcc); there is no CPU fallback.")
    lib = ctypes.CDLL(path)
    P = ctypes.POINTER
    lib.sphrt_version.restype = c_int
    lib.sphrt_error_string.restype = ctypes.c_char_p
    lib.sphrt_error_string.argtypes = [c_int]
    lib.sphrt_last_cuda_error.restype = c_int
    lib.sphrt_launch_count.restype = ctypes.c_ulonglong
    lib.sphrt_launch_count.argtypes = []
    lib.sphrt_scratch_bytes.restype = c_int64
    lib.sphrt_scratch_bytes.argtypes = [P(Grid), c_int64]
    lib.sphrt_plan_cache_bytes.restype = c_int64
    lib.sphrt_plan_cache_bytes.argtypes = [P(Grid), c_int64]
    vol_args = [P(Grid), P(Rays), c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p]
    lib.sphrt_forward.restype = c_int
    lib.sphrt_forward.argtypes = vol_args
    lib.sphrt_adjoint.restype = c_int
    lib.sphrt_adjoint.argtypes = vol_args
    lib.sphrt_count_segments.restype = c_int
    lib.sphrt_count_segments.argtypes = [P(Grid), P(Rays), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.sphrt_trace_segments.restype = c_int
    lib.sphrt_trace_segments.argtypes = [P(Grid), P(Rays), c_int64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int64, c_void_p]
    lib.sphrt_trace_events.restype = c_int
    lib.sphrt_trace_events.argtypes = [P(Grid), P(Rays), c_int64, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int64, c_void_p]
    lib.sphrt_store_segments.restype = c_int
    lib.sphrt_store_segments.argtypes = [P(Grid), P(Rays), c_void_p, c_void_p, c_void_p, c_int,
                                         c_void_p, c_int64, c_void_p]
    stored_args = [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_int64,
                   c_int64, c_void_p, c_void_p]
    lib.sphrt_forward_stored.restype = c_int
    lib.sphrt_forward_stored.argtypes = stored_args
    lib.sphrt_adjoint_stored.restype = c_int
    lib.sphrt_adjoint_stored.argtypes = stored_args
    lib.sphrt_pack_count.restype = c_int
    lib.sphrt_pack_count.argtypes = [P(Grid), P(Rays), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.sphrt_pack_segments.restype = c_int
    lib.sphrt_pack_segments.argtypes = [P(Grid), P(Rays), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    packed_args = [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int64, c_int64,
                   c_int64, c_void_p, c_void_p]
    lib.sphrt_forward_packed.restype = c_int
    lib.sphrt_forward_packed.argtypes = packed_args
    lib.sphrt_adjoint_packed.restype = c_int
    lib.sphrt_adjoint_packed.argtypes = packed_args
    lib.sphrt_normal.restype = c_int
    lib.sphrt_normal.argtypes = [P(Grid), P(Rays), c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p,
                                 c_void_p, ctypes.c_double, c_void_p, c_void_p, c_int64, c_void_p]
    lib.sphrt_candidates.restype = c_int
    lib.sphrt_candidates.argtypes = [P(Grid), P(Rays), c_void_p, c_void_p, c_void_p, c_void_p]
    lib.sphrt_walker_kind.restype = c_int
    lib.sphrt_walker_kind.argtypes = [P(Grid), c_int64]
    lib.sphrt_find_starts.restype = c_int
    lib.sphrt_find_starts.argtypes = [P(Grid), c_int64, c_void_p, c_void_p, c_void_p]
    lib.sphrt_export_rays.restype = c_int
    lib.sphrt_export_rays.argtypes = [P(Rays), c_void_p, c_void_p]
    _lib = lib
    return lib


def check(code):
    if code == 0:
        return
    lib = load()
    msg = lib.sphrt_error_string(code).decode()
    if code == -4:
        msg += f" (cudaError {lib.sphrt_last_cuda_error()})"
    raise SphrtError(f"libsphrt: {msg}")
