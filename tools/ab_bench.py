"""A/B of two builds of libsphrt.so on the bench workload (subset of observations), raw C-ABI calls
so that builds with different struct layouts / entry points can be compared."""
import ctypes, os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, sph_raytracer_b200 as S
from sph_raytracer_b200 import _cabi
n_obs = int(sys.argv[1]); libs = sys.argv[2:]
grid, geoms, _ = bench.workload("c5", n_obs)
os.environ["SPHRT_AUTOTUNE"] = "0"
op = S.Operator(grid, S.ViewGeomCollection(*geoms))
x = torch.rand(tuple(grid.shape), device="cuda"); y = torch.rand(op._rays.n, device="cuda")
out = torch.empty(op._rays.n, device="cuda"); grad = torch.zeros_like(x)
_, nseg = op.segment_counts()
for rep in range(2):
    for path in libs:
        lib = ctypes.CDLL(path)
        lib.sphrt_scratch_bytes.restype = ctypes.c_int64
        lib.sphrt_scratch_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        nb = lib.sphrt_scratch_bytes(ctypes.byref(op._tables.c), op._rays.n)
        scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
        args = lambda src, dst: (ctypes.byref(op._tables.c), ctypes.byref(op._rays.c), ctypes.c_void_p(src.data_ptr()), 0, ctypes.c_int64(1), ctypes.c_int64(0),
                                 ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(scratch.data_ptr()), ctypes.c_int64(nb), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        tf, ta = [], []
        for k in range(5):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record(); assert lib.sphrt_forward(*args(x, out)) == 0; b.record(); grad.zero_(); assert lib.sphrt_adjoint(*args(y, grad)) == 0; c.record()
            torch.cuda.synchronize()
            if k >= 2: tf.append(a.elapsed_time(b)); ta.append(b.elapsed_time(c))
        print(f"{path:24s} obs {n_obs} fwd {np.median(tf):8.2f} adj {np.median(ta):8.2f} ms  seg/s {2 * nseg / (np.median(tf) + np.median(ta)) * 1e3:.3e}  checksum {float(out.double().sum()):.6f}", flush=True)
