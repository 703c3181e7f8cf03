"""Does the adjoint/forward time ratio of the config-5 workload depend on the camera latitude?"""
import os, sys, math, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, sph_raytracer_b200 as S
grid, geoms, _ = bench.workload("c5", 1000)
gen = torch.Generator().manual_seed(0)
theta = torch.rand(1000, generator=gen, dtype=torch.float64) * 2 * math.pi
phi = (0.2 + 0.6 * torch.rand(1000, generator=gen, dtype=torch.float64)) * math.pi
def run(idx, label):
    op = S.Operator(grid, S.ViewGeomCollection(*[geoms[i] for i in idx]))
    shape = tuple(grid.shape)
    x = torch.rand(shape, device="cuda"); y = torch.rand(op._rays.shape, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf, ta = [], []
    for k in range(5):
        a, b, c = ev(), ev(), ev()
        a.record(); op._apply_forward(x); b.record(); op._apply_adjoint(y, shape); c.record()
        torch.cuda.synchronize()
        if k >= 2: tf.append(a.elapsed_time(b)); ta.append(b.elapsed_time(c))
    _, nseg = op.segment_counts()
    print(f"{label:28s} n={len(idx):3d} seg {nseg:.3e} fwd {np.median(tf):7.2f} adj {np.median(ta):7.2f}  adj/fwd {np.median(ta)/np.median(tf):.3f}  ns/seg fwd {np.median(tf)*1e6/nseg:.4f} adj {np.median(ta)*1e6/nseg:.4f}", flush=True)
order = torch.argsort((phi - math.pi / 2).abs())
run(order[:24].tolist(), "equatorial cameras")
run(order[-24:].tolist(), "high-latitude cameras")
run(list(range(24)), "first 24")
run(list(range(24)) * 2, "first 24 twice (48)")
run(list(range(96)), "first 96")
