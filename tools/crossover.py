"""Where does the cooperative walker start to pay off?  Cubic grids, one 512x512 pinhole detector."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sph_raytracer_b200 as S
for n in (32, 50, 64, 80, 100, 128, 160, 200):
    grid = S.SphericalGrid((n, n, n), size_r=(0.5, 5))
    geom = sum(S.ConeRectGeom((512, 512), p, fov=(7, 7)) for p in ((80.0, 10.0, 20.0), (-30.0, 70.0, -40.0), (10.0, -60.0, 55.0), (-50.0, -50.0, 5.0), (3.0, 31.0, -85.0), (70.0, 40.0, 40.0), (-80.0, 5.0, 30.0), (21.0, 17.0, 88.0)))
    op = S.Operator(grid, geom)
    x = torch.rand(tuple(grid.shape), device="cuda"); y = torch.rand(op._rays.shape, device="cuda")
    _, nseg = op.segment_counts()
    line = f"n={n:4d} segs/ray {nseg / op._rays.n:6.1f}"
    for mode in ("tracked", "coop"):
        os.environ["SPHRT_WALKER"] = mode
        ts = []
        for k in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); op._apply_forward(x); op._apply_adjoint(y, x.shape); b.record(); torch.cuda.synchronize()
            if k >= 2: ts.append(a.elapsed_time(b))
        line += f"  {mode} {np.median(ts):8.3f} ms"
    os.environ.pop("SPHRT_WALKER")
    print(line, flush=True)
