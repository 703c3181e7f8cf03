import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sph_raytracer_b200 as S
n = 200
grid = S.SphericalGrid((n, n, n), size_r=(0.5, 5))
for p in ((80.0, 10.0, 20.0), (-30.0, 70.0, -40.0), (0.0, 30.0, -85.0), (20.0, 20.0, 88.0)):
    op = S.Operator(grid, S.ConeRectGeom((512, 512), p, fov=(7, 7)))
    _, _, cnt, stats = op.events(max_ev=8)
    c, tot = op.segment_counts()
    print(p, "stats", stats, "rays", op._rays.n, "hits", int((c > 0).sum()), "segs", tot)
