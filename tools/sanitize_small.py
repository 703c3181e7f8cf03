"""Small run of every kernel family for compute-sanitizer (memcheck): cooperative walker with its
fallback passes, per-thread walker, stored operator, event export."""
import os, sys, math, types, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sph_raytracer_b200 as S
grid = S.SphericalGrid(shape=(200, 180, 360), size_r=(3, 25))
g = S.ConeRectGeom((512, 512), (99.0, 10.0, 0.2), fov=(30, 30))          # equatorial camera: all passes busy
geom = S.ViewGeom(g.ray_starts.expand(48, 64, 3), g.rays[100:484:8, ::8].contiguous())
op = S.Operator(grid, geom)
x = torch.rand(tuple(grid.shape), device="cuda"); y = torch.rand(48, 64, device="cuda")
print(op.walker(), op.events(max_ev=64)[3])
for mode in ("coop", "tracked"):
    os.environ["SPHRT_WALKER"] = mode
    a = op._apply_forward(x); b = op._apply_adjoint(y, x.shape)
    print(mode, float(a.sum()), float(b.sum()))
os.environ.pop("SPHRT_WALKER")
st = S.Operator(grid, geom, store_segments=True)
print("stored", float(st._apply_forward(x).sum()), float(st._apply_adjoint(y, x.shape).sum()))
small = S.SphericalGrid((20, 16, 24))
gs = sum(S.ConeCircGeom((20, 10), (5 * math.cos(t), 5 * math.sin(t), 1), fov=(0, 45)) for t in (0.3, 2.0))
for mode in ("coop", "tracked"):
    os.environ["SPHRT_WALKER"] = mode
    o = S.Operator(small, gs)
    xs = torch.rand((5,) + tuple(small.shape), device="cuda")
    print(mode, float(o(xs).sum()), o.segment_counts()[1])
torch.cuda.synchronize()
print("done")
