import os, sys, math, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import sph_raytracer_b200 as S
small = S.SphericalGrid((20, 16, 24))
gs = sum(S.ConeCircGeom((20, 10), (5 * math.cos(t), 5 * math.sin(t), 1), fov=(0, 45)) for t in (0.3, 2.0))
xs = torch.rand((5,) + tuple(small.shape), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
out = {}
for mode in ("coop", "tracked"):
    os.environ["SPHRT_WALKER"] = mode
    o = S.Operator(small, gs)
    out[mode] = o(xs)
d = (out["coop"] - out["tracked"]).abs()
print("shape", tuple(d.shape), "max diff", float(d.max()))
idx = (d > 1e-5).nonzero()
print("differing entries (channel, obs, row, col):", idx.tolist()[:40])
_, _, cnt, stats = o.events(max_ev=8)
print("stats", stats, "events count of differing rays", [int(cnt.reshape(2, 20, 10)[i[1], i[2], i[3]]) for i in idx.tolist()[:10]])
