import os, sys, numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sph_raytracer_b200 as S
from oracle import oracle as O
from helpers import canonical_from_events
g = torch.Generator().manual_seed(3)
n = 400000
xs = torch.randn(n, 3, generator=g, dtype=torch.float64)
xs = xs / xs.norm(dim=-1, keepdim=True) * (30 + 60 * torch.rand(n, 1, generator=g, dtype=torch.float64))
d = -xs + 12 * torch.randn(n, 3, generator=g, dtype=torch.float64)
grid = S.SphericalGrid((64, 181, 91), size_r=(0, 20))
idx = [74397, 77896, 229459, 377999]
geom = S.ViewGeom(xs[idx], d[idx])
ms = 800
x_, d_, _ = O.broadcast_rays(geom.ray_starts, geom.rays)
og = O.OracleGrid(grid)
regs, lens = O.trace(og, x_, d_)
v0, s0, c0 = O.canonical(og, regs, lens, max_seg=ms)
op = S.Operator(grid, geom)
vox, ln, cnt, stats = op.events()
cc, cv, cl = canonical_from_events(vox.cpu().numpy(), ln.cpu().numpy(), cnt.cpu().numpy(), ms)
sv, sl, sc = op.segments(ms)
sv = sv.reshape(-1, ms, 3).cpu().numpy(); sl = sl.reshape(-1, ms).cpu().numpy(); sc = sc.reshape(-1).cpu().numpy()
print("stats", stats)
for r in range(4):
    print("ray", idx[r], "x", xs[idx[r]].tolist(), "d", d[idx[r]].tolist())
    print("  counts oracle", c0[r], "coop", cc[r], "tracked", sc[r])
    shown = 0
    for j in range(max(c0[r], cc[r], sc[r])):
        a = (tuple(int(v) for v in v0[r, j]), float(s0[r, j])) if j < c0[r] else None
        b = (tuple(int(v) for v in cv[r, j]), float(cl[r, j])) if j < cc[r] else None
        t = (tuple(int(v) for v in sv[r, j]), float(sl[r, j])) if j < sc[r] else None
        okb = a and b and a[0] == b[0] and abs(a[1] - b[1]) < 1e-9
        okt = a and t and a[0] == t[0] and abs(a[1] - t[1]) < 1e-9
        if not (okb and okt) and shown < 6:
            print("   seg", j, "oracle", a, "| coop", b, "OK" if okb else "BAD", "| tracked", t, "OK" if okt else "BAD"); shown += 1
