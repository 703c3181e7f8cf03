import os, sys, numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sph_raytracer_b200 as S
from oracle import oracle as O
from helpers import canonical_from_events
seed, ray = 1, 2102
gen = torch.Generator().manual_seed(seed)
n = 6000
xs = torch.randn(n, 3, generator=gen, dtype=torch.float64)
xs = xs / xs.norm(dim=-1, keepdim=True) * (3 + 4 * torch.rand(n, 1, generator=gen, dtype=torch.float64))
d = -xs + 1.2 * torch.randn(n, 3, generator=gen, dtype=torch.float64)
grid = S.SphericalGrid((12, 40, 17), spacing="log", size_r=(0.1, 2.9))
geom = S.ViewGeom(xs[ray:ray+1], d[ray:ray+1])
print("x", xs[ray].tolist(), "d", d[ray].tolist())
ms = 200
x_, d_, _ = O.broadcast_rays(geom.ray_starts, geom.rays)
og = O.OracleGrid(grid)
regs, lens = O.trace(og, x_, d_)
v0, s0, c0 = O.canonical(og, regs, lens, max_seg=ms)
print("oracle count", c0[0])
op = S.Operator(grid, geom)
vox, ln, cnt, stats = op.events()
print("stats", stats, "events", int(cnt[0]))
k = int(cnt[0])
ev = vox[0, :k].cpu().numpy(); el = ln[0, :k].cpu().numpy()
cc, cv, cl = canonical_from_events(vox.cpu().numpy(), ln.cpu().numpy(), cnt.cpu().numpy(), ms)
sv, sl, sc = op.segments(ms)
sv = sv.reshape(-1, ms, 3).cpu().numpy(); sl = sl.reshape(-1, ms).cpu().numpy(); sc = sc.reshape(-1).cpu().numpy()
print("coop count", cc[0], "tracked count", sc[0])
for j in range(max(c0[0], cc[0])):
    a = (tuple(v0[0, j]), s0[0, j]) if j < c0[0] else None
    b = (tuple(cv[0, j]), cl[0, j]) if j < cc[0] else None
    t = (tuple(sv[0, j]), sl[0, j]) if j < sc[0] else None
    flag = "" if (a and b and a[0] == b[0] and abs(a[1] - b[1]) < 1e-9) else "  <<<"
    print(j, "oracle", a, "coop", b, "tracked", t, flag)
print("raw events:")
for j in range(k): print(j, ev[j], el[j])
