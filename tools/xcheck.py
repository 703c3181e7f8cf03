"""Cross-check of the two GPU walkers at scale: float64 volume, per-ray integrals and the gradient
of whole config-5 observations (and of random rays on random grids) must agree to rounding."""
import os, sys, math, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, sph_raytracer_b200 as S

def compare(grid, geom, label):
    op = S.Operator(grid, geom)
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(tuple(grid.shape), generator=gen, device="cuda", dtype=torch.float64)
    y = torch.rand(op._rays.shape, generator=gen, device="cuda", dtype=torch.float64)
    res = {}
    for mode in ("tracked", "coop"):
        os.environ["SPHRT_WALKER"] = mode
        res[mode] = (op._apply_forward(x), op._apply_adjoint(y, x.shape))
    os.environ.pop("SPHRT_WALKER")
    f0, g0 = res["tracked"]; f1, g1 = res["coop"]
    df = (f0 - f1).abs(); rel = df / f0.abs().clamp_min(1e-3)
    bad = (rel > 1e-11).nonzero()
    dg = (g0 - g1).abs().max().item()
    _, _, _, stats = op.events(max_ev=8)
    print(f"{label:36s} rays {op._rays.n:8d} stats {stats}  max rel diff fwd {rel.max().item():.2e}  bad rays {len(bad)}  max |dgrad| {dg:.2e} (max grad {g0.abs().max().item():.1f})", flush=True)
    if len(bad): print("   first bad", bad[:5].tolist(), f0[tuple(bad[0])].item(), f1[tuple(bad[0])].item())

grid, geoms, _ = bench.workload("c5", 1000)
gen = torch.Generator().manual_seed(0)
theta = torch.rand(1000, generator=gen, dtype=torch.float64) * 2 * math.pi
phi = (0.2 + 0.6 * torch.rand(1000, generator=gen, dtype=torch.float64)) * math.pi
order = torch.argsort((phi - math.pi / 2).abs())
for idx, lab in ((order[:2].tolist(), "c5 equatorial cameras"), (order[-2:].tolist(), "c5 high-latitude cameras"),
                 ([500, 501], "c5 cameras 500,501")):
    compare(grid, S.ViewGeomCollection(*[geoms[i] for i in idx]), lab)
g = torch.Generator().manual_seed(3)
n = 400000
xs = torch.randn(n, 3, generator=g, dtype=torch.float64)
xs = xs / xs.norm(dim=-1, keepdim=True) * (30 + 60 * torch.rand(n, 1, generator=g, dtype=torch.float64))
d = -xs + 12 * torch.randn(n, 3, generator=g, dtype=torch.float64)
for shape, kw in (((120, 90, 180), dict(size_r=(2, 25))), ((64, 181, 91), dict(size_r=(0, 20))),
                  ((200, 64, 360), dict(size_r=(5, 28), spacing="log"))):
    compare(S.SphericalGrid(shape, **kw), S.ViewGeom(xs, d), f"random rays, grid {shape}")
