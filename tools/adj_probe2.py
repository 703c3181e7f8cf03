import os, sys, math, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, sph_raytracer_b200 as S
grid, geoms, _ = bench.workload("c5", 1000)
gen = torch.Generator().manual_seed(0)
theta = torch.rand(1000, generator=gen, dtype=torch.float64) * 2 * math.pi
phi = (0.2 + 0.6 * torch.rand(1000, generator=gen, dtype=torch.float64)) * math.pi
order = torch.argsort((phi - math.pi / 2).abs())
which = sys.argv[1]
idx = order[:8].tolist() if which == "eq" else order[-8:].tolist()
op = S.Operator(grid, S.ViewGeomCollection(*[geoms[i] for i in idx]))
print(which, "phi", [round(float(phi[i]) / math.pi, 3) for i in idx])
_, _, _, stats = op.events(max_ev=8)
print("stats (coop, listed, handed, failed)", stats, "of", op._rays.n)
shape = tuple(grid.shape)
x = torch.rand(shape, device="cuda"); y = torch.rand(op._rays.shape, device="cuda")
for k in range(2):
    op._apply_forward(x); op._apply_adjoint(y, shape)
torch.cuda.synchronize()
