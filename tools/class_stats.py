import os, sys
sys.path.insert(0, "/root/repo")
os.environ["SPHRT_AUTOTUNE"]="0"
import torch, bench, sph_raytracer_b200 as S
for name in ["c2","c1","c4"]:
    grid, geoms, _ = bench.workload(name, None)
    geom = S.ViewGeomCollection(*geoms) if len(geoms) > 1 or grid.dynamic else geoms[0]
    op = S.Operator(grid, geom)
    vox, lens, count, stats = op.events(max_ev=8)
    print(name, "rays", op._rays.n, "stats (coop, listed, handed, failed)", stats, "count>=0:", int((count>=0).sum()), "miss(0):", int((count==0).sum()))
