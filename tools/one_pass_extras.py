#!/usr/bin/env python3
"""One warm-up and one measured pass of the fused normal step and of the packed stored operator
(for ncu captures):  python tools/one_pass_extras.py c5:16"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import sph_raytracer_b200 as S  # noqa: E402

os.environ.setdefault("SPHRT_AUTOTUNE", "0")
spec = sys.argv[1] if len(sys.argv) > 1 else "c5:16"
name, _, obs = spec.partition(":")
grid, geoms, _ = bench.workload(name, int(obs) if obs else None)
geom = S.ViewGeomCollection(*geoms) if len(geoms) > 1 or grid.dynamic else geoms[0]
op = S.Operator(grid, geom)
pk = S.Operator(grid, geom, store_segments="packed")
shape = tuple(grid.shape)
x = torch.rand(shape, device="cuda")
y = torch.rand(op._rays.shape, device="cuda")
for k in range(2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(); op.normal(x, b=y, scale=0.5); ev[1].record()
    pk._apply_forward(x); ev[2].record(); pk._apply_adjoint(y, shape); ev[3].record()
    torch.cuda.synchronize()
    print(f"{spec} pass {k}: normal {ev[0].elapsed_time(ev[1]):.3f} ms  packed fwd {ev[1].elapsed_time(ev[2]):.3f} ms  "
          f"packed adj {ev[2].elapsed_time(ev[3]):.3f} ms")
