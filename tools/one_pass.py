#!/usr/bin/env python3
"""One warm-up and one measured forward + adjoint pass of a workload (for ncu captures).

    SPHRT_COOP_GEN=2 python tools/one_pass.py c5:2
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import sph_raytracer_b200 as S  # noqa: E402

os.environ.setdefault("SPHRT_AUTOTUNE", "0")
spec = sys.argv[1] if len(sys.argv) > 1 else "c5:2"
name, _, obs = spec.partition(":")
grid, geoms, _ = bench.workload(name, int(obs) if obs else None)
geom = S.ViewGeomCollection(*geoms) if len(geoms) > 1 or grid.dynamic else geoms[0]
op = S.Operator(grid, geom)
shape = tuple(grid.shape)
x = torch.rand(shape, device="cuda")
y = torch.rand(op._rays.shape, device="cuda")
for k in range(2):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(); op._apply_forward(x); b.record(); op._apply_adjoint(y, shape); c.record()
    torch.cuda.synchronize()
    print(f"{spec} pass {k}: fwd {a.elapsed_time(b):.3f} ms  adj {b.elapsed_time(c):.3f} ms")
