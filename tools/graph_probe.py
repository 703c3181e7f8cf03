#!/usr/bin/env python3
"""Is a forward + adjoint pass capturable in a CUDA graph, and what does replaying it save on the
small configurations?  python tools/graph_probe.py [c2] [c1] [c4]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, sph_raytracer_b200 as S

for name in (sys.argv[1:] or ["c2"]):
    grid, geoms, _ = bench.workload(name, None)
    geom = S.ViewGeomCollection(*geoms) if len(geoms) > 1 or grid.dynamic else geoms[0]
    op = S.Operator(grid, geom)
    shape = tuple(grid.shape)
    x = torch.rand(shape, device="cuda"); y = torch.rand(op._rays.shape, device="cuda")
    for _ in range(3):                                   # warm-up: autotune done, plan cache filled
        out = op._apply_forward(x); g = op._apply_adjoint(y, shape)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_g = op._apply_forward(x); g_g = op._apply_adjoint(y, shape)
    graph.replay(); torch.cuda.synchronize()
    same = torch.equal(out_g, out) and torch.allclose(g_g, g, rtol=1e-5, atol=1e-6 * float(g.abs().max()))
    def timed(fn, n=200):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
    eager = timed(lambda: (op._apply_forward(x), op._apply_adjoint(y, shape)))
    replay = timed(graph.replay)
    print(f"{name}: graph replay equals eager: {same};  wall time per forward + adjoint: eager {eager:.3f} ms, graph replay {replay:.3f} ms")
