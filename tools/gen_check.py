#!/usr/bin/env python3
"""Second-generation warp walker (csrc/sphrt_warp.cuh) against the first (csrc/sphrt_coop.cuh):
event export, integrals and gradients in float64 on the BASELINE configurations and on random
grids, then kernel timings of both.

    python tools/gen_check.py [c5:2] [c2] [c1] [c4] [rand]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import sph_raytracer_b200 as S  # noqa: E402


def setgen(gen):
    os.environ["SPHRT_COOP_GEN"] = str(gen)


def compare(op, shape, label, events=True):
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(shape, generator=gen, device="cuda", dtype=torch.float64)
    w = torch.rand(op._rays.shape, generator=gen, device="cuda", dtype=torch.float64)
    res = {}
    for g in (1, 2):
        setgen(g)
        res[g] = (op._apply_forward(x), op._apply_adjoint(w, shape))
    dy = (res[1][0] - res[2][0]).abs().max().item()
    sy = res[1][0].abs().max().item()
    dg = (res[1][1] - res[2][1]).abs().max().item()
    sg = res[1][1].abs().max().item()
    ok = dy <= 1e-12 * max(sy, 1) and dg <= 1e-11 * max(sg, 1)
    msg = f"{label:10s} rays {op._rays.n:9d}  fwd max|d| {dy:.2e} (of {sy:.2e})  adj max|d| {dg:.2e} (of {sg:.2e})"
    if events and op._rays.n <= 300000:
        ev = {}
        for g in (1, 2):
            setgen(g)
            ev[g] = op.events()
        c1, c2 = ev[1][2], ev[2][2]
        same_c = torch.equal(c1, c2)
        k = torch.arange(ev[1][0].shape[1], device="cuda")[None, :] < c1.clamp(min=0)[:, None]
        # compare only events with a non-zero length (the voxel of a zero-length event is arbitrary)
        l1, l2 = ev[1][1], ev[2][1]
        live = k & ((l1 > 0) | (l2 > 0))
        same_v = bool(((ev[1][0] == ev[2][0]).all(-1) | ~live).all())
        same_l = bool(((l1 == l2) | ~k).all())
        msg += f"  events: counts {'==' if same_c else '!='} vox {'==' if same_v else '!='} len {'==' if same_l else '!='}"
        msg += f"  stats gen1 {ev[1][3]} gen2 {ev[2][3]}"
        ok = ok and same_c and same_v and same_l
    print(("ok   " if ok else "FAIL ") + msg, flush=True)
    return ok


def timing(op, shape, label, reps=7):
    x = torch.rand(shape, device="cuda")
    y = torch.rand(op._rays.shape, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for g in (1, 2):
        setgen(g)
        tf, ta = [], []
        for k in range(reps + 2):
            a, b, c = ev(), ev(), ev()
            a.record(); op._apply_forward(x); b.record(); op._apply_adjoint(y, shape); c.record()
            torch.cuda.synchronize()
            if k >= 2:
                tf.append(a.elapsed_time(b)); ta.append(b.elapsed_time(c))
        print(f"     {label:10s} gen{g}: fwd {np.median(tf):8.3f} ms  adj {np.median(ta):8.3f} ms", flush=True)


def main():
    os.environ["SPHRT_WALKER"] = "coop"
    os.environ["SPHRT_AUTOTUNE"] = "0"
    specs = sys.argv[1:] or ["c1", "c2", "c4", "c5:2", "rand"]
    all_ok = True
    for spec in specs:
        if spec == "rand":
            gen = torch.Generator().manual_seed(0)
            n = 20000
            xs = torch.randn(n, 3, generator=gen, dtype=torch.float64)
            xs = xs / xs.norm(dim=-1, keepdim=True) * (3 + 4 * torch.rand(n, 1, generator=gen, dtype=torch.float64))
            d = -xs + 1.2 * torch.randn(n, 3, generator=gen, dtype=torch.float64)
            for grid in (S.SphericalGrid((20, 16, 24), size_r=(0, 2.5)),
                         S.SphericalGrid((30, 31, 33), size_r=(0.5, 2)),
                         S.SphericalGrid((12, 40, 17), spacing="log", size_r=(0.1, 2.9)),
                         S.SphericalGrid((64, 181, 91), size_r=(0, 2.9)),
                         S.SphericalGrid((9, 12, 64), size_r=(1, 2), size_e=(0.4, 2.6))):
                op = S.Operator(grid, S.ViewGeom(xs, d))
                all_ok &= compare(op, tuple(grid.shape), "rand" + str(tuple(grid.shape)))
            continue
        name, _, obs = spec.partition(":")
        grid, geoms, _ = bench.workload(name, int(obs) if obs else None)
        geom = S.ViewGeomCollection(*geoms) if len(geoms) > 1 or grid.dynamic else geoms[0]
        op = S.Operator(grid, geom)
        shape = tuple(grid.shape)
        all_ok &= compare(op, shape, spec)
        # several chunks per launch
        os.environ["SPHRT_WARP_CHUNK"] = "40000"
        op2 = S.Operator(grid, geom)
        all_ok &= compare(op2, shape, spec + "/chunk", events=False)
        os.environ.pop("SPHRT_WARP_CHUNK")
        timing(op, shape, spec)
    print("ALL OK" if all_ok else "SOME FAILED")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
