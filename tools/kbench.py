#!/usr/bin/env python3
"""Kernel-level timing of forward / adjoint (CUDA events, per call) for tuning experiments.

    SPHRT_LIB=path/to/variant.so python tools/kbench.py [c5:8] [c2] [c1]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import sph_raytracer_b200 as S  # noqa: E402


def run(spec, reps=7):
    name, _, obs = spec.partition(":")
    grid, geoms, _ = bench.workload(name, int(obs) if obs else None)
    geom = S.ViewGeomCollection(*geoms) if len(geoms) > 1 or grid.dynamic else geoms[0]
    store = os.environ.get('SPHRT_STORE')            # 1 (8 B/segment list) | packed | auto
    op = S.Operator(grid, geom, store_segments=(store if store in ('packed', 'auto') else bool(store)))
    shape = tuple(grid.shape)
    x = torch.rand(shape, device="cuda")
    y = torch.rand(op._rays.shape, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tf, ta = [], []
    for k in range(reps + 2):
        a, b, c = ev(), ev(), ev()
        a.record(); op._apply_forward(x); b.record(); op._apply_adjoint(y, shape); c.record()
        torch.cuda.synchronize()
        if k >= 2:
            tf.append(a.elapsed_time(b)); ta.append(b.elapsed_time(c))
    if os.environ.get('SPHRT_NORMAL'):           # the fused normal-equations step next to fwd + adj
        tn = []
        for k in range(reps + 2):
            a, b = ev(), ev()
            a.record(); op.normal(x, b=y, scale=0.5); b.record()
            torch.cuda.synchronize()
            if k >= 2:
                tn.append(a.elapsed_time(b))
        print(f"{'':40s} {spec:8s} fused normal step {np.median(tn):8.3f} ms (min {min(tn):.3f})  vs fwd + adj {np.median(tf) + np.median(ta):.3f} ms")
    _, nseg = op.segment_counts()
    mode = os.environ.get('SPHRT_LIB', 'default') + (' stored %.2f GB' % (op.stored_bytes / 1e9) if op.stored_bytes else '')
    print(f"{mode:40s} {spec:8s} fwd {np.median(tf):8.3f} ms (min {min(tf):.3f})  "
          f"adj {np.median(ta):8.3f} ms (min {min(ta):.3f})  seg/s fwd+adj {2 * nseg / (np.median(tf) + np.median(ta)) * 1e3:.3e}")


if __name__ == "__main__":
    for spec in (sys.argv[1:] or ["c5:8", "c2"]):
        run(spec)
