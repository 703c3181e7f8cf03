#!/usr/bin/env python3
"""Aggregate `ncu --page source --print-source cuda,sass --csv` output per source REGION of the
second-generation warp walker (phases of warp_walk_ray, the kernel loop, the exact primitives).
usage: ncu_regions.py file.csv [kernel-substring] [segments]"""
import csv, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
nseg = float(sys.argv[3]) if len(sys.argv) > 3 else 0
import os
HERE = os.path.dirname(os.path.abspath(__file__))
WARP = os.path.join(HERE, "..", "sph_raytracer_b200", "csrc", "sphrt_warp.cuh")


def warp_regions():
    """Line ranges of the phases of warp_walk_ray, found by their marker comments."""
    marks = [("__device__ __forceinline__ int warp_corank", "co-rank (both stages)"),
             ("__device__ __forceinline__ int warp_walk_ray", "per-ray setup"),
             ("// ---- elevation:", "elevation candidates (list build)"),
             ("// ---- azimuth:", "azimuth candidates (list build)"),
             ("// ---- stage 1:", "stage 1: merge E+A"),
             ("// ---- radial list", "radial list"),
             ("// ---- stage 2:", "stage 2: merge R+EA, segments"),
             ("}  // namespace sphrt", None)]
    lines = open(WARP).read().split("\n")
    pos = []
    for text, name in marks:
        for k, l in enumerate(lines):
            if text in l:
                pos.append((k + 1, name)); break
    return [("sphrt_warp.cuh", pos[k][0], pos[k + 1][0] - 1, pos[k][1]) for k in range(len(pos) - 1)]


REGIONS = warp_regions() + [
    ("sphrt_warp.cuh", 0, 100000, "shared-memory helpers (ld/st.shared asm)"),
    ("sphrt_walker.cuh", 154, 230, "cone primitives (quadratic, roots, region, shadow)"),
    ("sphrt_walker.cuh", 231, 255, "plane_event (seam)"),
    ("sphrt_walker.cuh", 60, 90, "fma/dot/norm helpers"),
    ("sphrt_fast.cuh", 0, 400, "FastAzimuth helpers"),
    ("sphrt_kernels.cuh", 120, 420, "sink (gather / RED, reduction, store)"),
    ("sphrt_kernels.cuh", 80, 112, "stage_tables"),
    ("sphrt_kernels.cuh", 421, 100000, "kernel loop (hand-out, record load)"),
]
rows = list(csv.reader(open(path)))
cur_file = cur_fn = None; hdr = None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0]))
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10 or not r[0].isdigit(): continue
    ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    line = int(r[0])
    name = f"other {cur_file}"
    for f, a, b, n in REGIONS:
        if f == cur_file and a <= line <= b: name = n; break
    try:
        v = agg[cur_fn][name]; v[0] += int(r[ti]); v[1] += int(r[si]); v[2] += int(r[ii])
    except ValueError:
        pass
for fn, regs in agg.items():
    if want not in fn: continue
    tt = sum(v[0] for v in regs.values()); ts = sum(v[1] for v in regs.values()); tw = sum(v[2] for v in regs.values())
    print(f"== {fn[:100]}\n   thread-inst {tt:.4g} warp-inst {tw:.4g} samples {ts}" + (f"  thread-inst/segment {tt / nseg:.1f}" if nseg else ""))
    for n, v in sorted(regs.items(), key=lambda kv: -kv[1][1]):
        extra = f"  {v[0] / nseg:6.1f} thr-inst/seg" if nseg else ""
        print(f"  {100 * v[1] / max(ts, 1):5.1f}% samples  {100 * v[2] / max(tw, 1):5.1f}% warp-inst  lanes {v[0] / max(v[2], 1):4.1f}{extra}  {n}")
