#!/usr/bin/env python3
"""SASS listing in address order with source line, executed count and stall samples, from
`ncu --page source --print-source cuda,sass --csv` output.
usage: ncu_sass.py file.csv kernel-substring [file:lo-hi]   (optional source-line filter)"""
import csv, sys
path, want = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
cur_file = cur_fn = None; hdr = None; out = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or want not in cur_fn: continue
    if r[0].isdigit(): cur_line = int(r[0]); continue
    if r[0] != "" or not r[2].startswith("0x"): continue
    ai = hdr.index("Address"); si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    ti = hdr.index("Thread Instructions Executed")
    try:
        out[int(r[ai], 16)] = (cur_file, cur_line, r[ai + 1], int(r[ii]), int(r[si]), int(r[ti]))
    except ValueError:
        pass
if flt:
    f, rng = flt.split(":"); lo, hi = map(int, rng.split("-"))
tot = sum(v[3] for v in out.values())
for a in sorted(out):
    f_, line, sass, n, s, t = out[a]
    if flt and not (f_ == f and lo <= line <= hi): continue
    print(f"{a & 0xfffff:05x} {f_[6:-4]:>8}:{line:<4d} {n:>10d} {100*n/tot:5.2f}% s{s:<5d} l{t/max(n,1):4.1f} {sass[:90]}")
