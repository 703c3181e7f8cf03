#!/usr/bin/env python3
"""Aggregate `ncu --page source --print-source cuda,sass --csv` output per source line.
usage: ncu_lines.py file.csv [kernel-substring] [top-n]"""
import csv, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
rows = list(csv.reader(open(path)))
cur_file = cur_fn = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])   # (fn,file,line) -> [thread_inst, samples, inst, src]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] in ("", "-") or not r[0].isdigit(): continue
    ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    try:
        a = agg[(cur_fn, cur_file, int(r[0]))]
        a[0] += int(r[ti]); a[1] += int(r[si]); a[2] += int(r[ii]); a[3] = r[1]
    except ValueError:
        pass
fns = sorted({k[0] for k in agg})
for fn in fns:
    if want not in fn: continue
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    tot_t = sum(v[0] for _, v in items); tot_s = sum(v[1] for _, v in items); tot_i = sum(v[2] for _, v in items)
    print(f"== {fn[:110]}\n   thread-inst {tot_t:.4g}  warp-inst {tot_i:.4g}  samples {tot_s}")
    for (f_, file, line), v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
        print(f"{100*v[1]/max(tot_s,1):5.1f}%s {100*v[0]/max(tot_t,1):5.1f}%t {100*v[2]/max(tot_i,1):5.1f}%w  avg_thr {v[0]/max(v[2],1):4.1f}  {file}:{line:<4d} {v[3].strip()[:100]}")
