#!/usr/bin/env python3
"""Registers / stack / spills per kernel from a verbose build: python tools/ptxas_info.py [filter] [nvcc -D flags...]"""
import re, subprocess, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sph_raytracer_b200 import build as B
flt = sys.argv[1] if len(sys.argv) > 1 else ""
flags = sys.argv[2:]
out = "/tmp/ptxas_info.so" if flags or os.environ.get("PTXAS_TMP") else None
import io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    B.build(force=True, verbose=True, extra_flags=flags, out=out)
txt = buf.getvalue()
cur = None
for line in txt.split("\n"):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur); cur = re.sub(r"\(sphrt_grid.*", "", cur)
        continue
    if cur and "spill" in line: spill = line.strip(); continue
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        if flt in cur: print(f"{cur[:70]:70s} regs {m.group(1):>3s}  {spill}")
        cur = None
