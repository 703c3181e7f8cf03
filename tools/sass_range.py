#!/usr/bin/env python3
"""Static SASS instruction count of a source-line range of a kernel (in-tree library):
    python tools/sass_range.py forward|adjoint|api <mangled-kernel-substring> <file.cuh> <lo> <hi>
Counts the instructions between the first and the last instruction attributed to lines lo..hi."""
import os, re, subprocess, sys, tempfile
tu, pat, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.environ.get("SPHRT_LIB") or os.path.join(root, "sph_raytracer_b200", "libsphrt.so")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith("sphrt_" + tu)][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
inside = False; cur = (None, 0); rows = []
for line in dis.split("\n"):
    if line.startswith(".text."):
        inside = pat in line
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        rows.append((int(m.group(1), 16), cur, m.group(2).strip()))
hit = [i for i, (a, (f, l), t) in enumerate(rows) if f == fname and lo <= l <= hi]
if not hit:
    sys.exit("no instruction attributed to that range")
a, b = hit[0], hit[-1]
print(f"{b - a + 1} instructions between {rows[a][0]:#x} and {rows[b][0]:#x} ({len(hit)} of them attributed to {fname}:{lo}-{hi}); kernel total {len(rows)}")
if os.environ.get("SASS_DUMP"):
    for addr, (f, l), t in rows[a:b + 1]:
        print(f"{addr:05x} {f[6:-4]:>8}:{l:<4d} {t}")
