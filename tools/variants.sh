#!/bin/bash
# build tuning variants of libsphrt.so into variants/ : tools/variants.sh name "-DFLAG=..." [name2 "flags2" ...]
mkdir -p variants
while [ $# -ge 2 ]; do
  python - "$1" $2 <<'PY' &
import sys
sys.path.insert(0, ".")
from sph_raytracer_b200 import build
print(build.build(force=True, extra_flags=sys.argv[2:], out="variants/%s.so" % sys.argv[1]))
PY
  shift 2
done
wait
ls -la variants/
