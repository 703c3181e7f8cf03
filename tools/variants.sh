#!/bin/bash
# build tuning variants of libsphrt.so into variants/ : tools/variants.sh name "-DFLAG=..." [name2 "flags2" ...]
mkdir -p variants
while [ $# -ge 2 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -shared -diag-suppress 177 $2 -o variants/$1.so sph_raytracer_b200/csrc/sphrt_kernels.cu &
  shift 2
done
wait
ls -la variants/
