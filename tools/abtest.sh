#!/bin/bash
# A/B timing of tuning variants on one box: tools/abtest.sh spec name1 name2 ...
spec=$1; shift
for rep in 1; do
  for v in "$@"; do
    SPHRT_WALKER=coop SPHRT_LIB=variants/$v.so python tools/kbench.py $spec 2>&1 | tail -1
  done
done
