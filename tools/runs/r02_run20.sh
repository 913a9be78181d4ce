#!/bin/bash
# single-copy factor: weighted backward work units; two accumulator sets (prealps_b200/lib) against one (build/var_lib)
set -u
out=gpurun_out; mkdir -p $out
for lib in prealps_b200/lib build/var_lib; do
  echo "== $lib"
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
SUBDOMAINS=1 PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 64 1 2 2>&1 | tail -n 75 | grep bwd | tail -n 17
PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 128 1 2 2>&1 | tail -n 75 | grep bwd | tail -n 18
