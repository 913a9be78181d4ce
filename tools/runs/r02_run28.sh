#!/bin/bash
# block-Jacobi sweep: ring depth D (k-blocks in flight per warp) and CTAs per SM: 4/2 (default), 2/3, 2/2
set -u
for lib in prealps_b200/lib build/bjs_d2_o3 build/bjs_d2_o2; do
  echo "== $lib"
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
done
