#!/bin/bash
# shared-memory panel ring: stages x k-blocks per stage, input tile rows, CTAs per SM
#   default 4 x 2, 32 rows, 2 CTAs | a: 3 x 2, 16 rows, 3 CTAs | b: 8 x 1, 32, 2 | c: 2 x 4, 32, 2 | d: 6 x 1, 16 rows, 3 CTAs
set -u
for lib in prealps_b200/lib build/ring_a build/ring_b build/ring_c build/ring_d; do
  echo "== $lib"
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
done
