#!/bin/bash
# A/B: two-copy factor (build/old_lib, HEAD 9bb4a17) against the single-copy factor, one 64^3 block (what a GPU holds at N = 8)
set -u
out=gpurun_out; mkdir -p $out
for lib in build/old_lib prealps_b200/lib; do
  echo "== $lib"
  PREALPS_B200_LIBDIR=$PWD/$lib SUBDOMAINS=1 PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 64 1 2 > $out/r02_prof64_$(basename $(dirname $lib))_$(basename $lib).txt 2>&1
  tail -n 75 $out/r02_prof64_$(basename $(dirname $lib))_$(basename $lib).txt | grep -v asm | tail -n 52
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
done
