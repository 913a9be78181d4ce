#!/bin/bash
# no tiny-panel kernel at all (PREALPS_BJ_NOTINY=1) against folding the tiny panels into the regular launch whenever a level has other panels
set -u
for cfg in "PREALPS_BJ_TINYFOLD=1000000" "PREALPS_BJ_NOTINY=1"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels " | cut -c1-60
  env $cfg timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels " | cut -c1-60
  env $cfg timeout 300 python tools/variants.py 128 8 16 2>&1 | grep " levels " | cut -c1-60
done
