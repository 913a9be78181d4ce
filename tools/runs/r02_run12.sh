#!/bin/bash
set -u
for cq in 8 10 12 16; do
  echo "== chunk = $cq q"
  PREALPS_BJ_CHUNKQ=$cq timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  PREALPS_BJ_CHUNKQ=$cq timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
