#!/bin/bash
set -u
for a in 2 3 4 6; do
  echo "== split_kb = clamp(q * $a / 8, 32, 256)"
  PREALPS_BJ_SPLITA=$a timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  PREALPS_BJ_SPLITA=$a timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
