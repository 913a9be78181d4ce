#!/bin/bash
set -u
for sk in 512 256 128 1024; do
  echo "== splitk $sk"
  PREALPS_BJ_SPLITK=$sk timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  PREALPS_BJ_SPLITK=$sk timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
