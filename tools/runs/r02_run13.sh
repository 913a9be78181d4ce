#!/bin/bash
set -u
for tf in 512 2048 8192 32768; do
  echo "== tiny fold = $tf"
  PREALPS_BJ_TINYFOLD=$tf timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  PREALPS_BJ_TINYFOLD=$tf timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
