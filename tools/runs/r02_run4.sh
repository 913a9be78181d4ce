#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
export PREALPS_BJ_LEVELS=1
for leaf in 32 48 64 96; do
  for rx in "default" "0.05:256" "0.1:256"; do
    echo "== leaf=$leaf relax_big=$rx"
    PREALPS_BJ_LEAF=$leaf timeout 300 python tools/variants.py 128 8 8 "$rx" 2>&1 | grep " levels " 
    PREALPS_BJ_LEAF=$leaf timeout 300 python tools/variants.py 64 1 8 "$rx" 2>&1 | grep " levels "
  done
done
