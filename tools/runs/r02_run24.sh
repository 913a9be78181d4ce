#!/bin/bash
# leaves of the forest read the caller's block directly (no level-0 assembly launch): A/B, then the whole GPU suite, bench, ncu
set -u
out=gpurun_out; mkdir -p $out
for v in 1 0; do
  if [ $v = 1 ]; then export PREALPS_BJ_NO_LEAF_DIRECT=1; else unset PREALPS_BJ_NO_LEAF_DIRECT; fi
  echo "== leaf levels through the assembly kernel: $v"
  timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
  timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
done
unset PREALPS_BJ_NO_LEAF_DIRECT
bash tools/runs/r02_run17.sh
