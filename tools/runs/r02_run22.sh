#!/bin/bash
# factor with the transposed copy (default when memory allows) against the single copy (PREALPS_BJ_COPIES=1)
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t22_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 2 $out/r02_t22_kernels.log
for c in 2 1; do
  echo "== copies $c"
  PREALPS_BJ_COPIES=$c timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  PREALPS_BJ_COPIES=$c timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
  PREALPS_BJ_COPIES=$c timeout 300 python tools/variants.py 128 8 16 2>&1 | grep " levels "
done
