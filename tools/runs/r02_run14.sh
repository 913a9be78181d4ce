#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" 2>&1 | tail -n 2
for v in "" "PREALPS_BJ_NOLPT=1"; do
  echo "== $v"
  env $v timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  env $v timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
