#!/bin/bash
# sweep kernel streaming across several panels per warp (PREALPS_BJ_PW = most panels per warp; 1 = off)
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t34_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 2 $out/r02_t34_kernels.log
for pw in 1 2 4 8; do
  echo "== at most $pw panels per warp"
  PREALPS_BJ_PW=$pw timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
PREALPS_BJ_PW=4 timeout 300 python tools/variants.py 128 8 16 2>&1 | grep " levels "
PREALPS_BJ_PW=1 timeout 300 python tools/variants.py 128 8 16 2>&1 | grep " levels "
timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
