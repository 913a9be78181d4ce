#!/bin/bash
# block-Jacobi sweeps: a warp streams across several panels (PREALPS_BJ_PW = most panels per warp)
set -u
out=gpurun_out; mkdir -p $out
PREALPS_BJ_PW_FORCE=3 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t23_kernels.log 2>&1; echo "kernel tests (3 panels per warp forced) rc=$?"; tail -n 2 $out/r02_t23_kernels.log
for pw in 1 2 4 8; do
  echo "== at most $pw panels per warp"
  PREALPS_BJ_PW=$pw timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
PREALPS_BJ_PW=4 timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
PREALPS_BJ_PW=4 PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 128 1 2 > $out/r02_prof128_pw4.txt 2>&1;  tail -n 75 $out/r02_prof128_pw4.txt | grep -v asm | tail -n 52
