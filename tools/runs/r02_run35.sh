#!/bin/bash
# assembly of the levels with long gather lists: one warp per column (PREALPS_BJ_ASM_WIDE=1, default) against 4 lanes per column (=0)
set -u
out=gpurun_out; mkdir -p $out
PREALPS_BJ_ASM_WIDE=2 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t35_kernels.log 2>&1; echo "kernel tests (wide assembly on every level) rc=$?"; tail -n 2 $out/r02_t35_kernels.log
for wide in 0 1; do
  echo "== PREALPS_BJ_ASM_WIDE=$wide"
  PREALPS_BJ_ASM_WIDE=$wide timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
  PREALPS_BJ_ASM_WIDE=$wide timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
done
