#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t1.log 2>&1; echo "kernel tests rc=$?"
tail -n 3 $out/r02_t1.log
for cq in 8 4 2 1; do
  echo "== chunk = $cq q"
  PREALPS_BJ_CHUNKQ=$cq timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
  PREALPS_BJ_CHUNKQ=$cq timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
done
