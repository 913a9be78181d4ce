#!/bin/bash
# SpMM row phase: gathers in flight per lane (unroll) against resident CTAs (minimum blocks per SM); block-Jacobi tests with the
# interval-allocated update matrices
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t26_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 2 $out/r02_t26_kernels.log
for lib in prealps_b200/lib build/spmm_u8_b5 build/spmm_u8_b4 build/spmm_u7_b6 build/spmm_u2_b8; do
  echo "== $lib"
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/spmm_sweep.py 128 8,16 2>&1 | cut -c1-140
done
timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
