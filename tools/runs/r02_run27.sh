#!/bin/bash
# SpMM row-loop unroll at 8 CTAs/SM: 4 columns per lane (7-point) unroll 2 (default now) / 1 (a) / 3 (b); 2 columns per lane (27-point) unroll 4 (default) / 2 (c) / 8 (d)
set -u
for lib in prealps_b200/lib build/spmm_a build/spmm_b build/spmm_c build/spmm_d; do
  echo "== $lib"
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/spmm_sweep.py 128 8,16,32 2>&1 | cut -c1-150
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "spmm" 2>&1 | tail -n 2
