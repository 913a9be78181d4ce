#!/bin/bash
# block-Jacobi sweep: panel data through a shared-memory ring fed by cp.async.bulk (prealps_b200/lib) against the register ring (build/base_lib)
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t29_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 2 $out/r02_t29_kernels.log
for lib in build/base_lib prealps_b200/lib; do
  echo "== $lib"
  for t in 8 16 4; do PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 128 8 $t 2>&1 | grep " levels "; done
  PREALPS_B200_LIBDIR=$PWD/$lib timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
done
PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 128 1 2 > $out/r02_prof128_smem_ring.txt 2>&1;  tail -n 72 $out/r02_prof128_smem_ring.txt | grep -v asm | tail -n 50
