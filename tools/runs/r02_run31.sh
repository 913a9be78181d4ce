#!/bin/bash
# shared-memory panel ring 2 x 4 KB as the default; tiny panels with one bulk copy
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t31_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 2 $out/r02_t31_kernels.log
for t in 8 16 32 4 1; do timeout 300 python tools/variants.py 128 8 $t 2>&1 | grep " levels "; done
timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
PREALPS_BJ_COPIES=1 timeout 300 python tools/variants.py 128 8 8 2>&1 | grep " levels "
PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 128 1 2 > $out/r02_prof128_smem_ring.txt 2>&1;  tail -n 72 $out/r02_prof128_smem_ring.txt | grep "fwd" | tail -n 28
