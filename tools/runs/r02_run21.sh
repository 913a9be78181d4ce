#!/bin/bash
# single-copy factor, backward A fragments read straight from the forward panels (permuted lane addresses, no shuffles)
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t21_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 2 $out/r02_t21_kernels.log
timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
for t in 8 16 32 1; do timeout 300 python tools/variants.py 128 8 $t 2>&1 | grep " levels "; done
SUBDOMAINS=1 PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 64 1 2 2>&1 | tail -n 75 | grep bwd | tail -n 17
PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 128 1 2 > $out/r02_prof128_single_copy.txt 2>&1;  tail -n 75 $out/r02_prof128_single_copy.txt | grep bwd | tail -n 18
