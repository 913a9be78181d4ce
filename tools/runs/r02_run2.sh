#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t1.log 2>&1; echo "kernel tests rc=$?"
tail -n 3 $out/r02_t1.log
timeout 300 python tools/variants.py 64 1 8 > $out/r02_var_n64.log 2>&1; echo "variants 64 rc=$?"
grep -v METIS $out/r02_var_n64.log | tail -n 3
PREALPS_BJ_DFPROF=1 timeout 300 python tools/variants.py 64 1 8 2>&1 | grep "dataflow apply" | tail -n 2
timeout 600 python tools/variants.py 128 8 8 > $out/r02_var_n128.log 2>&1; echo "variants 128 rc=$?"
grep -v METIS $out/r02_var_n128.log | tail -n 3
PREALPS_BJ_DFPROF=1 timeout 300 python tools/variants.py 128 8 8 2>&1 | grep "dataflow apply" | tail -n 2
