#!/bin/bash
# dataflow apply: parity first (bounded), then A/B against the level-by-level launches
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi or spmm" > $out/r02_t1.log 2>&1; echo "kernel tests rc=$?"
tail -n 5 $out/r02_t1.log
timeout 300 python tools/variants.py 64 1 8 "default,0.05:256" > $out/r02_var_n64.log 2>&1; echo "variants 64 rc=$?"
grep -v METIS $out/r02_var_n64.log | tail -n 6
timeout 600 python tools/variants.py 128 8 8 "default,0.05:256" > $out/r02_var_n128.log 2>&1; echo "variants 128 rc=$?"
grep -v METIS $out/r02_var_n128.log | tail -n 6
timeout 900 python -m pytest tests/test_gpu_ecg.py -q -m gpu -x -s > $out/r02_t2.log 2>&1; echo "ecg tests rc=$?"
grep -E "max relative|passed|failed|error" $out/r02_t2.log | tail -n 12
