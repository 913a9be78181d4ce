#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu -x -s > $out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "max relative|passed|failed|error" $out/r02_gpu_tests.log | tail -n 8
timeout 900 python bench.py > $out/r02_bench_n1.json 2> $out/r02_bench_n1.err; echo "bench rc=$?"
cut -c1-3000 $out/r02_bench_n1.json
tail -n 3 $out/r02_bench_n1.err
