#!/bin/bash
# single-copy factor (one set of panels for both sweeps): kernel tests, per-level profile, apply timings, bench
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "block_jacobi" > $out/r02_t18_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 3 $out/r02_t18_kernels.log
for t in 8 16 32 4 1; do timeout 300 python tools/variants.py 128 8 $t 2>&1 | grep " levels "; done
timeout 300 python tools/variants.py 64 1 8 2>&1 | grep " levels "
PREALPS_BJ_PROFILE=1 timeout 300 python tools/profile_apply.py 128 1 2 > $out/r02_prof128_single_copy.txt 2>&1; tail -n 72 $out/r02_prof128_single_copy.txt | head -n 71
timeout 900 python bench.py > $out/r02_bench_n1_single_copy.json 2> $out/r02_bench_n1_single_copy.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1_single_copy.json"))
print("it/s %.1f ms/step %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, "roofline frac %.3f" % d["roofline"]["frac"], d["e2e"]["iterations"], d["e2e"]["final_res"])
PY
