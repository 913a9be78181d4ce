#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu -s > $out/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "max relative|passed|failed|rror" $out/r02_gpu_tests.log | tail -n 6
timeout 900 python bench.py > $out/r02_bench_n1.json 2> $out/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n1.json"))
print("it/s %.1f ms/step %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, "roofline frac %.3f stored %.3f" % (d["roofline"]["frac"], d["roofline"]["achieved_counting_stored_bytes"] / d["roofline"]["peak"]), d["clocks"], d["setup"], d["cpu_baseline"]["value"])
PY
bash tools/ncu_capture.sh > $out/ncu_capture.log 2>&1; tail -n 4 $out/ncu_capture.log
