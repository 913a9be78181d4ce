#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for w in 1; do for sh in 0 3 4; do
  PREALPS_SPMM_WIDE=$w PREALPS_SPMM_SHAPE=$sh timeout 300 python tools/spmm_sweep.py 128 8,16,32 > $out/r02_spmm_w${w}_shape$sh.jsonl 2> $out/r02_spmm_w${w}_shape$sh.err
  python - <<PY
import json
print("wide $w shape $sh:", " | ".join("%s t=%d %.1f us %.3f" % (d["operator"].split()[0], d["t"], d["us"], d["frac_of_measured_peak"]) for d in map(json.loads, open("$out/r02_spmm_w${w}_shape$sh.jsonl"))))
PY
done; done
