#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "spmm" > $out/r02_t1.log 2>&1; echo "spmm tests rc=$?"; tail -n 2 $out/r02_t1.log
for sh in 0 1 2 3 4; do
  PREALPS_SPMM_SHAPE=$sh timeout 300 python tools/spmm_sweep.py 128 8,16,32 > $out/r02_spmm_shape$sh.jsonl 2> $out/r02_spmm_shape$sh.err
  python - <<PY
import json
print("shape $sh:", " | ".join("%s t=%d %.1f us %.3f" % (d["operator"].split()[0], d["t"], d["us"], d["frac_of_measured_peak"]) for d in map(json.loads, open("$out/r02_spmm_shape$sh.jsonl"))))
PY
done
python tools/spmm_once.py 1 128 8 3 > $out/spmm_once.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'spmm_bulk' -s 2 -c 1 -o $out/r02_prof_spmm27_t8 -f python tools/spmm_once.py 1 128 8 3 > $out/ncu_spmm8.log 2>&1
python tools/spmm_once.py 1 128 32 3 >> $out/spmm_once.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'spmm_bulk' -s 2 -c 1 -o $out/r02_prof_spmm27_t32 -f python tools/spmm_once.py 1 128 32 3 > $out/ncu_spmm32.log 2>&1
python tools/spmm_once.py 0 128 8 3 >> $out/spmm_once.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'spmm_bulk' -s 2 -c 1 -o $out/r02_prof_spmm7_t8 -f python tools/spmm_once.py 0 128 8 3 > $out/ncu_spmm7.log 2>&1
cat $out/spmm_once.log; ls -la $out/*.ncu-rep | tail -n 4
