#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -n 3
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $out/r02_multi2.log 2>&1; echo "multi tests rc=$?"
tail -n 5 $out/r02_multi2.log
run() { PREALPS_B200_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu-baseline "$@"; }
run > $out/r02_bench_n2.json 2> $out/r02_bench_n2.err; echo "bench n2 rc=$?"
PREALPS_SPMM_OVERLAP=1 run > $out/r02_bench_n2_overlap.json 2> $out/r02_bench_n2_overlap.err; echo "bench n2 overlap rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n2", "r02_bench_n2_overlap"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], "iters", d["e2e"]["iterations"],
              {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d["setup"])
    except Exception as e:
        print(f, "failed", e)
PY
grep "setup:" $out/r02_bench_n2.err | head -n 14
