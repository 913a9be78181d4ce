#!/bin/bash
# one 8-GPU session at HEAD: multi-GPU parity at world 8, the headline bench at N = 8 and N = 4, BASELINE config 5 (Poisson 256^3) on FOUR GPUs
# (two 128^3 blocks per GPU: fits with one copy of the factor and the interval-allocated update matrices)
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l; nproc
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "8]" > $out/r02_multi8.log 2>&1; echo "multi tests rc=$?"; tail -n 3 $out/r02_multi8.log
run() { np=$1; name=$2; shift 2; PREALPS_B200_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $np --no-cpu-baseline "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc=$?"; }
run 8 r02_bench_n8
run 4 r02_bench_n4
run 4 r02_config5_poisson256_t8_n4 --grid 256 --steps 20 --warmup 3
python - <<'PY'
import json
for f in ("r02_bench_n8", "r02_bench_n4", "r02_config5_poisson256_t8_n4"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.1f it/s, %d iterations in %.3f s, true relres %.2e" % (d["e2e"]["value"], d["e2e"]["iterations"], d["e2e"]["time_to_solution_s"], d["e2e"]["true_relres"]),
              {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, "roofline %.3f" % d["roofline"]["frac"], {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["setup"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
grep -h "ABORT\|rror" $out/r02_config5_poisson256_t8_n4.err | head -n 3
