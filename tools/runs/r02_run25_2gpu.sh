#!/bin/bash
# two GPUs at HEAD: multi-GPU parity tests, the headline bench at N = 2, and BASELINE config 5 (Poisson 256^3) on TWO GPUs, which only
# fits with one copy of the block-Jacobi factor (4 x 128^3 blocks per GPU)
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -n 3; nproc; free -g | head -n 2
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $out/r02_multi2.log 2>&1; echo "multi tests rc=$?"
tail -n 3 $out/r02_multi2.log
run() { PREALPS_B200_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu-baseline "$@"; }
run > $out/r02_bench_n2.json 2> $out/r02_bench_n2.err; echo "bench n2 rc=$?"
run --grid 256 --steps 20 --warmup 3 > $out/r02_config5_poisson256_t8_n2.json 2> $out/r02_config5_poisson256_t8_n2.err; echo "config 5 on 2 GPUs rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_n2", "r02_config5_poisson256_t8_n2"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.1f" % d["e2e"]["value"], "iters", d["e2e"]["iterations"], "tts %.3f" % d["e2e"]["time_to_solution_s"],
              {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, d["setup"])
    except Exception as e:
        print(f, "failed", e)
PY
grep "setup:" $out/r02_bench_n2.err | head -n 14
tail -n 5 $out/r02_config5_poisson256_t8_n2.err
nvidia-smi --query-gpu=memory.used --format=csv | head -n 3
