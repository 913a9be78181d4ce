#!/bin/bash
# one 8-GPU session: multi-GPU parity at world 8, the headline bench at N = 8 (with and without the overlapped halo exchange),
# BASELINE config 4 (elasticity 96^3, t = 16, -r 1, 16 subdomains) and config 5 (Poisson 256^3, t = 8, one subdomain per GPU)
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l; nproc; free -g | head -n 2
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "8]" > $out/r02_multi8.log 2>&1; echo "multi tests rc=$?"; tail -n 4 $out/r02_multi8.log
run() { name=$1; shift; PREALPS_B200_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc=$?"; }
run r02_bench_n8
PREALPS_SPMM_OVERLAP=1 run r02_bench_n8_overlap
run r02_config4_elasticity96_t16_r1 --operator elasticity --grid 96 --enl 16 --nsub 16 --bs-red 1 --steps 20 --warmup 3
run r02_config4_elasticity96_t16_r0 --operator elasticity --grid 96 --enl 16 --nsub 16 --bs-red 0 --steps 20 --warmup 3
run r02_config5_poisson256_t8 --grid 256 --steps 20 --warmup 3
python - <<'PY'
import json
for f in ("r02_bench_n8", "r02_bench_n8_overlap", "r02_config4_elasticity96_t16_r1", "r02_config4_elasticity96_t16_r0", "r02_config5_poisson256_t8"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "it/s %.1f" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.1f it/s, %d iterations in %.3f s, true relres %.2e" % (d["e2e"]["value"], d["e2e"]["iterations"], d["e2e"]["time_to_solution_s"], d["e2e"]["true_relres"]),
              {k: round(v["ms"], 4) for k, v in d["kernels"].items()}, "roofline %.3f" % d["roofline"]["frac"], {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["setup"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
grep -h "setup:" $out/r02_bench_n8.err | sort | uniq -c | sort -rn | head -n 6
tail -n 3 $out/r02_config5_poisson256_t8.err
