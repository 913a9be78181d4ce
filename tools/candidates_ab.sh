#!/bin/bash
# tools/candidates_ab.sh -- first GPU call of the next round: the opt-in kernels written after the last GPU session of
# round 1 (none of them has run on a B200 yet), each against the default, all in ONE gpurun call on one GPU:
#     gpurun --timeout 900 -- 'bash tools/candidates_ab.sh'
# 1. parity of the candidates (bit-identical to the defaults), 2. A/B timings.  Whatever wins becomes the default
# (and its environment switch is inverted or dropped); whatever loses is deleted and recorded under profiles/.
set -u
out=gpurun_out
mkdir -p $out
PREALPS_TEST_CANDIDATES=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ecg.py -q -m gpu -k "candidate or harness" \
    > $out/cand_tests.log 2>&1
echo "candidate tests rc=$?" | tee -a $out/cand_tests.log
# SpMM: default vs lean row phase (27-point and 7-point 128^3, t = 1..32; only t >= 8 differ)
python tools/spmm_sweep.py 128 > $out/cand_spmm_default.jsonl 2> $out/cand_spmm_default.err
for nb in 1 2 4; do  # gathers in flight per lane
  PREALPS_SPMM_LEAN=$nb python tools/spmm_sweep.py 128 > $out/cand_spmm_lean$nb.jsonl 2> $out/cand_spmm_lean$nb.err
done
PREALPS_SPMM_BULK=1 python tools/spmm_sweep.py 128 > $out/cand_spmm_bulk.jsonl 2> $out/cand_spmm_bulk.err
# block-Jacobi apply, 8 subdomains of 64^3 on this GPU (the N=1 bench) and ONE 64^3 subdomain (what each GPU holds at N=8)
python tools/variants.py 128 8 8 > $out/cand_bj.log 2>&1
python tools/variants.py 64 1 8 >> $out/cand_bj.log 2>&1
# whole iterations with everything on
python bench.py --no-cpu-baseline > $out/cand_bench_default.json 2> $out/cand_bench_default.err
PREALPS_SPMM_LEAN=1 PREALPS_BJ_ASM_PREFETCH=1 PREALPS_BJ_GRAPH=1 PREALPS_BJ_BOTTOM=4 python bench.py --no-cpu-baseline > $out/cand_bench_all.json 2> $out/cand_bench_all.err
grep -h '"t": 8,' $out/cand_spmm_*.jsonl; tail -n 3 $out/cand_tests.log; cat $out/cand_bj.log | grep -v METIS | tail -n 20
# multi-GPU candidates need `gpurun --gpus 2` (or 8):
#   PREALPS_TEST_CANDIDATES=1 python -m pytest tests/test_gpu_multi.py -q -m gpu -k overlapped
#   for v in "" PREALPS_SPMM_OVERLAP=1 PREALPS_BJ_GRAPH=1 PREALPS_BJ_BOTTOM=6 "PREALPS_SPMM_OVERLAP=1 PREALPS_BJ_GRAPH=1 PREALPS_BJ_ASM_PREFETCH=1 PREALPS_BJ_BOTTOM=6"; do
#     env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8; done
