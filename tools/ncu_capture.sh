#!/bin/bash
# tools/ncu_capture.sh -- the ncu evidence under profiles/ (run on the GPU box through gpurun, one GPU).
# Every profiled command is first run without ncu and must exit 0; numbers printed under ncu are never bench values.
set -u
out=gpurun_out
mkdir -p $out
# 1. launch list of 4 ECG iterations (+ the initial block-Jacobi apply and SpMM): per-launch durations
python tools/profile_apply.py 128 3 4 > $out/p_iter.log 2>&1 || { echo "iteration run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep|assemble|spmm|gram2|ortho_update|update_z|reduce_partials|split_rhs|halo|fro2' \
    -c 900 --csv --log-file $out/launches_iter.csv python tools/profile_apply.py 128 3 4 > $out/ncu_iter.log 2>&1
# 2. DRAM traffic of every launch of one block-Jacobi apply
python tools/profile_apply.py 128 1 1 > $out/p_bj.log 2>&1 || { echo "apply run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'sweep|assemble' \
    -c 600 --csv --log-file $out/traffic_bj.csv python tools/profile_apply.py 128 1 1 > $out/ncu_bj.log 2>&1
# 3. full capture of the forward sweep of three big levels (launch 10-12 of sweep_kernel)
ncu --set full --clock-control none --import-source on -k regex:'sweep_kernel' -s 10 -c 3 -o $out/r02_prof_sweep -f \
    python tools/profile_apply.py 128 1 1 > $out/ncu_sw.log 2>&1
ls -la $out/*.csv $out/*.ncu-rep
