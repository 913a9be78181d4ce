#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
# the native harness (preAlps half of the reference's test_bench_spmm.c / test_bench_bjacobi.c)
timeout 600 prealps_b200/bin/bench_kernels > $out/r02_bench_kernels.log 2>&1; echo "bench_kernels rc=$?"; tail -n 12 $out/r02_bench_kernels.log
# per-level profile with the new assembly kernel, 8 blocks and 1 block per GPU
cat > /tmp/prof1.py <<'PY'
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from prealps_b200 import capi
n, nsub = int(sys.argv[1]), int(sys.argv[2])
assert capi.lib.preAlps_b200_OperatorBuildStencil(0, n, nsub, 0, nsub) == 0
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
ms = C.c_float()
capi.lib.preAlps_b200_BenchKernel(1, 8, 5, 1, C.byref(ms))
print("apply: %.3f ms" % ms.value, file=sys.stderr)
os.environ["PREALPS_BJ_PROFILE"] = "1"
capi.lib.preAlps_b200_BenchKernel(1, 8, 1, 1, C.byref(ms))
PY
timeout 300 python /tmp/prof1.py 64 1 > $out/r02_prof_levels_n64.txt 2>&1
timeout 300 python /tmp/prof1.py 128 8 > $out/r02_prof_levels_n128.txt 2>&1
grep -E "apply:|total" $out/r02_prof_levels_n64.txt $out/r02_prof_levels_n128.txt | tail -n 8
bash tools/ncu_capture.sh > $out/ncu_capture.log 2>&1; tail -n 6 $out/ncu_capture.log
