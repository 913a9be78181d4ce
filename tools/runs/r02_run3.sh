#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
cat > /tmp/prof1.py <<'PY'
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from prealps_b200 import capi
n, nsub = int(sys.argv[1]), int(sys.argv[2])
assert capi.lib.preAlps_b200_OperatorBuildStencil(0, n, nsub, 0, nsub) == 0
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
ms = C.c_float()
os.environ["PREALPS_BJ_LEVELS"] = "1"
capi.lib.preAlps_b200_BenchKernel(1, 8, 3, 1, C.byref(ms))
os.environ["PREALPS_BJ_PROFILE"] = "1"
capi.lib.preAlps_b200_BenchKernel(1, 8, 1, 1, C.byref(ms))
PY
timeout 300 python /tmp/prof1.py 64 1 > $out/r02_prof_n64.txt 2>&1
tail -n 62 $out/r02_prof_n64.txt
