"""tools/spmm_once.py -- a handful of SpMM launches on one synthetic operator, for ncu:
    python tools/spmm_once.py [kind = 1 (27-point)] [n = 128] [t = 8] [reps = 3]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prealps_b200 import capi  # noqa: E402
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
t = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
assert capi.lib.preAlps_b200_OperatorBuildStencil(kind, n, 8, 0, 8) == 0
ms = C.c_float()
capi.lib.preAlps_b200_BenchKernel(0, t, reps, 1, C.byref(ms))
print("kind %d n %d t %d: %.1f us" % (kind, n, t, ms.value * 1e3))
