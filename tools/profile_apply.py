"""tools/profile_apply.py -- small driver for ncu: build the bench operator, then run a few calls of one kernel class.
    python tools/profile_apply.py [n] [what: 0 spmm | 1 block-Jacobi | 2 dense | 3 iterations] [reps] [t]
STENCIL=0|1|2 picks the operator, SUBDOMAINS=S the number of METIS subdomains (default 8; SUBDOMAINS=1 with n=64 is what one
GPU of an 8-GPU run of the 128^3 problem holds).
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prealps_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
what = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t = int(sys.argv[4]) if len(sys.argv) > 4 else 8
kind = int(os.environ.get("STENCIL", "0"))
S = int(os.environ.get("SUBDOMAINS", "8"))
assert capi.lib.preAlps_b200_OperatorBuildStencil(kind, n, S, 0, S) == 0
if what != 0:
    assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
ms = C.c_float()
if what == 3:
    M, m = C.c_int(), C.c_int()
    capi.lib.preAlps_OperatorGetSizes(C.byref(M), C.byref(m))
    rhs = capi.driver_rhs(m.value)
    L = C.c_longlong()
    capi.lib.preAlps_b200_BenchIterations(t, C.c_double(1e-8), 0, capi.dp(rhs), 3, reps, C.byref(ms), C.byref(L))
    print("iterations: %.3f ms each, %d launches" % (ms.value / reps, L.value))
else:
    capi.lib.preAlps_b200_BenchKernel(what, t, reps, 1, C.byref(ms))
    name = ["spmm_bytes_t%d" % t, "bj_bytes_t%d" % t, None][what]
    b = capi.stat(name) if name else 0
    print("what=%d t=%d: %.4f ms per call, %.1f GB/s" % (what, t, ms.value, b / ms.value / 1e6 if b else 0))
