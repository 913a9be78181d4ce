"""tools/variants.py -- time the block-Jacobi apply on the bench operator (Poisson n^3, `nsub` subdomains on this GPU):
    python tools/variants.py [n = 128] [0] [nsub = 8]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prealps_b200 import capi  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
nsub = int(sys.argv[3]) if len(sys.argv) > 3 else 8
assert capi.lib.preAlps_b200_OperatorBuildStencil(0, n, nsub, 0, nsub) == 0
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
for t in (8,):
    for v in variants:
        os.environ["PREALPS_BJ_VARIANT"] = str(v)
        ms = C.c_float()
        capi.lib.preAlps_b200_BenchKernel(1, t, 5, 1, C.byref(ms))
        b = capi.stat("bj_bytes_t%d" % t)
        print("variant %d t=%d: %.3f ms  %.1f GB/s" % (v, t, ms.value, b / ms.value / 1e6), flush=True)
