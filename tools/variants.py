"""tools/variants.py -- time the block-Jacobi apply on the bench operator (Poisson n^3, `nsub` subdomains on this GPU) under
the environment switches of bj_solve.cu: PREALPS_BJ_BOTTOM (read when the factor is created: the factor is re-created for
each value) x {PREALPS_BJ_ASM_PREFETCH, PREALPS_BJ_GRAPH} (read per apply):
    python tools/variants.py [n = 128] [nsub = 8] [t = 8] [bottom levels, comma separated = 0,4,6,8]
(nsub = 1 with n = 64 is what one GPU of an 8-GPU run of the 128^3 problem holds.)"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prealps_b200 import capi  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t = int(sys.argv[3]) if len(sys.argv) > 3 else 8
bottoms = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 4, 6, 8]
SWITCHES = [(), ("PREALPS_BJ_ASM_PREFETCH",), ("PREALPS_BJ_GRAPH",), ("PREALPS_BJ_ASM_PREFETCH", "PREALPS_BJ_GRAPH")]
assert capi.lib.preAlps_b200_OperatorBuildStencil(0, n, nsub, 0, nsub) == 0
for lc in bottoms:
    os.environ.pop("PREALPS_BJ_BOTTOM", None)
    if lc:
        os.environ["PREALPS_BJ_BOTTOM"] = str(lc)
    assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
    for on in SWITCHES:
        for name in ("PREALPS_BJ_ASM_PREFETCH", "PREALPS_BJ_GRAPH"):
            os.environ.pop(name, None)
        for name in on:
            os.environ[name] = "1"
        ms = C.c_float()
        capi.lib.preAlps_b200_BenchKernel(1, t, 20, 1, C.byref(ms))
        b = capi.stat("bj_bytes_t%d" % t)
        print("n=%d nsub=%d bottom=%d %-48s t=%d: %.3f ms  %.1f GB/s" % (n, nsub, lc, " ".join(on) or "default", t, ms.value, b / ms.value / 1e6),
              flush=True)
