"""tools/variants.py -- time the block-Jacobi apply on the bench operator (Poisson n^3, `nsub` subdomains on this GPU) for the
supernode relaxation settings given as "relax_big:relax_big_cols" pairs (read when the factor is created); other switches
(PREALPS_BJ_COPIES, PREALPS_BJ_LEAF, PREALPS_BJ_SPLITA, PREALPS_BJ_CHUNKQ, ...) come from the environment:
    python tools/variants.py [n = 128] [nsub = 8] [t = 8] [relax settings, comma separated = default]
(nsub = 1 with n = 64 is what one GPU of an 8-GPU run of the 128^3 problem holds.)"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prealps_b200 import capi  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 8
t = int(sys.argv[3]) if len(sys.argv) > 3 else 8
relax = sys.argv[4].split(",") if len(sys.argv) > 4 else ["default"]
assert capi.lib.preAlps_b200_OperatorBuildStencil(0, n, nsub, 0, nsub) == 0
for rx in relax:
    for k in ("PREALPS_BJ_RELAX_BIG", "PREALPS_BJ_RELAX_BIG_COLS"):
        os.environ.pop(k, None)
    if rx != "default":
        os.environ["PREALPS_BJ_RELAX_BIG"], os.environ["PREALPS_BJ_RELAX_BIG_COLS"] = rx.split(":")
    assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
    for mode in ("levels",):
        ms = C.c_float()
        capi.lib.preAlps_b200_BenchKernel(1, t, 20, 1, C.byref(ms))
        b, ex = capi.stat("bj_stored_bytes_t%d" % t), capi.stat("bj_bytes_t%d" % t)
        print("n=%d nsub=%d relax=%-10s %-9s t=%d: %.3f ms  stored %.2f GB -> %.1f GB/s   algorithmic (exact nnz(L)) %.2f GB -> %.1f GB/s, levels %d"
              % (n, nsub, rx, mode, t, ms.value, b / 1e9, b / ms.value / 1e6, ex / 1e9, ex / ms.value / 1e6, int(capi.stat("bj_levels"))), flush=True)
