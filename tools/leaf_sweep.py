"""tools/leaf_sweep.py -- block-Jacobi apply time vs supernode relaxation settings (re-creates the factor each time)"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prealps_b200 import capi  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
assert capi.lib.preAlps_b200_OperatorBuildStencil(0, n, 8, 0, 8) == 0
for leaf, relax in ((32, 0.2), (48, 0.2), (64, 0.2), (96, 0.2), (64, 0.3), (32, 0.3)):
    os.environ["PREALPS_BJ_LEAF"] = str(leaf)
    os.environ["PREALPS_BJ_RELAX"] = str(relax)
    assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
    ms = C.c_float()
    capi.lib.preAlps_b200_BenchKernel(1, 8, 5, 1, C.byref(ms))
    b = capi.stat("bj_bytes_t8")
    print("leaf %3d relax %.2f: %.3f ms  %.1f GB/s  stored %.0fM exact %.0fM supernodes %d levels %d analysis %.1fs factor %.2fs"
          % (leaf, relax, ms.value, b / ms.value / 1e6, capi.stat("bj_nnz_stored") / 1e6, capi.stat("bj_nnz_exact") / 1e6,
             capi.stat("bj_supernodes"), capi.stat("bj_levels"), capi.stat("bj_analysis_s"), capi.stat("bj_factor_s")), flush=True)
