"""The C entry points bench.py drives (operator from the native stencil generator, preAlps_b200_BenchIterations,
preAlps_b200_BenchKernel for the three kernel classes, a whole solve) on the CPU emulation of the stack; run by
tests/test_full_emul.py in a process of its own."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import build_bj_emul  # noqa: E402

os.environ["PREALPS_B200_LIBDIR"] = build_bj_emul.build_full()
from prealps_b200 import capi  # noqa: E402

assert capi.lib.preAlps_b200_OperatorBuildStencil(0, 8, 8, 0, 8) == 0
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
M, m = C.c_int(), C.c_int()
capi.lib.preAlps_OperatorGetSizes(C.byref(M), C.byref(m))
rhs = capi.driver_rhs(m.value)
ms, launches = C.c_float(), C.c_longlong()
assert capi.lib.preAlps_b200_BenchIterations(8, C.c_double(1e-8), 0, capi.dp(rhs), 3, 6, C.byref(ms), C.byref(launches)) == 0
kernel_ms = []
for what in (0, 1, 2):
    assert capi.lib.preAlps_b200_BenchKernel(what, 8, 2, 1, C.byref(ms)) == 0
    kernel_ms.append(float(ms.value))
sol, hist, info = capi.solve(rhs, 8, 1e-8)
print(json.dumps({"M": M.value, "launches": int(launches.value), "kernel_ms": kernel_ms, "iter": info.iter, "true_relres": info.true_relres,
                  "spmm_bytes": capi.stat("spmm_bytes_t8"), "bj_bytes": capi.stat("bj_bytes_t8")}))
