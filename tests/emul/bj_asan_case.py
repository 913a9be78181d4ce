"""Run under AddressSanitizer by tests/test_bj_emul.py::test_kernels_under_address_sanitizer (LD_PRELOAD=libasan):
factorisation + sweeps of a small block, twice (same bits)."""
import ctypes as C
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import build_bj_emul  # noqa: E402
import gen_matrices  # noqa: E402

lib = C.CDLL(build_bj_emul.build(asan=True))
lib.emul_ctx_create.restype = C.c_void_p
ctx = C.c_void_p(lib.emul_ctx_create())
A = gen_matrices.poisson7(5).tocsr()
n = A.shape[0]
U = sp.triu(A, format="csr")
U.sort_indices()
rp_, ci_, v_ = U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data.copy()
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))  # noqa: E731
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
rp, ci, vv = (C.POINTER(C.c_int) * 1)(ip(rp_)), (C.POINTER(C.c_int) * 1)(ip(ci_)), (C.POINTER(C.c_double) * 1)(dp(v_))
cuts = np.array([0, n], dtype=np.int32)
out = []
for lc in (0, 1):
    bj = C.c_void_p()
    assert lib.pcu_bj_create(ctx, 1, ip(cuts), rp, ci, vv, C.byref(bj)) == 0
    for t in (1, 8):
        B = np.random.default_rng(t).standard_normal((n, t))
        X = np.zeros((n, t))
        assert lib.pcu_bj_apply(bj, dp(B), t, dp(X), t, t) == 0
        out.append(X)
    lib.pcu_bj_destroy(bj)
assert np.array_equal(out[0], out[2]) and np.array_equal(out[1], out[3])
print("asan case ok")
