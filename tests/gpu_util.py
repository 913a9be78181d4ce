"""helpers shared by the -m gpu tests"""
import ctypes as C
import os

import numpy as np

import gen_matrices
import restate
from conftest import GOLDEN
from prealps_b200 import capi


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    A = gen_matrices.build(g["gen"], g["N"]).tocsr()
    return g, A


def build_single_process(A, S, bj=True, scale=1, parts=None):
    """operator (+ block-Jacobi) for all S subdomains in this one process / GPU"""
    A = A.tocsr()
    A.sort_indices()
    rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    pp = capi.ip(np.ascontiguousarray(parts, dtype=np.int32)) if parts is not None else None
    rc = capi.lib.preAlps_b200_OperatorBuildCSR(A.shape[0], capi.ip(rp), capi.ip(ci), capi.dp(v), S, 0, S, scale, pp)
    assert rc == 0
    if bj:
        assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0


class Dev:
    """raw access to the C ABI of libprealps_cuda for kernel-level tests"""

    def __init__(self):
        self.ctx = C.c_void_p()
        assert capi.cuda.pcu_ctx_create(0, C.byref(self.ctx)) == 0, capi.cuda.pcu_last_error()

    def up(self, a):
        a = np.ascontiguousarray(a)
        p = capi.cuda.pcu_malloc(self.ctx, C.c_size_t(max(a.nbytes, 8)))
        assert p
        assert capi.cuda.pcu_h2d(self.ctx, C.c_void_p(p), a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes)) == 0
        return C.c_void_p(p)

    def zeros(self, n, dtype=np.float64):
        return self.up(np.zeros(n, dtype=dtype))

    def down(self, p, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        assert capi.cuda.pcu_d2h(self.ctx, out.ctypes.data_as(C.c_void_p), p, C.c_size_t(out.nbytes)) == 0
        return out

    def free(self, *ps):
        for p in ps:
            capi.cuda.pcu_free(self.ctx, p)
