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


def decoupled_block_case():
    """Orthomin + ADAPT_BS with a rank drop in the MIDDLE of a solve: subdomain 0 is a diagonal block of the matrix that no
    other subdomain touches, so after the first iteration its column of the enlarged residual is zero up to rounding, the
    new search directions have numerical rank 3 of 4 and dpstrf drops one (ecg.c:375-391); the other three subdomains
    (a 6^3 Poisson grid cut by METIS) keep iterating.  Returns (matrix, parts, S, t, tol)."""
    import scipy.sparse as sp
    A1, A2 = gen_matrices.poisson7(3).tocsr(), gen_matrices.poisson7(6).tocsr()
    B = sp.block_diag([A1, A2]).tocsr()
    parts = np.concatenate([np.zeros(A1.shape[0], np.int32), 1 + restate.kway_parts(restate.sym_scale(A2), 3)]).astype(np.int32)
    return B, parts, 4, 4, 1e-10


def check_rank_drop_solve(B, parts, S, t, tol, sol, hist, info, bs):
    """the library against the numpy restatement of the same algorithm, and against the equations themselves"""
    P = restate.Partitioned(B, S, parts=parts)
    ref = restate.ecg_solve(P, t, tol, ortho=1, rrqr=True)
    n = min(len(hist), len(ref["res_hist"]))
    first = int(np.argmax(ref["bs_hist"] < t))
    assert ref["bs_hist"][first] < t and first <= 2                     # the drop happens right after the first iteration ...
    assert np.array_equal(bs[:first + 1], ref["bs_hist"][:first + 1])  # ... at the same iteration in the library
    assert len(hist) > first + 5                                        # and the solve goes on for a while with fewer directions
    assert np.allclose(hist[:first + 1], ref["res_hist"][:first + 1], rtol=1e-10, atol=0)
    # after the drop the discarded direction is rounding noise that two implementations resolve differently: the
    # histories agree loosely, the iteration counts within 2, and both end at the requested tolerance
    assert abs(info.iter - ref["iter"]) <= 2
    assert np.allclose(hist[:n - 2], ref["res_hist"][:n - 2], rtol=0.2, atol=0)
    assert hist[-1] <= tol * info.normb and info.true_relres < tol * np.sqrt(t)
    sol_ref = np.linalg.solve(P.Ap.toarray(), np.concatenate(ref["rhs"]))
    assert np.linalg.norm(sol - sol_ref) <= 1e-8 * np.linalg.norm(sol_ref)
