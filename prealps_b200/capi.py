"""ctypes binding of libprealps_b200.so / libprealps_cuda.so (tests and bench only)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PREALPS_B200_LIBDIR: tests/emul loads the CPU emulation of the CUDA kernels (tests/_build/emul_lib) through the same binding;
# the product libraries live in prealps_b200/lib and have no CPU path
LIBDIR = os.environ.get("PREALPS_B200_LIBDIR") or os.path.join(_HERE, "lib")


def _load(name):
    path = os.path.join(LIBDIR, name)
    if not os.path.exists(path):
        raise ImportError(
            "%s is missing: run `make` (or `python -c 'import __graft_entry__ as g; g.build()'`) at the repo root. "
            "prealps_b200 has no Python/CPU fallback for its CUDA hot path." % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


mpishim = _load("libmpishim.so")
cuda = _load("libprealps_cuda.so")
lib = _load("libprealps_b200.so")

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)


class Info(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("nnz", C.c_int), ("m", C.c_int), ("n", C.c_int),
                ("lnnz", C.c_int), ("blockSize", C.c_int), ("format", C.c_int), ("structure", C.c_int)]


class MatCSR(C.Structure):
    _fields_ = [("info", Info), ("rowPtr", c_int_p), ("colInd", c_int_p), ("val", c_double_p)]

    def arrays(self):
        m, nnz = self.info.m, self.info.lnnz
        rp = np.ctypeslib.as_array(self.rowPtr, shape=(m + 1,)).copy()
        ci = np.ctypeslib.as_array(self.colInd, shape=(max(nnz, 1),))[:nnz].copy()
        v = np.ctypeslib.as_array(self.val, shape=(max(nnz, 1),))[:nnz].copy()
        return rp, ci, v


class InfoDense(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("m", C.c_int), ("n", C.c_int), ("lda", C.c_int),
                ("nval", C.c_int), ("stor_type", C.c_int)]


class MatDense(C.Structure):
    _fields_ = [("val", c_double_p), ("info", InfoDense)]


ROW_MAJOR, COL_MAJOR = 0, 1


class SolveInfo(C.Structure):
    _fields_ = [("iter", C.c_int), ("res", C.c_double), ("normb", C.c_double), ("true_relres", C.c_double),
                ("t_solve", C.c_double), ("t_dev_ms", C.c_double), ("nhist", C.c_int), ("stopped", C.c_int)]


def ip(a):
    return a.ctypes.data_as(c_int_p)


def dp(a):
    return a.ctypes.data_as(c_double_p)


def csr_from_arrays(rowPtr, colInd, val, symmetric=True):
    """Build a MatCSR that BORROWS the numpy arrays (keep them alive)."""
    A = MatCSR()
    m = len(rowPtr) - 1
    A.info.M = A.info.m = A.info.N = A.info.n = m
    A.info.nnz = A.info.lnnz = int(rowPtr[-1])
    A.info.blockSize = 1
    A.info.structure = 1 if symmetric else 0
    A.rowPtr, A.colInd, A.val = ip(rowPtr), ip(colInd), dp(val)
    return A


lib.preAlps_b200_Stat.restype = C.c_double
lib.preAlps_b200_Stat.argtypes = [C.c_char_p]
cuda.pcu_last_error.restype = C.c_char_p
cuda.pcu_malloc.restype = C.c_void_p
cuda.pcu_malloc.argtypes = [C.c_void_p, C.c_size_t]
cuda.pcu_host_alloc.restype = C.c_void_p
cuda.pcu_spmm_bytes.restype = C.c_double
cuda.pcu_bj_bytes.restype = C.c_double
cuda.pcu_bj_stored_bytes.restype = C.c_double
cuda.pcu_bj_stat.restype = C.c_double
cuda.pcu_launch_count.restype = C.c_int64
cuda.pcu_ctx_stream.restype = C.c_void_p


def stat(name):
    return lib.preAlps_b200_Stat(name.encode())


def device_count():
    return cuda.pcu_device_count()


def operator_arrays():
    """Host copies of the current operator's integer maps (bit-exact parity targets)."""
    A = MatCSR()
    lib.preAlps_OperatorGetA(C.byref(A))
    out = {}
    out["A_rowPtr"], out["A_colInd"], out["A_val"] = A.arrays()
    for nm, fn in (("rowPos", lib.preAlps_OperatorGetRowPosPtr), ("colPos", lib.preAlps_OperatorGetColPosPtr),
                   ("dep", lib.preAlps_OperatorGetDepPtr)):
        p, n = c_int_p(), C.c_int()
        fn(C.byref(p), C.byref(n))
        out[nm] = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value].copy()
    p, n = c_int_p(), C.c_int()
    if lib.preAlps_b200_GetPerm(C.byref(p), C.byref(n)) == 0:
        out["perm"] = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value].copy()
    lib.preAlps_b200_GetHalo(C.byref(p), C.byref(n))
    out["halo"] = np.ctypeslib.as_array(p, shape=(max(n.value, 1),))[:n.value].copy()
    M, m = C.c_int(), C.c_int()
    lib.preAlps_OperatorGetSizes(C.byref(M), C.byref(m))
    out["M"], out["m"] = M.value, m.value
    return out


def diag_block(b):
    D = MatCSR()
    if lib.preAlps_b200_GetDiagBlock(b, C.byref(D)) != 0:
        raise RuntimeError("no diagonal block %d" % b)
    return D.arrays()


def driver_rhs(m):
    rhs = np.empty(m, dtype=np.float64)
    lib.preAlps_b200_DriverRhs(dp(rhs))
    return rhs


def solve(rhs, t, tol, max_iter=1000, ortho=0, bs_red=0, max_hist=2000):
    """One full solve through the RCI API with host buffers in and out."""
    sol = np.empty_like(rhs)
    hist = np.zeros(max_hist)
    info = SolveInfo()
    lib.preAlps_b200_Solve(C.c_int(t), C.c_double(tol), C.c_int(max_iter), C.c_int(ortho), C.c_int(bs_red),
                           dp(rhs), dp(sol), dp(hist), C.c_int(max_hist), C.byref(info))
    return sol, hist[:min(info.nhist, max_hist)].copy(), info


def last_block_sizes():
    """ecg.bs after every iteration of the last solve()"""
    buf = np.zeros(4096, dtype=np.int32)
    n = lib.preAlps_b200_LastBlockSizes(ip(buf), C.c_int(buf.size))
    return buf[:min(n, buf.size)].copy()


def block_operator_host(X):
    """AX = A X for a host (m x t) array through preAlps_BlockOperator (staged through HBM)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    m, t = X.shape
    AX = np.zeros_like(X)
    a, b = MatDense(), MatDense()
    for s, arr in ((a, X), (b, AX)):
        s.val = dp(arr)
        s.info.M, s.info.N, s.info.m, s.info.n, s.info.lda, s.info.nval, s.info.stor_type = m, t, m, t, t, m * t, ROW_MAJOR
    lib.preAlps_BlockOperator(C.byref(a), C.byref(b))
    return AX


def block_jacobi_host(B):
    B = np.ascontiguousarray(B, dtype=np.float64)
    m, t = B.shape
    Z = np.zeros_like(B)
    a, b = MatDense(), MatDense()
    for s, arr in ((a, B), (b, Z)):
        s.val = dp(arr)
        s.info.M, s.info.N, s.info.m, s.info.n, s.info.lda, s.info.nval, s.info.stor_type = m, t, m, t, t, m * t, ROW_MAJOR
    lib.preAlps_BlockJacobiApply(C.byref(a), C.byref(b))
    return Z
