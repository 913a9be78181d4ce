"""The block-Jacobi factorisation and sweeps (prealps_b200/csrc/bj_factor.cu, bj_solve.cu) executed on the CPU by tests/emul:
one pthread per CUDA thread, mma.sync.m8n8k4.f64 as a warp-wide exchange, cp.async / griddepcontrol as plain copies / no-ops.
The whole pcu_bj_create + pcu_bj_apply path (host planning, numeric multifrontal Cholesky, panel packing, level-scheduled
forward and backward sweeps) against a direct sparse solve, on problems small enough for a few seconds of emulation.
TEST INFRASTRUCTURE: compiled here into tests/_build, never part of the product libraries; timing, cache hints and the
asynchrony of cp.async / programmatic dependent launch are out of its reach and stay with the -m gpu tests."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import gen_matrices
from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import build_bj_emul  # noqa: E402


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build_bj_emul.build())
    lib.emul_ctx_create.restype = C.c_void_p
    lib.pcu_last_error.restype = C.c_char_p
    return lib, C.c_void_p(lib.emul_ctx_create())


def factor(emul, A, nblk):
    lib, ctx = emul
    n = A.shape[0]
    cuts = np.linspace(0, n, nblk + 1).astype(np.int32)
    blocks = [A[cuts[b]:cuts[b + 1], cuts[b]:cuts[b + 1]].tocsr() for b in range(nblk)]
    keep = []
    for Bk in blocks:
        U = sp.triu(Bk, format="csr")
        U.sort_indices()
        keep.append((U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data.copy()))
    rp = (C.POINTER(C.c_int) * nblk)(*[ip(k[0]) for k in keep])
    ci = (C.POINTER(C.c_int) * nblk)(*[ip(k[1]) for k in keep])
    vv = (C.POINTER(C.c_double) * nblk)(*[dp(k[2]) for k in keep])
    bj = C.c_void_p()
    rc = lib.pcu_bj_create(ctx, nblk, ip(cuts), rp, ci, vv, C.byref(bj))
    return rc, bj, cuts, blocks


def direct(blocks, cuts, B):
    return np.vstack([spla.splu(Bk.tocsc()).solve(B[cuts[b]:cuts[b + 1]]) for b, Bk in enumerate(blocks)])


@pytest.mark.parametrize("gen,N,nblk,ts,copies,pw", [("poisson7", 5, 1, (1, 8), 2, 0), ("poisson7", 6, 2, (16, 4, 8), 1, 0),
                                                       ("stencil27", 4, 1, (3, 8, 2), 2, 2), ("poisson7", 9, 1, (8, 32, 3), 1, 0),
                                                       ("poisson7", 9, 1, (8, 16, 1), 2, 3)])
def test_factor_and_solve_match_direct_solver(emul, gen, N, nblk, ts, copies, pw, monkeypatch):
    """copies = 2: the factor keeps the transposed panels; copies = 1: the backward sweep reads the forward panels tile by
    tile (bj.h).  poisson7 9^3 has supernodes wider and taller than one 32 x 32 tile.  pw > 0 forces that many panels per
    warp (the sweep kernel streams across them; the planner only does this on levels with tens of thousands of panels)."""
    monkeypatch.setenv("PREALPS_BJ_COPIES", str(copies))
    if pw:
        monkeypatch.setenv("PREALPS_BJ_PW_FORCE", str(pw))
    lib, ctx = emul
    A = getattr(gen_matrices, gen)(N).tocsr()
    n = A.shape[0]
    rc, bj, cuts, blocks = factor(emul, A, nblk)
    assert rc == 0, lib.pcu_last_error()
    for t in ts:
        ld = t if (t % 2 == 0 or t == 1) else t + 1
        B = np.random.default_rng(t).standard_normal((n, ld))
        X = np.full((n, ld), np.nan)
        assert lib.pcu_bj_apply(bj, dp(B), ld, dp(X), ld, t) == 0, lib.pcu_last_error()
        ref = direct(blocks, cuts, B[:, :t])
        assert np.linalg.norm(X[:, :t] - ref) <= 1e-12 * np.linalg.norm(ref)
        # in place, and reproducible
        B2 = B.copy()
        assert lib.pcu_bj_apply(bj, dp(B2), ld, dp(B2), ld, t) == 0
        assert np.array_equal(B2[:, :t], X[:, :t])
    lib.pcu_bj_destroy(bj)


def test_indefinite_block_is_rejected(emul):
    lib, ctx = emul
    A = gen_matrices.poisson7(4).tolil()
    A[5, 5] = -3.0
    rc, bj, cuts, blocks = factor(emul, A.tocsr(), 1)
    assert rc == 2 and b"not positive definite" in lib.pcu_last_error()


def test_kernels_under_address_sanitizer():
    """the same emulation built with -fsanitize=address: "device" buffers are host allocations with red zones, so an
    out-of-bounds access of a kernel (factorisation, sweeps) aborts the run"""
    import subprocess
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan is not available")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "bj_asan_case.py")], env=env, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "asan case ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
