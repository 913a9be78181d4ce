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


def test_update_matrices_share_their_workspace(emul):
    """the (h-w)^2 update matrices of the multifrontal factorisation live from their supernode's level to their parent's only and
    get first-fit offsets over that interval (bj_factor.cu): the workspace is smaller than the sum of all of them, and the
    factor + solve still match a direct solver (an overlap of two live matrices would corrupt the factor)"""
    lib, ctx = emul
    lib.pcu_bj_stat.restype = C.c_double
    A = gen_matrices.poisson7(11).tocsr()
    n = A.shape[0]
    rc, bj, cuts, blocks = factor(emul, A, 1)
    assert rc == 0, lib.pcu_last_error()
    ws = lib.pcu_bj_stat(bj, 10)
    # all update matrices at once: recompute from the symbolic structure the library reports
    N = n
    U = sp.triu(A, format="csr"); U.sort_indices()
    perm = np.zeros(N, np.int32); nsuper = C.c_int(); sn_col = np.zeros(N + 1, np.int32); sn_rowptr = np.zeros(N + 1, np.int64)
    sn_rows = np.zeros(400 * N, np.int32); sn_parent = np.zeros(N, np.int32); sn_level = np.zeros(N, np.int32); st = np.zeros(4)
    assert lib.pcu_bj_analyze(N, ip(U.indptr.astype(np.int32)), ip(U.indices.astype(np.int32)), 1, ip(perm), C.byref(nsuper), ip(sn_col),
                              sn_rowptr.ctypes.data_as(C.POINTER(C.c_longlong)), ip(sn_rows), C.c_longlong(400 * N), ip(sn_parent),
                              ip(sn_level), dp(st)) == 0
    ns = nsuper.value
    w = np.diff(sn_col[:ns + 1]).astype(np.int64); h = np.diff(sn_rowptr[:ns + 1]).astype(np.int64)
    total = 8.0 * float(((h - w) ** 2).sum())
    assert 0 < ws < 0.8 * total, (ws, total)
    B = np.random.default_rng(0).standard_normal((n, 4))
    X = np.full((n, 4), np.nan)
    assert lib.pcu_bj_apply(bj, dp(B), 4, dp(X), 4, 4) == 0
    ref = direct(blocks, cuts, B)
    assert np.linalg.norm(X - ref) <= 1e-12 * np.linalg.norm(ref)
    lib.pcu_bj_destroy(bj)


def test_panel_offsets_closed_form(tmp_path):
    """bj.h: panel_cum(w, p), the doubles stored in front of slice p of a supernode (what lets the backward sweep find the forward
    panels without an offset table), against the plain sum of the slice sizes for every width up to 700 and 60 slices"""
    import subprocess
    exe = str(tmp_path / "panel_cum_check")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "tests", "emul"), "-I" + os.path.join(ROOT, "prealps_b200", "csrc"),
                    "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "emul", "panel_cum_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "panel_cum ok" in out.stdout, out.stdout


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
