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


@pytest.mark.parametrize("gen,N,nblk,ts", [("poisson7", 5, 1, (1, 8)), ("poisson7", 6, 2, (4, 16)), ("stencil27", 4, 1, (3, 8))])
def test_factor_and_solve_match_direct_solver(emul, gen, N, nblk, ts):
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


def test_assemble_prefetch_candidate_is_bit_identical(emul, monkeypatch):
    """PREALPS_BJ_ASM_PREFETCH=1 (assemble_kernel<T, true>): static data fetched before the dependency wait, same bits"""
    lib, ctx = emul
    A = gen_matrices.poisson7(6).tocsr()
    n = A.shape[0]
    rc, bj, cuts, blocks = factor(emul, A, 2)
    assert rc == 0
    for t in (1, 8):
        B = np.random.default_rng(10 + t).standard_normal((n, t))
        out = []
        for flag in (None, "1"):
            if flag is None:
                monkeypatch.delenv("PREALPS_BJ_ASM_PREFETCH", raising=False)
            else:
                monkeypatch.setenv("PREALPS_BJ_ASM_PREFETCH", flag)
            X = np.zeros((n, t))
            assert lib.pcu_bj_apply(bj, dp(B), t, dp(X), t, t) == 0
            out.append(X)
        assert np.array_equal(out[0], out[1])
        ref = direct(blocks, cuts, B)
        assert np.linalg.norm(out[1] - ref) <= 1e-12 * np.linalg.norm(ref)
    lib.pcu_bj_destroy(bj)


def test_indefinite_block_is_rejected(emul):
    lib, ctx = emul
    A = gen_matrices.poisson7(4).tolil()
    A[5, 5] = -3.0
    rc, bj, cuts, blocks = factor(emul, A.tocsr(), 1)
    assert rc == 2 and b"not positive definite" in lib.pcu_last_error()


def test_graph_candidate_replays_the_same_chain(emul, monkeypatch):
    """PREALPS_BJ_GRAPH=1: first use of an argument tuple runs directly, the second captures + launches, later ones replay;
    growing the work vectors (a wider solve) drops the captured graphs"""
    lib, ctx = emul
    lib.emul_graph_replay_count.restype = C.c_longlong
    lib.emul_launch_count.restype = C.c_longlong
    A = gen_matrices.poisson7(5).tocsr()
    n = A.shape[0]
    rc, bj, cuts, blocks = factor(emul, A, 1)
    assert rc == 0
    B4 = np.random.default_rng(4).standard_normal((n, 4))
    X4 = np.zeros((n, 4))
    monkeypatch.delenv("PREALPS_BJ_GRAPH", raising=False)
    assert lib.pcu_bj_apply(bj, dp(B4), 4, dp(X4), 4, 4) == 0
    base, l0 = X4.copy(), lib.emul_launch_count(ctx)
    assert lib.pcu_bj_apply(bj, dp(B4), 4, dp(X4), 4, 4) == 0
    per_apply = lib.emul_launch_count(ctx) - l0
    monkeypatch.setenv("PREALPS_BJ_GRAPH", "1")
    r0 = lib.emul_graph_replay_count()
    for k in range(4):
        X4[:] = 0
        l0 = lib.emul_launch_count(ctx)
        assert lib.pcu_bj_apply(bj, dp(B4), 4, dp(X4), 4, 4) == 0, lib.pcu_last_error()
        assert np.array_equal(X4, base)
        assert lib.emul_launch_count(ctx) - l0 == per_apply  # the launch counter keeps counting kernels
        replays = lib.emul_graph_replay_count() - r0
        assert (replays == 0) if k == 0 else (replays > 0)
        r0 = lib.emul_graph_replay_count()
    # a wider solve re-allocates the work vectors: the old graph must not be replayed into freed memory
    B16 = np.random.default_rng(16).standard_normal((n, 16))
    X16 = np.zeros((n, 16))
    for k in range(3):
        assert lib.pcu_bj_apply(bj, dp(B16), 16, dp(X16), 16, 16) == 0
    ref = direct(blocks, cuts, B16)
    assert np.linalg.norm(X16 - ref) <= 1e-12 * np.linalg.norm(ref)
    for k in range(3):
        X4[:] = 0
        assert lib.pcu_bj_apply(bj, dp(B4), 4, dp(X4), 4, 4) == 0
        assert np.array_equal(X4, base)
    lib.pcu_bj_destroy(bj)


@pytest.mark.parametrize("nblk,t", [(2, 8), (1, 3)])
def test_bottom_of_forest_candidate_is_bit_identical(emul, monkeypatch, nblk, t):
    """PREALPS_BJ_BOTTOM=Lc: levels [0, Lc) as one forward and one backward launch, a CTA per subtree; same operation order
    per panel and per gather list as the level-by-level kernels, far fewer launches"""
    lib, ctx = emul
    lib.emul_launch_count.restype = C.c_longlong
    A = gen_matrices.poisson7(6).tocsr()
    n = A.shape[0]
    ld = t if t % 2 == 0 else t + 1
    B = np.random.default_rng(t).standard_normal((n, ld))
    out, launches = {}, {}
    for Lc in (0, 2, 99):
        if Lc:
            monkeypatch.setenv("PREALPS_BJ_BOTTOM", str(Lc))
        else:
            monkeypatch.delenv("PREALPS_BJ_BOTTOM", raising=False)
        rc, bj, cuts, blocks = factor(emul, A, nblk)
        assert rc == 0, lib.pcu_last_error()
        X = np.full((n, ld), np.nan)
        l0 = lib.emul_launch_count(ctx)
        assert lib.pcu_bj_apply(bj, dp(B), ld, dp(X), ld, t) == 0, lib.pcu_last_error()
        launches[Lc] = lib.emul_launch_count(ctx) - l0
        out[Lc] = X[:, :t].copy()
        B2 = B.copy()  # in place
        assert lib.pcu_bj_apply(bj, dp(B2), ld, dp(B2), ld, t) == 0
        assert np.array_equal(B2[:, :t], out[Lc])
        lib.pcu_bj_destroy(bj)
    ref = direct(blocks, cuts, B[:, :t])
    assert np.linalg.norm(out[0] - ref) <= 1e-12 * np.linalg.norm(ref)
    assert np.array_equal(out[2], out[0]) and np.array_equal(out[99], out[0])
    assert launches[99] == 2 and launches[99] < launches[2] < launches[0]


def test_kernels_under_address_sanitizer():
    """the same emulation built with -fsanitize=address: "device" buffers are host allocations with red zones, so an
    out-of-bounds access of a kernel (factorisation, sweeps, bottom-of-the-forest launch) aborts the run"""
    import subprocess
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan is not available")
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "bj_asan_case.py")], env=env, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "asan case ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
