"""The SpMM kernels of prealps_b200/csrc/spmm_kernels.cuh executed on the CPU by tests/emul (one pthread per CUDA thread):
index logic of every template instantiation against scipy; the bulk-staging kernels and the local/halo split of the
overlapped product bit for bit against the LDG + STS staging kernel.  TEST
INFRASTRUCTURE: the emulation is compiled here into tests/_build and is not part of the product libraries; what it cannot
see (PTX semantics, alignment traps beyond the asserted ones, timing) is left to the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

import gen_matrices
from conftest import ROOT

BUILD = os.path.join(ROOT, "tests", "_build")
SRC = os.path.join(ROOT, "tests", "emul", "spmm_emul.cpp")
DEPS = [SRC, os.path.join(ROOT, "tests", "emul", "cuda_shim.h"), os.path.join(ROOT, "prealps_b200", "csrc", "spmm_kernels.cuh")]


@pytest.fixture(scope="module")
def emul():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libspmm_emul.so")
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in DEPS):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", SRC, "-o", so])
    return C.CDLL(so)


def aligned(shape, align=64):
    n = int(np.prod(shape))
    raw = np.zeros(n + align // 8, dtype=np.float64)
    off = (-raw.ctypes.data % align) // 8
    return raw[off:off + n].reshape(shape)


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def operator(gen, N, nh, seed=3):
    A = getattr(gen_matrices, gen)(N).tocsr()
    m = A.shape[0]
    B = sp.random(m, nh, density=0.03, random_state=seed, format="csr")
    Aext = sp.hstack([A, B]).tocsr()
    Aext.sort_indices()
    return Aext, m, Aext.indptr.astype(np.int32), Aext.indices.astype(np.int32), Aext.data.astype(np.float64)


def run(emul, fn, lean, cpl, m, rp, ci, v, X, H, t, ld):
    Y = aligned((m, ld))
    Y[:] = np.nan
    args = [lean, cpl, m, ip(rp), ip(ci), dp(v), dp(X), ld, dp(H), dp(Y), ld, t]
    if fn == "split":
        nb = C.c_int()
        rc = emul.emul_spmm_split(*args, C.byref(nb))
        assert nb.value > 0
    else:
        rc = emul.emul_spmm(*args)
    assert rc == 0, rc
    return Y


def inputs(m, nh, t, ld, seed):
    rng = np.random.default_rng(seed)
    X, H = aligned((m, ld)), aligned((nh, t))
    X[:] = rng.standard_normal((m, ld))
    H[:] = rng.standard_normal((nh, t))
    return X, H


@pytest.mark.parametrize("t,cpl", [(1, 1), (2, 2), (4, 2), (8, 2), (8, 4), (16, 2), (16, 4), (32, 2), (32, 4), (3, 0), (12, 0), (8, 0)])
def test_default_kernels_match_scipy(emul, t, cpl):
    """validates the emulation itself on the kernels the GPU tests have already checked"""
    Aext, m, rp, ci, v = operator("stencil27", 7, 29)
    ld = t if (t % 2 == 0 or t == 1) else t + 1
    X, H = inputs(m, 29, t, ld, t)
    Y = run(emul, "merged", 0, cpl, m, rp, ci, v, X, H, t, ld)
    ref = Aext @ np.vstack([X[:, :t], H])
    assert np.allclose(Y[:, :t], ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())


def test_long_and_empty_rows(emul):
    rng = np.random.default_rng(0)
    m = 300
    A = sp.random(m, m, density=0.01, random_state=1, format="lil")
    A[7, :] = rng.standard_normal(m)
    A[11, :] = 0
    big = sp.random(1, 5000, density=0.9, random_state=2, format="csr")  # longer than the staging buffer
    A = sp.vstack([sp.hstack([A.tocsr(), sp.csr_matrix((m, 5000 - m))]), big]).tocsr()
    A.sort_indices()
    mm, nh = A.shape[0], A.shape[1] - A.shape[0]
    rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    for t, cpl in ((8, 2), (8, 4), (4, 2), (6, 0)):
        X, H = inputs(mm, nh, t, t, t)
        Y = run(emul, "merged", 0, cpl, mm, rp, ci, v, X, H, t, t)
        ref = A @ np.vstack([X, H])
        assert np.allclose(Y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    # the host never picks the bulk-staging kernel when a row block exceeds the staging buffer
    X, H = inputs(mm, nh, 8, 8, 1)
    Y = aligned((mm, 8))
    assert emul.emul_spmm(8, 4, mm, ip(rp), ip(ci), dp(v), dp(X), 8, dp(H), dp(Y), 8, 8) == 3


@pytest.mark.parametrize("gen,N", [("poisson7", 11), ("stencil27", 7)])
@pytest.mark.parametrize("t", [8, 16, 32])
@pytest.mark.parametrize("cpl", [2, 4])
def test_bulk_staging_is_bit_identical(emul, gen, N, t, cpl):
    """spmm_bulk_kernel (cp.async.bulk staging, the default from t = 8 up) against spmm_kernel (LDG + STS staging)"""
    Aext, m, rp, ci, v = operator(gen, N, 41)
    X, H = inputs(m, 41, t, t, t + cpl)
    base = run(emul, "merged", 0, cpl, m, rp, ci, v, X, H, t, t)
    ref = Aext @ np.vstack([X, H])
    assert np.allclose(base, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
    for lean in (8,):  # cp.async.bulk staging, HALO = true
        Y = run(emul, "merged", lean, cpl, m, rp, ci, v, X, H, t, t)
        assert np.array_equal(Y, base), (lean, np.abs(Y - base).max())


@pytest.mark.parametrize("t,cpl", [(8, 2), (8, 4), (16, 4), (32, 2)])
def test_bulk_candidate_without_halo(emul, t, cpl):
    """spmm_bulk_kernel<T, CPL, false>: one process (no column >= m); every alignment of a row block's first entry"""
    A = gen_matrices.poisson7(9).tolil()
    A[3, :] = 0  # an empty row and rows of odd length move the chunk starts over all residues mod 4
    A[5, 100:111] = 1.5
    A = A.tocsr()
    A.sort_indices()
    m = A.shape[0]
    rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    X, H = inputs(m, 1, t, t, 3 * t)
    base = run(emul, "merged", 0, cpl, m, rp, ci, v, X, H, t, t)
    assert np.allclose(base, A @ X, rtol=1e-13, atol=1e-13)
    Y = run(emul, "merged", 9, cpl, m, rp, ci, v, X, H, t, t)
    assert np.array_equal(Y, base)


@pytest.mark.parametrize("t,cpl,lean", [(1, 1, 0), (4, 2, 0), (8, 4, 0), (12, 0, 0), (8, 4, 9), (16, 2, 9), (32, 4, 9)])
def test_local_halo_split_is_bit_identical(emul, t, cpl, lean):
    """PREALPS_SPMM_OVERLAP=1: local kernel on the entries with column < m, then halo_add_kernel continues the FMA chains"""
    Aext, m, rp, ci, v = operator("poisson7", 11, 37)
    ld = t if (t % 2 == 0 or t == 1) else t + 1
    X, H = inputs(m, 37, t, ld, 7 * t)
    merged = run(emul, "merged", 0, cpl, m, rp, ci, v, X, H, t, ld)
    split = run(emul, "split", lean, cpl, m, rp, ci, v, X, H, t, ld)
    assert np.array_equal(split[:, :t], merged[:, :t])


def test_halo_pack(emul):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((500, 10))
    idx = rng.permutation(500)[:77].astype(np.int32)
    out = np.zeros((77, 9))
    assert emul.emul_halo_pack(dp(X), 10, 9, ip(idx), 77, dp(out)) == 0
    assert np.array_equal(out, X[idx, :9])
