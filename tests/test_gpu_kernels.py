"""Kernel-level parity on the GPU through the C ABI of libprealps_cuda (include/prealps_cuda.h)."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import gen_matrices
from gpu_util import Dev
from prealps_b200 import capi

pytestmark = pytest.mark.gpu
cu = capi.cuda


@pytest.fixture(scope="module")
def dev():
    return Dev()


@pytest.mark.parametrize("t", [1, 2, 3, 4, 8, 12, 16, 32])
def test_spmm_matches_scipy(dev, t):
    A = gen_matrices.stencil27(9).tocsr()
    m = A.shape[0]
    nh = 37  # pretend the last 37 columns live in a halo buffer
    rng = np.random.default_rng(t)
    B = sp.random(m, nh, density=0.02, random_state=3, format="csr")
    Aext = sp.hstack([A, B]).tocsr()
    Aext.sort_indices()
    op = C.c_void_p()
    assert cu.pcu_spmm_create(dev.ctx, m, nh, capi.ip(Aext.indptr.astype(np.int32)), capi.ip(Aext.indices.astype(np.int32)),
                              capi.dp(Aext.data), C.byref(op)) == 0, cu.pcu_last_error()
    ld = t if (t % 2 == 0 or t == 1) else t + 1
    X = rng.standard_normal((m, ld))
    H = rng.standard_normal((nh, t))
    # one fake neighbour (rank 0 itself is never contacted: pack only)
    assert cu.pcu_spmm_set_halo(op, 0, None, capi.ip(np.zeros(1, np.int32)), None, capi.ip(np.array([nh], np.int32))) != 0
    assert cu.pcu_spmm_set_halo(op, 1, capi.ip(np.array([0], np.int32)), capi.ip(np.array([0, 5], np.int32)),
                                capi.ip(np.array([1, 3, 5, 7, 9], np.int32)), capi.ip(np.array([0, nh], np.int32))) == 0
    cu.pcu_spmm_halo_buffer.restype = C.c_void_p
    hb = cu.pcu_spmm_halo_buffer(op, t)
    assert hb
    assert cu.pcu_h2d(dev.ctx, C.c_void_p(hb), H.ctypes.data_as(C.c_void_p), C.c_size_t(H.nbytes)) == 0
    dX, dY = dev.up(X), dev.zeros(m * ld)
    assert cu.pcu_spmm_apply(op, dX, ld, dY, ld, t) == 0, cu.pcu_last_error()
    Y = dev.down(dY, (m, ld))[:, :t]
    ref = Aext @ np.vstack([X[:, :t], H])
    assert np.allclose(Y, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
    # halo pack gathers the requested rows
    pk, nr = C.c_void_p(), C.c_int()
    assert cu.pcu_spmm_halo_pack(op, dX, ld, t, C.byref(pk), C.byref(nr)) == 0
    packed = dev.down(pk, (5, t))
    assert np.array_equal(packed, X[[1, 3, 5, 7, 9], :t])
    dev.free(dX, dY)
    cu.pcu_spmm_destroy(op)


@pytest.mark.parametrize("gen,N", [("poisson7", 14), ("stencil27", 9)])
def test_spmm_bulk_staging_is_bit_identical(dev, gen, N, monkeypatch):
    """spmm_bulk_kernel (cp.async.bulk staging, the default from t = 8 up) keeps the mapping and the summation order of
    spmm_kernel (PREALPS_SPMM_BULK=0): same bits"""
    A = getattr(gen_matrices, gen)(N).tocsr()
    m = A.shape[0]
    nh = 53
    B = sp.random(m, nh, density=0.03, random_state=5, format="csr")
    Aext = sp.hstack([A, B]).tocsr()
    Aext.sort_indices()
    out = {}
    for bulk in ("0", "1"):
        monkeypatch.setenv("PREALPS_SPMM_BULK", bulk)
        op = C.c_void_p()
        assert cu.pcu_spmm_create(dev.ctx, m, nh, capi.ip(Aext.indptr.astype(np.int32)), capi.ip(Aext.indices.astype(np.int32)),
                                  capi.dp(Aext.data), C.byref(op)) == 0, cu.pcu_last_error()
        assert cu.pcu_spmm_set_halo(op, 1, capi.ip(np.array([0], np.int32)), capi.ip(np.array([0, 0], np.int32)),
                                    capi.ip(np.zeros(1, np.int32)), capi.ip(np.array([0, nh], np.int32))) == 0
        cu.pcu_spmm_halo_buffer.restype = C.c_void_p
        for t in (8, 16, 32):
            r2 = np.random.default_rng(t)
            X, H = r2.standard_normal((m, t)), r2.standard_normal((nh, t))
            hb = cu.pcu_spmm_halo_buffer(op, t)
            assert cu.pcu_h2d(dev.ctx, C.c_void_p(hb), H.ctypes.data_as(C.c_void_p), C.c_size_t(H.nbytes)) == 0
            dX, dY = dev.up(X), dev.zeros(m * t)
            assert cu.pcu_spmm_apply(op, dX, t, dY, t, t) == 0, cu.pcu_last_error()
            out[bulk, t] = dev.down(dY, (m, t))
            ref = Aext @ np.vstack([X, H])
            assert np.allclose(out[bulk, t], ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
            dev.free(dX, dY)
        cu.pcu_spmm_destroy(op)
    for t in (8, 16, 32):
        assert np.array_equal(out["0", t], out["1", t])


def test_spmm_long_rows_and_empty_rows(dev):
    rng = np.random.default_rng(0)
    m = 300
    A = sp.random(m, m, density=0.01, random_state=1, format="lil")
    A[7, :] = rng.standard_normal(m)           # dense row
    A[11, :] = 0                               # empty row
    big = sp.random(1, 5000, density=0.9, random_state=2, format="csr")  # a row longer than the staging buffer
    A = sp.vstack([sp.hstack([A.tocsr(), sp.csr_matrix((m, 5000 - m))]), big]).tocsr()
    A.sort_indices()
    mm, nc = A.shape
    nh = nc - mm
    op = C.c_void_p()
    assert cu.pcu_spmm_create(dev.ctx, mm, nh, capi.ip(A.indptr.astype(np.int32)), capi.ip(A.indices.astype(np.int32)),
                              capi.dp(A.data), C.byref(op)) == 0, cu.pcu_last_error()
    for t in (1, 4, 6, 8):
        ld = t
        X = rng.standard_normal((mm, ld))
        H = rng.standard_normal((nh, t))
        assert cu.pcu_spmm_set_halo(op, 1, capi.ip(np.array([0], np.int32)), capi.ip(np.array([0, 0], np.int32)),
                                    capi.ip(np.zeros(1, np.int32)), capi.ip(np.array([0, nh], np.int32))) == 0
        cu.pcu_spmm_halo_buffer.restype = C.c_void_p
        hb = cu.pcu_spmm_halo_buffer(op, t)
        cu.pcu_h2d(dev.ctx, C.c_void_p(hb), H.ctypes.data_as(C.c_void_p), C.c_size_t(H.nbytes))
        dX, dY = dev.up(X), dev.zeros(mm * ld)
        assert cu.pcu_spmm_apply(op, dX, ld, dY, ld, t) == 0, cu.pcu_last_error()
        Y = dev.down(dY, (mm, ld))
        ref = A @ np.vstack([X, H])
        assert np.allclose(Y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
        dev.free(dX, dY)
    cu.pcu_spmm_destroy(op)


@pytest.mark.parametrize("t", [1, 2, 3, 4, 8, 12, 16, 32])
def test_gram_ortho_update_z(dev, t):
    rng = np.random.default_rng(10 + t)
    m = 5003
    ld = t if (t % 2 == 0 or t == 1) else t + 1
    P, AP, R, X, Z, Pp = (rng.standard_normal((m, ld)) for _ in range(6))
    # make AP^T P SPD: AP = P W with W SPD
    W = rng.standard_normal((t, t)); W = W @ W.T + t * np.eye(t)
    AP[:, :t] = P[:, :t] @ W
    dP, dAP, dR, dX, dZ, dPp = (dev.up(a) for a in (P, AP, R, X, Z, Pp))
    small = dev.zeros(8 * t * t + 16)
    sm = small.value
    G, Gpr, U, al, b1, b2, rr = (C.c_void_p(sm + 8 * k * t * t) for k in range(7))
    st = dev.zeros(4, np.int32)
    assert cu.pcu_gram2(dev.ctx, m, t, dAP, ld, dP, ld, G, dP, ld, dR, ld, Gpr) == 0, cu.pcu_last_error()
    Gh = dev.down(G, (t, t)).T  # column-major on the device
    Gprh = dev.down(Gpr, (t, t)).T
    Gref, Gprref = AP[:, :t].T @ P[:, :t], P[:, :t].T @ R[:, :t]
    assert np.allclose(Gh, Gref, rtol=1e-12, atol=1e-12 * np.abs(Gref).max())
    assert np.allclose(Gprh, Gprref, rtol=1e-12, atol=1e-11 * np.abs(Gprref).max())
    # deterministic: a second run gives identical bits
    assert cu.pcu_gram2(dev.ctx, m, t, dAP, ld, dP, ld, G, dP, ld, dR, ld, Gpr) == 0
    assert np.array_equal(dev.down(G, (t, t)).T, Gh)
    assert cu.pcu_ortho_update(dev.ctx, m, t, G, Gpr, dP, ld, dAP, ld, dX, ld, dR, ld, U, al, rr, st) == 0, cu.pcu_last_error()
    Uref = sla.cholesky(np.triu(Gh) + np.triu(Gh, 1).T, lower=False)
    Uh = dev.down(U, (t, t)).T
    assert np.allclose(np.triu(Uh), Uref, rtol=1e-11, atol=1e-12)
    Pn = sla.solve_triangular(Uref, P[:, :t].T, trans="T", lower=False).T
    APn = sla.solve_triangular(Uref, AP[:, :t].T, trans="T", lower=False).T
    alpha = Pn.T @ R[:, :t]                      # the reference's (P U^-1)^T R
    alh = dev.down(al, (t, t)).T
    assert np.allclose(alh, alpha, rtol=1e-9, atol=1e-10 * np.abs(alpha).max())
    Xn, Rn = X[:, :t] + Pn @ alpha, R[:, :t] - APn @ alpha
    scale = lambda a: 1e-10 * np.abs(a).max()
    assert np.allclose(dev.down(dP, (m, ld))[:, :t], Pn, rtol=1e-10, atol=scale(Pn))
    assert np.allclose(dev.down(dAP, (m, ld))[:, :t], APn, rtol=1e-10, atol=scale(APn))
    assert np.allclose(dev.down(dX, (m, ld))[:, :t], Xn, rtol=1e-9, atol=scale(Xn))
    Rh = dev.down(dR, (m, ld))[:, :t]
    assert np.allclose(Rh, Rn, rtol=1e-9, atol=scale(Rn))
    assert abs(dev.down(rr, (1,))[0] - np.sum(Rh ** 2)) <= 1e-12 * np.sum(Rh ** 2)
    assert dev.down(st, (1,), np.int32)[0] == 0
    # beta and the Z update (Orthodir, ref ecg.c:510-517)
    APh, Pnh = dev.down(dAP, (m, ld)), dev.down(dP, (m, ld))
    assert cu.pcu_gram2(dev.ctx, m, t, dAP, ld, dZ, ld, b1, dPp, ld, dZ, ld, b2) == 0
    b1h, b2h = dev.down(b1, (t, t)).T, dev.down(b2, (t, t)).T
    assert np.allclose(b1h, APh[:, :t].T @ Z[:, :t], rtol=1e-11, atol=1e-11 * np.abs(b1h).max())
    assert cu.pcu_update_z(dev.ctx, m, t, dZ, ld, dP, ld, t, b1, dPp, ld, t, b2) == 0, cu.pcu_last_error()
    Zn = Z[:, :t] - Pnh[:, :t] @ b1h - Pp[:, :t] @ b2h
    assert np.allclose(dev.down(dZ, (m, ld))[:, :t], Zn, rtol=1e-10, atol=scale(Zn))
    # not SPD -> status flag, like dpotrf's info
    bad = -np.eye(t)
    dbad = dev.up(np.asfortranarray(bad))
    assert cu.pcu_ortho_update(dev.ctx, m, t, dbad, Gpr, dP, ld, dAP, ld, None, ld, None, ld, U, al, rr, st) == 0
    assert dev.down(st, (1,), np.int32)[0] == 1
    dev.free(dP, dAP, dR, dX, dZ, dPp, small, st, dbad)


def test_split_sum_fro(dev):
    rng = np.random.default_rng(5)
    m, t = 1000, 8
    rhs = rng.standard_normal(m)
    col = (np.arange(m) // 125 % t).astype(np.int32)
    dr, dc, dR, ds, dn = dev.up(rhs), dev.up(col), dev.zeros(m * t), dev.zeros(m), dev.zeros(2)
    assert cu.pcu_split_rhs(dev.ctx, m, t, dr, dc, dR, t) == 0
    R = dev.down(dR, (m, t))
    ref = np.zeros((m, t)); ref[np.arange(m), col] = rhs
    assert np.array_equal(R, ref)
    assert cu.pcu_sum_columns(dev.ctx, m, t, dR, t, ds) == 0
    assert np.array_equal(dev.down(ds, (m,)), rhs)
    assert cu.pcu_fro2(dev.ctx, m, t, dR, t, dn) == 0
    assert abs(dev.down(dn, (1,))[0] - np.sum(rhs ** 2)) < 1e-12 * np.sum(rhs ** 2)
    dev.free(dr, dc, dR, ds, dn)


@pytest.mark.parametrize("gen,N,nblk", [("poisson7", 8, 1), ("poisson7", 12, 3), ("stencil27", 9, 2), ("poisson7", 20, 2)])
@pytest.mark.parametrize("t", [1, 4, 8, 16])
@pytest.mark.parametrize("copies", [2, 1, 23])
def test_block_jacobi_factor_and_solve(dev, gen, N, nblk, t, copies, monkeypatch):
    """pcu_bj_create + pcu_bj_apply against a direct sparse solve of every diagonal block; with the transposed copy of the
    panels (copies = 2) and with the backward sweep reading the forward panels tile by tile (copies = 1)"""
    import scipy.sparse.linalg as spla
    if copies == 23:   # two copies, three panels per warp forced (the sweep kernel streams across a warp's panels) and the
        copies = 2     # warp-per-column assembly on every level
        monkeypatch.setenv("PREALPS_BJ_PW_FORCE", "3")
        monkeypatch.setenv("PREALPS_BJ_ASM_WIDE", "2")
    monkeypatch.setenv("PREALPS_BJ_COPIES", str(copies))
    A = getattr(gen_matrices, gen)(N).tocsr()
    n = A.shape[0]
    cuts = np.linspace(0, n, nblk + 1).astype(np.int32)
    blocks = [A[cuts[b]:cuts[b + 1], cuts[b]:cuts[b + 1]].tocsr() for b in range(nblk)]
    ups = []
    for Bk in blocks:
        U = sp.triu(Bk, format="csr"); U.sort_indices(); ups.append(U)
    keep = [(U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data.copy()) for U in ups]
    rp = (C.POINTER(C.c_int) * nblk)(*[capi.ip(k[0]) for k in keep])
    ci = (C.POINTER(C.c_int) * nblk)(*[capi.ip(k[1]) for k in keep])
    vv = (C.POINTER(C.c_double) * nblk)(*[capi.dp(k[2]) for k in keep])
    bj = C.c_void_p()
    assert cu.pcu_bj_create(dev.ctx, nblk, capi.ip(cuts), rp, ci, vv, C.byref(bj)) == 0, cu.pcu_last_error()
    rng = np.random.default_rng(t)
    ld = t
    B = rng.standard_normal((n, ld))
    dB, dX = dev.up(B), dev.zeros(n * ld)
    assert cu.pcu_bj_apply(bj, dB, ld, dX, ld, t) == 0, cu.pcu_last_error()
    X = dev.down(dX, (n, ld))
    ref = np.vstack([spla.splu(Bk.tocsc()).solve(B[cuts[b]:cuts[b + 1]]) for b, Bk in enumerate(blocks)])
    err = np.linalg.norm(X - ref) / np.linalg.norm(ref)
    assert err < 1e-11, err
    # in place, and bit-reproducible
    assert cu.pcu_bj_apply(bj, dB, ld, dB, ld, t) == 0
    assert np.array_equal(dev.down(dB, (n, ld)), X)
    assert cu.pcu_bj_stat(bj, 0) > 0 and cu.pcu_bj_stat(bj, 1) >= cu.pcu_bj_stat(bj, 0)
    assert cu.pcu_bj_stat(bj, 9) == copies
    dev.free(dB, dX)
    cu.pcu_bj_destroy(bj)


@pytest.mark.parametrize("copies", [2, 1])
def test_block_jacobi_long_panels_cut_across_ctas(dev, copies, monkeypatch):
    """a block large enough for separators beyond 1024 columns: exercises the inter-CTA split of long panels"""
    import scipy.sparse.linalg as spla
    monkeypatch.setenv("PREALPS_BJ_COPIES", str(copies))
    A = gen_matrices.poisson7(36).tocsr()
    n = A.shape[0]
    U = sp.triu(A, format="csr"); U.sort_indices()
    k = (U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data.copy())
    rp = (C.POINTER(C.c_int) * 1)(capi.ip(k[0])); ci = (C.POINTER(C.c_int) * 1)(capi.ip(k[1]))
    vv = (C.POINTER(C.c_double) * 1)(capi.dp(k[2]))
    bj = C.c_void_p()
    assert cu.pcu_bj_create(dev.ctx, 1, capi.ip(np.array([0, n], np.int32)), rp, ci, vv, C.byref(bj)) == 0, cu.pcu_last_error()
    rng = np.random.default_rng(3)
    lu = spla.splu(A.tocsc())
    for t in (8, 4, 16, 32):
        B = rng.standard_normal((n, t))
        dB, dX = dev.up(B), dev.zeros(n * t)
        for rep in range(2):  # the second apply re-uses the arrival counters
            assert cu.pcu_bj_apply(bj, dB, t, dX, t, t) == 0, cu.pcu_last_error()
            X = dev.down(dX, (n, t))
            ref = lu.solve(B)
            assert np.linalg.norm(X - ref) / np.linalg.norm(ref) < 1e-11
        dev.free(dB, dX)
    cu.pcu_bj_destroy(bj)


def test_block_jacobi_rejects_indefinite_block(dev):
    A = (gen_matrices.poisson7(5) - 7.0 * sp.eye(125)).tocsr()
    U = sp.triu(A, format="csr"); U.sort_indices()
    k = (U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data.copy())
    rp = (C.POINTER(C.c_int) * 1)(capi.ip(k[0])); ci = (C.POINTER(C.c_int) * 1)(capi.ip(k[1]))
    vv = (C.POINTER(C.c_double) * 1)(capi.dp(k[2]))
    bj = C.c_void_p()
    assert cu.pcu_bj_create(dev.ctx, 1, capi.ip(np.array([0, 125], np.int32)), rp, ci, vv, C.byref(bj)) != 0
    assert b"positive definite" in cu.pcu_last_error()
