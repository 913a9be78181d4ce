"""Host C layer (libprealps_b200.so) integer pipeline against the reference's golden vectors, bit for bit.
No GPU needed: these entry points never touch the device."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import gen_matrices
from conftest import GOLDEN, golden_cases
from prealps_b200 import capi

lib = capi.lib


def _pipeline(g):
    """load -> scale -> k-way -> perm -> permute, all through the library's C functions"""
    A = gen_matrices.build(g["gen"], g["N"])
    S = int(g["S"])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "A.mtx")
        gen_matrices.write_mtx(path, A)
        G = capi.MatCSR()
        assert lib.pa_load_mtx(path.encode(), C.byref(G), 0) == 0
    assert lib.pa_sym_scale(C.byref(G)) == 0
    M = G.info.m
    parts = np.zeros(M, dtype=np.int32)
    assert lib.pa_kway_parts(C.byref(G), S, capi.ip(parts)) == 0
    posB = np.zeros(S + 1, dtype=np.int32)
    perm = np.zeros(M, dtype=np.int32)
    lib.pa_parts_to_perm(M, capi.ip(parts), S, capi.ip(posB), capi.ip(perm))
    P = capi.MatCSR()
    assert lib.pa_permute_sym(C.byref(G), capi.ip(perm), C.byref(P)) == 0
    return G, P, parts, posB, perm


@pytest.mark.parametrize("name", golden_cases())
def test_pipeline_bit_exact(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    S = int(g["S"])
    G, P, parts, posB, perm = _pipeline(g)
    rp, ci, v = G.arrays()
    assert G.info.structure == 1  # symmetric file: structure stays SYMMETRIC after expansion
    assert np.array_equal(rp, g["S_rowPtr"]) and np.array_equal(ci, g["S_colInd"]) and np.array_equal(v, g["S_val"])
    assert np.array_equal(perm, g["perm"])
    assert np.array_equal(posB, g["posB"])
    for r in range(S):
        B = capi.MatCSR()
        lib.pa_row_panel(C.byref(P), int(posB[r]), int(posB[r + 1]), C.byref(B))
        rp, ci, v = B.arrays()
        assert np.array_equal(rp, g["r%d_A_rowPtr" % r])
        assert np.array_equal(ci, g["r%d_A_colInd" % r])
        assert np.array_equal(v, g["r%d_A_val" % r])
        assert B.info.M == int(g["M"]) and B.info.m == posB[r + 1] - posB[r]
        cp, n = capi.c_int_p(), C.c_int()
        lib.pa_col_block_pos(C.byref(B), capi.ip(posB), S, C.byref(cp), C.byref(n))
        colPos = np.ctypeslib.as_array(cp, shape=(n.value,)).copy()
        assert np.array_equal(colPos, g["r%d_colPos" % r])
        dpp, nd = capi.c_int_p(), C.c_int()
        lib.pa_comm_dep(cp, B.info.m, S, r, r + 1, C.byref(dpp), C.byref(nd))
        dep = np.ctypeslib.as_array(dpp, shape=(max(nd.value, 1),))[:nd.value].copy()
        assert np.array_equal(dep, g["r%d_dep" % r])
        D = capi.MatCSR()
        lib.pa_diag_block(C.byref(B), capi.ip(posB), cp, S, r, 0, B.info.m, C.byref(D))
        rp, ci, v = D.arrays()
        assert np.array_equal(rp, g["r%d_D_rowPtr" % r])
        assert np.array_equal(ci, g["r%d_D_colInd" % r])
        assert np.array_equal(v, g["r%d_D_val" % r])
        # halo map: sorted unique off-range columns; local renumbering is consistent
        hp, nh, cl = capi.c_int_p(), C.c_int(), capi.c_int_p()
        lib.pa_halo_map(C.byref(B), int(posB[r]), int(posB[r + 1]), C.byref(hp), C.byref(nh), C.byref(cl))
        halo = np.ctypeslib.as_array(hp, shape=(max(nh.value, 1),))[:nh.value].copy()
        gcols = g["r%d_A_colInd" % r]
        off = gcols[(gcols < posB[r]) | (gcols >= posB[r + 1])]
        assert np.array_equal(halo, np.unique(off))
        loc = np.ctypeslib.as_array(cl, shape=(len(gcols),)).copy()
        m = B.info.m
        back = np.where(loc < m, loc + posB[r], halo[np.clip(loc - m, 0, max(len(halo) - 1, 0))] if len(halo) else 0)
        assert np.array_equal(back, gcols)


def test_stencil_generators_match_python():
    for kind, gen in ((0, gen_matrices.poisson7), (1, gen_matrices.stencil27)):
        A = gen(6)
        G = capi.MatCSR()
        assert lib.pa_stencil_csr(kind, 6, C.byref(G)) == 0
        rp, ci, v = G.arrays()
        assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices) and np.array_equal(v, A.data)


def test_elasticity_generator_matches_python():
    """kind 2 = Q1 linear elasticity (BASELINE config 4): same values to rounding, exactly symmetric"""
    import scipy.sparse as sp
    for N in (3, 6):
        A = gen_matrices.elasticity3d(N, N, N)
        G = capi.MatCSR()
        assert lib.pa_stencil_csr(2, N, C.byref(G)) == 0
        rp, ci, v = G.arrays()
        B = sp.csr_matrix((v, ci, rp), shape=A.shape)
        assert abs(A - B).max() <= 1e-12 * abs(A).max()
        assert abs(B - B.T).max() == 0.0
        # couplings that vanish analytically are not stored (the python generator keeps them as rounding noise)
        tol = 1e-9 * abs(A).max()
        Ab = A.copy(); Ab.data[np.abs(Ab.data) < tol] = 0.0; Ab.eliminate_zeros(); Ab.sort_indices()
        assert np.array_equal(rp, Ab.indptr) and np.array_equal(ci, Ab.indices)
        assert np.abs(v).min() > tol


def test_loader_general_and_zero_based(tmp_path):
    A = gen_matrices.poisson7(4)
    p = tmp_path / "g.mtx"
    gen_matrices.write_mtx(str(p), A, symmetric=False)
    G = capi.MatCSR()
    assert lib.pa_load_mtx(str(p).encode(), C.byref(G), 0) == 0
    rp, ci, v = G.arrays()
    assert G.info.structure == 0
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices) and np.array_equal(v, A.data)
    # 0-based file: first triple has a zero index (ref: cplm_matcsr.c:178-184)
    coo = A.tocoo()
    q = tmp_path / "z.mtx"
    with open(q, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (A.shape[0], A.shape[1], coo.nnz))
        for i, j, x in zip(coo.row, coo.col, coo.data):
            f.write("%d %d %.17g\n" % (i, j, x))
    Z = capi.MatCSR()
    assert lib.pa_load_mtx(str(q).encode(), C.byref(Z), 0) == 0
    rp, ci, v = Z.arrays()
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices)


def test_unsymmetric_pattern_is_symmetrised_for_metis():
    # general-format input goes through the A + A^T pattern (ref: CPLM_MatCSRSymStruct)
    A = gen_matrices.poisson7(5)
    rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    G1 = capi.csr_from_arrays(rp, ci, v, symmetric=True)
    G0 = capi.csr_from_arrays(rp, ci, v, symmetric=False)
    p1 = np.zeros(A.shape[0], dtype=np.int32)
    p0 = np.zeros(A.shape[0], dtype=np.int32)
    lib.pa_kway_parts(C.byref(G1), 4, capi.ip(p1))
    lib.pa_kway_parts(C.byref(G0), 4, capi.ip(p0))
    assert np.array_equal(p0, p1)
    assert set(np.unique(p1)) == {0, 1, 2, 3}
