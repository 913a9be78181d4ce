"""The t x t host algebra behind ADAPT_BS (prealps_b200/host/pa_ecg.c) against numpy/scipy: Cholesky, triangular
inverse, left singular vectors by one-sided Jacobi (stands in for LAPACKE_dgesvd + dgeqrf/dormqr, ref: ecg.c:455-479),
pivoted Cholesky (LAPACKE_dpstrf, ref: ecg.c:375)."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg as sla

from prealps_b200 import capi

lib = capi.lib


def F(a):
    return np.asfortranarray(a, dtype=np.float64)


@pytest.mark.parametrize("n", [1, 3, 8, 16, 32])
def test_cholesky_and_triangular_inverse(n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n + 5, n))
    A = F(M.T @ M + 0.1 * np.eye(n))
    U = A.copy(order="F")
    assert lib.pa_h_chol_upper(n, capi.dp(U), n) == 0
    Uref = sla.cholesky(A, lower=False)
    assert np.allclose(np.triu(U), Uref, rtol=1e-12, atol=1e-13)
    Ui = F(np.zeros((n, n)))
    lib.pa_h_triu_inv(n, capi.dp(U), n, capi.dp(Ui), n)
    assert np.allclose(np.triu(U) @ Ui, np.eye(n), atol=1e-11)
    # not positive definite: the failing pivot (1-based) is returned like dpotrf's info
    B = F(np.eye(n)); B[n - 1, n - 1] = -1.0
    assert lib.pa_h_chol_upper(n, capi.dp(B), n) == n


@pytest.mark.parametrize("t,n", [(1, 4), (3, 8), (8, 8), (5, 16), (16, 16), (32, 32)])
def test_left_singular_vectors(t, n):
    rng = np.random.default_rng(100 * t + n)
    # graded singular values, like alpha = P^T R near convergence
    Us, _ = np.linalg.qr(rng.standard_normal((t, t)))
    Vs, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sig = 10.0 ** (-np.arange(t, dtype=float))
    A = F(Us @ np.diag(sig) @ Vs[:, :t].T)
    sv = np.zeros(t); Q = F(np.zeros((t, t))); rows = np.zeros(t * n)
    lib.pa_h_left_svd(t, n, capi.dp(A), t, capi.dp(sv), capi.dp(Q), capi.dp(rows))
    # LAPACK's singular values carry an absolute error of a few eps * sigma_max; one-sided Jacobi is relatively accurate
    assert np.allclose(sv, np.linalg.svd(A, compute_uv=False), rtol=1e-10, atol=1e-15 * sig[0])
    assert np.allclose(sv[:min(t, 12)], sig[:min(t, 12)], rtol=1e-9)
    assert np.all(np.diff(sv) <= 0)
    assert np.allclose(Q.T @ Q, np.eye(t), atol=1e-12)
    R = rows.reshape(t, n)
    assert np.allclose(R, Q.T @ A, atol=1e-13)                      # rows = Q^T A
    assert np.allclose(np.linalg.norm(R, axis=1), sv, rtol=1e-12)   # mutually orthogonal rows of norm sigma_i
    G = R @ R.T
    assert np.allclose(G - np.diag(np.diag(G)), 0.0, atol=1e-12 * sv[0] ** 2)


def test_left_singular_vectors_of_rank_deficient_block():
    A = F(np.zeros((4, 6))); A[0, 0] = 2.0; A[1, 1] = 1e-12
    sv = np.zeros(4); Q = F(np.zeros((4, 4))); rows = np.zeros(24)
    lib.pa_h_left_svd(4, 6, capi.dp(A), 4, capi.dp(sv), capi.dp(Q), capi.dp(rows))
    assert np.allclose(sv, [2.0, 1e-12, 0.0, 0.0])


@pytest.mark.parametrize("n", [2, 4, 8, 16])
def test_pivoted_cholesky_matches_dpstrf(n):
    rng = np.random.default_rng(7 * n)
    M = rng.standard_normal((n + 3, n)) * (10.0 ** -rng.integers(0, 4, n))[None, :]
    A = F(M.T @ M)
    W = A.copy(order="F")
    piv = np.zeros(n, dtype=np.int32)
    rank = lib.pa_h_pivoted_chol(n, capi.dp(W), n, capi.ip(piv), C.c_double(-1.0))
    Uf, pref, rref, info = sla.lapack.dpstrf(np.triu(A), lower=0, tol=-1.0)
    assert rank == rref == n
    assert np.array_equal(piv, pref - 1)
    assert np.allclose(np.triu(W), np.triu(Uf), rtol=1e-10, atol=1e-14)
    U = np.triu(W)
    assert np.allclose(U.T @ U, A[np.ix_(piv, piv)], rtol=1e-11, atol=1e-14)
    # a rank-deficient matrix stops at the numerical rank
    B = F(M[:, :n - 1].T @ M[:, :n - 1]); B2 = F(np.zeros((n, n))); B2[:n - 1, :n - 1] = B; B2[n - 1, n - 1] = 0.0
    piv2 = np.zeros(n, dtype=np.int32)
    assert lib.pa_h_pivoted_chol(n, capi.dp(B2), n, capi.ip(piv2), C.c_double(-1.0)) == n - 1
