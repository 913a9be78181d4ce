"""oracle/restate.py -- TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's ECG + block-Jacobi path.

Plain numpy/scipy restatement of what NLAFET/preAlps computes on the path
examples/test_ecg_prealps_op.c drives, each function citing the reference lines it follows.
It is pinned against golden vectors produced by the UNMODIFIED reference sources
(tests/golden/*.npz, made by tests/golden/make_golden.py from oracle/_ref/ecg_dump_ref):
tests/test_oracle.py checks every integer map bit-for-bit and the residual history to 1e-10.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product never does.

The one piece that is not the reference's own code is the block solve: the reference calls MKL
PARDISO (cplm_kernels.c:741-853); any exact sparse Cholesky/LU gives the same M^-1 up to rounding
(SURVEY.md H3-v), here scipy's SuperLU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_METIS_A = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"
_libc = C.CDLL("libc.so.6")
_libc.rand.restype = C.c_int
RAND_MAX = 2147483647


def _metis_lib():
    so = os.path.join(_HERE, "_build", "libmetis_oracle.so")
    if not os.path.exists(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", os.path.join(_HERE, "metis_oracle.c"), _METIS_A,
                               "-lm", "-o", so])
    return C.CDLL(so)


def load_mtx(path):
    """CPLM_LoadMatrixMarket (cplm_matcsr.c:96-243): real general|symmetric coordinate, 0/1-base autodetect,
    symmetric files expanded to full storage; rows sorted by column."""
    with open(path) as f:
        hdr = f.readline().split()
        assert hdr[0].lower() == "%%matrixmarket" and hdr[2].lower() == "coordinate" and hdr[3].lower() == "real"
        sym = hdr[4].lower() == "symmetric"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        M, N, nnz = (int(x) for x in line.split())
        data = np.loadtxt(f, ndmin=2)
    i, j, v = data[:, 0].astype(np.int64), data[:, 1].astype(np.int64), data[:, 2]
    if not (i[0] == 0 or j[0] == 0):
        i, j = i - 1, j - 1
    if sym:
        off = i != j
        i, j, v = np.concatenate([i, j[off]]), np.concatenate([j, i[off]]), np.concatenate([v, v[off]])
    A = sp.csr_matrix((v, (i, j)), shape=(M, N))
    A.sort_indices()
    return A, sym


def sym_scale(A):
    """CPLM_MatCSRSymRACScaling (cplm_matcsr.c:1461-1554): a_ij <- (r_i * a_ij) * r_j, r_i = sqrt(1/max_j|a_ij|)."""
    A = A.tocsr().copy()
    r = np.zeros(A.shape[0])
    np.maximum.at(r, np.repeat(np.arange(A.shape[0]), np.diff(A.indptr)), np.abs(A.data))
    r = np.sqrt(1.0 / r)
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    A.data = (r[rows] * A.data) * r[A.indices]
    return A


def kway_parts(A, S):
    """CPLM_metisKwayOrdering -> CPLM_MatCSRDelDiag -> callKway (cplm_v0_matcsr.c:114-167,
    cplm_matcsr_core.c:325-375,394-457): adjacency = pattern without the diagonal, rows sorted."""
    if S == 1:
        return np.zeros(A.shape[0], dtype=np.int32)
    P = (A + A.T).tocsr()
    P.sort_indices()
    rows = np.repeat(np.arange(P.shape[0]), np.diff(P.indptr))
    keep = P.indices != rows
    adj = P.indices[keep].astype(np.int32)
    cnt = np.bincount(rows[keep], minlength=P.shape[0])
    xadj = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    parts = np.zeros(P.shape[0], dtype=np.int32)
    rc = _metis_lib().oracle_kway(C.c_int(P.shape[0]), xadj.ctypes.data_as(C.c_void_p), adj.ctypes.data_as(C.c_void_p),
                                  C.c_int(S), parts.ctypes.data_as(C.c_void_p))
    assert rc == 1
    return parts


def parts_to_perm(parts, S):
    """CPLM_getBlockPosition + CPLM_getIntPermArray (cplm_v0_metis_utils.c:197-222,22-43)."""
    posB = np.concatenate([[0], np.cumsum(np.bincount(parts, minlength=S))]).astype(np.int32)
    perm = np.argsort(parts, kind="stable").astype(np.int32)
    return posB, perm


def permute(A, perm):
    """CPLM_MatCSRPermute(A, B, perm, perm, PERMUTE) (cplm_v0_matcsr.c:941-1022): B = P A P^T, rows sorted."""
    B = A.tocsr()[perm][:, perm].tocsr()
    B.sort_indices()
    return B


def col_block_pos(A_panel, rowPos):
    """CPLM_MatCSRGetColBlockPos (cplm_v0_matcsr.c:175-227)."""
    S = len(rowPos) - 1
    m = A_panel.shape[0]
    cp = np.zeros(m * S + 1, dtype=np.int32)
    for i in range(m):
        cols = A_panel.indices[A_panel.indptr[i]:A_panel.indptr[i + 1]]
        cp[i * S:i * S + S + 1] = A_panel.indptr[i] + np.searchsorted(cols, rowPos, side="left")
        cp[i * S] = A_panel.indptr[i]
        cp[i * S + S] = A_panel.indptr[i + 1]
    return cp


def comm_dep(colPos, m, S, me):
    """CPLM_MatCSRGetCommDep (cplm_v0_matcsr.c:234-273)."""
    cnt = np.zeros(S, dtype=np.int64)
    for j in range(S):
        cnt[j] = np.sum(colPos[np.arange(m) * S + j + 1] - colPos[np.arange(m) * S + j])
    return np.array([j for j in range(S) if cnt[j] and j != me], dtype=np.int32)


def diag_block_upper(A_panel, rowPos, me):
    """CPLM_MatCSRGetDiagBlock(..., SYMMETRIC) (cplm_v0_matcsr.c:287-389): upper triangle incl. diagonal,
    block-local columns."""
    c0, c1 = rowPos[me], rowPos[me + 1]
    D = A_panel[:, c0:c1].tocsr()
    D = sp.triu(D, k=0, format="csr")
    D.sort_indices()
    return D


def driver_rhs(sizes):
    """test_ecg_prealps_op.c:172-184: per rank srand(0), rhs[i] = rand()/RAND_MAX, global norm,
    rhs[i] /= norm for i >= 1 (rhs[0] of every rank stays unscaled)."""
    out, nb = [], 0.0
    for m in sizes:
        _libc.srand(0)
        r = np.array([_libc.rand() / RAND_MAX for _ in range(m)])
        part = 0.0
        for x in r:
            part += x ** 2
        nb += part
        out.append(r)
    nb = np.sqrt(nb)
    for r in out:
        r[1:] /= nb
    return out


class Partitioned:
    """Everything preAlps_OperatorBuild leaves on the S ranks (operator.c:38-134)."""

    def __init__(self, A, S, scale=True, parts=None):
        self.S = S
        self.A_scaled = sym_scale(A) if scale else A.tocsr().copy()
        self.parts = kway_parts(self.A_scaled, S) if parts is None else np.asarray(parts, dtype=np.int32)
        self.posB, self.perm = parts_to_perm(self.parts, S)
        self.Ap = permute(self.A_scaled, self.perm)
        self.rowPos = self.posB
        self.panels = [self.Ap[self.rowPos[r]:self.rowPos[r + 1]].tocsr() for r in range(S)]
        for p in self.panels:
            p.sort_indices()

    def colPos(self, r):
        return col_block_pos(self.panels[r], self.rowPos)

    def dep(self, r):
        return comm_dep(self.colPos(r), self.panels[r].shape[0], self.S, r)

    def diag(self, r):
        return diag_block_upper(self.panels[r], self.rowPos, r)

    def block_solvers(self):
        sol = []
        for r in range(self.S):
            U = self.diag(r)
            full = (U + sp.triu(U, k=1).T).tocsc()
            sol.append(spla.splu(full))
        return sol


def ecg_solve_adapt(P, t, tol, max_iter=1000, rhs=None):
    """_preAlps_ECGIterateOdir with bs_red = ADAPT_BS (ecg.c:402-530, reduction at :445-497) inside the driver
    loop of test_ecg_prealps_op.c:203-223.  The reference keeps V = [slot0 | slot1] and AV as two slots of nrhs
    columns each (ecg.c:126-136); the live search directions are the first bs columns of slot0, the columns
    [bs, nrhs) of slot0 hold every direction discarded so far (they stay in the A-orthogonalisation), slot1 holds
    the previous directions.  beta = AV[:, :kbs]^T Z and Z -= V[:, :kbs] beta use kbs = nrhs + (block size before
    the last reduction) columns (ecg.c:491-496); the end-of-iteration copies move bs columns (ecg.c:521-523).
    Returns the block-size history as well."""
    import scipy.linalg as sla
    S, rowPos, M = P.S, P.rowPos, P.Ap.shape[0]
    sizes = [rowPos[r + 1] - rowPos[r] for r in range(S)]
    if rhs is None:
        rhs = driver_rhs(sizes)
    if S < t:
        raise ValueError("Enlarging factor must be lower than the number of processors")
    sl = [slice(rowPos[r], rowPos[r + 1]) for r in range(S)]
    lus = P.block_solvers()
    A = P.Ap
    nrhs = t

    def gsum(f):
        acc = f(0)
        for r in range(1, S):
            acc = acc + f(r)
        return acc

    def prec(B):
        Z = np.empty_like(B)
        for r in range(S):
            Z[sl[r]] = lus[r].solve(B[sl[r]])
        return Z

    normb = np.sqrt(gsum(lambda r: float(np.sum(rhs[r] ** 2))))
    R = np.zeros((M, nrhs))
    for r in range(S):
        R[sl[r], r % nrhs] = rhs[r]
    X = np.zeros((M, nrhs))
    V = np.zeros((M, 2 * nrhs)); AV = np.zeros((M, 2 * nrhs))
    V[:, :nrhs] = prec(R)
    AV[:, :nrhs] = A @ V[:, :nrhs]
    bs, kbs = nrhs, 2 * nrhs
    cut = tol * normb / np.sqrt(nrhs)                                    # ecg.c:420
    hist, bs_hist, it = [], [], 0
    while True:
        tt = bs                                                          # t = P->info.n
        Pk, AP = V[:, :tt], AV[:, :tt]                                   # views: updates are in place
        G = gsum(lambda r: AP[sl[r]].T @ Pk[sl[r]])
        U = sla.cholesky(np.triu(G) + np.triu(G, 1).T, lower=False)
        Pk[:] = sla.solve_triangular(U, Pk.T, trans="T", lower=False).T
        AP[:] = sla.solve_triangular(U, AP.T, trans="T", lower=False).T
        alpha = gsum(lambda r: Pk[sl[r]].T @ R[sl[r]])                   # tt x nrhs
        Us, sv, _ = np.linalg.svd(alpha, full_matrices=True)             # ecg.c:455 (dgesvd 'O': U overwrites)
        t1 = 0
        for sgm in sv[:tt]:
            if sgm > cut:
                t1 += 1
            else:
                break                                                    # ecg.c:460-464
        if 0 < t1 < nrhs and t1 < tt:                                    # ecg.c:467
            Q = Us                                                       # geqrf of an orthogonal matrix: Q = U diag(+-1)
            alpha = (Q.T @ alpha)[:t1]                                   # ecg.c:474,483
            Pk[:] = Pk @ Q                                               # ecg.c:476-479 (all tt columns rotated,
            AP[:] = AP @ Q                                               #  the last tt - t1 stay in the slot)
            bs, kbs = t1, tt + nrhs                                      # ecg.c:491-492
        X = X + V[:, :bs] @ alpha                                        # ecg.c:500-501
        R = R - AV[:, :bs] @ alpha
        it += 1
        res = np.sqrt(np.trace(gsum(lambda r: R[sl[r]].T @ R[sl[r]])))
        hist.append(res); bs_hist.append(bs)
        if not (res > normb * tol and it < max_iter and bs > 0):         # ecg.c:264
            break
        Z = prec(AV[:, :bs])                                             # test_ecg_prealps_op.c:219
        beta = gsum(lambda r: AV[sl[r], :kbs].T @ Z[sl[r]])              # ecg.c:510-514
        Z = Z - V[:, :kbs] @ beta                                        # ecg.c:517
        V[:, nrhs:nrhs + bs] = V[:, :bs]                                 # ecg.c:521-523
        AV[:, nrhs:nrhs + bs] = AV[:, :bs]
        V[:, :bs] = Z
        AV[:, :bs] = A @ V[:, :bs]                                       # test_ecg_prealps_op.c:212
    sol = X.sum(axis=1)
    b = np.concatenate(rhs)
    return {"iter": it, "res_hist": np.array(hist), "bs_hist": np.array(bs_hist), "sol": sol, "normb": normb,
            "true_relres": np.linalg.norm(b - A @ sol) / np.linalg.norm(b), "rhs": rhs}


def ecg_solve_fused_adapt(P, t, tol, max_iter=1000, rhs=None):
    """_preAlps_ECGIterateOdirFused with bs_red = ADAPT_BS (ecg.c:532-658, reduction at :593-641) inside the loop of
    test_ecg_bench_fused.c:245-259.  Same slot layout as ecg_solve_adapt; the Gram products are taken before the
    normalisation by U = chol(AP^T P) and the residual test lags one iteration."""
    import scipy.linalg as sla
    S, rowPos, M = P.S, P.rowPos, P.Ap.shape[0]
    sizes = [rowPos[r + 1] - rowPos[r] for r in range(S)]
    if rhs is None:
        rhs = driver_rhs(sizes)
    sl = [slice(rowPos[r], rowPos[r + 1]) for r in range(S)]
    lus = P.block_solvers()
    A = P.Ap
    nrhs = t

    def gsum(f):
        acc = f(0)
        for r in range(1, S):
            acc = acc + f(r)
        return acc

    def prec(B):
        Z = np.empty_like(B)
        for r in range(S):
            Z[sl[r]] = lus[r].solve(B[sl[r]])
        return Z

    normb = np.sqrt(gsum(lambda r: float(np.sum(rhs[r] ** 2))))
    R = np.zeros((M, nrhs))
    for r in range(S):
        R[sl[r], r % nrhs] = rhs[r]
    X = np.zeros((M, nrhs))
    V = np.zeros((M, 2 * nrhs)); AV = np.zeros((M, 2 * nrhs))
    V[:, :nrhs] = prec(R)
    bs, kbs = nrhs, 2 * nrhs
    cut = tol * normb / np.sqrt(nrhs)
    hist, bs_hist, it = [], [], 0
    rsolve = lambda U, B: sla.solve_triangular(U, B.T, trans="T", lower=False).T   # B U^-1
    while True:
        tt = bs
        AV[:, :tt] = A @ V[:, :tt]                                       # test_ecg_bench_fused.c:252
        Z = prec(AV[:, :tt])                                             # :253
        Pk, AP = V[:, :tt], AV[:, :tt]
        alpha = gsum(lambda r: Pk[sl[r]].T @ R[sl[r]])                   # ecg.c:557 (tt x nrhs)
        beta = gsum(lambda r: AV[sl[r], :kbs].T @ Z[sl[r]])              # ecg.c:558 (kbs x tt)
        mu = gsum(lambda r: AP[sl[r]].T @ Pk[sl[r]])                     # ecg.c:559
        rtr = gsum(lambda r: R[sl[r]].T @ R[sl[r]])                      # ecg.c:560
        res = np.sqrt(np.trace(rtr))
        hist.append(res)
        conv = res < tol * normb or it > max_iter                        # ecg.c:571
        U = sla.cholesky(np.triu(mu) + np.triu(mu, 1).T, lower=False)    # ecg.c:577
        Pk[:] = rsolve(U, Pk); AP[:] = rsolve(U, AP)                     # ecg.c:580-581
        beta = rsolve(U, beta); Z = rsolve(U, Z)                         # ecg.c:582-583
        alpha = sla.solve_triangular(U, alpha, trans="T", lower=False)   # ecg.c:584
        beta[:tt] = sla.solve_triangular(U, beta[:tt], trans="T", lower=False)   # ecg.c:586 (first t rows)
        Z = Z - V[:, :kbs] @ beta                                        # ecg.c:590
        Us, sv, _ = np.linalg.svd(alpha, full_matrices=True)             # ecg.c:600
        t1 = 0
        for sgm in sv[:tt]:
            if sgm > cut:
                t1 += 1
            else:
                break
        if 0 < t1 < nrhs and t1 < tt:                                    # ecg.c:609
            Q = Us
            alpha = (Q.T @ alpha)[:t1]
            Pk[:] = Pk @ Q; AP[:] = AP @ Q                               # ecg.c:617-620
            Z = (Z @ Q)[:, :t1]                                          # ecg.c:621-622,631
            bs, kbs = t1, tt + nrhs                                      # ecg.c:635-636
        bs_hist.append(bs)                                               # what the driver sees after Iterate
        X = X + V[:, :bs] @ alpha                                        # ecg.c:644-645
        R = R - AV[:, :bs] @ alpha
        it += 1
        V[:, nrhs:nrhs + bs] = V[:, :bs]                                 # ecg.c:651-653
        AV[:, nrhs:nrhs + bs] = AV[:, :bs]
        V[:, :bs] = Z
        if conv:
            break
    sol = X.sum(axis=1)
    b = np.concatenate(rhs)
    return {"iter": it, "res_hist": np.array(hist), "bs_hist": np.array(bs_hist), "sol": sol, "normb": normb,
            "true_relres": np.linalg.norm(b - A @ sol) / np.linalg.norm(b), "rhs": rhs}


def ecg_solve(P, t, tol, max_iter=1000, ortho=0, rhs=None, rrqr=False):
    """_preAlps_ECGIterateOdir / Omin with the driver loop (ecg.c:98-171,223-271,289-530;
    test_ecg_prealps_op.c:203-223), NO_BS_RED.  Global arrays; reductions summed over ranks in rank order
    like the oracle's MPI shim.  rrqr=True (Orthomin only) adds the ADAPT_BS branch of ecg.c:360-393 -- the
    rank-revealing Cholesky QR of the new directions.  The reference is consistent up to and including its first rank
    drop (pinned by the golden poisson7_n4_s4_t4_omin_adapt_rankdrop); afterwards it forms P^T P with the stale column
    count of P's info (ecg.c:366 after the copy of nrhs columns at :357).  From there this follows what the reduction
    stands for: every iteration the t new directions Z - P beta are orthonormalised again and `rank` of them are kept.
    Returns bs_hist (block size after each iteration) as well."""
    S, rowPos, M = P.S, P.rowPos, P.Ap.shape[0]
    sizes = [rowPos[r + 1] - rowPos[r] for r in range(S)]
    if rhs is None:
        rhs = driver_rhs(sizes)
    if S < t:
        raise ValueError("Enlarging factor must be lower than the number of processors")
    sl = [slice(rowPos[r], rowPos[r + 1]) for r in range(S)]
    lus = P.block_solvers()
    A = P.Ap

    def gsum(f):  # sum over ranks of a local t x t product, rank order
        acc = f(0)
        for r in range(1, S):
            acc = acc + f(r)
        return acc

    def prec(B):
        Z = np.empty_like(B)
        for r in range(S):
            Z[sl[r]] = lus[r].solve(B[sl[r]])
        return Z

    normb = np.sqrt(gsum(lambda r: float(np.sum(rhs[r] ** 2))))
    R = np.zeros((M, t))
    for r in range(S):
        R[sl[r], r % t] = rhs[r]          # ecg.c:162, _preAlps_ECGSplit
    X = np.zeros((M, t))
    Pk = prec(R)
    AP = A @ Pk
    Pp, APp = np.zeros((M, t)), np.zeros((M, t))
    hist, it = [], 0
    import scipy.linalg as sla
    if ortho == 2:
        # _preAlps_ECGIterateOdirFused (ecg.c:532-658) inside the loop of test_ecg_bench_fused.c:245-259
        rsolve = lambda U, B: sla.solve_triangular(U, B.T, trans="T", lower=False).T   # B U^-1
        while True:
            Z = prec(AP)
            alpha = gsum(lambda r: Pk[sl[r]].T @ R[sl[r]])                          # ecg.c:557
            b1 = gsum(lambda r: AP[sl[r]].T @ Z[sl[r]])                             # ecg.c:558 (beta = AV^T Z)
            b2 = gsum(lambda r: APp[sl[r]].T @ Z[sl[r]])
            mu = gsum(lambda r: AP[sl[r]].T @ Pk[sl[r]])                            # ecg.c:559
            rtr = gsum(lambda r: R[sl[r]].T @ R[sl[r]])                             # ecg.c:560
            res = np.sqrt(np.trace(rtr))                                            # ecg.c:566-569 (lags one iteration)
            hist.append(res)
            conv = res < tol * normb or it > max_iter                               # ecg.c:571
            U = sla.cholesky(np.triu(mu) + np.triu(mu, 1).T, lower=False)           # ecg.c:577
            Pk, AP, Z = rsolve(U, Pk), rsolve(U, AP), rsolve(U, Z)                  # ecg.c:580-583
            b1, b2 = rsolve(U, b1), rsolve(U, b2)                                   # ecg.c:582 (whole 2t x t beta)
            alpha = sla.solve_triangular(U, alpha, trans="T", lower=False)          # ecg.c:584
            b1 = sla.solve_triangular(U, b1, trans="T", lower=False)                # ecg.c:586
            Z = Z - Pk @ b1 - Pp @ b2                                               # ecg.c:590
            X = X + Pk @ alpha                                                      # ecg.c:644-645
            R = R - AP @ alpha
            it += 1
            Pp, APp, Pk = Pk, AP, Z                                                 # ecg.c:651-653
            if conv:
                break
            AP = A @ Pk
        sol = X.sum(axis=1)
        b = np.concatenate(rhs)
        return {"iter": it, "res_hist": np.array(hist), "sol": sol, "normb": normb,
                "true_relres": np.linalg.norm(b - A @ sol) / np.linalg.norm(b), "rhs": rhs}
    bs_hist = []
    while True:
        G = gsum(lambda r: AP[sl[r]].T @ Pk[sl[r]])                     # ecg.c:425-428
        U = sla.cholesky(np.triu(G) + np.triu(G, 1).T, lower=False)      # ecg.c:431 ('U' triangle)
        Pk = sla.solve_triangular(U, Pk.T, trans="T", lower=False).T     # P U^-1, ecg.c:434
        AP = sla.solve_triangular(U, AP.T, trans="T", lower=False).T
        alpha = gsum(lambda r: Pk[sl[r]].T @ R[sl[r]])                   # ecg.c:438-442
        X = X + Pk @ alpha                                               # ecg.c:500-501
        R = R - AP @ alpha
        it += 1
        res = np.sqrt(np.trace(gsum(lambda r: R[sl[r]].T @ R[sl[r]])))   # ecg.c:250-261
        hist.append(res)
        bs_hist.append(Pk.shape[1])
        if not (res > normb * tol and it < max_iter and Pk.shape[1] > 0):   # ecg.c:264
            break
        if ortho == 0:
            Z = prec(AP)                                                 # test_ecg_prealps_op.c:219
            b1 = gsum(lambda r: AP[sl[r]].T @ Z[sl[r]])                  # ecg.c:510-514 (beta = AV^T Z)
            b2 = gsum(lambda r: APp[sl[r]].T @ Z[sl[r]])
            Z = Z - Pk @ b1 - Pp @ b2                                    # ecg.c:517
            Pp, APp, Pk = Pk, AP, Z                                      # ecg.c:521-523
        else:
            Z = prec(R)                                                  # test_ecg_prealps_op.c:217
            b = gsum(lambda r: AP[sl[r]].T @ Z[sl[r]])                   # ecg.c:347-352
            Pk = Z - Pk @ b                                              # ecg.c:354-358
            if rrqr:
                C = gsum(lambda r: Pk[sl[r]].T @ Pk[sl[r]])              # ecg.c:366-372
                Uf, piv, rank, info = sla.lapack.dpstrf(np.triu(C), lower=0, tol=-1.0)   # ecg.c:375
                Pk = sla.solve_triangular(np.triu(Uf)[:rank, :rank], Pk[:, piv[:rank] - 1].T, trans="T", lower=False).T  # ecg.c:380-391
                # ecg.c:391 sets ecg->bs = rank: the driver sees it at the next stopping test (bs_hist above)
        AP = A @ Pk
    sol = X.sum(axis=1)                                                  # ecg.c:674
    b = np.concatenate(rhs)
    true_rel = np.linalg.norm(b - A @ sol) / np.linalg.norm(b)
    return {"iter": it, "res_hist": np.array(hist), "sol": sol, "normb": normb, "true_relres": true_rel,
            "rhs": rhs, "bs_hist": np.array(bs_hist, dtype=np.int32)}
