"""tools/bj_layout_stats.py -- host-only: per-level storage of the block-Jacobi factor of ONE n^3 7-point block under
the symbolic options (PREALPS_BJ_LEAF / PREALPS_BJ_RELAX): supernodes, columns, update rows sum(h-w), dense trapezoid
entries and the doubles the 32-row panel layout of bj.h stores for both sweeps.
    python tools/bj_layout_stats.py [n = 64]"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import gen_matrices  # noqa: E402
import scipy.sparse as sp  # noqa: E402
from prealps_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
A = sp.triu(gen_matrices.poisson7(n)).tocsr()
A.sort_indices()
N = A.shape[0]
rp, ci = A.indptr.astype(np.int32), A.indices.astype(np.int32)
perm = np.zeros(N, np.int32); nsuper = C.c_int(); sn_col = np.zeros(N + 1, np.int32); sn_rowptr = np.zeros(N + 1, np.int64)
cap = 400 * N
sn_rows = np.zeros(cap, np.int32); sn_parent = np.zeros(N, np.int32); sn_level = np.zeros(N, np.int32); st = np.zeros(4)
rc = capi.cuda.pcu_bj_analyze(N, capi.ip(rp), capi.ip(ci), 1, capi.ip(perm), C.byref(nsuper), capi.ip(sn_col),
                              sn_rowptr.ctypes.data_as(C.POINTER(C.c_longlong)), capi.ip(sn_rows), C.c_longlong(cap),
                              capi.ip(sn_parent), capi.ip(sn_level), capi.dp(st))
assert rc == 0, rc
ns = nsuper.value
w = np.diff(sn_col[:ns + 1]).astype(np.int64); h = np.diff(sn_rowptr[:ns + 1]).astype(np.int64); lev = sn_level[:ns]
trap = w * (w + 1) // 2 + (h - w) * w
def fwd_doubles(w_, h_):
    tot = 0
    for p in range((h_ + 31) // 32):
        k = min(w_, 32 * p + 32); k = (k + 3) & ~3
        tot += 32 * k
    return tot
def bwd_doubles(w_, h_):
    tot = 0
    for p in range((w_ + 31) // 32):
        k = h_ - 32 * p; k = (k + 3) & ~3
        tot += 32 * k
    return tot
fd = np.array([fwd_doubles(int(a), int(b)) for a, b in zip(w, h)]); bd = np.array([bwd_doubles(int(a), int(b)) for a, b in zip(w, h)])
print("n=%d rows=%d supernodes=%d levels=%d nnzL exact=%.1fM stored(trapezoid)=%.1fM fwd panels=%.1fM bwd panels=%.1fM nu=%.2fM"
      % (n, N, ns, int(st[2]), st[0] / 1e6, st[1] / 1e6, fd.sum() / 1e6, bd.sum() / 1e6, (h - w).sum() / 1e6))
print("lev   nsn     cols    sum(h-w)   trap(M)   fwd(M)   bwd(M)  fwd/trap  mean_w  mean_h")
for l in range(int(st[2])):
    m = lev == l
    print("%3d %6d %8d %10d %9.2f %8.2f %8.2f %8.2f %7.1f %7.1f" % (l, m.sum(), w[m].sum(), (h - w)[m].sum(), trap[m].sum() / 1e6,
          fd[m].sum() / 1e6, bd[m].sum() / 1e6, fd[m].sum() / max(trap[m].sum(), 1), w[m].mean(), h[m].mean()))
# what a panel layout with a per-panel row count (multiples of G rows instead of always 32) would store
for G in (8, 16):
    def fwdG(w_, h_):
        tot = 0
        for p in range((h_ + 31) // 32):
            k = min(w_, 32 * p + 32); k = (k + 3) & ~3
            r = min(32, h_ - 32 * p); r = (r + G - 1) // G * G
            tot += r * k
        return tot
    def bwdG(w_, h_):
        tot = 0
        for p in range((w_ + 31) // 32):
            k = h_ - 32 * p; k = (k + 3) & ~3
            r = min(32, w_ - 32 * p); r = (r + G - 1) // G * G
            tot += r * k
        return tot
    fg = np.array([fwdG(int(a), int(b)) for a, b in zip(w, h)]); bg = np.array([bwdG(int(a), int(b)) for a, b in zip(w, h)])
    print("row granularity %2d: fwd panels %.1fM bwd panels %.1fM (sum %.1fM vs exact x2 %.1fM = %.3fx)"
          % (G, fg.sum() / 1e6, bg.sum() / 1e6, (fg.sum() + bg.sum()) / 1e6, 2 * st[0] / 1e6, (fg.sum() + bg.sum()) / (2 * st[0])))
