"""tools/spmm_trace.py -- where does a CTA of spmm_kernel spend its life?  Per-CTA clock64() stamps
(PREALPS_SPMM_TRACE=1): start -> row block known -> col/val staged -> done.  Poisson N^3, one block, no halo."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["PREALPS_SPMM_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_matrices  # noqa: E402
from prealps_b200 import capi  # noqa: E402

cu = capi.cuda
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
t = int(sys.argv[2]) if len(sys.argv) > 2 else 8
A = gen_matrices.poisson7(N).tocsr()
m = A.shape[0]
rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
mode = sys.argv[3] if len(sys.argv) > 3 else "real"
rows_of = np.repeat(np.arange(m, dtype=np.int32), np.diff(rp))
if mode == "self":        # every entry reads the row's own X row: no gather traffic beyond the X stream itself
    ci = rows_of.copy()
elif mode == "near":      # entries read rows r-3..r+3 (all inside the CTA's own 128 rows or next to them)
    ci = np.clip(rows_of + (np.arange(ci.size, dtype=np.int32) - rp[rows_of]) - 3, 0, m - 1).astype(np.int32)
ctx = C.c_void_p()
assert cu.pcu_ctx_create(0, C.byref(ctx)) == 0
op = C.c_void_p()
assert cu.pcu_spmm_create(ctx, m, 0, capi.ip(rp), capi.ip(ci), capi.dp(v), C.byref(op)) == 0, cu.pcu_last_error()
cu.pcu_malloc.restype = C.c_void_p
X = C.c_void_p(cu.pcu_malloc(ctx, C.c_size_t(m * t * 8)))
Y = C.c_void_p(cu.pcu_malloc(ctx, C.c_size_t(m * t * 8)))
x = np.random.default_rng(0).standard_normal(m * t)
assert cu.pcu_h2d(ctx, X, capi.dp(x), C.c_size_t(m * t * 8)) == 0
ms_all = []
for rep in range(5):
    cu.pcu_flush_l2(ctx)
    cu.pcu_timer_start(ctx, 0)
    assert cu.pcu_spmm_apply(op, X, t, Y, t, t) == 0, cu.pcu_last_error()
    cu.pcu_timer_stop(ctx, 0)
    ms = C.c_float()
    cu.pcu_timer_elapsed_ms(ctx, 0, C.byref(ms))
    ms_all.append(ms.value)
cu.pcu_sync(ctx)
print("mode %s: kernel %.1f us (with tracing stores)" % (mode, 1e3 * sorted(ms_all)[2]))
nb = (m + 31) // 32
buf = np.zeros(4 * nb, dtype=np.int64)
cu.pcu_spmm_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
n = cu.pcu_spmm_trace(op, buf.ctypes.data_as(C.c_void_p), nb)
tr = buf[:4 * n].reshape(n, 4)
d = np.diff(tr, axis=1).astype(np.float64)
clk = 1.9  # GHz, SM clock under load
print("CTAs %d; mean cycles (us at %.1f GHz): descriptor chain %.0f (%.2f), staging + barrier %.0f (%.2f), rows %.0f (%.2f), total %.0f (%.2f)"
      % (n, clk, d[:, 0].mean(), d[:, 0].mean() / clk / 1e3, d[:, 1].mean(), d[:, 1].mean() / clk / 1e3, d[:, 2].mean(),
         d[:, 2].mean() / clk / 1e3, d.sum(axis=1).mean(), d.sum(axis=1).mean() / clk / 1e3))
for q in (10, 50, 90):
    print("  p%d: %s" % (q, np.percentile(d, q, axis=0).round(0)))
