"""One ECG + block-Jacobi solve on the CPU emulation of the whole stack (tests/_build/emul_lib: every .cu of libprealps_cuda
compiled against tests/emul/cuda_emul.h + the plain-C host layer).  Run by tests/test_full_emul.py in a process of its own
(the binding loads its libraries once):   python full_solve_case.py <golden case | poisson7:N:S:t:tol>
Prints one JSON line; the residual history as hex floats (bit-for-bit comparisons between kernel variants)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import build_bj_emul  # noqa: E402

os.environ["PREALPS_B200_LIBDIR"] = build_bj_emul.build_full()
import numpy as np  # noqa: E402
import gen_matrices  # noqa: E402
from prealps_b200 import capi  # noqa: E402

spec = sys.argv[1]
parts = None
if spec == "rankdrop":  # tests/gpu_util.py: decoupled_block_case
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import gpu_util
    A, parts, S, t, tol = gpu_util.decoupled_block_case()
    ortho, bs_red = 1, 1
elif spec.startswith("poisson7:"):
    _, N, S, t, tol = spec.split(":")
    N, S, t, tol, ortho, bs_red = int(N), int(S), int(t), float(tol), 0, 0
    A = gen_matrices.poisson7(N).tocsr()
else:
    g = np.load(os.path.join(ROOT, "tests", "golden", spec + ".npz"))
    A = gen_matrices.build(g["gen"], g["N"]).tocsr()
    S, t, tol, ortho = int(g["S"]), int(g["t"]), float(g["tol"]), int(g["ortho"])
    bs_red = int(g["bs_red"]) if "bs_red" in g.files else 0
A.sort_indices()
rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
assert capi.lib.preAlps_b200_OperatorBuildCSR(A.shape[0], capi.ip(rp), capi.ip(ci), capi.dp(v), S, 0, S, 1,
                                               capi.ip(parts) if parts is not None else None) == 0
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
arr = capi.operator_arrays()
rhs = capi.driver_rhs(arr["m"])
sol, hist, info = capi.solve(rhs, t, tol, ortho=ortho, bs_red=bs_red)
if spec == "rankdrop":
    gpu_util.check_rank_drop_solve(A, parts, S, t, tol, sol, hist, info, capi.last_block_sizes())
print(json.dumps({"iter": info.iter, "res": info.res, "true_relres": info.true_relres, "hist": [float(x).hex() for x in hist],
                  "sol_sum": float(np.sum(sol)).hex(), "launches": int(capi.stat("launches")),
                  "bs_hist": [int(b) for b in capi.last_block_sizes()[:len(hist)]]}))
capi.lib.preAlps_OperatorFree()
