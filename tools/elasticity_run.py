"""tools/elasticity_run.py -- BASELINE config 4 in reduced size on one GPU: Q1 linear elasticity (3 dof/node, up to 81
non-zeros per row), ECG t=16 + block Jacobi with and without the adaptive reduction of the search directions (-r 1).
    python tools/elasticity_run.py [nodes per side = 40] [t = 16] [subdomains = 16]
Prints one JSON line per solve."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))  # the matrix generator only (test infrastructure, not the solver)
import gen_matrices  # noqa: E402
from prealps_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
t = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 16
t0 = time.time()
A = gen_matrices.elasticity3d(n, n, n).tocsr()
A.sort_indices()
t_gen = time.time() - t0
rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
t0 = time.time()
assert capi.lib.preAlps_b200_OperatorBuildCSR(A.shape[0], capi.ip(rp), capi.ip(ci), capi.dp(v), S, 0, S, 1, None) == 0
t_part = time.time() - t0
t0 = time.time()
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
t_bj = time.time() - t0
rhs = capi.driver_rhs(capi.operator_arrays()["m"])
for bs_red in (0, 1, 0, 1):
    sol, hist, info = capi.solve(rhs, t, 1e-8, max_iter=2000, bs_red=bs_red)
    bs = capi.last_block_sizes()
    print(json.dumps({"operator": "elasticity3d %d^3 nodes" % n, "rows": int(A.shape[0]), "nnz_per_row": round(A.nnz / A.shape[0], 1),
                      "t": t, "subdomains": S, "bs_red": bs_red, "iterations": info.iter, "time_to_solution_s": round(info.t_solve, 4),
                      "true_relres": info.true_relres, "block_columns_through_spmm_and_bj": int(bs[:info.iter].sum()),
                      "final_bs": int(bs[info.iter - 1]), "setup_s": {"generate": round(t_gen, 1), "partition": round(t_part, 1),
                                                                      "block_jacobi": round(t_bj, 1)}}), flush=True)
