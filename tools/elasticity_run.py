"""tools/elasticity_run.py -- BASELINE config 4 in reduced size on one GPU: Q1 linear elasticity (3 dof/node, up to 81
non-zeros per row), ECG t=16 + block Jacobi with and without the adaptive reduction of the search directions (-r 1).
    python tools/elasticity_run.py [nodes per side = 40] [t = 16] [subdomains = 16] [repetitions = 2]
Prints one JSON line per solve."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prealps_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
t = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 16
# the native generator (pa_stencil_csr kind 2, checked against oracle/gen_matrices.py: elasticity3d on the CPU)
t0 = time.time()
assert capi.lib.preAlps_b200_OperatorBuildStencil(2, n, S, 0, S) == 0
t_part = time.time() - t0

arr = capi.operator_arrays()
rows, nnz = int(arr["m"]), int(len(arr["A_colInd"]))
t0 = time.time()
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
t_bj = time.time() - t0
rhs = capi.driver_rhs(rows)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
for bs_red in (0, 1) * reps:
    sol, hist, info = capi.solve(rhs, t, 1e-8, max_iter=2000, bs_red=bs_red)
    bs = capi.last_block_sizes()
    print(json.dumps({"operator": "elasticity3d %d^3 nodes" % n, "rows": rows, "nnz_per_row": round(nnz / rows, 1),
                      "t": t, "subdomains": S, "bs_red": bs_red, "iterations": info.iter, "time_to_solution_s": round(info.t_solve, 4),
                      "true_relres": info.true_relres, "block_columns_through_spmm_and_bj": int(bs[:info.iter].sum()),
                      "final_bs": int(bs[info.iter - 1]), "setup_s": {"generate_and_partition": round(t_part, 1),
                                                                      "block_jacobi": round(t_bj, 1)}}), flush=True)
