"""Multi-GPU path (one process per GPU, NCCL halo exchange + all-reduce).  Skipped on a 1-GPU box."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from prealps_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import ctypes as C, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
from prealps_b200 import capi
import gen_matrices
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); capi.lib.preAlps_b200_SetDevice(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = (C.c_ubyte * 128)(); assert capi.lib.preAlps_b200_NcclUniqueId(buf) == 0
    uid = torch.tensor(list(buf), dtype=torch.uint8)
uid = uid.cuda(); dist.broadcast(uid, 0)
assert capi.lib.preAlps_b200_InitNccl(world, rank, bytes(uid.cpu().tolist())) == 0
g = np.load(os.path.join(%(root)r, "tests", "golden", %(case)r + ".npz"))
S, t, tol = int(g["S"]), int(g["t"]), float(g["tol"])
A = gen_matrices.build(g["gen"], g["N"]).tocsr(); A.sort_indices()
rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
per = S // world
assert capi.lib.preAlps_b200_OperatorBuildCSR(A.shape[0], capi.ip(rp), capi.ip(ci), capi.dp(v), S, rank * per, (rank + 1) * per, 1, None) == 0
assert capi.lib.preAlps_b200_BlockJacobiCreate() == 0
arr = capi.operator_arrays()
rhs = capi.driver_rhs(arr["m"])
ref_rhs = np.concatenate([g["r%%d_rhs" %% r] for r in range(rank * per, (rank + 1) * per)])
sol, hist, info = capi.solve(rhs, t, tol, ortho=int(g["ortho"]))
ref_sol = np.concatenate([g["r%%d_sol" %% r] for r in range(rank * per, (rank + 1) * per)])
# ||b|| is summed over the subdomains in subdomain order on every process (pa_sum_over_subdomains): the scaled right-hand
# side is the golden run's, bit for bit, on any number of GPUs
out = {"rank": rank, "rhs_equal": bool(np.array_equal(rhs, ref_rhs)), "iter": info.iter, "ref_iter": int(g["iter"]),
       "hist_dev": float(np.max(np.abs(hist[:len(g["res_hist"])] - g["res_hist"][:len(hist)]) / g["res_hist"][:len(hist)])),
       "sol_dev": float(np.linalg.norm(sol - ref_sol) / np.linalg.norm(ref_sol)), "true": info.true_relres,
       "nhalo": len(arr["halo"]), "dep": arr["dep"].tolist()}
open(os.path.join(%(out)r, "result_%%d.json" %% rank), "w").write(json.dumps(out))
capi.lib.preAlps_OperatorFree(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", ["poisson7_n12_s8_t8_odir", "poisson7_n10_s8_t2_omin"])
def test_nccl_solve_matches_reference(world, case, tmp_path):
    if capi.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "run.py"
    script.write_text(SCRIPT % {"root": ROOT, "case": case, "out": str(tmp_path)})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    res = [json.load(open(str(tmp_path / ("result_%d.json" % r)))) for r in range(world)]
    for r in res:
        assert r["rhs_equal"]
        assert abs(r["iter"] - r["ref_iter"]) <= 1
        assert r["hist_dev"] < 1e-6
        assert r["sol_dev"] < 1e-7
        assert r["nhalo"] > 0 and len(r["dep"]) > 0


@pytest.mark.parametrize("world", [2, 8])
def test_nccl_solve_with_overlapped_halo_exchange(world, tmp_path):
    """PREALPS_SPMM_OVERLAP=1: halo exchange on a second stream next to the local part of the product, halo entries of
    the boundary rows added afterwards (same summation order): same iterations and history as the reference"""
    if capi.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "run.py"
    script.write_text(SCRIPT % {"root": ROOT, "case": "poisson7_n12_s8_t8_odir", "out": str(tmp_path)})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), str(script)],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, PREALPS_SPMM_OVERLAP="1"))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    res = [json.load(open(str(tmp_path / ("result_%d.json" % r)))) for r in range(world)]
    for r in res:
        assert abs(r["iter"] - r["ref_iter"]) <= 1
        assert r["hist_dev"] < 1e-6
        assert r["sol_dev"] < 1e-7
