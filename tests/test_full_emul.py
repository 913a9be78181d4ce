"""The whole stack on the CPU emulation (tests/emul): the plain-C host layer (libprealps_b200) on top of every CUDA kernel of
libprealps_cuda compiled against a miniature CUDA runtime -- operator build, block-Jacobi factorisation, ECG iterations --
against a golden run of the reference and against the numpy restatement.  TEST INFRASTRUCTURE: logic and data flow only; the product libraries
have no CPU path and the -m gpu tests remain the parity tests proper."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import gen_matrices
import restate
from conftest import GOLDEN, ROOT

CASE = os.path.join(ROOT, "tests", "emul", "full_solve_case.py")
CANDIDATE_VARS = ("PREALPS_SPMM_BULK", "PREALPS_SPMM_OVERLAP")


def solve(spec, **switches):
    env = {k: v for k, v in os.environ.items() if k not in CANDIDATE_VARS and k != "PREALPS_B200_LIBDIR"}
    env.update(switches)
    out = subprocess.run([sys.executable, CASE, spec], env=env, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    r["hist"] = np.array([float.fromhex(x) for x in r["hist"]])
    return r


def test_whole_solve_and_staging_variants():
    """t = 8 so that the bulk-staging SpMM kernel engages; it keeps the operation order of the kernel it replaced, so the
    residual history of the whole solve is the same bits with PREALPS_SPMM_BULK=0"""
    spec = "poisson7:8:8:8:1e-6"
    base = solve(spec)
    P = restate.Partitioned(gen_matrices.poisson7(8).tocsr(), 8)
    ref = restate.ecg_solve(P, 8, 1e-6)
    assert abs(base["iter"] - ref["iter"]) <= 1
    n = min(len(base["hist"]), len(ref["res_hist"]))
    assert np.allclose(base["hist"][:n], ref["res_hist"][:n], rtol=1e-6)
    a = solve(spec, PREALPS_SPMM_BULK="0")
    assert a["iter"] == base["iter"] and np.array_equal(a["hist"], base["hist"]) and a["sol_sum"] == base["sol_sum"]


def test_orthomin_adapt_bs_rank_drop():
    """Orthomin + ADAPT_BS when dpstrf drops a direction (ecg.c:375-391; round 1 aborted there): the golden from the
    reference, where the drop is the last thing that happens, and a solve that continues on 3 of 4 directions, checked
    inside the case script against the numpy restatement and a dense solve (tests/gpu_util.py: check_rank_drop_solve)"""
    name = "poisson7_n4_s4_t4_omin_adapt_rankdrop"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    r = solve(name)
    assert r["iter"] == int(g["iter"]) and r["bs_hist"] == g["bs_hist"].tolist() and r["bs_hist"][-1] == 3
    n = len(r["hist"])
    assert np.allclose(r["hist"][:n - 1], g["res_hist"][:n - 1], rtol=1e-7)
    r = solve("rankdrop")
    assert r["bs_hist"][0] == 4 and r["bs_hist"][1] == 3 and r["iter"] > 8


def test_unchanged_reference_driver_on_the_emulated_stack(tmp_path):
    """examples/test_ecg_prealps_op.c of the reference, compiled unchanged (prealps_b200/bin), one process per subdomain over
    the MPI shim, with the emulated libraries in front of its RUNPATH: the multi-process path (halo plan, pack kernel,
    boundary rows through host MPI, all-reduces) against the reference's golden run"""
    exe = os.path.join(ROOT, "prealps_b200", "bin", "test_ecg_prealps_op")
    if not os.path.exists(exe):
        pytest.skip("driver binary not built (reference tree was absent at build time)")
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_bj_emul
    libdir = build_bj_emul.build_full()
    name = "poisson7_n8_s4_t4_odir"
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    mtx = str(tmp_path / "A.mtx")
    gen_matrices.write_mtx(mtx, gen_matrices.build(g["gen"], g["N"]))
    env = {k: v for k, v in os.environ.items() if k not in CANDIDATE_VARS}
    env.update(LD_LIBRARY_PATH=libdir, MPISHIM_NP=str(int(g["S"])))
    out = subprocess.run([exe, "-e", str(int(g["t"])), "-m", mtx, "-o", str(int(g["ortho"])), "-r", "0", "-t", repr(float(g["tol"]))],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    it = int([l for l in out.stdout.splitlines() if "iter:" in l][0].split(":")[1])
    res = float([l for l in out.stdout.splitlines() if "res :" in l][0].split(":")[1])
    assert it == int(g["iter"])
    assert res == pytest.approx(float(g["res"]), rel=1e-5)


def test_bench_entry_points_on_the_emulated_stack():
    """what bench.py calls underneath (it cannot run here itself: it needs torch.cuda for the clocks and the barriers)"""
    env = {k: v for k, v in os.environ.items() if k not in CANDIDATE_VARS and k != "PREALPS_B200_LIBDIR"}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "bench_entry_case.py")], env=env, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["M"] == 512 and r["launches"] > 0 and all(x > 0 for x in r["kernel_ms"])
    assert r["true_relres"] < 1e-7 and 5 < r["iter"] < 40
    assert r["spmm_bytes"] == 3200 * 12 + 513 * 4 + 2 * 512 * 8 * 8 and r["bj_bytes"] > 0  # SURVEY.md 8(d); 7-point 8^3: nnz = 7*512 - 6*64
