"""tests/golden/make_golden.py -- regenerate the golden vectors from the UNMODIFIED reference.

Runs oracle/_ref/ecg_dump_ref (the reference sources of /root/reference compiled by
oracle/Makefile against the shims) on small synthetic operators and packs everything the
parity tests compare against into tests/golden/*.npz:
  scaled matrix, METIS permutation / posB, per-rank row panels, colPos, dep, diagonal blocks,
  the driver's right-hand side, the first block-Jacobi apply and SpMM, residual history,
  iteration count, solution, true residual.
Only this script needs /root/reference; the .npz files travel with the repo.

    python tests/golden/make_golden.py [case names ...]
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_matrices  # noqa: E402

CASES = [
    # name, generator, N, np (= subdomains), t, ortho (-o), tol
    ("poisson7_n8_s4_t4_odir", "poisson7", 8, 4, 4, 0, 1e-8),
    ("poisson7_n12_s8_t8_odir", "poisson7", 12, 8, 8, 0, 1e-8),
    ("poisson7_n10_s8_t2_omin", "poisson7", 10, 8, 2, 1, 1e-6),
    ("stencil27_n8_s4_t2_odir", "stencil27", 8, 4, 2, 0, 1e-8),
    ("poisson7_n9_s6_t3_odir", "poisson7", 9, 6, 3, 0, 1e-7),
    ("poisson7_n8_s4_t4_fused", "poisson7", 8, 4, 4, 2, 1e-8),
    ("poisson7_n12_s8_t8_fused", "poisson7", 12, 8, 8, 2, 1e-8),
    # ADAPT_BS (-r 1): the block size shrinks while the iteration goes on (bs_hist)
    ("poisson7_n12_s8_t8_odir_adapt", "poisson7", 12, 8, 8, 0, 1e-8, 1),
    ("elasticity3d_766_s8_t8_odir_adapt", "elasticity3d", (7, 6, 6), 8, 8, 0, 1e-8, 1),
    ("elasticity3d_655_s4_t4_odir_adapt", "elasticity3d", (6, 5, 5), 4, 4, 0, 1e-8, 1),
    ("elasticity3d_655_s4_t4_omin_adapt", "elasticity3d", (6, 5, 5), 4, 4, 1, 1e-8, 1),
    ("elasticity3d_766_s8_t8_fused_adapt", "elasticity3d", (7, 6, 6), 8, 8, 2, 1e-8, 1),
    # Orthomin + ADAPT_BS where dpstrf drops a direction (ecg.c:375-391): 64 unknowns, the enlarged Krylov space runs out
    ("poisson7_n4_s4_t4_omin_adapt_rankdrop", "poisson7", 4, 4, 4, 1, 1e-13, 1),
]


def load(d, name, dtype):
    return np.fromfile(os.path.join(d, name), dtype=dtype)


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "ecg_dump_ref")
    if not os.path.exists(exe):
        sys.exit("build the oracle first: make -C oracle")
    only = sys.argv[1:]
    for case in CASES:
        name, gen, N, S, t, ortho, tol = case[:7]
        bs_red = case[7] if len(case) > 7 else 0
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            mtx = os.path.join(d, "A.mtx")
            A = getattr(gen_matrices, gen)(*N) if isinstance(N, tuple) else getattr(gen_matrices, gen)(N)
            gen_matrices.write_mtx(mtx, A)
            env = dict(os.environ, MPISHIM_NP=str(S))
            subprocess.run([exe, "-m", mtx, "-e", str(t), "-o", str(ortho), "-r", str(bs_red), "-t", repr(tol), "-d", d],
                           check=True, env=env, stdout=subprocess.DEVNULL)
            out = {"gen": gen, "N": np.array(N), "S": S, "t": t, "ortho": ortho, "tol": tol, "bs_red": bs_red}
            summ = json.load(open(os.path.join(d, "summary.json")))
            for k in ("iter", "res", "normb", "true_relres", "M"):
                out[k] = summ[k]
            out["res_hist"] = np.array(summ["res_hist"])
            out["bs_hist"] = np.array(summ["bs_hist"], dtype=np.int32)
            out["perm"] = load(d, "perm.i32", np.int32)
            out["posB"] = load(d, "posB.i32", np.int32)
            out["S_rowPtr"] = load(d, "S_rowPtr.i32", np.int32)
            out["S_colInd"] = load(d, "S_colInd.i32", np.int32)
            out["S_val"] = load(d, "S_val.f64", np.float64)
            for r in range(S):
                for arr, dt in (("rowPos", np.int32), ("colPos", np.int32), ("dep", np.int32),
                                ("A_rowPtr", np.int32), ("A_colInd", np.int32), ("A_val", np.float64),
                                ("D_rowPtr", np.int32), ("D_colInd", np.int32), ("D_val", np.float64),
                                ("rhs", np.float64), ("sol", np.float64), ("P1", np.float64), ("AP1", np.float64)):
                    out["r%d_%s" % (r, arr)] = load(d, "r%d_%s.%s" % (r, arr, "i32" if dt == np.int32 else "f64"), dt)
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
            print(name, "iter", summ["iter"], "res", summ["res"], "true", summ["true_relres"])


if __name__ == "__main__":
    main()
