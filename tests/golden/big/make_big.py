"""tests/golden/big/make_big.py -- pin BASELINE.json's full-size configurations to the UNMODIFIED reference.

Runs oracle/_ref/ecg_dump_ref (reference sources of /root/reference + shims, oracle/Makefile) as S ranks x 1 thread
over mpishim on the synthetic operator written as a MatrixMarket file -- the call sequence of
/root/reference/examples/test_ecg_prealps_op.c:158-223 -- and keeps what the parity tests and bench.py compare with:
iteration count, residual history, block-size history, ||b||, true residual, the reference's wall-clock times and
checksums of the METIS permutation / rowPos (the partition the GPU run must reproduce bit for bit).

    python tests/golden/big/make_big.py poisson7_n128_s8_t8 [...]      (needs /root/reference only through oracle/_ref)
"""
import hashlib
import json
import os
import platform
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_matrices  # noqa: E402

CASES = {
    # name: generator, N, S (= ranks), t, ortho (-o), bs_red (-r), tol
    "poisson7_n32_s8_t8": ("poisson7", 32, 8, 8, 0, 0, 1e-8),
    "poisson7_n64_s8_t8": ("poisson7", 64, 8, 8, 0, 0, 1e-8),
    "poisson7_n128_s8_t8": ("poisson7", 128, 8, 8, 0, 0, 1e-8),          # BASELINE configs[1]
    "poisson7_n64_s8_t8_fused": ("poisson7", 64, 8, 8, 2, 0, 1e-8),
    "stencil27_n48_s8_t8": ("stencil27", 48, 8, 8, 0, 0, 1e-8),
    "elasticity_n24_s16_t16_adapt": ("elasticity3d", 24, 16, 16, 0, 1, 1e-8),  # BASELINE configs[3] shape, reduced grid
    "elasticity_n32_s16_t16_adapt": ("elasticity3d", 32, 16, 16, 0, 1, 1e-8),
}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def run(name):
    gen, n, S, t, ortho, bs_red, tol = CASES[name]
    exe = os.path.join(ROOT, "oracle", "_ref", "ecg_dump_ref")
    A = getattr(gen_matrices, gen)(n) if gen != "elasticity3d" else gen_matrices.elasticity3d(n, n, n)
    with tempfile.TemporaryDirectory() as d:
        mtx = os.path.join(d, "A.mtx")
        gen_matrices.write_mtx(mtx, A)
        t0 = time.time()
        subprocess.run([exe, "-m", mtx, "-e", str(t), "-o", str(ortho), "-r", str(bs_red), "-t", repr(tol), "-d", d, "-q"],
                       check=True, env=dict(os.environ, MPISHIM_NP=str(S)))
        wall = time.time() - t0
        s = json.load(open(os.path.join(d, "summary.json")))
        perm = np.fromfile(os.path.join(d, "perm.i32"), dtype=np.int32)
        posB = np.fromfile(os.path.join(d, "posB.i32"), dtype=np.int32)
    s.pop("matrix")
    s.update({"case": name, "generator": gen, "n": n, "rows": int(A.shape[0]), "nnz": int(A.nnz),
              "perm_sha256": hashlib.sha256(perm.tobytes()).hexdigest(), "posB": posB.tolist(),
              "wall_s": wall, "cores": min(S, os.cpu_count() or 1), "ranks": S, "cpu": cpu_model(),
              "how": "oracle/_ref/ecg_dump_ref = unmodified reference sources, MKL -> OpenBLAS + plain-C pardiso shim, "
                     "METIS = CUDA toolkit libmetis_static.a, MPI -> mpishim (fork + shared memory)"})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".json")
    with open(out, "w") as f:
        json.dump(s, f, indent=1)
    print("%s: iter %d, res %.3e, true relres %.3e, t_solve %.1f s (op %.1f, prec %.1f), factor %.1f s, wall %.0f s"
          % (name, s["iter"], s["res"], s["true_relres"], s["t_solve"], s["t_op"], s["t_prec"], s["t_factor"], wall), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["poisson7_n32_s8_t8", "poisson7_n64_s8_t8"]):
        run(nm)
