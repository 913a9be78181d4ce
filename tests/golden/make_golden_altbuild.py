"""tests/golden/make_golden_altbuild.py -- golden vectors for the two alternative builders of the reference API,
preAlps_OperatorRHSBuild (operator.c:136-268) and preAlps_OperatorBuildNoPerm (operator.c:271-308), from the UNMODIFIED
reference: tests/alt_build_dump.c is compiled against the reference's own headers and oracle/_ref/libprealps_ref.so
(make -C oracle first), run over the MPI shim, and its dumps packed into tests/golden/alt/*.npz.
Only this script needs /root/reference.

    python tests/golden/make_golden_altbuild.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_matrices  # noqa: E402

CASES = [("poisson7_n7_s4", "poisson7", 7, 4), ("stencil27_n6_s3", "stencil27", 6, 3), ("elasticity3d_544_s4", "elasticity3d", (5, 4, 4), 4)]


def rhs_of(M):
    """a right-hand side with a recognisable pattern, written like the files preAlps_doubleVector_load reads"""
    return np.sin(0.37 * np.arange(M)) + 2.0


def main():
    lib = os.path.join(ROOT, "oracle", "_ref", "libprealps_ref.so")
    if not os.path.exists(lib):
        sys.exit("build the oracle first: make -C oracle")
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "alt_build_dump_ref")
        inc = ["-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "mpishim")] + \
              ["-I" + os.path.join(REF, p) for p in ("utils/cplm_core", "utils/cplm_light", "utils/cplm_v0", "utils",
                                                      "src/preconditioners", "src/solvers")]
        subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-w", "-DAdd_", "-DMKLACTIVATE", "-DUSE_MKL"] + inc +
                              [os.path.join(ROOT, "tests", "alt_build_dump.c"), "-o", exe, lib,
                               "-Wl,-rpath," + os.path.dirname(lib), "-lm"])
        for name, gen, N, S in CASES:
            out = os.path.join(d, name)
            os.makedirs(out)
            A = gen_matrices.build(gen, N)
            mtx = os.path.join(out, "A.mtx")
            gen_matrices.write_mtx(mtx, A)
            rhs = rhs_of(A.shape[0])
            rhsf = os.path.join(out, "rhs.txt")
            with open(rhsf, "w") as f:
                f.write("%% rhs of %s\n%d 1\n" % (name, A.shape[0]))
                for x in rhs:
                    f.write("%.17g\n" % x)
            env = dict(os.environ, MPISHIM_NP=str(S))
            subprocess.run([exe, "rhs", mtx, rhsf, out], check=True, env=env, stdout=subprocess.DEVNULL)
            subprocess.run([exe, "noperm", out], check=True, env=env, stdout=subprocess.DEVNULL)
            pack = {"gen": gen, "N": np.array(N), "S": S}
            for r in range(S):
                for nm, dt in (("A_rowPtr", np.int32), ("A_colInd", np.int32), ("A_val", np.float64), ("rhs", np.float64),
                               ("rhs_rowPos", np.int32), ("rhs_colPos", np.int32), ("rhs_dep", np.int32),
                               ("noperm_sizes", np.int32), ("noperm_A_colInd", np.int32), ("noperm_rowPos", np.int32),
                               ("noperm_colPos", np.int32), ("noperm_dep", np.int32)):
                    ext = "i32" if dt == np.int32 else "f64"
                    pack["r%d_%s" % (r, nm)] = np.fromfile(os.path.join(out, "r%d_%s.%s" % (r, nm, ext)), dtype=dt)
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", "alt", name + ".npz"), **pack)
            print(name, "rows per rank", [len(pack["r%d_rhs" % r]) for r in range(S)])


if __name__ == "__main__":
    main()
