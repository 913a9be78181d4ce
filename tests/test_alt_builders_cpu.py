"""preAlps_OperatorRHSBuild and preAlps_OperatorBuildNoPerm (ref: utils/operator.c:136-308) as real processes over the
MPI shim, device stage off, against golden vectors from the unmodified reference (tests/golden/make_golden_altbuild.py).
RHSBuild scales in line (a_ij / sqrt(r_i r_j), one divide -- not the two multiplies of CPLM_MatCSRSymRACScaling), leaves
the right-hand side unscaled, permutes it with the METIS permutation and scatters it."""
import os
import subprocess

import numpy as np
import pytest

import gen_matrices
from conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALT = os.path.join(GOLDEN, "alt")


def rhs_of(M):  # same as tests/golden/make_golden_altbuild.py
    return np.sin(0.37 * np.arange(M)) + 2.0


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("alt")
    out = str(d / "alt_build_dump")
    lib = os.path.join(ROOT, "prealps_b200", "lib")
    subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "mpishim"),
                           os.path.join(ROOT, "tests", "alt_build_dump.c"), "-o", out, "-L" + lib, "-lprealps_b200",
                           "-lprealps_cuda", "-lmpishim", "-Wl,-rpath," + lib])
    return out


@pytest.mark.parametrize("name", sorted(f[:-4] for f in os.listdir(ALT) if f.endswith(".npz")))
def test_rhs_build_and_build_no_perm(exe, name, tmp_path):
    g = np.load(os.path.join(ALT, name + ".npz"))
    S = int(g["S"])
    A = gen_matrices.build(g["gen"], g["N"])
    mtx = str(tmp_path / "A.mtx")
    gen_matrices.write_mtx(mtx, A)
    rhsf = str(tmp_path / "rhs.txt")
    with open(rhsf, "w") as f:
        f.write("%% rhs of %s\n%d 1\n" % (name, A.shape[0]))
        for x in rhs_of(A.shape[0]):
            f.write("%.17g\n" % x)
    env = dict(os.environ, MPISHIM_NP=str(S), PREALPS_B200_HOST_ONLY="1")
    subprocess.run([exe, "rhs", mtx, rhsf, str(tmp_path)], check=True, env=env, stdout=subprocess.DEVNULL, timeout=120)
    subprocess.run([exe, "noperm", str(tmp_path)], check=True, env=env, stdout=subprocess.DEVNULL, timeout=120)

    def rd(r, nm, dt=np.int32):
        return np.fromfile(str(tmp_path / ("r%d_%s.%s" % (r, nm, "i32" if dt == np.int32 else "f64"))), dtype=dt)

    for r in range(S):
        # RHSBuild: panel (pattern and in-line scaled values), maps and the permuted, scattered right-hand side: bit-exact
        for nm in ("A_rowPtr", "A_colInd", "rhs_rowPos", "rhs_colPos", "rhs_dep"):
            assert np.array_equal(rd(r, nm), g["r%d_%s" % (r, nm)]), nm
        assert np.array_equal(rd(r, "A_val", np.float64), g["r%d_A_val" % r])
        assert np.array_equal(rd(r, "rhs", np.float64), g["r%d_rhs" % r])
        # BuildNoPerm on those panels: same sizes and maps, columns still global
        for nm in ("noperm_sizes", "noperm_A_colInd", "noperm_rowPos", "noperm_colPos", "noperm_dep"):
            assert np.array_equal(rd(r, nm), g["r%d_%s" % (r, nm)]), nm
