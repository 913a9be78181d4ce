"""N > 1 host logic on the CPU: the reference-style distributed build (rank 0 loads / partitions / ships
panels over MPI, every rank builds colPos, dep and the halo plan) run as real processes over the MPI shim,
with the device stage switched off.  Checked against the reference's golden vectors and a numpy halo plan."""
import os
import subprocess

import numpy as np
import pytest

import gen_matrices
from conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("halo")
    out = str(d / "halo_plan_dump")
    lib = os.path.join(ROOT, "prealps_b200", "lib")
    subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "mpishim"),
                           os.path.join(ROOT, "tests", "halo_plan_dump.c"), "-o", out, "-L" + lib, "-lprealps_b200",
                           "-lprealps_cuda", "-lmpishim", "-Wl,-rpath," + lib])
    return out


@pytest.mark.parametrize("name", ["poisson7_n8_s4_t4_odir", "poisson7_n12_s8_t8_odir", "poisson7_n9_s6_t3_odir",
                                  "stencil27_n8_s4_t2_odir"])
def test_distributed_build_and_halo_plan(exe, name, tmp_path):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    S = int(g["S"])
    A = gen_matrices.build(g["gen"], g["N"])
    mtx = str(tmp_path / "A.mtx")
    gen_matrices.write_mtx(mtx, A)
    env = dict(os.environ, MPISHIM_NP=str(S), PREALPS_B200_HOST_ONLY="1")
    subprocess.run([exe, mtx, str(tmp_path)], check=True, env=env, stdout=subprocess.DEVNULL, timeout=120)

    def rd(r, nm):
        return np.fromfile(str(tmp_path / ("r%d_%s.i32" % (r, nm))), dtype=np.int32)

    posB = g["posB"]
    halos = []
    for r in range(S):  # what every rank holds after the MPI distribution, bit-exact against the reference
        assert np.array_equal(rd(r, "rowPos"), g["r%d_rowPos" % r])
        assert np.array_equal(rd(r, "colPos"), g["r%d_colPos" % r])
        assert np.array_equal(rd(r, "dep"), g["r%d_dep" % r])
        assert np.array_equal(rd(r, "A_rowPtr"), g["r%d_A_rowPtr" % r])
        assert np.array_equal(rd(r, "A_colInd"), g["r%d_A_colInd" % r])
        cols = g["r%d_A_colInd" % r]
        halo = np.unique(cols[(cols < posB[r]) | (cols >= posB[r + 1])])
        assert np.array_equal(rd(r, "halo"), halo)
        halos.append(halo)
    owner = lambda c: np.searchsorted(posB, c, side="right") - 1
    for r in range(S):  # the plan: boundary rows only, symmetric between sender and receiver
        nbr, sp, si, rp = rd(r, "nbr"), rd(r, "send_ptr"), rd(r, "send_idx"), rd(r, "recv_ptr")
        need_from = {q: halos[r][owner(halos[r]) == q] for q in range(S) if q != r}
        wanted_by = {q: halos[q][owner(halos[q]) == r] for q in range(S) if q != r}
        exp_nbr = [q for q in range(S) if q != r and (len(need_from[q]) or len(wanted_by[q]))]
        assert nbr.tolist() == exp_nbr
        assert np.array_equal(nbr, g["r%d_dep" % r])  # structurally symmetric matrix: same as the reference's dep
        for k, q in enumerate(nbr):
            assert np.array_equal(si[sp[k]:sp[k + 1]] + posB[r], wanted_by[q])
            assert np.array_equal(halos[r][rp[k]:rp[k + 1]], need_from[q])
        # the reference would ship m rows to every neighbour; the plan ships the boundary only
        m = posB[r + 1] - posB[r]
        assert sp[-1] <= m * len(nbr)


def test_mpishim_collectives(tmp_path):
    src = r'''
#include <stdio.h>
#include <mpi.h>
int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int r, n; MPI_Comm_rank(MPI_COMM_WORLD, &r); MPI_Comm_size(MPI_COMM_WORLD, &n);
  double x[3] = {r + 1.0, 2.0 * r, 1.0}; MPI_Allreduce(MPI_IN_PLACE, x, 3, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  int v = r == 0 ? 42 : 0; MPI_Bcast(&v, 1, MPI_INT, 0, MPI_COMM_WORLD);
  int all[16]; MPI_Allgather(&r, 1, MPI_INT, all, 1, MPI_INT, MPI_COMM_WORLD);
  int ok = x[0] == n * (n + 1) / 2.0 && x[1] == (double)n * (n - 1) && x[2] == n && v == 42;
  for (int i = 0; i < n; ++i) ok = ok && all[i] == i;
  /* ring with non-blocking receives from any source */
  int tok = r, got = -1; MPI_Request rq; MPI_Status st;
  MPI_Irecv(&got, 1, MPI_INT, MPI_ANY_SOURCE, 5, MPI_COMM_WORLD, &rq);
  MPI_Send(&tok, 1, MPI_INT, (r + 1) % n, 5, MPI_COMM_WORLD);
  MPI_Wait(&rq, &st);
  int cnt; MPI_Get_count(&st, MPI_INT, &cnt);
  ok = ok && got == (r + n - 1) % n && st.MPI_SOURCE == got && cnt == 1;
  int allok; MPI_Allreduce(&ok, &allok, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
  if (r == 0) printf("%s\n", allok ? "OK" : "FAIL");
  MPI_Finalize();
  return 0;
}'''
    c = tmp_path / "t.c"
    c.write_text(src)
    exe = str(tmp_path / "t")
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "mpishim"), str(c), os.path.join(ROOT, "mpishim", "mpishim.c"),
                           "-o", exe, "-lpthread"])
    for n in (1, 2, 5, 8):
        out = subprocess.run([exe], env=dict(os.environ, MPISHIM_NP=str(n)), capture_output=True, text=True, timeout=60)
        assert out.stdout.strip() == "OK", (n, out.stdout, out.stderr)


def test_reference_arm_under_torchrun_world2(tmp_path):
    """bench.py --impl reference launched like the driver does for N = 2 (gloo is enough on the CPU):
    rank 0 alone runs and prints ONE JSON line, the other rank exits 0 without work."""
    import json
    import sys
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ecg_dump_ref")):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3", "--ref-n", "16"],
                         capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["unit"] == "iterations/s"
    assert j["cpu_baseline"]["kind"] == "reference" and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["value"] > 0


def test_gloo_world2_subdomain_split():
    """the uniform consecutive split of subdomains over processes that bench.py and the library agree on,
    exercised through a real 2-process gloo group"""
    import sys
    code = r'''
import os, torch, torch.distributed as dist
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
S = 8
per = S // w
lo, hi = r * per, (r + 1) * per
t = torch.tensor([lo, hi])
out = [torch.zeros(2, dtype=torch.long) for _ in range(w)]
dist.all_gather(out, t)
cover = sorted(sum([list(range(int(a), int(b))) for a, b in out], []))
assert cover == list(range(S)), cover
assert [int((p * S) // w) for p in range(w + 1)] == [0] + [int(b) for _, b in out]
mx = torch.tensor([float(r + 1)]); dist.all_reduce(mx, op=dist.ReduceOp.MAX); assert mx.item() == w
dist.destroy_process_group()
print("ok", r)
'''
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        path = f.name
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", path], capture_output=True, text=True,
                         timeout=300)
    os.unlink(path)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2
