"""oracle/gen_matrices.py -- TEST INFRASTRUCTURE ONLY.

Synthetic SPD operators of BASELINE.json's configs, written as MatrixMarket
files so that the unmodified reference (oracle/_ref) and the B200 library read
byte-identical input:

  poisson7(N)   7-point Laplacian on an N^3 grid, lexicographic, a_ii=6, a_ij=-1
  stencil27(N)  27-point stencil, a_ii=26, a_ij=-1 for the 26 neighbours

`write_mtx` emits "coordinate real symmetric" (lower triangle, 1-based), the form
CPLM_LoadMatrixMarket expands itself (/root/reference/utils/cplm_light/cplm_matcsr.c:96-243).
"""
import sys
import numpy as np
import scipy.sparse as sp


def _grid_stencil(N, offsets, diag):
    idx = np.arange(N ** 3, dtype=np.int64).reshape(N, N, N)  # idx[z, y, x], x fastest
    rows, cols = [np.arange(N ** 3, dtype=np.int64)], [np.arange(N ** 3, dtype=np.int64)]
    vals = [np.full(N ** 3, float(diag))]
    for dz, dy, dx in offsets:
        zs = slice(max(0, -dz), N - max(0, dz)); zd = slice(max(0, dz), N - max(0, -dz))
        ys = slice(max(0, -dy), N - max(0, dy)); yd = slice(max(0, dy), N - max(0, -dy))
        xs = slice(max(0, -dx), N - max(0, dx)); xd = slice(max(0, dx), N - max(0, -dx))
        r = idx[zs, ys, xs].ravel(); c = idx[zd, yd, xd].ravel()
        rows.append(r); cols.append(c); vals.append(np.full(r.size, -1.0))
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(N ** 3, N ** 3))
    A.sort_indices()
    return A


def poisson7(N):
    offs = [(0, 0, 1), (0, 0, -1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)]
    return _grid_stencil(N, offs, 6.0)


def stencil27(N):
    offs = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
            if (dz, dy, dx) != (0, 0, 0)]
    return _grid_stencil(N, offs, 26.0)


def write_mtx(path, A, symmetric=True):
    A = sp.coo_matrix(A)
    if symmetric:
        keep = A.row >= A.col
        r, c, v = A.row[keep], A.col[keep], A.data[keep]
    else:
        r, c, v = A.row, A.col, A.data
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real %s\n" % ("symmetric" if symmetric else "general"))
        f.write("%d %d %d\n" % (A.shape[0], A.shape[1], r.size))
        np.savetxt(f, np.column_stack([r + 1, c + 1, v]), fmt="%d %d %.17g")


if __name__ == "__main__":
    kind, N, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    write_mtx(out, {"poisson7": poisson7, "stencil27": stencil27}[kind](N))
