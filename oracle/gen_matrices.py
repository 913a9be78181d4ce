"""oracle/gen_matrices.py -- TEST INFRASTRUCTURE ONLY.

Synthetic SPD operators of BASELINE.json's configs, written as MatrixMarket
files so that the unmodified reference (oracle/_ref) and the B200 library read
byte-identical input:

  poisson7(N)   7-point Laplacian on an N^3 grid, lexicographic, a_ii=6, a_ij=-1
  stencil27(N)  27-point stencil, a_ii=26, a_ij=-1 for the 26 neighbours

  elasticity3d(nx, ny, nz)  trilinear (Q1) hexahedral finite elements for 3D linear elasticity on a grid of
                nx*ny*nz nodes, 3 dof per node (up to 81 non-zeros per row), the face x = 0 clamped (those dof
                are eliminated), Young's modulus jumping between layers of elements ("var" like the reference's
                matrix/elasticity3d_12x10x10_var.mtx, which is not shipped) -- the operator shape of BASELINE config 4

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


def _hex_stiffness(E, nu):
    """24x24 stiffness of the unit-cube Q1 element, 2x2x2 Gauss points, dof order (node, component)."""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu)); mu = E / (2 * (1 + nu))
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] += 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    corners = np.array([[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1)], dtype=float)  # x fastest
    g = 0.5 + np.array([-1.0, 1.0]) / (2 * np.sqrt(3.0))
    K = np.zeros((24, 24))
    for gz in g:
        for gy in g:
            for gx in g:
                pt = np.array([gx, gy, gz])
                dN = np.zeros((8, 3))
                for a in range(8):
                    f = [(pt[d] if corners[a, d] == 1 else 1 - pt[d]) for d in range(3)]
                    s = [(1.0 if corners[a, d] == 1 else -1.0) for d in range(3)]
                    dN[a] = [s[0] * f[1] * f[2], f[0] * s[1] * f[2], f[0] * f[1] * s[2]]
                B = np.zeros((6, 24))
                for a in range(8):
                    dx, dy, dz = dN[a]
                    B[:, 3 * a:3 * a + 3] = [[dx, 0, 0], [0, dy, 0], [0, 0, dz], [dy, dx, 0], [0, dz, dy], [dz, 0, dx]]
                K += B.T @ D @ B / 8.0
    return K


def elasticity3d(nx, ny, nz, contrast=1e3, nu=0.3):
    node = np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)  # x fastest
    K1 = _hex_stiffness(1.0, nu)
    ez, ey, ex = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    ez, ey, ex = ez.ravel(), ey.ravel(), ex.ravel()
    E = np.where((ez // 2) % 2 == 0, 1.0, contrast)  # layers of two elements in z
    conn = np.stack([node[ez + dz, ey + dy, ex + dx] for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)], axis=1)
    dof = (3 * conn[:, :, None] + np.arange(3)[None, None, :]).reshape(-1, 24)
    rows = np.repeat(dof, 24, axis=1).ravel()
    cols = np.tile(dof, (1, 24)).ravel()
    vals = (E[:, None] * K1.ravel()[None, :]).ravel()
    n = 3 * nx * ny * nz
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    free = np.setdiff1d(np.arange(n), (3 * node[:, :, 0].ravel()[:, None] + np.arange(3)).ravel())
    A = A[free][:, free].tocsr()
    A = ((A + A.T) * 0.5).tocsr()  # exact symmetry of the stored values
    A.sum_duplicates(); A.sort_indices()
    return A


def build(gen, N):
    """operator of a golden case: N is the grid size, or (nx, ny, nz) for elasticity3d"""
    N = np.atleast_1d(np.asarray(N))
    return globals()[str(gen)](*(int(x) for x in N))


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
