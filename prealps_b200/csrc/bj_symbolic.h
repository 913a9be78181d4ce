// bj_symbolic.h -- host-side symbolic analysis for the block-Jacobi sparse Cholesky.
//
// Replaces the analysis half of MKL PARDISO phase 12 as the reference drives it
// (reference: utils/cplm_light/cplm_kernels.c:741-783 with iparm[1]=2, i.e. METIS
// nested dissection, called from src/preconditioners/block_jacobi.c:54).
// Integer-only work: fill-reducing ordering (METIS_NodeND), elimination tree,
// column counts, relaxed supernodes, supernodal row structures, tree levels.
// The numeric factorisation and the triangular solves run on the GPU
// (bj_factor.cu, bj_solve.cu) over the layout produced here.
#pragma once
#include <cstdint>
#include <vector>

namespace pcu {

struct Symbolic {
  int n = 0;                      // block order
  std::vector<int> perm;          // perm[new] = old (ND ordering composed with the etree postorder)
  std::vector<int> iperm;         // iperm[old] = new
  int nsuper = 0;
  std::vector<int> sn_col;        // nsuper+1: supernode s owns columns [sn_col[s], sn_col[s+1])
  std::vector<int64_t> sn_rowptr; // nsuper+1: offsets into sn_rows
  std::vector<int> sn_rows;       // row structure of each supernode, sorted; the first w entries are its own columns
  std::vector<int> sn_parent;     // supernodal elimination tree (-1 = root)
  std::vector<int> sn_level;      // 0 = leaves ... (a node's level is 1 + max level of its children)
  int nlevels = 0;
  std::vector<int> col2sn;        // n: supernode of each column
  int64_t nnzL_exact = 0;         // sum of exact column counts (no relaxation)
  double ordering_seconds = 0.0;  // time spent in METIS_NodeND
  int64_t nnzL_stored = 0;        // sum over supernodes of w*(w+1)/2 + (h-w)*w (dense trapezoids)
  double flops = 0;               // sum of squared column counts of the stored structure
};

struct SymbolicOptions {
  int leaf_cols = 48;       // merge a whole elimination subtree into one dense supernode if it has <= this many columns
  double relax_zero = 0.2;  // merge a last child into its parent if the fraction of explicit zeros stays below this
  int relax_small = 16;     // ... or if both are narrower than this
  int relax_big_cols = 256;      // merged supernodes wider than this are held to relax_big instead: a wide supernode streams
  double relax_big = 0.05;       // at full rate anyway, its explicit zeros are pure extra traffic (measured on B200, 8 blocks of
                                 // 64^3: stored bytes 19.19 -> 18.34 GB, apply 3.76 -> 3.69 ms; profiles/r02_candidates_ab.md)
  bool use_metis = true;    // false: natural ordering (tests)
};

// A is the upper triangle (diagonal included) of an SPD matrix in 0-based CSR, columns sorted,
// exactly what CPLM_MatCSRGetDiagBlock(..., SYMMETRIC) hands to PARDISO in the reference
// (reference: utils/cplm_v0/cplm_v0_matcsr.c:287-389).
// Returns 0 on success, <0 on a malformed matrix / METIS failure.
int analyze(int n, const int* rowPtr, const int* colInd, const SymbolicOptions& opt, Symbolic* out);

}  // namespace pcu
