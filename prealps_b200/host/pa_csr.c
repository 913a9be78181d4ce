/*
 * pa_csr.c -- host-side integer pipeline of the operator: MatrixMarket loader,
 * symmetric scaling, METIS k-way partition, permutation, row panels, column-block
 * positions, neighbour list, diagonal blocks and the new halo maps.
 *
 * Everything here must be BIT-EXACT against the reference given the same parts[]
 * (SURVEY.md 8c); each function names the reference routine it reproduces.  The code
 * is written from the behaviour, not from the reference sources: counting sorts instead
 * of qsort/quicksort (same result for duplicate-free input), O(M) instead of O(S*M)
 * permutation build.
 */
#include "pa_internal.h"

#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

/* METIS 5 from the CUDA toolkit's libmetis_static.a: idx_t = int64, real_t = float */
int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt,
                        int64_t* vsize, int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec,
                        int64_t* options, int64_t* edgecut, int64_t* part);

/* ------------------------------------------------------------------ small API of cplm_types.h */
void CPLM_FAbort(const char* fun, const char* format, ...) {
  va_list ap;
  va_start(ap, format);
  fprintf(stderr, "[ABORTING] in %s: ", fun);
  vfprintf(stderr, format, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  fflush(stderr);
  MPI_Abort(MPI_COMM_WORLD, 1);
  exit(1);
}

int CPLM_MatDenseSetInfo(CPLM_Mat_Dense_t* A, int M, int N, int m, int n, CPLM_storage_type_t storage) {
  A->info.M = M; A->info.N = N; A->info.m = m; A->info.n = n;
  A->info.lda = (storage == COL_MAJOR) ? m : n;
  A->info.nval = m * n;
  A->info.stor_type = storage;
  return 0;
}

void CPLM_MatCSRFree(CPLM_Mat_CSR_t* A) {
  if (!A) return;
  free(A->rowPtr); free(A->colInd); free(A->val);
  A->rowPtr = NULL; A->colInd = NULL; A->val = NULL;
  A->info.m = A->info.lnnz = 0;
}

void* pa_xmalloc(size_t n) {
  void* p = malloc(n ? n : 1);
  if (!p) CPLM_Abort("out of host memory (%zu bytes)", n);
  return p;
}
void* pa_xcalloc(size_t n, size_t s) {
  void* p = calloc(n ? n : 1, s ? s : 1);
  if (!p) CPLM_Abort("out of host memory (%zu bytes)", n * s);
  return p;
}

/* sort (col,val) pairs of one row by column: insertion sort for short rows, heap-free merge otherwise */
static void sort_row(int* c, double* v, int n) {
  if (n < 2) return;
  int sorted = 1;
  for (int i = 1; i < n; ++i) if (c[i - 1] > c[i]) { sorted = 0; break; }
  if (sorted) return;
  if (n <= 64) {
    for (int i = 1; i < n; ++i) {
      int ci = c[i]; double vi = v ? v[i] : 0.0; int j = i - 1;
      while (j >= 0 && c[j] > ci) { c[j + 1] = c[j]; if (v) v[j + 1] = v[j]; --j; }
      c[j + 1] = ci; if (v) v[j + 1] = vi;
    }
    return;
  }
  /* bottom-up merge sort (stable) */
  int* tc = (int*)pa_xmalloc(sizeof(int) * (size_t)n);
  double* tv = v ? (double*)pa_xmalloc(sizeof(double) * (size_t)n) : NULL;
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int a = lo, b = mid, k = lo;
      while (a < mid && b < hi) {
        if (c[b] < c[a]) { tc[k] = c[b]; if (v) tv[k] = v[b]; ++b; } else { tc[k] = c[a]; if (v) tv[k] = v[a]; ++a; }
        ++k;
      }
      while (a < mid) { tc[k] = c[a]; if (v) tv[k] = v[a]; ++a; ++k; }
      while (b < hi) { tc[k] = c[b]; if (v) tv[k] = v[b]; ++b; ++k; }
    }
    memcpy(c, tc, sizeof(int) * (size_t)n);
    if (v) memcpy(v, tv, sizeof(double) * (size_t)n);
  }
  free(tc); free(tv);
}

/* COO (0-based) -> CSR with rows sorted by column */
static void coo_to_csr(int M, int64_t nnz, const int* ri, const int* ci, const double* vv, CPLM_Mat_CSR_t* A) {
  A->rowPtr = (int*)pa_xcalloc((size_t)M + 1, sizeof(int));
  A->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)nnz);
  A->val = (double*)pa_xmalloc(sizeof(double) * (size_t)nnz);
  for (int64_t k = 0; k < nnz; ++k) A->rowPtr[ri[k] + 1]++;
  for (int i = 0; i < M; ++i) A->rowPtr[i + 1] += A->rowPtr[i];
  int* fill = (int*)pa_xmalloc(sizeof(int) * (size_t)(M > 0 ? M : 1));
  memcpy(fill, A->rowPtr, sizeof(int) * (size_t)M);
  for (int64_t k = 0; k < nnz; ++k) { int p = fill[ri[k]]++; A->colInd[p] = ci[k]; A->val[p] = vv[k]; }
  free(fill);
  for (int i = 0; i < M; ++i) sort_row(A->colInd + A->rowPtr[i], A->val + A->rowPtr[i], A->rowPtr[i + 1] - A->rowPtr[i]);
}

/* ------------------------------------------------------------------ loader
 * ref: CPLM_LoadMatrixMarket, utils/cplm_light/cplm_matcsr.c:96-243 (+ CPLM_MatCSRUnsymStruct,
 * cplm_matcsr_core.c:243-315; CPLM_MatCSRChkDiag, cplm_matcsr.c:257): "matrix coordinate real
 * general|symmetric", comment lines, size line, triples; the file is 0-based iff the first triple
 * has a zero index; a symmetric file is expanded to full storage (info.structure stays SYMMETRIC);
 * a missing diagonal entry aborts. */
int pa_load_mtx(const char* filename, CPLM_Mat_CSR_t* A, int verbose) {
  if (verbose) printf("Load of %s ...\n", filename);
  FILE* f = fopen(filename, "r");
  if (!f) CPLM_Abort("Impossible to open the file %s", filename);
  char line[1100], t0[64] = "", t1[64] = "", t2[64] = "", t3[64] = "", t4[64] = "";
  if (!fgets(line, sizeof line, f)) CPLM_Abort("Empty file %s", filename);
  sscanf(line, "%63s %63s %63s %63s %63s", t0, t1, t2, t3, t4);
  if (strcasecmp(t0, "%%MatrixMarket") || strcasecmp(t1, "matrix") || strcasecmp(t2, "coordinate") ||
      strcasecmp(t3, "real") || (strcasecmp(t4, "general") && strcasecmp(t4, "symmetric"))) {
    fclose(f);
    CPLM_Abort("Only sparse real < symmetric | general > matrix are currently supported.\nHere is %s", t4);
  }
  const int sym = !strcasecmp(t4, "symmetric");
  do {
    if (!fgets(line, sizeof line, f)) break;
  } while (line[0] == '%');
  int M = 0, N = 0, nnz = 0;
  sscanf(line, "%d%d%d", &M, &N, &nnz);
  if (M < 1 || N < 1 || nnz < 1 || (long long)nnz > (long long)M * N) {
    fprintf(stderr, "[LoadMatrixMarket] Error: Invalid matrix dimensions.\n");
    fclose(f);
    MPI_Abort(MPI_COMM_WORLD, 1);
  }
  int* ri = (int*)pa_xmalloc(sizeof(int) * (size_t)nnz * (sym ? 2 : 1));
  int* ci = (int*)pa_xmalloc(sizeof(int) * (size_t)nnz * (sym ? 2 : 1));
  double* vv = (double*)pa_xmalloc(sizeof(double) * (size_t)nnz * (sym ? 2 : 1));
  int base0 = 0;
  for (int k = 0; k < nnz; ++k) {
    if (fscanf(f, "%d%d%lf", &ri[k], &ci[k], &vv[k]) != 3) CPLM_Abort("%s: truncated at entry %d", filename, k);
    if (k == 0 && (ri[0] == 0 || ci[0] == 0)) { printf("0-based detected\n"); base0 = 1; }
    if (!base0) { ri[k]--; ci[k]--; }
    if (ri[k] < 0 || ri[k] >= M || ci[k] < 0 || ci[k] >= N) CPLM_Abort("%s: entry %d out of range", filename, k);
  }
  fclose(f);
  int64_t tot = nnz;
  if (sym) {
    if (M != N) { fprintf(stderr, "matrix is not square\n"); exit(1); }
    for (int k = 0; k < nnz; ++k)
      if (ri[k] != ci[k]) { ri[tot] = ci[k]; ci[tot] = ri[k]; vv[tot] = vv[k]; ++tot; }
  }
  memset(&A->info, 0, sizeof A->info);
  coo_to_csr(M, tot, ri, ci, vv, A);
  free(ri); free(ci); free(vv);
  A->info.M = A->info.m = M;
  A->info.N = A->info.n = N;
  A->info.nnz = A->info.lnnz = (int)tot;
  A->info.blockSize = 1;
  A->info.format = FORMAT_CSR;
  A->info.structure = sym ? SYMMETRIC : UNSYMMETRIC;
  if (pa_check_diag(A)) CPLM_Abort("Diagonal is not set correctly");
  return 0;
}

int pa_check_diag(const CPLM_Mat_CSR_t* A) {
  for (int i = 0; i < A->info.m; ++i) {
    int found = 0;
    for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) if (A->colInd[p] == i) { found = 1; break; }
    if (!found) return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------ scaling
 * ref: CPLM_MatCSRSymRACScaling, cplm_matcsr.c:1461-1554: r_i = sqrt(1/max_j|a_ij|),
 * a_ij <- (r_i * a_ij) * r_j evaluated in that order; rcmin == 0 aborts. */
int pa_sym_scale(CPLM_Mat_CSR_t* A) {
  const int m = A->info.m;
  double* R = (double*)pa_xmalloc(sizeof(double) * (size_t)m);
  for (int i = 0; i < m; ++i) {
    double r = 0.0;
    for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) { double a = fabs(A->val[p]); if (a > r) r = a; }
    R[i] = r;
  }
  double rcmin = R[0];
  for (int i = 1; i < m; ++i) if (R[i] < rcmin) rcmin = R[i];
  if (rcmin == 0.) { free(R); CPLM_Abort("Impossible to scale the matrix, rcmin=0"); }
  for (int i = 0; i < m; ++i) R[i] = sqrt(1.0 / R[i]);
  for (int i = 0; i < m; ++i)
    for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) A->val[p] = R[i] * A->val[p] * R[A->colInd[p]];
  free(R);
  return 0;
}

/* ------------------------------------------------------------------ partition
 * ref: CPLM_metisKwayOrdering -> CPLM_MatCSRDelDiag / CPLM_MatCSRSymStruct -> callKway
 * (cplm_v0_matcsr.c:114-167, cplm_matcsr_core.c:120-233,325-375,394-457):
 * adjacency = pattern of A (+A^T when unsymmetric) without the diagonal, rows sorted;
 * METIS_PartGraphKway(nvtxs, ncon=1, xadj, adjncy, NULL.., nparts, NULL, NULL, options=NULL). */
int pa_kway_parts(const CPLM_Mat_CSR_t* A, int S, int* parts) {
  const int M = A->info.m;
  if (S == 1) { for (int i = 0; i < M; ++i) parts[i] = 0; return 0; }
  int64_t* xadj = (int64_t*)pa_xcalloc((size_t)M + 1, sizeof(int64_t));
  int64_t* adj = NULL;
  if (A->info.structure == SYMMETRIC) {
    for (int i = 0; i < M; ++i) {
      int has = 0;
      for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) if (A->colInd[p] == i) has = 1;
      if (!has) { fprintf(stderr, "Error, no diagonal value on row %d\n", i); free(xadj); return 1; }
      xadj[i + 1] = xadj[i] + (A->rowPtr[i + 1] - A->rowPtr[i] - 1);
    }
    adj = (int64_t*)pa_xmalloc(sizeof(int64_t) * (size_t)xadj[M]);
    int64_t q = 0;
    for (int i = 0; i < M; ++i)
      for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) if (A->colInd[p] != i) adj[q++] = A->colInd[p];
  } else {
    /* union of the pattern and its transpose, diagonal removed, each row sorted and unique */
    int* cnt = (int*)pa_xcalloc((size_t)M + 1, sizeof(int));
    for (int i = 0; i < M; ++i)
      for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) {
        int j = A->colInd[p];
        if (j != i) { cnt[i + 1]++; cnt[j + 1]++; }
      }
    for (int i = 0; i < M; ++i) cnt[i + 1] += cnt[i];
    int* tmp = (int*)pa_xmalloc(sizeof(int) * (size_t)cnt[M]);
    int* fill = (int*)pa_xmalloc(sizeof(int) * (size_t)M);
    memcpy(fill, cnt, sizeof(int) * (size_t)M);
    for (int i = 0; i < M; ++i)
      for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) {
        int j = A->colInd[p];
        if (j != i) { tmp[fill[i]++] = j; tmp[fill[j]++] = i; }
      }
    adj = (int64_t*)pa_xmalloc(sizeof(int64_t) * (size_t)(cnt[M] > 0 ? cnt[M] : 1));
    int64_t q = 0;
    for (int i = 0; i < M; ++i) {
      int n = cnt[i + 1] - cnt[i];
      sort_row(tmp + cnt[i], NULL, n);
      for (int k = 0; k < n; ++k)
        if (k == 0 || tmp[cnt[i] + k] != tmp[cnt[i] + k - 1]) adj[q++] = tmp[cnt[i] + k];
      xadj[i + 1] = q;
    }
    free(cnt); free(tmp); free(fill);
  }
  int64_t nv = M, ncon = 1, np = S, objval = 0;
  int64_t* p64 = (int64_t*)pa_xmalloc(sizeof(int64_t) * (size_t)M);
  int rc = METIS_PartGraphKway(&nv, &ncon, xadj, adj, NULL, NULL, NULL, &np, NULL, NULL, NULL, &objval, p64);
  free(xadj); free(adj);
  if (rc != 1) { fprintf(stderr, "METIS_PartGraphKway failed (%d)\n", rc); free(p64); exit(1); }
  for (int i = 0; i < M; ++i) parts[i] = (int)p64[i];
  free(p64);
  return 0;
}

/* ref: CPLM_getBlockPosition + CPLM_getIntPermArray, cplm_v0_metis_utils.c:197-222,22-43:
 * posB = exclusive prefix sums of the part sizes; perm[new] = old, parts in id order and the
 * original order kept inside a part. */
void pa_parts_to_perm(int M, const int* parts, int S, int* posB, int* perm) {
  if (S < 1) return;
  for (int s = 0; s <= S; ++s) posB[s] = 0;
  for (int i = 0; i < M; ++i) posB[parts[i] + 1]++;
  for (int s = 0; s < S; ++s) posB[s + 1] += posB[s];
  int* fill = (int*)pa_xmalloc(sizeof(int) * (size_t)S);
  memcpy(fill, posB, sizeof(int) * (size_t)S);
  for (int i = 0; i < M; ++i) perm[fill[parts[i]]++] = i;
  free(fill);
}

/* ref: CPLM_MatCSRPermute(A, B, perm, perm, PERMUTE), cplm_v0_matcsr.c:941-1022:
 * B = P A P^T, row new <- row perm[new], columns through the inverse permutation, rows re-sorted. */
int pa_permute_sym(const CPLM_Mat_CSR_t* A, const int* perm, CPLM_Mat_CSR_t* B) {
  const int m = A->info.m;
  int* iperm = (int*)pa_xmalloc(sizeof(int) * (size_t)m);
  for (int i = 0; i < m; ++i) iperm[perm[i]] = i;
  B->info = A->info;
  B->rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)m + 1));
  B->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)A->info.lnnz);
  B->val = (double*)pa_xmalloc(sizeof(double) * (size_t)A->info.lnnz);
  B->rowPtr[0] = 0;
  for (int i = 0; i < m; ++i) {
    const int o = perm[i], n = A->rowPtr[o + 1] - A->rowPtr[o];
    int* bc = B->colInd + B->rowPtr[i];
    double* bv = B->val + B->rowPtr[i];
    for (int k = 0; k < n; ++k) { bc[k] = iperm[A->colInd[A->rowPtr[o] + k]]; bv[k] = A->val[A->rowPtr[o] + k]; }
    sort_row(bc, bv, n);
    B->rowPtr[i + 1] = B->rowPtr[i] + n;
  }
  free(iperm);
  return 0;
}

/* rows [r0, r1) of P A P^T without forming the rest of it: what pa_permute_sym + pa_row_panel give, entry for entry (a
 * process that owns a few subdomains of a 16.8 M-row operator permutes 1/8 of it) */
int pa_permute_panel(const CPLM_Mat_CSR_t* A, const int* perm, int r0, int r1, CPLM_Mat_CSR_t* B) {
  const int m = A->info.m, lm = r1 - r0;
  int* iperm = (int*)pa_xmalloc(sizeof(int) * (size_t)m);
  for (int i = 0; i < m; ++i) iperm[perm[i]] = i;
  long long lnnz = 0;
  for (int i = r0; i < r1; ++i) lnnz += A->rowPtr[perm[i] + 1] - A->rowPtr[perm[i]];
  B->info = A->info;
  B->info.M = A->info.m;
  B->info.nnz = A->info.lnnz;
  B->info.m = lm;
  B->info.lnnz = (int)lnnz;
  B->rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)lm + 1));
  B->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)(lnnz > 0 ? lnnz : 1));
  B->val = (double*)pa_xmalloc(sizeof(double) * (size_t)(lnnz > 0 ? lnnz : 1));
  B->rowPtr[0] = 0;
  for (int i = 0; i < lm; ++i) {
    const int o = perm[r0 + i], n = A->rowPtr[o + 1] - A->rowPtr[o];
    int* bc = B->colInd + B->rowPtr[i];
    double* bv = B->val + B->rowPtr[i];
    for (int k = 0; k < n; ++k) { bc[k] = iperm[A->colInd[A->rowPtr[o] + k]]; bv[k] = A->val[A->rowPtr[o] + k]; }
    sort_row(bc, bv, n);
    B->rowPtr[i + 1] = B->rowPtr[i] + n;
  }
  free(iperm);
  return 0;
}

/* ref: CPLM_MatCSRGetRowPanel, cplm_v0_matcsr.c:655-720: rows [r0, r1) with global columns,
 * rowPtr rebased; info.M <- parent's m, info.nnz <- parent's lnnz. */
int pa_row_panel(const CPLM_Mat_CSR_t* A, int r0, int r1, CPLM_Mat_CSR_t* B) {
  const int lm = r1 - r0, off = A->rowPtr[r0], lnnz = A->rowPtr[r1] - off;
  B->info = A->info;
  B->info.M = A->info.m;
  B->info.nnz = A->info.lnnz;
  B->info.m = lm;
  B->info.lnnz = lnnz;
  B->rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)lm + 1));
  B->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)lnnz);
  B->val = (double*)pa_xmalloc(sizeof(double) * (size_t)lnnz);
  for (int i = 0; i <= lm; ++i) B->rowPtr[i] = A->rowPtr[r0 + i] - off;
  memcpy(B->colInd, A->colInd + off, sizeof(int) * (size_t)lnnz);
  memcpy(B->val, A->val + off, sizeof(double) * (size_t)lnnz);
  return 0;
}

/* ref: CPLM_MatCSRGetColBlockPos, cplm_v0_matcsr.c:175-227: colPos[i*S + j] = first index in
 * colInd of row i whose column is >= rowPos[j] (j >= 1), colPos[i*S] = rowPtr[i], colPos[m*S] = lnnz. */
int pa_col_block_pos(const CPLM_Mat_CSR_t* A, const int* rowPos, int S, int** colPos_out, int* n_out) {
  const int m = A->info.m;
  int* cp = (int*)pa_xmalloc(sizeof(int) * ((size_t)m * S + 1));
  cp[0] = 0;
  for (int i = 0; i < m; ++i) {
    int blk = 0;
    for (int p = A->rowPtr[i]; p < A->rowPtr[i + 1]; ++p) {
      const int c = A->colInd[p];
      while (c >= rowPos[blk + 1]) { ++blk; cp[(size_t)i * S + blk] = p; }
    }
    for (int k = blk + 1; k <= S; ++k) cp[(size_t)i * S + k] = A->rowPtr[i + 1];
  }
  *colPos_out = cp;
  *n_out = m * S + 1;
  return 0;
}

/* ref: CPLM_MatCSRGetCommDep, cplm_v0_matcsr.c:234-273: column blocks j (outside [lo,hi)) holding at
 * least one entry, ascending.  The reference aborts when the list is empty; the caller decides. */
int pa_comm_dep(const int* colPos, int m, int S, int lo, int hi, int** dep_out, int* ndep) {
  long long* cnt = (long long*)pa_xcalloc((size_t)S, sizeof(long long));
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < S; ++j) cnt[j] += colPos[(size_t)i * S + j + 1] - colPos[(size_t)i * S + j];
  int* dep = (int*)pa_xmalloc(sizeof(int) * (size_t)S);
  int n = 0;
  for (int j = 0; j < S; ++j) if (cnt[j] && (j < lo || j >= hi)) dep[n++] = j;
  free(cnt);
  *dep_out = dep;
  *ndep = n;
  return 0;
}

/* ref: CPLM_MatCSRGetDiagBlock(..., SYMMETRIC) + CPLM_MatCSRGetDiagIndOfPanel,
 * cplm_v0_matcsr.c:287-389,510-569: for subdomain blk whose rows are panel rows [pr0, pr1), keep per
 * row the entries from the diagonal to the end of its own column block, columns shifted to be
 * block-local => upper triangle including the diagonal. */
int pa_diag_block(const CPLM_Mat_CSR_t* A, const int* rowPos, const int* colPos, int S, int blk, int pr0, int pr1,
                  CPLM_Mat_CSR_t* D) {
  const int lm = pr1 - pr0, c0 = rowPos[blk];
  int* dpos = (int*)pa_xmalloc(sizeof(int) * (size_t)(lm > 0 ? lm : 1));
  int sum = 0;
  for (int i = 0; i < lm; ++i) {
    const size_t r = (size_t)(pr0 + i);
    const int b = colPos[r * S + blk], e = colPos[r * S + blk + 1];
    int d = -1;
    for (int p = b; p < e; ++p) if (A->colInd[p] - c0 == i) { d = p; break; }
    if (d < 0) { free(dpos); CPLM_Abort("row %d of diagonal block %d has no diagonal entry", i, blk); }
    dpos[i] = d;
    sum += e - d;
  }
  memset(&D->info, 0, sizeof D->info);
  D->info.M = lm; D->info.N = A->info.n; D->info.nnz = A->info.lnnz;
  D->info.m = lm; D->info.n = rowPos[blk + 1] - rowPos[blk]; D->info.lnnz = sum;
  D->info.blockSize = 1; D->info.format = FORMAT_CSR; D->info.structure = UNSYMMETRIC;
  D->rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)lm + 1));
  D->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)(sum > 0 ? sum : 1));
  D->val = (double*)pa_xmalloc(sizeof(double) * (size_t)(sum > 0 ? sum : 1));
  D->rowPtr[0] = 0;
  int q = 0;
  for (int i = 0; i < lm; ++i) {
    const int e = colPos[(size_t)(pr0 + i) * S + blk + 1];
    for (int p = dpos[i]; p < e; ++p) { D->colInd[q] = A->colInd[p] - c0; D->val[q] = A->val[p]; ++q; }
    D->rowPtr[i + 1] = q;
  }
  free(dpos);
  return 0;
}

/* New (no reference counterpart: the reference ships whole blocks, cplm_matdense.c:90-109):
 * halo = sorted unique global columns outside [g0, g1); colLoc = panel columns renumbered to
 * [0, m) for own rows and m + position-in-halo otherwise. */
int pa_halo_map(const CPLM_Mat_CSR_t* A, int g0, int g1, int** halo_out, int* nhalo_out, int** colLoc_out) {
  const int lnnz = A->info.lnnz, m = A->info.m;
  int* tmp = (int*)pa_xmalloc(sizeof(int) * (size_t)(lnnz > 0 ? lnnz : 1));
  int n = 0;
  for (int p = 0; p < lnnz; ++p) { const int c = A->colInd[p]; if (c < g0 || c >= g1) tmp[n++] = c; }
  sort_row(tmp, NULL, n);
  int nh = 0;
  for (int k = 0; k < n; ++k) if (k == 0 || tmp[k] != tmp[k - 1]) tmp[nh++] = tmp[k];
  int* halo = (int*)pa_xmalloc(sizeof(int) * (size_t)(nh > 0 ? nh : 1));
  memcpy(halo, tmp, sizeof(int) * (size_t)nh);
  free(tmp);
  int* cl = (int*)pa_xmalloc(sizeof(int) * (size_t)(lnnz > 0 ? lnnz : 1));
  for (int p = 0; p < lnnz; ++p) {
    const int c = A->colInd[p];
    if (c >= g0 && c < g1) cl[p] = c - g0;
    else {
      int lo = 0, hi = nh;
      while (lo < hi) { int mid = (lo + hi) / 2; if (halo[mid] < c) lo = mid + 1; else hi = mid; }
      cl[p] = m + lo;
    }
  }
  *halo_out = halo; *nhalo_out = nh; *colLoc_out = cl;
  return 0;
}

/* synthetic operators of SURVEY.md 8(d): N^3 grid, lexicographic (x fastest), Dirichlet by truncation.
 * kind 0: a_ii = 6, -1 for the 6 face neighbours; kind 1: a_ii = 26, -1 for the 26 neighbours. */
/* kind 2: trilinear (Q1) hexahedral elements for 3D linear elasticity on N^3 nodes, 3 dof per node (up to 81
 * non-zeros per row), the face x = 0 clamped (those dof eliminated), Young's modulus 1 or 1e3 in layers of two
 * elements along z, nu = 0.3 -- the operator of BASELINE config 4, same definition as
 * oracle/gen_matrices.py: elasticity3d (tests/test_host_integer.py compares the two). */
static void hex_stiffness(double nu, double K[24][24]) {
  const double lam = nu / ((1 + nu) * (1 - 2 * nu)), mu = 1.0 / (2 * (1 + nu));
  double D[6][6] = {{0}};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) D[i][j] = lam;
  for (int i = 0; i < 3; ++i) D[i][i] += 2 * mu;
  for (int i = 3; i < 6; ++i) D[i][i] = mu;
  const double g[2] = {0.5 - 1.0 / (2 * sqrt(3.0)), 0.5 + 1.0 / (2 * sqrt(3.0))};
  for (int i = 0; i < 24; ++i) for (int j = 0; j < 24; ++j) K[i][j] = 0.0;
  for (int gz = 0; gz < 2; ++gz) for (int gy = 0; gy < 2; ++gy) for (int gx = 0; gx < 2; ++gx) {
    const double pt[3] = {g[gx], g[gy], g[gz]};
    double B[6][24] = {{0}};
    for (int a = 0; a < 8; ++a) {  /* corner a = (x fastest): (a&1, (a>>1)&1, a>>2) */
      const int cx[3] = {a & 1, (a >> 1) & 1, a >> 2};
      double f[3], sg[3];
      for (int d = 0; d < 3; ++d) { f[d] = cx[d] ? pt[d] : 1 - pt[d]; sg[d] = cx[d] ? 1.0 : -1.0; }
      const double dx = sg[0] * f[1] * f[2], dy = f[0] * sg[1] * f[2], dz = f[0] * f[1] * sg[2];
      B[0][3 * a] = dx; B[1][3 * a + 1] = dy; B[2][3 * a + 2] = dz;
      B[3][3 * a] = dy; B[3][3 * a + 1] = dx;
      B[4][3 * a + 1] = dz; B[4][3 * a + 2] = dy;
      B[5][3 * a] = dz; B[5][3 * a + 2] = dx;
    }
    for (int i = 0; i < 24; ++i)
      for (int j = 0; j < 24; ++j) {
        double v = 0.0;
        for (int p = 0; p < 6; ++p) for (int q = 0; q < 6; ++q) v += B[p][i] * D[p][q] * B[q][j];
        K[i][j] += v / 8.0;
      }
  }
  for (int i = 0; i < 24; ++i) for (int j = i + 1; j < 24; ++j) { const double v = 0.5 * (K[i][j] + K[j][i]); K[i][j] = K[j][i] = v; }
}

static int elasticity_csr(int N, CPLM_Mat_CSR_t* A) {
  const long long nfree_nodes = (long long)(N - 1) * N * N;
  const long long M = 3 * nfree_nodes;
  if (N < 2) CPLM_Abort("elasticity operator needs at least 2 nodes per side");
  if (M * 81 > 2147483647LL) CPLM_Abort("elasticity operator too large for 32-bit indices");
  double K[24][24];
  hex_stiffness(0.3, K);
  /* free node id of node (x >= 1, y, z): nodes are numbered x fastest and every (y, z) line loses its x = 0 node */
#define FREE_ID(x, y, z) ((((long long)(z) * N + (y)) * (N - 1)) + ((x) - 1))
  double kmax = 0.0;
  for (int i = 0; i < 24; ++i) for (int j = 0; j < 24; ++j) if (fabs(K[i][j]) > kmax) kmax = fabs(K[i][j]);
  const double tiny = 1e-12 * kmax;  /* couplings that vanish analytically come out as rounding noise: not stored */
  A->rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)M + 1));
  A->colInd = NULL; A->val = NULL;
  long long nnz = 0;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      A->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)nnz);
      A->val = (double*)pa_xmalloc(sizeof(double) * (size_t)nnz);
    }
    long long pos = 0;
    for (int z = 0; z < N; ++z) for (int y = 0; y < N; ++y) for (int x = 1; x < N; ++x) {
      const long long r = 3 * FREE_ID(x, y, z);
      for (int ci = 0; ci < 3; ++ci) {
        if (pass == 0) A->rowPtr[r + ci] = (int)pos;
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx, yy = y + dy, zz = z + dz;
          if (xx < 1 || xx >= N || yy < 0 || yy >= N || zz < 0 || zz >= N) continue;
          double row[3] = {0.0, 0.0, 0.0};
          double emax = 0.0;
          /* elements (ex, ey, ez) that contain both nodes, ascending element index */
          for (int ez = z - 1; ez <= z; ++ez) for (int ey = y - 1; ey <= y; ++ey) for (int ex = x - 1; ex <= x; ++ex) {
            if (ex < 0 || ex >= N - 1 || ey < 0 || ey >= N - 1 || ez < 0 || ez >= N - 1) continue;
            if (xx < ex || xx > ex + 1 || yy < ey || yy > ey + 1 || zz < ez || zz > ez + 1) continue;
            const double E = ((ez / 2) % 2 == 0) ? 1.0 : 1e3;
            if (E > emax) emax = E;
            const int a = (x - ex) + 2 * (y - ey) + 4 * (z - ez), b = (xx - ex) + 2 * (yy - ey) + 4 * (zz - ez);
            for (int cj = 0; cj < 3; ++cj) row[cj] += E * K[3 * a + ci][3 * b + cj];
          }
          const long long cfree = 3 * FREE_ID(xx, yy, zz);
          for (int cj = 0; cj < 3; ++cj) {
            if (!(fabs(row[cj]) > tiny * emax)) continue;
            if (pass == 1) { A->colInd[pos] = (int)(cfree + cj); A->val[pos] = row[cj]; }
            ++pos;
          }
        }
      }
    }
    if (pass == 0) { nnz = pos; A->rowPtr[M] = (int)nnz; }
  }
#undef FREE_ID
  memset(&A->info, 0, sizeof A->info);
  A->info.M = A->info.m = A->info.N = A->info.n = (int)M;
  A->info.nnz = A->info.lnnz = (int)nnz;
  A->info.blockSize = 1;
  A->info.format = FORMAT_CSR;
  A->info.structure = SYMMETRIC;
  return 0;
}

int pa_stencil_csr(int kind, int N, CPLM_Mat_CSR_t* A) {
  if (kind == 2) return elasticity_csr(N, A);
  const long long M = (long long)N * N * N;
  if (M > 2000000000LL) CPLM_Abort("stencil too large for 32-bit indices");
  const int reach = 1;
  long long nnz = 0;
  A->rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)M + 1));
  /* count */
  for (int z = 0; z < N; ++z) for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
    int c = 0;
    for (int dz = -reach; dz <= reach; ++dz) for (int dy = -reach; dy <= reach; ++dy) for (int dx = -reach; dx <= reach; ++dx) {
      if (kind == 0 && (abs(dx) + abs(dy) + abs(dz)) > 1) continue;
      const int xx = x + dx, yy = y + dy, zz = z + dz;
      if (xx < 0 || xx >= N || yy < 0 || yy >= N || zz < 0 || zz >= N) continue;
      ++c;
    }
    A->rowPtr[((long long)z * N + y) * N + x] = c;
  }
  { long long s = 0; for (long long i = 0; i < M; ++i) { int c = A->rowPtr[i]; A->rowPtr[i] = (int)s; s += c; } nnz = s; }
  if (nnz > 2147483647LL) CPLM_Abort("stencil has more than 2^31-1 non-zeros");
  A->rowPtr[M] = (int)nnz;
  A->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)nnz);
  A->val = (double*)pa_xmalloc(sizeof(double) * (size_t)nnz);
  const double diag = kind == 0 ? 6.0 : 26.0;
  for (int z = 0; z < N; ++z) for (int y = 0; y < N; ++y) for (int x = 0; x < N; ++x) {
    const long long i = ((long long)z * N + y) * N + x;
    int p = A->rowPtr[i];
    for (int dz = -reach; dz <= reach; ++dz) for (int dy = -reach; dy <= reach; ++dy) for (int dx = -reach; dx <= reach; ++dx) {
      if (kind == 0 && (abs(dx) + abs(dy) + abs(dz)) > 1) continue;
      const int xx = x + dx, yy = y + dy, zz = z + dz;
      if (xx < 0 || xx >= N || yy < 0 || yy >= N || zz < 0 || zz >= N) continue;
      A->colInd[p] = (int)(((long long)zz * N + yy) * N + xx);
      A->val[p] = (dx == 0 && dy == 0 && dz == 0) ? diag : -1.0;
      ++p;
    }
  }
  memset(&A->info, 0, sizeof A->info);
  A->info.M = A->info.m = A->info.N = A->info.n = (int)M;
  A->info.nnz = A->info.lnnz = (int)nnz;
  A->info.blockSize = 1;
  A->info.format = FORMAT_CSR;
  A->info.structure = SYMMETRIC;
  return 0;
}
