/*
 * oracle/mkl_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Implements the MKL entry points that the unmodified preAlps reference
 * sources call on the ECG + block-Jacobi path, so that those sources can be
 * compiled and run here (MKL itself is not installed):
 *   - cblas_* / LAPACKE_* are forwarded to OpenBLAS as bundled with scipy
 *     (identical C interfaces, symbols prefixed scipy_);
 *   - mkl_dcsrmm / mkl_dcsrmv follow the Sparse BLAS level-2/3 semantics used at
 *     /root/reference/utils/cplm_light/cplm_kernels.c:644-664 (4-array CSR,
 *     offsets relative to pntrb[0]; matdescra[3]=='F' => 1-based column indices
 *     and column-major dense operands, 'C' => 0-based and row-major);
 *   - mkl_?omatcopy / mkl_domatadd / mkl_calloc / mkl_free are plain C.
 * pardiso() lives in oracle/pardiso_shim.c.
 */
#include "shim/mkl.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- OpenBLAS (scipy build: every public symbol carries a scipy_ prefix) */
extern void scipy_cblas_dgemm(int, int, int, int, int, int, double, const double*, int, const double*,
                              int, double, double*, int);
extern void scipy_cblas_dtrsm(int, int, int, int, int, int, int, double, const double*, int, double*, int);
extern void scipy_cblas_dtrmm(int, int, int, int, int, int, int, double, const double*, int, double*, int);
extern void scipy_cblas_dgemv(int, int, int, int, double, const double*, int, const double*, int,
                              double, double*, int);
extern int scipy_LAPACKE_dpotrf(int, char, int, double*, int);
extern int scipy_LAPACKE_dpstrf(int, char, int, double*, int, int*, int*, double);
extern int scipy_LAPACKE_dlapmt(int, int, int, int, double*, int, int*);
extern int scipy_LAPACKE_dgesvd(int, char, char, int, int, double*, int, double*, double*, int,
                                double*, int, double*);
extern int scipy_LAPACKE_dgeqrf(int, int, int, double*, int, double*);
extern int scipy_LAPACKE_dormqr(int, char, char, int, int, int, const double*, int, const double*,
                                double*, int);
extern int scipy_LAPACKE_dorgqr(int, int, int, int, double*, int, const double*);
extern int scipy_LAPACKE_dtrtrs(int, char, char, char, int, int, const double*, int, double*, int);
extern double scipy_LAPACKE_dlange(int, char, int, int, const double*, int);
extern void scipy_openblas_set_num_threads(int);

void cblas_dgemm(CBLAS_LAYOUT l, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int M, int N, int K,
                 double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                 double* C, int ldc) {
  scipy_cblas_dgemm(l, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
void cblas_dtrsm(CBLAS_LAYOUT l, CBLAS_SIDE s, CBLAS_UPLO u, CBLAS_TRANSPOSE t, CBLAS_DIAG d, int M,
                 int N, double alpha, const double* A, int lda, double* B, int ldb) {
  scipy_cblas_dtrsm(l, s, u, t, d, M, N, alpha, A, lda, B, ldb);
}
void cblas_dtrmm(CBLAS_LAYOUT l, CBLAS_SIDE s, CBLAS_UPLO u, CBLAS_TRANSPOSE t, CBLAS_DIAG d, int M,
                 int N, double alpha, const double* A, int lda, double* B, int ldb) {
  scipy_cblas_dtrmm(l, s, u, t, d, M, N, alpha, A, lda, B, ldb);
}
void cblas_dgemv(CBLAS_LAYOUT l, CBLAS_TRANSPOSE t, int M, int N, double alpha, const double* A,
                 int lda, const double* X, int incX, double beta, double* Y, int incY) {
  scipy_cblas_dgemv(l, t, M, N, alpha, A, lda, X, incX, beta, Y, incY);
}
int LAPACKE_dpotrf(int l, char uplo, int n, double* a, int lda) { return scipy_LAPACKE_dpotrf(l, uplo, n, a, lda); }
int LAPACKE_dpstrf(int l, char uplo, int n, double* a, int lda, int* piv, int* rank, double tol) {
  return scipy_LAPACKE_dpstrf(l, uplo, n, a, lda, piv, rank, tol);
}
int LAPACKE_dlapmt(int l, int f, int m, int n, double* x, int ldx, int* k) { return scipy_LAPACKE_dlapmt(l, f, m, n, x, ldx, k); }
int LAPACKE_dgesvd(int l, char ju, char jvt, int m, int n, double* a, int lda, double* s, double* u,
                   int ldu, double* vt, int ldvt, double* superb) {
  return scipy_LAPACKE_dgesvd(l, ju, jvt, m, n, a, lda, s, u, ldu, vt, ldvt, superb);
}
int LAPACKE_dgeqrf(int l, int m, int n, double* a, int lda, double* tau) { return scipy_LAPACKE_dgeqrf(l, m, n, a, lda, tau); }
int LAPACKE_dormqr(int l, char side, char trans, int m, int n, int k, const double* a, int lda,
                   const double* tau, double* c, int ldc) {
  return scipy_LAPACKE_dormqr(l, side, trans, m, n, k, a, lda, tau, c, ldc);
}
int LAPACKE_dorgqr(int l, int m, int n, int k, double* a, int lda, const double* tau) {
  return scipy_LAPACKE_dorgqr(l, m, n, k, a, lda, tau);
}
int LAPACKE_dtrtrs(int l, char uplo, char trans, char diag, int n, int nrhs, const double* a, int lda,
                   double* b, int ldb) {
  return scipy_LAPACKE_dtrtrs(l, uplo, trans, diag, n, nrhs, a, lda, b, ldb);
}
double LAPACKE_dlange(int l, char norm, int m, int n, const double* a, int lda) {
  return scipy_LAPACKE_dlange(l, norm, m, n, a, lda);
}

/* ---- memory / threading */
void* mkl_malloc(size_t size, int align) {
  void* p = NULL;
  if (align < (int)sizeof(void*)) align = sizeof(void*);
  if (posix_memalign(&p, (size_t)align, size ? size : 1)) return NULL;
  return p;
}
void* mkl_calloc(size_t num, size_t size, int align) {
  void* p = mkl_malloc(num * size, align);
  if (p) memset(p, 0, num * size);
  return p;
}
void mkl_free(void* p) { free(p); }
void MKL_Set_Num_Threads(int n) { scipy_openblas_set_num_threads(n); }

/* ---- out-of-place / in-place scaled copies (no transposition is ever requested on this path) */
static int is_col(char ordering) { return ordering == 'C' || ordering == 'c'; }

void mkl_domatcopy(char ordering, char trans, size_t rows, size_t cols, double alpha, const double* A,
                   size_t lda, double* B, size_t ldb) {
  int tr = !(trans == 'N' || trans == 'n');
  size_t outer = is_col(ordering) ? cols : rows, inner = is_col(ordering) ? rows : cols;
  if (!tr) {
    for (size_t o = 0; o < outer; ++o)
      for (size_t i = 0; i < inner; ++i) B[o * ldb + i] = alpha * A[o * lda + i];
  } else {
    for (size_t o = 0; o < outer; ++o)
      for (size_t i = 0; i < inner; ++i) B[i * ldb + o] = alpha * A[o * lda + i];
  }
}

void mkl_dimatcopy(char ordering, char trans, size_t rows, size_t cols, double alpha, double* AB,
                   size_t lda, size_t ldb) {
  if (!(trans == 'N' || trans == 'n')) {
    fprintf(stderr, "[mkl_shim] mkl_dimatcopy: transposition not implemented\n");
    abort();
  }
  size_t outer = is_col(ordering) ? cols : rows, inner = is_col(ordering) ? rows : cols;
  /* ldb <= lda on this path (ecg.c:483 compacts t rows to t1 rows); moving
   * forward through memory therefore never overwrites unread input */
  for (size_t o = 0; o < outer; ++o)
    for (size_t i = 0; i < inner; ++i) AB[o * ldb + i] = alpha * AB[o * lda + i];
}

void mkl_domatadd(char ordering, char transa, char transb, size_t m, size_t n, double alpha,
                  const double* A, size_t lda, double beta, const double* B, size_t ldb, double* C,
                  size_t ldc) {
  int col = is_col(ordering);
  int ta = !(transa == 'N' || transa == 'n'), tb = !(transb == 'N' || transb == 'n');
  for (size_t i = 0; i < m; ++i)
    for (size_t j = 0; j < n; ++j) {
      double a = col ? (ta ? A[i * lda + j] : A[j * lda + i]) : (ta ? A[j * lda + i] : A[i * lda + j]);
      double b = col ? (tb ? B[i * ldb + j] : B[j * ldb + i]) : (tb ? B[j * ldb + i] : B[i * ldb + j]);
      if (col) C[j * ldc + i] = alpha * a + beta * b; else C[i * ldc + j] = alpha * a + beta * b;
    }
}

/* ---- Sparse BLAS */
void mkl_dcsrmm(const char* transa, const MKL_INT* m_, const MKL_INT* n_, const MKL_INT* k_,
                const double* alpha_, const char* matdescra, const double* val, const MKL_INT* indx,
                const MKL_INT* pntrb, const MKL_INT* pntre, const double* b, const MKL_INT* ldb_,
                const double* beta_, double* c, const MKL_INT* ldc_) {
  (void)k_;
  if (!(*transa == 'N' || *transa == 'n') || !(matdescra[0] == 'G' || matdescra[0] == 'g')) {
    fprintf(stderr, "[mkl_shim] mkl_dcsrmm: only general, non-transposed A is implemented\n");
    abort();
  }
  const int m = *m_, n = *n_, ldb = *ldb_, ldc = *ldc_;
  const double alpha = *alpha_, beta = *beta_;
  const int fortran = (matdescra[3] == 'F' || matdescra[3] == 'f');
  const int base = fortran ? 1 : 0;
  const int p0 = m > 0 ? pntrb[0] : 0;
  double* acc = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j) acc[j] = 0.0;
    for (int p = pntrb[i] - p0; p < pntre[i] - p0; ++p) {
      const double a = val[p];
      const int col = indx[p] - base;
      if (fortran) for (int j = 0; j < n; ++j) acc[j] += a * b[(size_t)j * ldb + col];
      else         for (int j = 0; j < n; ++j) acc[j] += a * b[(size_t)col * ldb + j];
    }
    for (int j = 0; j < n; ++j) {
      double* cij = fortran ? &c[(size_t)j * ldc + i] : &c[(size_t)i * ldc + j];
      *cij = (beta == 0.0) ? alpha * acc[j] : alpha * acc[j] + beta * (*cij);
    }
  }
  free(acc);
}

void mkl_dcsrmv(const char* transa, const MKL_INT* m_, const MKL_INT* k_, const double* alpha_,
                const char* matdescra, const double* val, const MKL_INT* indx, const MKL_INT* pntrb,
                const MKL_INT* pntre, const double* x, const double* beta_, double* y) {
  (void)k_;
  if (!(*transa == 'N' || *transa == 'n')) { fprintf(stderr, "[mkl_shim] mkl_dcsrmv: transposed A\n"); abort(); }
  const int m = *m_;
  const int base = (matdescra[3] == 'F' || matdescra[3] == 'f') ? 1 : 0;
  const int p0 = m > 0 ? pntrb[0] : 0;
  for (int i = 0; i < m; ++i) {
    double s = 0.0;
    for (int p = pntrb[i] - p0; p < pntre[i] - p0; ++p) s += val[p] * x[indx[p] - base];
    y[i] = (*beta_ == 0.0) ? (*alpha_) * s : (*alpha_) * s + (*beta_) * y[i];
  }
}
