/*
 * oracle/shim/mkl.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Minimal stand-in for Intel MKL's umbrella header so that the UNMODIFIED
 * preAlps reference sources under /root/reference compile here (MKL is not
 * installed).  Dense BLAS/LAPACKE calls are forwarded to the OpenBLAS that
 * ships inside scipy (symbols carry a scipy_ prefix); the MKL-only entry points
 * (mkl_dcsrmm, mkl_?omatcopy, mkl_calloc, pardiso, ...) are implemented in
 * plain C in oracle/mkl_shim.c and oracle/pardiso_shim.c.
 */
#ifndef ORACLE_SHIM_MKL_H
#define ORACLE_SHIM_MKL_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int MKL_INT;
typedef void* _MKL_DSS_HANDLE_t;
#define lapack_int int

typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef CBLAS_LAYOUT CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;

#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102

void cblas_dgemm(CBLAS_LAYOUT, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int M, int N, int K, double alpha,
                 const double* A, int lda, const double* B, int ldb, double beta, double* C, int ldc);
void cblas_dtrsm(CBLAS_LAYOUT, CBLAS_SIDE, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int M, int N,
                 double alpha, const double* A, int lda, double* B, int ldb);
void cblas_dtrmm(CBLAS_LAYOUT, CBLAS_SIDE, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int M, int N,
                 double alpha, const double* A, int lda, double* B, int ldb);
void cblas_dgemv(CBLAS_LAYOUT, CBLAS_TRANSPOSE, int M, int N, double alpha, const double* A, int lda,
                 const double* X, int incX, double beta, double* Y, int incY);

int LAPACKE_dpotrf(int layout, char uplo, int n, double* a, int lda);
int LAPACKE_dpstrf(int layout, char uplo, int n, double* a, int lda, int* piv, int* rank, double tol);
int LAPACKE_dlapmt(int layout, int forwrd, int m, int n, double* x, int ldx, int* k);
int LAPACKE_dgesvd(int layout, char jobu, char jobvt, int m, int n, double* a, int lda, double* s,
                   double* u, int ldu, double* vt, int ldvt, double* superb);
int LAPACKE_dgeqrf(int layout, int m, int n, double* a, int lda, double* tau);
int LAPACKE_dormqr(int layout, char side, char trans, int m, int n, int k, const double* a, int lda,
                   const double* tau, double* c, int ldc);
int LAPACKE_dorgqr(int layout, int m, int n, int k, double* a, int lda, const double* tau);
int LAPACKE_dtrtrs(int layout, char uplo, char trans, char diag, int n, int nrhs, const double* a,
                   int lda, double* b, int ldb);
double LAPACKE_dlange(int layout, char norm, int m, int n, const double* a, int lda);

void* mkl_malloc(size_t size, int align);
void* mkl_calloc(size_t num, size_t size, int align);
void  mkl_free(void* p);
void  MKL_Set_Num_Threads(int n);
#define mkl_set_num_threads MKL_Set_Num_Threads

void mkl_domatcopy(char ordering, char trans, size_t rows, size_t cols, double alpha,
                   const double* A, size_t lda, double* B, size_t ldb);
void mkl_dimatcopy(char ordering, char trans, size_t rows, size_t cols, double alpha,
                   double* AB, size_t lda, size_t ldb);
void mkl_domatadd(char ordering, char transa, char transb, size_t m, size_t n, double alpha,
                  const double* A, size_t lda, double beta, const double* B, size_t ldb,
                  double* C, size_t ldc);

void mkl_dcsrmm(const char* transa, const MKL_INT* m, const MKL_INT* n, const MKL_INT* k,
                const double* alpha, const char* matdescra, const double* val, const MKL_INT* indx,
                const MKL_INT* pntrb, const MKL_INT* pntre, const double* b, const MKL_INT* ldb,
                const double* beta, double* c, const MKL_INT* ldc);
void mkl_dcsrmv(const char* transa, const MKL_INT* m, const MKL_INT* k, const double* alpha,
                const char* matdescra, const double* val, const MKL_INT* indx, const MKL_INT* pntrb,
                const MKL_INT* pntre, const double* x, const double* beta, double* y);

void pardiso(_MKL_DSS_HANDLE_t pt, const MKL_INT* maxfct, const MKL_INT* mnum, const MKL_INT* mtype,
             const MKL_INT* phase, const MKL_INT* n, const void* a, const MKL_INT* ia,
             const MKL_INT* ja, MKL_INT* perm, const MKL_INT* nrhs, MKL_INT* iparm,
             const MKL_INT* msglvl, void* b, void* x, MKL_INT* error);
void pardisoinit(_MKL_DSS_HANDLE_t pt, const MKL_INT* mtype, MKL_INT* iparm);

#ifdef __cplusplus
}
#endif
#endif
