/*
 * ecg.h -- Enlarged Conjugate Gradient, reverse-communication interface.
 * Field order and types of preAlps_ECG_t are ABI: the reference driver allocates the
 * struct on its stack, fills comm/globPbSize/locPbSize/maxIter/enlFac/tol/ortho_alg/bs_red
 * and reads R, P, AP, Z, iter, res, bs and the timers (ref: src/solvers/ecg.h:45-100,
 * examples/test_ecg_prealps_op.c:187-226).
 *
 * B200 edition: every block (X, R, P, AP, Z, ...) is a ROW_MAJOR device array owned by
 * the library; `work` is the device pool, `iwork` a host array.  The protocol is the
 * reference's (ref: src/solvers/ecg.c:173-286, manual section 3.2):
 *   Initialize -> rci 0; caller computes P = M^{-1} R, AP = A P
 *   Iterate(rci 0): consumes AP, updates X and R, sets rci 1
 *   StoppingCriterion: res = ||R||_F, *stop = !(res > normb*tol && iter < maxIter && bs > 0)
 *   caller computes Z = M^{-1} AP (Orthodir) or M^{-1} R (Orthomin)
 *   Iterate(rci 1): consumes Z, builds the next P, sets rci 0; caller computes AP = A P
 *   Finalize: solution = sum of the columns of X, frees everything.
 */
#ifndef ECG_H
#define ECG_H

#include <mpi.h>
#include "cplm_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { ORTHOMIN, ORTHODIR, ORTHODIR_FUSED } preAlps_ECG_Ortho_Alg_t;
typedef enum { ADAPT_BS, NO_BS_RED } preAlps_ECG_Block_Size_Red_t;

typedef struct {
  double* b;                 /* right-hand side (unused, as in the reference) */

  CPLM_Mat_Dense_t* X;       /* approximate solution, m x t */
  CPLM_Mat_Dense_t* R;       /* residual, m x t */
  CPLM_Mat_Dense_t* V;       /* search directions: aliases P (P_prev is a separate buffer here) */
  CPLM_Mat_Dense_t* AV;      /* A*V: aliases AP */
  CPLM_Mat_Dense_t* Z;       /* preconditioned block */
  CPLM_Mat_Dense_t* alpha;   /* t x t, column-major, device */
  CPLM_Mat_Dense_t* beta;    /* 2t x t, column-major, device */

  CPLM_Mat_Dense_t* P;       /* user-interface handles */
  CPLM_Mat_Dense_t* AP;
  double* R_p;
  double* P_p;
  double* AP_p;
  double* Z_p;

  double* work;              /* device pool */
  int* iwork;

  double normb;
  double res;
  int iter;
  int bs;
  int kbs;

  int globPbSize;
  int locPbSize;
  int maxIter;
  int enlFac;
  double tol;
  preAlps_ECG_Ortho_Alg_t ortho_alg;
  preAlps_ECG_Block_Size_Red_t bs_red;
  MPI_Comm comm;

  double tot_t;              /* timers, seconds (host wall-clock around the enqueued work) */
  double comm_t;
  double trsm_t;
  double gemm_t;
  double potrf_t;
  double pstrf_t;
  double lapmt_t;
  double gesvd_t;
  double geqrf_t;
  double ormqr_t;
  double copy_t;
} preAlps_ECG_t;

int preAlps_ECGInitialize(preAlps_ECG_t* ecg, double* rhs, int* rci_request);
int preAlps_ECGIterate(preAlps_ECG_t* ecg, int* rci_request);
int preAlps_ECGStoppingCriterion(preAlps_ECG_t* ecg, int* stop);
int preAlps_ECGFinalize(preAlps_ECG_t* ecg, double* solution);
void preAlps_ECGPrint(preAlps_ECG_t* ecg, int verbosity);

/* "private" entry points of the reference, kept because other drivers call Reset/WrapUp */
int _preAlps_ECGMalloc(preAlps_ECG_t* ecg);
int _preAlps_ECGReset(preAlps_ECG_t* ecg, double* rhs, int* rci_request);
int _preAlps_ECGWrapUp(preAlps_ECG_t* ecg, double* solution);
void _preAlps_ECGFree(preAlps_ECG_t* ecg);
int _preAlps_ECGSplit(double* x, CPLM_Mat_Dense_t* XSplit, int colIndex);  /* ref: ecg.h:214 */
int _preAlps_ECGIterateOmin(preAlps_ECG_t* ecg, int* rci_request);
int _preAlps_ECGIterateOdir(preAlps_ECG_t* ecg, int* rci_request);
int _preAlps_ECGIterateOdirFused(preAlps_ECG_t* ecg, int* rci_request);

#ifdef __cplusplus
}
#endif
#endif
