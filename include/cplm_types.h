/*
 * cplm_types.h -- the CPaLAMeM-light data types that cross the preAlps API of the
 * ECG + block-Jacobi path, with the binary layout of the reference so that callers
 * compiled against NLAFET/preAlps (examples/test_ecg_prealps_op.c) link unchanged:
 *   CPLM_Mat_CSR_t / CPLM_Info_t   ref: utils/cplm_core/cplm_matcsr_struct.h:49-73
 *   CPLM_Mat_Dense_t / Info_Dense  ref: utils/cplm_light/cplm_matdense.h:21-39
 *   CPLM_IVector_t                 ref: utils/cplm_v0/cplm_v0_ivector.h:21-25
 *   CPLM_DVector_t                 ref: utils/cplm_v0/cplm_v0_dvector.h:16-19
 *   timing no-op macros + stepN    ref: utils/cplm_core/cplm_timing.h:4-48
 * In this library a CPLM_Mat_Dense_t handed out by preAlps_ECGInitialize is
 * ROW_MAJOR and its val points to DEVICE memory (the driver never looks inside,
 * ref: examples/test_ecg_prealps_op.c:203-223).
 */
#ifndef PREALPS_B200_CPLM_TYPES_H
#define PREALPS_B200_CPLM_TYPES_H

#include <stddef.h>
#include <mpi.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { FORMAT_CSR, FORMAT_BCSR, FORMAT_BCSR_VAR } CPLM_Mat_CSR_format_t;
typedef enum { UNSYMMETRIC, SYMMETRIC } Struct_Type;
typedef enum { AVOID_PERMUTE, PERMUTE } Choice_permutation;

typedef struct {
  int M, N, nnz;      /* global rows, columns, non-zeros */
  int m, n, lnnz;     /* local rows, columns, non-zeros */
  int blockSize;
  CPLM_Mat_CSR_format_t format;
  Struct_Type structure;
} CPLM_Info_t;

typedef struct {
  CPLM_Info_t info;
  int* rowPtr;
  int* colInd;
  double* val;
} CPLM_Mat_CSR_t;

#define CPLM_MatCSRNULL() { .info = { .M = 0, .N = 0, .nnz = 0, .m = 0, .n = 0, .lnnz = 0, .blockSize = 0, \
                                      .format = FORMAT_CSR, .structure = UNSYMMETRIC },                    \
                            .rowPtr = NULL, .colInd = NULL, .val = NULL }

typedef enum { ROW_MAJOR, COL_MAJOR } CPLM_storage_type_t;

typedef struct {
  int M, N;           /* global shape */
  int m, n;           /* local shape */
  int lda;
  int nval;
  CPLM_storage_type_t stor_type;
} CPLM_Info_Dense_t;

typedef struct {
  double* val;
  CPLM_Info_Dense_t info;
} CPLM_Mat_Dense_t;

#define CPLM_MatDenseNULL() { .val = NULL, .info = { .M = 0, .N = 0, .m = 0, .n = 0, .lda = 0, .nval = 0, \
                                                     .stor_type = ROW_MAJOR } }

typedef struct { int* val; int nval; int size; } CPLM_IVector_t;
#define CPLM_IVectorNULL() { .val = NULL, .nval = 0, .size = 0 }
typedef struct { double* val; int nval; } CPLM_DVector_t;

/* lda = m (COL_MAJOR) or n (ROW_MAJOR), nval = m*n  (ref: utils/cplm_light/cplm_matdense.c:124-135) */
int CPLM_MatDenseSetInfo(CPLM_Mat_Dense_t* A, int M, int N, int m, int n, CPLM_storage_type_t storage);
void CPLM_MatCSRFree(CPLM_Mat_CSR_t* A);
/* message + MPI_Abort(MPI_COMM_WORLD, 1)  (ref: utils/cplm_core/cplm_utils.c:42-59) */
void CPLM_FAbort(const char* fun, const char* format, ...);
#define CPLM_Abort(...) CPLM_FAbort(__func__, __VA_ARGS__)

/* instrumentation hooks of the reference expand to nothing without CPaLAMeM */
#ifndef CPLM_TIMING_H
#define CPLM_TIMING_H
#define CPLM_PUSH
#define CPLM_POP
#define CPLM_BEGIN_TIME
#define CPLM_END_TIME
#define CPLM_OPEN_TIMER
#define CPLM_CLOSE_TIMER
#define CPLM_TIC(a, b)
#define CPLM_TAC(a)
#define CPLM_SetEnv()
#define CPLM_printTimer(a)
#define CPLM_resetTimer()
enum { step1 = 1, step2, step3, step4, step5, step6, step7, step8, step9, step10, step11, step12, step13,
       step14, step15, step16, step17, step18, step19, step20, step21, step22, step23, step24, step25,
       step26, step27, step28, step29, step30 };
#endif

#ifdef __cplusplus
}
#endif
#endif
