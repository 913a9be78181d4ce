/*
 * operator.h -- the distributed linear operator A of the ECG solver, B200 edition.
 * Same entry points, argument meaning and error behaviour as the reference
 * (ref: utils/operator.h:50-110, utils/operator.c:38-393); underneath, the local row
 * panel lives in HBM and preAlps_BlockOperator runs the sm_100a CSR SpMM of
 * libprealps_cuda (include/prealps_cuda.h) after exchanging boundary rows only.
 */
#ifndef OPERATOR_H
#define OPERATOR_H

#include <mpi.h>
#include "cplm_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* rank 0 reads the MatrixMarket file, scales (symmetric max-scaling), partitions with METIS k-way
 * into one part per rank, permutes, ships the row panels; every rank then builds colPos, dep and
 * its device operator.  ref: utils/operator.c:38-134 */
int preAlps_OperatorBuild(const char* matrixFilename, MPI_Comm comm);
/* same with a right-hand side read from a text file and scattered.  ref: utils/operator.c:136-268 */
int preAlps_OperatorRHSBuild(const char* matrixFilename, const char* rhsFilename, double** rhs, MPI_Comm comm);
/* the matrix is already partitioned: locA = my row panel (global columns), idxRowBegin = rowPos.
 * ref: utils/operator.c:271-308 */
int preAlps_OperatorBuildNoPerm(CPLM_Mat_CSR_t* locA, int* idxRowBegin, int nbBlockPerProcs, MPI_Comm comm);
void preAlps_OperatorFree(void);
void preAlps_OperatorPrint(int rank);
int preAlps_OperatorGetSizes(int* M, int* m);
/* AX = A * X for an m x n block.  Device-resident ROW_MAJOR blocks (the ones owned by
 * preAlps_ECG_t) are used in place; host blocks of either storage are staged through HBM.
 * ref: utils/operator.c:334-351 */
int preAlps_BlockOperator(CPLM_Mat_Dense_t* X, CPLM_Mat_Dense_t* AX);
/* aliases of the library's host copies, valid until preAlps_OperatorFree (ref: operator.c:353-393).
 * Unlike the reference, A.colInd stays GLOBAL after the first product. */
int preAlps_OperatorGetA(CPLM_Mat_CSR_t* A);
int preAlps_OperatorGetRowPosPtr(int** rowPos, int* sizeRowPos);
int preAlps_OperatorGetColPosPtr(int** colPos, int* sizeColPos);
int preAlps_OperatorGetDepPtr(int** dep, int* sizeDep);

#ifdef __cplusplus
}
#endif
#endif
