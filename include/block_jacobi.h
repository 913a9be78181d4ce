/*
 * block_jacobi.h -- block-Jacobi preconditioner, one exact Cholesky solve per METIS
 * subdomain (ref: src/preconditioners/block_jacobi.h:47-65, block_jacobi.c:26-119).
 * MKL PARDISO is replaced by the device supernodal Cholesky and the level-scheduled
 * multi-RHS sweeps of libprealps_cuda (pcu_bj_create / pcu_bj_apply).
 */
#ifndef BLOCK_JACOBI_H
#define BLOCK_JACOBI_H

#include "cplm_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A: local row panel with global columns, rowPos/colPos as returned by the operator getters.
 * Extracts the upper triangle of the diagonal block (ref: cplm_v0_matcsr.c:287-389), orders,
 * analyses and factorises it.  Aborts like the reference if the block is not SPD. */
int preAlps_BlockJacobiCreate(CPLM_Mat_CSR_t* A, int* rowPos, int sizeRowPos, int* colPos, int sizeColPos);
/* solve in place for a single host vector (internal use in the reference, block_jacobi.c:65-91) */
int preAlps_BlockJacobiInitialize(CPLM_DVector_t* rhs);
/* B_out = M^{-1} A_in for an m x n block (device ROW_MAJOR in place, host blocks staged) */
int preAlps_BlockJacobiApply(CPLM_Mat_Dense_t* A_in, CPLM_Mat_Dense_t* B_out);
void preAlps_BlockJacobiFree(void);

#ifdef __cplusplus
}
#endif
#endif
