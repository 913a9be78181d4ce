/*
 * pa_block_jacobi.c -- preAlps_BlockJacobi* (ref: src/preconditioners/block_jacobi.c:26-119).
 * The diagonal block of every local subdomain is extracted exactly as the reference does for
 * PARDISO (upper triangle, block-local columns), then ordered / analysed on the host and
 * factorised on the device by pcu_bj_create; Apply is pcu_bj_apply.
 */
#include "pa_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

double* pa_stage_in(const CPLM_Mat_Dense_t* X);
void pa_stage_out(CPLM_Mat_Dense_t* Y);
void pa_ensure_stage(size_t doubles);

static void create_from_blocks(int nblk, CPLM_Mat_CSR_t* D, const int* blk_ptr) {
  pa_state_t* g = &pa_g;
  if (g->bj) { pcu_bj_destroy(g->bj); g->bj = NULL; }
  const int** rp = (const int**)pa_xmalloc(sizeof(int*) * (size_t)nblk);
  const int** ci = (const int**)pa_xmalloc(sizeof(int*) * (size_t)nblk);
  const double** vv = (const double**)pa_xmalloc(sizeof(double*) * (size_t)nblk);
  for (int b = 0; b < nblk; ++b) { rp[b] = D[b].rowPtr; ci[b] = D[b].colInd; vv[b] = D[b].val; }
  const int rc = pcu_bj_create(pa_ctx(), nblk, blk_ptr, rp, ci, vv, &g->bj);
  free(rp); free(ci); free(vv);
  if (rc != 0) CPLM_Abort("PARDISO Cholesky error: %d (%s)", rc, pcu_last_error());  /* ref: block_jacobi.c:59 */
}

int preAlps_b200_BlockJacobiCreate(void) {
  pa_state_t* g = &pa_g;
  if (!g->built) CPLM_Abort("preAlps_b200_BlockJacobiCreate called before the operator was built");
  if (g->diag) { for (int b = 0; b < g->bj_nblk; ++b) CPLM_MatCSRFree(&g->diag[b]); free(g->diag); }
  const int nblk = g->s_hi - g->s_lo;
  g->bj_nblk = nblk;
  g->diag = (CPLM_Mat_CSR_t*)pa_xcalloc((size_t)nblk, sizeof(CPLM_Mat_CSR_t));
  int* blk_ptr = (int*)pa_xmalloc(sizeof(int) * ((size_t)nblk + 1));
  for (int b = 0; b <= nblk; ++b) blk_ptr[b] = g->rowPos[g->s_lo + b] - g->g0;
  for (int b = 0; b < nblk; ++b)
    pa_diag_block(&g->A, g->rowPos, g->colPos, g->S, g->s_lo + b, blk_ptr[b], blk_ptr[b + 1], &g->diag[b]);
  create_from_blocks(nblk, g->diag, blk_ptr);
  free(blk_ptr);
  return 0;
}

int preAlps_BlockJacobiCreate(CPLM_Mat_CSR_t* A, int* rowPos, int sizeRowPos, int* colPos, int sizeColPos) {
  pa_state_t* g = &pa_g;
  (void)sizeColPos;
  if (g->built && A->rowPtr == g->A.rowPtr && rowPos == g->rowPos && colPos == g->colPos)
    return preAlps_b200_BlockJacobiCreate();  /* the driver's usual path: arguments alias the operator */
  /* foreign panel: one block, the one of this MPI rank (ref: block_jacobi.c:33-34, cplm_v0_matcsr.c:300) */
  int rank;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  const int S = sizeRowPos - 1;
  if (g->diag) { for (int b = 0; b < g->bj_nblk; ++b) CPLM_MatCSRFree(&g->diag[b]); free(g->diag); }
  g->bj_nblk = 1;
  g->diag = (CPLM_Mat_CSR_t*)pa_xcalloc(1, sizeof(CPLM_Mat_CSR_t));
  pa_diag_block(A, rowPos, colPos, S, rank, 0, A->info.m, &g->diag[0]);
  int blk_ptr[2] = {0, A->info.m};
  create_from_blocks(1, g->diag, blk_ptr);
  return 0;
}

int preAlps_b200_GetDiagBlock(int b, CPLM_Mat_CSR_t* D) {
  if (!pa_g.diag || b < 0 || b >= pa_g.bj_nblk) return 1;
  *D = pa_g.diag[b];
  return 0;
}

int preAlps_BlockJacobiApply(CPLM_Mat_Dense_t* A_in, CPLM_Mat_Dense_t* B_out) {
  pa_state_t* g = &pa_g;
  if (!g->bj) CPLM_Abort("preAlps_BlockJacobiApply called before preAlps_BlockJacobiCreate");
  if (!A_in || !A_in->val) CPLM_Abort(" wrong test 'A_in->val != NULL'");
  if (!B_out || !B_out->val) CPLM_Abort(" wrong test 'B_out->val != NULL'");
  const int t = A_in->info.n;
  const int dev_in = pa_is_device_block(A_in), dev_out = pa_is_device_block(B_out);
  const double* b; int ldb;
  if (dev_in) {
    if (A_in->info.stor_type != ROW_MAJOR) CPLM_Abort("device blocks must be ROW_MAJOR");
    b = A_in->val; ldb = A_in->info.lda;
  } else { b = pa_stage_in(A_in); ldb = t; }
  double* x; int ldx;
  if (dev_out) { x = B_out->val; ldx = B_out->info.lda; }
  else {
    /* ref: cplm_kernels.c:819-828 re-shapes the output like the input */
    if (B_out->info.m != A_in->info.m || B_out->info.n != t) {
      CPLM_MatDenseSetInfo(B_out, A_in->info.M, A_in->info.N, A_in->info.m, t, A_in->info.stor_type);
      B_out->val = (double*)realloc(B_out->val, sizeof(double) * (size_t)A_in->info.m * t);
    }
    pa_ensure_stage((size_t)A_in->info.m * t + 8);
    if (!dev_in) b = g->d_stage_in;
    x = g->d_stage_out; ldx = t;
  }
  pa_cuda_check(pcu_bj_apply(g->bj, b, ldb, x, ldx, t), "pcu_bj_apply");
  if (!dev_out) pa_stage_out(B_out);
  return 0;
}

int preAlps_BlockJacobiInitialize(CPLM_DVector_t* rhs) {
  CPLM_Mat_Dense_t b = CPLM_MatDenseNULL();
  CPLM_MatDenseSetInfo(&b, rhs->nval, 1, rhs->nval, 1, COL_MAJOR);
  b.val = rhs->val;
  return preAlps_BlockJacobiApply(&b, &b);  /* solution overwrites rhs (ref: block_jacobi.c:65-91) */
}

void preAlps_BlockJacobiFree(void) {
  pa_state_t* g = &pa_g;
  if (g->bj) { pcu_bj_destroy(g->bj); g->bj = NULL; }
  if (g->diag) { for (int b = 0; b < g->bj_nblk; ++b) CPLM_MatCSRFree(&g->diag[b]); free(g->diag); g->diag = NULL; }
  g->bj_nblk = 0;
}
