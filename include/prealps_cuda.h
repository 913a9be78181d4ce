/*
 * prealps_cuda.h -- C ABI of libprealps_cuda.so, the sm_100a device library
 * underneath the preAlps ECG + block-Jacobi API (operator.h, block_jacobi.h,
 * ecg.h in this directory).  Plain C: opaque handles, pointers and sizes only.
 *
 * Every entry point names the reference interface it replaces; "ref:" paths are
 * relative to the NLAFET/preAlps tree.
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on failure; pcu_last_error()
 *     returns a static description of the last failure on the calling thread;
 *   - dense blocks are ROW-MAJOR m x t on the device (t doubles contiguous per
 *     row), t = number of active columns (the reference's enlarging factor /
 *     current block size), ld = allocated row stride in doubles;
 *   - everything is enqueued on the context's stream; only the functions
 *     documented as synchronous wait for the device;
 *   - there is no CPU fallback: without a CUDA device pcu_ctx_create fails.
 */
#ifndef PREALPS_CUDA_H
#define PREALPS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pcu_ctx pcu_ctx;
typedef struct pcu_spmm pcu_spmm;
typedef struct pcu_bj pcu_bj;

/* ------------------------------------------------------------------ context */
int pcu_ctx_create(int device, pcu_ctx** out);
int pcu_ctx_destroy(pcu_ctx* ctx);
const char* pcu_last_error(void);
int pcu_device_count(void);
/* raw cudaStream_t of the context (for callers that time with their own events) */
void* pcu_ctx_stream(pcu_ctx* ctx);
int pcu_sync(pcu_ctx* ctx);                      /* synchronous */

void* pcu_malloc(pcu_ctx* ctx, size_t bytes);    /* device memory, NULL on failure */
int pcu_free(pcu_ctx* ctx, void* dptr);
int pcu_memset(pcu_ctx* ctx, void* dptr, int byte, size_t bytes);
int pcu_h2d(pcu_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);  /* synchronous */
int pcu_d2h(pcu_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);  /* synchronous */
int pcu_d2d(pcu_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);
/* the first ncols columns of a row-major m x * device block: dst <- src (replaces the column-range
 * mkl_domatcopy of ref: ecg.c:521-523 once the block size has been reduced) / dst <- 0 */
int pcu_copy_cols(pcu_ctx* ctx, int m, int ncols, double* dst, int ldd, const double* src, int lds);
int pcu_zero_cols(pcu_ctx* ctx, int m, int ncols, double* dst, int ldd);
void* pcu_host_alloc(size_t bytes);              /* pinned host memory */
int pcu_host_free(void* p);

/* 1 if p is device (or managed) memory, 0 if it is ordinary host memory */
int pcu_ptr_is_device(const void* p);
/* evict L2 between timed repetitions: overwrites a scratch buffer larger than the 126 MB L2 */
int pcu_flush_l2(pcu_ctx* ctx);

/* device-side timers (CUDA events on the context stream) */
int pcu_timer_start(pcu_ctx* ctx, int slot);     /* slot in [0,16) */
int pcu_timer_stop(pcu_ctx* ctx, int slot);
int pcu_timer_elapsed_ms(pcu_ctx* ctx, int slot, float* ms); /* synchronous */
/* number of kernels this library launched on ctx since creation */
int64_t pcu_launch_count(pcu_ctx* ctx);

/* ------------------------------------------------------- inter-GPU plumbing
 * One process per GPU.  The unique id is created on one rank, shipped to the
 * others by the caller (torch.distributed / MPI / a file) and handed in here.
 * Replaces MPI_Allreduce at ref: src/solvers/ecg.c:148,254,427,441,513,563 and
 * the MPI_Isend/Irecv halo of ref: utils/cplm_v0/cplm_v0_matmult_v2.c:184-276. */
#define PCU_NCCL_ID_BYTES 128
int pcu_nccl_unique_id(void* out_id128);
int pcu_ctx_init_nccl(pcu_ctx* ctx, int nranks, int rank, const void* id128);
int pcu_comm_size(pcu_ctx* ctx);
int pcu_comm_rank(pcu_ctx* ctx);
/* setup-time exchange of host int lists with neighbour ranks over NCCL (synchronous):
 * sends send_data[send_ptr[q]..send_ptr[q+1]) to nbr[q], receives recv_ptr-sliced data from it */
int pcu_exchange_ints(pcu_ctx* ctx, int nnbr, const int* nbr, const int* send_ptr, const int* send_data,
                      const int* recv_ptr, int* recv_data);
/* in-place sum of n doubles that live on the device (no-op for one rank) */
int pcu_allreduce_sum(pcu_ctx* ctx, double* dbuf, int n);
/* host array of n ints from rank `root` to every rank (staged through HBM, ncclBroadcast); no-op on one rank */
int pcu_bcast_ints(pcu_ctx* ctx, int* host_data, long long n, int root);

/* ------------------------------------------------------------- K1: CSR SpMM
 * Y (m x t) = A_loc * [X ; H].  A_loc has m rows; column c < m refers to row c
 * of X, column c >= m to row c-m of the halo buffer H (nhalo rows, filled by
 * pcu_spmm_halo_exchange or left empty when the process owns every column).
 * Replaces mkl_dcsrmm as called by CPLM_MatCSRKernelGenMatDenseMult
 * (ref: utils/cplm_light/cplm_kernels.c:620-671) from CPLM_MatCSRMatMult_v2
 * (ref: utils/cplm_v0/cplm_v0_matmult_v2.c:216-230,258-273). */
int pcu_spmm_create(pcu_ctx* ctx, int m, int nhalo, const int* rowPtr, const int* colInd,
                    const double* val, pcu_spmm** out);
int pcu_spmm_destroy(pcu_spmm* op);
/* Halo plan: for each of nnbr neighbour ranks, the local rows to send
 * (send_idx[send_ptr[q]..send_ptr[q+1]), ascending) and the slice of H that the
 * neighbour fills (recv_ptr[q]..recv_ptr[q+1]).  Boundary rows only -- the
 * reference ships the whole m x t block to every neighbour
 * (ref: utils/cplm_light/cplm_matdense.c:90-109). */
int pcu_spmm_set_halo(pcu_spmm* op, int nnbr, const int* nbr_rank, const int* send_ptr,
                      const int* send_idx, const int* recv_ptr);
/* pack boundary rows of X and exchange them over NCCL into the halo buffer */
int pcu_spmm_halo_exchange(pcu_spmm* op, const double* X, int ldx, int t);
/* pack only (device buffer of send rows, row-major, ld = t) -- for transports
 * other than NCCL (the MPI staging path of the host layer) */
int pcu_spmm_halo_pack(pcu_spmm* op, const double* X, int ldx, int t, double** packed_dev, int* nrows);
double* pcu_spmm_halo_buffer(pcu_spmm* op, int t);   /* device pointer to H (nhalo x t, ld = t) */
int pcu_spmm_apply(pcu_spmm* op, const double* X, int ldx, double* Y, int ldy, int t);
/* pcu_spmm_halo_exchange + pcu_spmm_apply in one call (NCCL transport): the exchange runs next to the local part of the
 * product, like MPI_Isend / diagonal block / MPI_Irecv of ref: cplm_v0_matmult_v2.c:182-276 (PREALPS_SPMM_OVERLAP=0 when the
 * operator is created: one after the other) */
int pcu_spmm_apply_exchange(pcu_spmm* op, const double* X, int ldx, double* Y, int ldy, int t);
/* algorithmic bytes of one apply at block width t */
double pcu_spmm_bytes(pcu_spmm* op, int t);

/* ------------------------------------------------ K2/K3: block-Jacobi solves
 * nblk diagonal blocks; block b owns local rows [blk_ptr[b], blk_ptr[b+1]) and is
 * given as the upper triangle (diagonal included) in 0-based CSR with
 * block-local, sorted column indices -- exactly what CPLM_MatCSRGetDiagBlock
 * hands to PARDISO (ref: utils/cplm_v0/cplm_v0_matcsr.c:287-389,
 * src/preconditioners/block_jacobi.c:48-54).  Create = ordering + symbolic on
 * the host (integer work) + numeric supernodal Cholesky on the device
 * (replaces pardiso phase 12, ref: utils/cplm_light/cplm_kernels.c:741-783).
 * Apply = level-scheduled multi-RHS forward/backward sweeps, X = A_bb^{-1} B
 * (replaces pardiso phase 33, ref: cplm_kernels.c:790-853).  B and X may alias. */
int pcu_bj_create(pcu_ctx* ctx, int nblk, const int* blk_ptr, const int* const* rowPtr,
                  const int* const* colInd, const double* const* val, pcu_bj** out);
int pcu_bj_destroy(pcu_bj* bj);
int pcu_bj_apply(pcu_bj* bj, const double* B, int ldb, double* X, int ldx, int t);
/* host-only symbolic analysis of one block (what pcu_bj_create runs per block), exposed for tests:
 * perm[new] = old (n), supernode column ranges sn_col (nsuper+1), row structures sn_rows with
 * offsets sn_rowptr (nsuper+1), parent and level per supernode; stats = {nnz exact, nnz stored,
 * levels, flops}.  Arrays must hold n+1 entries (sn_rows: rows_cap); returns -9 if rows_cap is too small. */
int pcu_bj_analyze(int n, const int* rowPtr, const int* colInd, int use_metis, int* perm, int* nsuper,
                   int* sn_col, long long* sn_rowptr, int* sn_rows, long long rows_cap, int* sn_parent,
                   int* sn_level, double* stats4);
/* statistics: 0 nnz(L) exact, 1 nnz stored (dense supernodes), 2 supernodes,
 * 3 tree levels, 4 bytes streamed per apply at width t=8, 5 factor flops,
 * 6 factor seconds (device), 7 analysis seconds (host), 8 kernel launches per apply */
double pcu_bj_stat(pcu_bj* bj, int which);
double pcu_bj_bytes(pcu_bj* bj, int t);   /* algorithmic bytes of one apply at width t: exact nnz(L), SURVEY.md 8(d) */
double pcu_bj_stored_bytes(pcu_bj* bj, int t);   /* bytes the sweeps move: stored panels (zeros and padding included) + vectors */

/* -------------------------------------- K4/K5/K6: fused tall-skinny kernels
 * All blocks row-major with leading dimension ld; small matrices (t x t) are
 * column-major on the device like the reference's alpha/beta.  Reductions are
 * two-stage with a fixed grid, hence run-to-run deterministic.
 *
 * pcu_gram2: G1 (t x t) = A1^T B1 and, in the same pass, G2 = A2^T B2 (second pair optional).
 *   G1/G2 are column-major and should be adjacent so that one all-reduce covers both.
 *   Replaces cblas_dgemm('T','N') via CPLM_MatDenseKernelMatDotProd at
 *   ref: src/solvers/ecg.c:250,311,330,425,438,510,557-560. */
int pcu_gram2(pcu_ctx* ctx, int m, int t, const double* A1, int lda1, const double* B1, int ldb1,
              double* G1, const double* A2, int lda2, const double* B2, int ldb2, double* G2);
/* pcu_ortho_update: one streaming pass doing the whole "rci_request == 0" half of an
 *   Orthodir iteration (ref: ecg.c:431-443,500-501) plus the residual norm of ecg.c:250-261:
 *     U = chol_upper(G) with G = AP^T P (reduced over ranks), P <- P U^{-1}, AP <- AP U^{-1},
 *     alpha = U^{-T} Gpr with Gpr = P_old^T R (== P_new^T R), X += P alpha, R -= AP alpha,
 *     rr[0] = local part of ||R_new||_F^2.
 *   U_out/alpha_out (t x t col-major) receive U and alpha; status_dev[0] != 0 if G is not SPD
 *   (the reference ignores dpotrf's return code in Orthodir, ecg.c:431, and aborts in Orthomin, :320).
 *   X/R may be NULL: then only P and AP are orthonormalised. */
int pcu_ortho_update(pcu_ctx* ctx, int m, int t, const double* G, const double* Gpr, double* P,
                     int ldp, double* AP, int ldap, double* X, int ldx, double* R, int ldr,
                     double* U_out, double* alpha_out, double* rr, int* status_dev);
/* pcu_update_xr: X += P alpha, R -= AP alpha (alpha t x t col-major), rr[0] = ||R||_F^2 local part.
 *   Replaces 2x dgemm at ref: ecg.c:500-501 (the un-fused form, used after a block-size reduction). */
int pcu_update_xr(pcu_ctx* ctx, int m, int t, const double* P, int ldp, const double* AP, int ldap,
                  const double* alpha, double* X, int ldx, double* R, int ldr, double* rr);
/* pcu_transform_update: ADAPT_BS after a reduction (ref: ecg.c:470-501): P <- P Wm and AP <- AP Wm in place, X += P_old Wx,
 *   R -= AP_old Wx, rr[0] = ||R||_F^2 local part, Wm / Wx general t x t column-major.  AP, X, R (and Wx) may be NULL.
 *   Replaces dormqr/dtrsm + 2x dgemm of the reference in one pass over the blocks. */
int pcu_transform_update(pcu_ctx* ctx, int m, int t, const double* Wm, const double* Wx, double* P, int ldp, double* AP,
                         int ldap, double* X, int ldx, double* R, int ldr, double* rr);
/* pcu_update_z: Z -= P beta1 + Pprev beta2 with beta1 (t1 x tz) and beta2 (t2 x tz) column-major,
 *   tight leading dimensions; Pprev/beta2 may be NULL with t2 = 0.
 *   Replaces dgemm at ref: ecg.c:517 (V = [P, P_prev], beta = [beta1; beta2]) and ecg.c:354. */
int pcu_update_z(pcu_ctx* ctx, int m, int tz, double* Z, int ldz, const double* P, int ldp, int t1,
                 const double* beta1, const double* Pprev, int ldpp, int t2, const double* beta2);
/* ORTHODIR_FUSED (ref: ecg.c:532-658): the small-matrix step U = chol_upper(mu), beta1 <- U^-T beta1 U^-1,
 * beta2 <- beta2 U^-1 (ref: ecg.c:577-587), and Z <- Z U^-1 with U = chol_upper(G) (ref: ecg.c:583) */
int pcu_fused_small(pcu_ctx* ctx, int t, const double* mu, double* beta1, double* beta2, double* U_out, int* status_dev);
int pcu_right_solve(pcu_ctx* ctx, int m, int t, const double* G, double* Z, int ldz);
/* sol[i] = sum_c X[i,c]; replaces dgemv at ref: ecg.c:674 (cplm_kernels.c:454-472) */
int pcu_sum_columns(pcu_ctx* ctx, int m, int t, const double* X, int ldx, double* sol_dev);
/* R[i, col_of_row[i]] = rhs[i], all else 0 (R0 = T(b), ref: ecg.c:201-221); col_of_row on device */
int pcu_split_rhs(pcu_ctx* ctx, int m, int t, const double* rhs_dev, const int* col_of_row_dev,
                  double* R, int ldr);
/* fro2[0] = sum of squares of an m x t block (local part) */
int pcu_fro2(pcu_ctx* ctx, int m, int t, const double* R, int ldr, double* fro2_dev);

#ifdef __cplusplus
}
#endif
#endif /* PREALPS_CUDA_H */
