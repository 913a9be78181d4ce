/*
 * prealps_b200.h -- extensions of the preAlps ECG + block-Jacobi API that the B200
 * library adds next to the reference's entry points (operator.h, block_jacobi.h, ecg.h).
 *
 * Virtual subdomains.  The reference hard-wires
 *   #METIS parts = #block-Jacobi blocks = #MPI ranks >= enlFac
 * (ref: utils/operator.c:45,88-93, src/solvers/ecg.c:178-183).  On a GPU box one wants
 * S = 8 subdomains on 1, 2, 4 or 8 GPUs with identical numerics.  Here a process owns the
 * CONSECUTIVE subdomains [s_lo, s_hi) of S: its rows are rowPos[s_lo]..rowPos[s_hi], its
 * block-Jacobi has s_hi - s_lo blocks, the column of T(r0) that a row feeds is
 * (subdomain id) % enlFac (ref: ecg.c:162), and the driver's right-hand side is generated
 * per subdomain (ref: examples/test_ecg_prealps_op.c:172-184).  With one subdomain per MPI
 * rank this is exactly the reference's model.
 */
#ifndef PREALPS_B200_H
#define PREALPS_B200_H

#include "cplm_types.h"
#include "ecg.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device / communicator selection (before any Build) */
int preAlps_b200_SetDevice(int device);              /* default: PREALPS_CUDA_DEVICE, else rank % #devices */
int preAlps_b200_NcclUniqueId(void* id128);          /* 128 bytes, create on one process */
int preAlps_b200_InitNccl(int nranks, int rank, const void* id128);   /* one process per GPU */

/* ---- operator construction from memory (every process passes the same global matrix).
 * Full CSR (both triangles), 0-based, rows sorted by column.  scale != 0 applies the reference's
 * symmetric max-scaling (ref: cplm_matcsr.c:1461-1554); the matrix is then partitioned into S
 * parts with METIS k-way, permuted (ref: operator.c:77-86) and the process keeps the row panel of
 * subdomains [s_lo, s_hi).  parts_in (length M, may be NULL) overrides METIS (tests). */
int preAlps_b200_OperatorBuildCSR(int M, const int* rowPtr, const int* colInd, const double* val,
                                  int S, int s_lo, int s_hi, int scale, const int* parts_in);
int preAlps_b200_OperatorBuildFile(const char* mtx, int S, int s_lo, int s_hi);
/* synthetic operators of BASELINE.json: kind 0 = 7-point Poisson, 1 = 27-point stencil, 2 = Q1 linear
 * elasticity on N^3 nodes (3 dof per node, the face x = 0 clamped) */
int preAlps_b200_OperatorBuildStencil(int kind, int N, int S, int s_lo, int s_hi);
/* block-Jacobi over all local subdomains of the current operator */
int preAlps_b200_BlockJacobiCreate(void);

/* ---- inspection (host copies; valid until preAlps_OperatorFree) */
int preAlps_b200_GetPartition(int* S, int* s_lo, int* s_hi);
int preAlps_b200_GetPerm(int** perm, int* n);                 /* perm[new] = old; rank that built it only */
int preAlps_b200_GetDiagBlock(int b, CPLM_Mat_CSR_t* D);       /* b-th local block, upper triangle */
int preAlps_b200_GetHalo(int** halo_cols, int* nhalo);        /* sorted global columns read from other processes */
/* halo plan at process granularity: neighbours (ascending), local rows sent to each, slices of the halo received */
int preAlps_b200_GetHaloPlan(int* nnbr, int** nbr, int** send_ptr, int** send_idx, int** recv_ptr);
double preAlps_b200_Stat(const char* name);                   /* "spmm_bytes_t8", "bj_bytes_t8", "bj_nnz_exact", ... */

/* ---- the driver's right-hand side, per subdomain: srand(0); rhs[i] = rand()/RAND_MAX;
 * global 2-norm; rhs[i] /= norm for i >= 1 of every subdomain (ref: test_ecg_prealps_op.c:172-184) */
int preAlps_b200_DriverRhs(double* rhs);

/* ---- whole solves through the RCI API, loop of test_ecg_prealps_op.c:203-223 in C */
typedef struct {
  int iter;                /* iterations of the last solve */
  double res, normb;       /* final ||R||_F and ||b|| */
  double true_relres;      /* ||b - A x|| / ||b|| recomputed with the operator */
  double t_solve;          /* wall-clock seconds, Initialize .. Finalize, host buffers in and out */
  double t_dev_ms;         /* device time of the same region (CUDA events) */
  int nhist;
  int stopped;
} preAlps_b200_SolveInfo;
int preAlps_b200_Solve(int enlFac, double tol, int maxIter, int ortho_alg, int bs_red, double* rhs,
                       double* sol, double* res_hist, int max_hist, preAlps_b200_SolveInfo* info);
/* block size (ecg.bs) after every iteration of the last preAlps_b200_Solve: shrinks with bs_red = 1 (ADAPT_BS);
 * copies at most `max` entries, returns the number recorded */
int preAlps_b200_LastBlockSizes(int* out, int max);
/* benchmark region: `warmup` then `steps` ECG iterations (restarting converged solves), device
 * resident, timed with CUDA events on the library stream.  ms_out = time of the `steps` iterations. */
int preAlps_b200_BenchIterations(int enlFac, double tol, int ortho_alg, double* rhs, int warmup, int steps,
                                 float* ms_out, long long* launches_out);
/* time `reps` calls of one kernel class on the current operator: what 0 = SpMM, 1 = block-Jacobi apply,
 * 2 = the dense ECG passes of one iteration; returns the mean ms per call */
int preAlps_b200_BenchKernel(int what, int t, int reps, int flush_l2, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif
