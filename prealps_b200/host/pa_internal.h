/* pa_internal.h -- internals shared by the host C layer (libprealps_b200).  Not installed. */
#ifndef PA_INTERNAL_H
#define PA_INTERNAL_H

#include <stddef.h>
#include <mpi.h>

#include "../../include/cplm_types.h"
#include "../../include/operator.h"
#include "../../include/block_jacobi.h"
#include "../../include/ecg.h"
#include "../../include/prealps_b200.h"
#include "../../include/prealps_cuda.h"

void* pa_xmalloc(size_t n);
void* pa_xcalloc(size_t n, size_t s);

/* pa_csr.c */
int pa_load_mtx(const char* filename, CPLM_Mat_CSR_t* A, int verbose);
int pa_check_diag(const CPLM_Mat_CSR_t* A);
int pa_sym_scale(CPLM_Mat_CSR_t* A);
int pa_kway_parts(const CPLM_Mat_CSR_t* A, int S, int* parts);
void pa_parts_to_perm(int M, const int* parts, int S, int* posB, int* perm);
int pa_permute_sym(const CPLM_Mat_CSR_t* A, const int* perm, CPLM_Mat_CSR_t* B);
int pa_permute_panel(const CPLM_Mat_CSR_t* A, const int* perm, int r0, int r1, CPLM_Mat_CSR_t* B);
int pa_row_panel(const CPLM_Mat_CSR_t* A, int r0, int r1, CPLM_Mat_CSR_t* B);
int pa_col_block_pos(const CPLM_Mat_CSR_t* A, const int* rowPos, int S, int** colPos_out, int* n_out);
int pa_comm_dep(const int* colPos, int m, int S, int lo, int hi, int** dep_out, int* ndep);
int pa_diag_block(const CPLM_Mat_CSR_t* A, const int* rowPos, const int* colPos, int S, int blk, int pr0, int pr1,
                  CPLM_Mat_CSR_t* D);
int pa_halo_map(const CPLM_Mat_CSR_t* A, int g0, int g1, int** halo_out, int* nhalo_out, int** colLoc_out);
int pa_stencil_csr(int kind, int N, CPLM_Mat_CSR_t* A);

/* how blocks travel between processes */
enum { PA_XPORT_NONE = 0, PA_XPORT_MPI = 1, PA_XPORT_NCCL = 2 };

typedef struct {
  int built;
  MPI_Comm comm;
  int xport;
  int nproc, rank;          /* processes taking part (MPI ranks or NCCL ranks) */
  int S, s_lo, s_hi;        /* subdomains: total, local range */
  int* sub_of_proc;         /* nproc+1: process p owns subdomains [sub_of_proc[p], sub_of_proc[p+1]) */
  int M, m;
  int g0, g1;               /* global row range of this process */
  CPLM_Mat_CSR_t A;         /* local row panel, GLOBAL columns (host) */
  int* rowPos; int nrowPos; /* S+1 */
  int* colPos; int ncolPos; /* m*S+1 */
  int* dep; int ndep;       /* neighbour subdomains outside the local range */
  int* perm; int nperm;     /* perm[new] = old (only where the partition was computed) */
  int* halo; int nhalo;     /* sorted global columns owned by other processes */
  /* halo plan (process granularity) */
  int nnbr; int* nbr;       /* neighbour processes, ascending */
  int* send_ptr; int* send_idx; int* recv_ptr;
  double* h_send; double* h_recv; int h_cap_t;   /* pinned staging for the MPI transport */
  /* device */
  pcu_ctx* ctx;
  pcu_spmm* spmm;
  pcu_bj* bj;
  int bj_nblk;
  CPLM_Mat_CSR_t* diag;     /* host copies of the local diagonal blocks (upper triangles) */
  int device;               /* requested device, -1 = automatic */
  int nccl_ready;
  /* staging blocks for host-resident operands */
  double* d_stage_in; double* d_stage_out; size_t stage_cap;
} pa_state_t;

extern pa_state_t pa_g;

pcu_ctx* pa_ctx(void);                          /* creates the context on first use; aborts without a GPU */
void pa_cuda_check(int rc, const char* what);   /* aborts with pcu_last_error() */
int pa_is_device_block(const CPLM_Mat_Dense_t* X);
/* sum n doubles that live on the device across processes (no-op for one process) */
double pa_sum_over_subdomains(const double* part_local);
void pa_allreduce_dev(double* dbuf, int n, double* comm_t);
double pa_wtime(void);

/* small dense algebra of the ADAPT_BS paths (pa_ecg.c), column-major, n <= 32; exported for the CPU tests */
int pa_h_chol_upper(int n, double* A, int lda);
void pa_h_triu_inv(int n, const double* U, int ldu, double* Ui, int ldi);
void pa_h_left_svd(int t, int n, const double* A, int lda, double* sv, double* Q, double* rows);
int pa_h_pivoted_chol(int n, double* A, int lda, int* piv, double tol);

#endif
