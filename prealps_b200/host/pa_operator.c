/*
 * pa_operator.c -- preAlps_Operator* (ref: utils/operator.c:38-393) over the device SpMM.
 *
 * Setup mirrors the reference: rank 0 loads, scales, partitions with METIS k-way, permutes
 * and ships row panels (operator.c:54-121); every process then builds colPos and dep
 * (operator.c:123-130).  New here: the halo plan (boundary rows only), the device upload and
 * the virtual-subdomain generalisation of include/prealps_b200.h.  The host CSR keeps its
 * GLOBAL column indices for good (the reference rewrites them in place on the first product,
 * cplm_v0_matmult_v2.c:52-61 -- SURVEY.md H7).
 */
#include "pa_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

pa_state_t pa_g = {.device = -1, .comm = MPI_COMM_WORLD};

double pa_wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

void pa_cuda_check(int rc, const char* what) {
  if (rc != 0) CPLM_Abort("%s failed: %s", what, pcu_last_error());
}

pcu_ctx* pa_ctx(void) {
  if (pa_g.ctx) return pa_g.ctx;
  const int ndev = pcu_device_count();
  if (ndev <= 0)
    CPLM_Abort("no CUDA device is visible: libprealps_b200 runs its hot path on the GPU only (there is no CPU fallback)");
  int dev = pa_g.device;
  if (dev < 0) {
    const char* e = getenv("PREALPS_CUDA_DEVICE");
    if (e) dev = atoi(e);
    else {
      int rank = pa_g.rank, inited = 0;
      MPI_Initialized(&inited);
      if (inited && !pa_g.nccl_ready) MPI_Comm_rank(MPI_COMM_WORLD, &rank);
      const char* lr = getenv("LOCAL_RANK");
      if (lr) rank = atoi(lr);
      dev = rank % ndev;
    }
  }
  pa_cuda_check(pcu_ctx_create(dev, &pa_g.ctx), "pcu_ctx_create");
  return pa_g.ctx;
}

int preAlps_b200_SetDevice(int device) { pa_g.device = device; return 0; }
int preAlps_b200_NcclUniqueId(void* id128) { return pcu_nccl_unique_id(id128); }

int preAlps_b200_InitNccl(int nranks, int rank, const void* id128) {
  pa_g.rank = rank;
  pa_g.nproc = nranks;
  pa_g.nccl_ready = 1;
  pcu_ctx* c = pa_ctx();
  if (pcu_ctx_init_nccl(c, nranks, rank, id128)) {
    fprintf(stderr, "preAlps_b200_InitNccl: %s\n", pcu_last_error());
    return 1;
  }
  return 0;
}

int pa_is_device_block(const CPLM_Mat_Dense_t* X) { return pcu_ptr_is_device(X->val); }

/* sum over ALL subdomains, in subdomain order, of one number per local subdomain (NCCL transport): every process fills
 * its own slots of a zeroed S-vector, the all-reduce adds zeros to them (exact), and everybody adds the S numbers in the
 * same order -- the order in which the reference's ranks are summed by the MPI shim of the golden runs.  The result does
 * not depend on the number of GPUs (a scalar all-reduce does: its tree changes the last bit of ||b||). */
double pa_sum_over_subdomains(const double* part_local) {
  pa_state_t* g = &pa_g;
  pcu_ctx* c = pa_ctx();
  const int S = g->S;
  double* h = (double*)pa_xcalloc((size_t)S, sizeof(double));
  for (int s = g->s_lo; s < g->s_hi; ++s) h[s] = part_local[s - g->s_lo];
  double* d = (double*)pcu_malloc(c, sizeof(double) * (size_t)S);
  if (!d) CPLM_Abort("device allocation failed: %s", pcu_last_error());
  pa_cuda_check(pcu_h2d(c, d, h, sizeof(double) * (size_t)S), "pcu_h2d");
  pa_cuda_check(pcu_allreduce_sum(c, d, S), "pcu_allreduce_sum");
  pa_cuda_check(pcu_d2h(c, h, d, sizeof(double) * (size_t)S), "pcu_d2h");
  pcu_free(c, d);
  double sum = 0.0;
  for (int s = 0; s < S; ++s) sum += h[s];
  free(h);
  return sum;
}

void pa_allreduce_dev(double* dbuf, int n, double* comm_t) {
  if (pa_g.nproc <= 1 || pa_g.xport == PA_XPORT_NONE) return;
  const double t0 = pa_wtime();
  if (pa_g.xport == PA_XPORT_NCCL) {
    pa_cuda_check(pcu_allreduce_sum(pa_g.ctx, dbuf, n), "pcu_allreduce_sum");
  } else {
    double* h = (double*)pa_xmalloc(sizeof(double) * (size_t)n);
    pa_cuda_check(pcu_d2h(pa_g.ctx, h, dbuf, sizeof(double) * (size_t)n), "pcu_d2h");
    MPI_Allreduce(MPI_IN_PLACE, h, n, MPI_DOUBLE, MPI_SUM, pa_g.comm);
    pa_cuda_check(pcu_h2d(pa_g.ctx, dbuf, h, sizeof(double) * (size_t)n), "pcu_h2d");
    free(h);
  }
  if (comm_t) *comm_t += pa_wtime() - t0;
}

/* ------------------------------------------------------------------ teardown */
void preAlps_OperatorFree(void) {
  pa_state_t* g = &pa_g;
  if (g->bj) { pcu_bj_destroy(g->bj); g->bj = NULL; }
  if (g->diag) { for (int b = 0; b < g->bj_nblk; ++b) CPLM_MatCSRFree(&g->diag[b]); free(g->diag); g->diag = NULL; g->bj_nblk = 0; }
  if (g->spmm) { pcu_spmm_destroy(g->spmm); g->spmm = NULL; }
  if (g->ctx) {
    if (g->d_stage_in) pcu_free(g->ctx, g->d_stage_in);
    if (g->d_stage_out) pcu_free(g->ctx, g->d_stage_out);
  }
  g->d_stage_in = g->d_stage_out = NULL; g->stage_cap = 0;
  if (g->h_send) pcu_host_free(g->h_send);
  if (g->h_recv) pcu_host_free(g->h_recv);
  g->h_send = g->h_recv = NULL; g->h_cap_t = 0;
  CPLM_MatCSRFree(&g->A);
  free(g->rowPos); free(g->colPos); free(g->dep); free(g->perm); free(g->halo); free(g->nbr);
  free(g->send_ptr); free(g->send_idx); free(g->recv_ptr); free(g->sub_of_proc);
  g->rowPos = g->colPos = g->dep = g->perm = g->halo = g->nbr = NULL;
  g->send_ptr = g->send_idx = g->recv_ptr = g->sub_of_proc = NULL;
  g->nrowPos = g->ncolPos = g->ndep = g->nperm = g->nhalo = g->nnbr = 0;
  g->built = 0;
  /* the device context (and the NCCL communicator) survive: they belong to the process */
}

/* ------------------------------------------------------------------ common tail of every Build */
static int owner_proc(const pa_state_t* g, int gcol) {
  int lo = 0, hi = g->S;  /* subdomain s with rowPos[s] <= gcol < rowPos[s+1] */
  while (hi - lo > 1) { int mid = (lo + hi) / 2; if (g->rowPos[mid] <= gcol) lo = mid; else hi = mid; }
  int plo = 0, phi = g->nproc;
  while (phi - plo > 1) { int mid = (plo + phi) / 2; if (g->sub_of_proc[mid] <= lo) plo = mid; else phi = mid; }
  return plo;
}

/* pa_g.A (panel, global columns), rowPos, S, s_lo, s_hi, nproc, rank, xport are set */
static int finish_build(void) {
  pa_state_t* g = &pa_g;
  g->M = g->rowPos[g->S];
  g->m = g->A.info.m;
  g->g0 = g->rowPos[g->s_lo];
  g->g1 = g->rowPos[g->s_hi];
  if (g->m != g->g1 - g->g0) CPLM_Abort("row panel has %d rows, partition says %d", g->m, g->g1 - g->g0);
  g->sub_of_proc = (int*)pa_xmalloc(sizeof(int) * ((size_t)g->nproc + 1));
  for (int p = 0; p <= g->nproc; ++p) g->sub_of_proc[p] = (int)((long long)p * g->S / g->nproc);
  if (g->sub_of_proc[g->rank] != g->s_lo || g->sub_of_proc[g->rank + 1] != g->s_hi)
    CPLM_Abort("process %d/%d must own subdomains [%d,%d) of %d (uniform consecutive split), got [%d,%d)", g->rank,
               g->nproc, g->sub_of_proc[g->rank], g->sub_of_proc[g->rank + 1], g->S, g->s_lo, g->s_hi);
  pa_col_block_pos(&g->A, g->rowPos, g->S, &g->colPos, &g->ncolPos);
  pa_comm_dep(g->colPos, g->m, g->S, g->s_lo, g->s_hi, &g->dep, &g->ndep);
  if (g->ndep == 0 && g->nproc > 1)
    CPLM_Abort("There is no dependencies between some blocks of A...");  /* ref: cplm_v0_matcsr.c:263 */
  int* colLoc = NULL;
  pa_halo_map(&g->A, g->g0, g->g1, &g->halo, &g->nhalo, &colLoc);
  /* PREALPS_B200_HOST_ONLY=1: stop before anything touches the device (CPU tests of the partition / halo plan) */
  const int host_only = getenv("PREALPS_B200_HOST_ONLY") != NULL;
  pcu_ctx* c = host_only ? NULL : pa_ctx();
  if (!host_only)
    pa_cuda_check(pcu_spmm_create(c, g->m, g->nhalo, g->A.rowPtr, colLoc, g->A.val, &g->spmm), "pcu_spmm_create");
  free(colLoc);
  if (host_only && g->xport == PA_XPORT_NCCL) CPLM_Abort("host-only mode needs the MPI transport");
  if (g->nproc > 1) {
    /* who needs how many rows from whom: need[p*nproc + q] = rows p reads from q */
    const int np = g->nproc;
    int* mine = (int*)pa_xcalloc((size_t)np, sizeof(int));
    for (int k = 0; k < g->nhalo; ++k) mine[owner_proc(g, g->halo[k])]++;
    int* need = (int*)pa_xcalloc((size_t)np * np, sizeof(int));
    if (g->xport == PA_XPORT_MPI) {
      MPI_Allgather(mine, np, MPI_INT, need, np, MPI_INT, g->comm);
    } else {
      double* tmp = (double*)pa_xcalloc((size_t)np * np, sizeof(double));
      for (int q = 0; q < np; ++q) tmp[(size_t)g->rank * np + q] = mine[q];
      double* d = (double*)pcu_malloc(c, sizeof(double) * (size_t)np * np);
      pa_cuda_check(d == NULL, "pcu_malloc");
      pa_cuda_check(pcu_h2d(c, d, tmp, sizeof(double) * (size_t)np * np), "pcu_h2d");
      pa_cuda_check(pcu_allreduce_sum(c, d, np * np), "pcu_allreduce_sum");
      pa_cuda_check(pcu_d2h(c, tmp, d, sizeof(double) * (size_t)np * np), "pcu_d2h");
      pcu_free(c, d);
      for (int i = 0; i < np * np; ++i) need[i] = (int)tmp[i];
      free(tmp);
    }
    g->nbr = (int*)pa_xmalloc(sizeof(int) * (size_t)np);
    g->nnbr = 0;
    for (int q = 0; q < np; ++q)
      if (q != g->rank && (need[(size_t)g->rank * np + q] || need[(size_t)q * np + g->rank])) g->nbr[g->nnbr++] = q;
    g->send_ptr = (int*)pa_xcalloc((size_t)g->nnbr + 1, sizeof(int));
    g->recv_ptr = (int*)pa_xcalloc((size_t)g->nnbr + 1, sizeof(int));
    for (int k = 0; k < g->nnbr; ++k) {
      g->recv_ptr[k + 1] = g->recv_ptr[k] + need[(size_t)g->rank * np + g->nbr[k]];
      g->send_ptr[k + 1] = g->send_ptr[k] + need[(size_t)g->nbr[k] * np + g->rank];
    }
    /* the halo is sorted by global column, hence already grouped by owner in ascending process order:
     * request list to neighbour k = halo[recv_ptr[k] .. recv_ptr[k+1]) */
    g->send_idx = (int*)pa_xmalloc(sizeof(int) * (size_t)(g->send_ptr[g->nnbr] > 0 ? g->send_ptr[g->nnbr] : 1));
    if (g->xport == PA_XPORT_MPI) {
      MPI_Request* rq = (MPI_Request*)pa_xmalloc(sizeof(MPI_Request) * (size_t)(g->nnbr > 0 ? g->nnbr : 1));
      for (int k = 0; k < g->nnbr; ++k)
        MPI_Isend(g->halo + g->recv_ptr[k], g->recv_ptr[k + 1] - g->recv_ptr[k], MPI_INT, g->nbr[k], 77, g->comm, &rq[k]);
      for (int k = 0; k < g->nnbr; ++k)
        MPI_Recv(g->send_idx + g->send_ptr[k], g->send_ptr[k + 1] - g->send_ptr[k], MPI_INT, g->nbr[k], 77, g->comm,
                 MPI_STATUS_IGNORE);
      MPI_Waitall(g->nnbr, rq, MPI_STATUSES_IGNORE);
      free(rq);
    } else {
      pa_cuda_check(pcu_exchange_ints(c, g->nnbr, g->nbr, g->recv_ptr, g->halo, g->send_ptr, g->send_idx),
                    "pcu_exchange_ints");
    }
    for (int i = 0; i < g->send_ptr[g->nnbr]; ++i) {
      g->send_idx[i] -= g->g0;  /* global -> local row */
      if (g->send_idx[i] < 0 || g->send_idx[i] >= g->m) CPLM_Abort("halo plan: neighbour asked for a row I do not own");
    }
    if (!host_only)
      pa_cuda_check(pcu_spmm_set_halo(g->spmm, g->nnbr, g->nbr, g->send_ptr, g->send_idx, g->recv_ptr), "pcu_spmm_set_halo");
    free(mine); free(need);
  } else if (g->nhalo > 0) {
    CPLM_Abort("single process but %d columns fall outside its rows", g->nhalo);
  }
  g->built = 1;
  return 0;
}

/* scale + partition + permute a global matrix held by this process; keeps rows of [s_lo, s_hi) */
static int partition_global(CPLM_Mat_CSR_t* G, int S, int s_lo, int s_hi, int scale, const int* parts_in) {
  pa_state_t* g = &pa_g;
  const int M = G->info.m;
  const int timing = getenv("PREALPS_B200_TIMING") != NULL;
  double t0 = pa_wtime(), t1;
#define PA_LAP(what) do { if (timing) { t1 = pa_wtime(); fprintf(stderr, "[prealps_b200] setup: %-28s %.3f s\n", what, t1 - t0); t0 = t1; } } while (0)
  if (scale) pa_sym_scale(G);
  PA_LAP("symmetric max-scaling");
  int* parts = (int*)pa_xmalloc(sizeof(int) * (size_t)M);
  if (parts_in) memcpy(parts, parts_in, sizeof(int) * (size_t)M);
  else if (g->xport == PA_XPORT_NCCL && g->nproc > 1) {
    /* one partition for everybody, like the reference (rank 0 partitions and ships the panels, ref: operator.c:54-121):
     * process 0 runs METIS, the others receive parts[] -- N concurrent copies of the same serial partitioning took 9.5 s
     * at N = 8 against 1.9 s alone, and nothing checked that they agreed */
    if (g->rank == 0 && pa_kway_parts(G, S, parts)) CPLM_Abort("METIS k-way partitioning failed");
    pa_cuda_check(pcu_bcast_ints(pa_ctx(), parts, M, 0), "pcu_bcast_ints");
  } else if (pa_kway_parts(G, S, parts)) CPLM_Abort("METIS k-way partitioning failed");
  PA_LAP("METIS k-way partition");
  g->rowPos = (int*)pa_xmalloc(sizeof(int) * ((size_t)S + 1));
  g->nrowPos = S + 1;
  g->perm = (int*)pa_xmalloc(sizeof(int) * (size_t)M);
  g->nperm = M;
  pa_parts_to_perm(M, parts, S, g->rowPos, g->perm);
  free(parts);
  PA_LAP("parts -> permutation");
  pa_permute_panel(G, g->perm, g->rowPos[s_lo], g->rowPos[s_hi], &g->A);  /* = rows of P A P^T, ref: operator.c:86,96-121 */
  CPLM_MatCSRFree(G);
  PA_LAP("symmetric permutation of the local rows");
#undef PA_LAP
  return 0;
}

/* ------------------------------------------------------------------ reference entry points */
static void send_panel(const CPLM_Mat_CSR_t* B, int dest, MPI_Comm comm) {
  /* ref: CPLM_MatCSRSend, cplm_matcsr.c:466-500: header, rowPtr, colInd, val */
  int hdr[9] = {B->info.M, B->info.N, B->info.nnz, B->info.m, B->info.n, B->info.lnnz, B->info.blockSize,
                (int)B->info.format, (int)B->info.structure};
  MPI_Send(hdr, 9, MPI_INT, dest, 0, comm);
  MPI_Send(B->rowPtr, B->info.m + 1, MPI_INT, dest, 1, comm);
  MPI_Send(B->colInd, B->info.lnnz, MPI_INT, dest, 2, comm);
  MPI_Send(B->val, B->info.lnnz, MPI_DOUBLE, dest, 3, comm);
}

static void recv_panel(CPLM_Mat_CSR_t* B, int src, MPI_Comm comm) {
  int hdr[9];
  MPI_Recv(hdr, 9, MPI_INT, src, 0, comm, MPI_STATUS_IGNORE);
  B->info.M = hdr[0]; B->info.N = hdr[1]; B->info.nnz = hdr[2]; B->info.m = hdr[3]; B->info.n = hdr[4];
  B->info.lnnz = hdr[5]; B->info.blockSize = hdr[6]; B->info.format = (CPLM_Mat_CSR_format_t)hdr[7];
  B->info.structure = (Struct_Type)hdr[8];
  B->rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)B->info.m + 1));
  B->colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)(B->info.lnnz > 0 ? B->info.lnnz : 1));
  B->val = (double*)pa_xmalloc(sizeof(double) * (size_t)(B->info.lnnz > 0 ? B->info.lnnz : 1));
  MPI_Recv(B->rowPtr, B->info.m + 1, MPI_INT, src, 1, comm, MPI_STATUS_IGNORE);
  MPI_Recv(B->colInd, B->info.lnnz, MPI_INT, src, 2, comm, MPI_STATUS_IGNORE);
  MPI_Recv(B->val, B->info.lnnz, MPI_DOUBLE, src, 3, comm, MPI_STATUS_IGNORE);
}

static int build_from_root(CPLM_Mat_CSR_t* G /* valid on rank 0 */, MPI_Comm comm, double** rhs_io, int rhs_len) {
  pa_state_t* g = &pa_g;
  if (g->built) preAlps_OperatorFree();
  int size, rank;
  MPI_Comm_size(comm, &size);
  MPI_Comm_rank(comm, &rank);
  g->comm = comm;
  g->nproc = size; g->rank = rank;
  g->S = size; g->s_lo = rank; g->s_hi = rank + 1;
  g->xport = size > 1 ? PA_XPORT_MPI : PA_XPORT_NONE;
  double* rhs_perm = NULL;
  if (rank == 0) {
    const int M = G->info.m;
    int* parts = (int*)pa_xmalloc(sizeof(int) * (size_t)M);
    if (pa_kway_parts(G, size, parts)) CPLM_Abort("METIS k-way partitioning failed");
    g->rowPos = (int*)pa_xmalloc(sizeof(int) * ((size_t)size + 1));
    g->nrowPos = size + 1;
    g->perm = (int*)pa_xmalloc(sizeof(int) * (size_t)M);
    g->nperm = M;
    pa_parts_to_perm(M, parts, size, g->rowPos, g->perm);
    free(parts);
    CPLM_Mat_CSR_t P = CPLM_MatCSRNULL();
    pa_permute_sym(G, g->perm, &P);
    CPLM_MatCSRFree(G);
    if (rhs_io && *rhs_io) {  /* ref: preAlps_doubleVector_permute, x_out[i] = b_in[p[i]] */
      rhs_perm = (double*)pa_xmalloc(sizeof(double) * (size_t)rhs_len);
      for (int i = 0; i < rhs_len; ++i) rhs_perm[i] = (*rhs_io)[g->perm[i]];
      free(*rhs_io);
      *rhs_io = NULL;
    }
    for (int dest = 1; dest < size; ++dest) {
      CPLM_Mat_CSR_t B = CPLM_MatCSRNULL();
      pa_row_panel(&P, g->rowPos[dest], g->rowPos[dest + 1], &B);
      send_panel(&B, dest, comm);
      CPLM_MatCSRFree(&B);
    }
    pa_row_panel(&P, g->rowPos[0], g->rowPos[1], &g->A);
    CPLM_MatCSRFree(&P);
  } else {
    recv_panel(&g->A, 0, comm);
    g->rowPos = (int*)pa_xmalloc(sizeof(int) * ((size_t)size + 1));
    g->nrowPos = size + 1;
  }
  MPI_Bcast(g->rowPos, size + 1, MPI_INT, 0, comm);
  if (rhs_io) {  /* ref: operator.c:241-252, plain scatter of the permuted right-hand side */
    int* cnt = (int*)pa_xmalloc(sizeof(int) * (size_t)size);
    int* dsp = (int*)pa_xmalloc(sizeof(int) * (size_t)size);
    for (int i = 0; i < size; ++i) { cnt[i] = g->rowPos[i + 1] - g->rowPos[i]; dsp[i] = g->rowPos[i]; }
    *rhs_io = (double*)pa_xmalloc(sizeof(double) * (size_t)cnt[rank]);
    MPI_Scatterv(rhs_perm, cnt, dsp, MPI_DOUBLE, *rhs_io, cnt[rank], MPI_DOUBLE, 0, comm);
    free(cnt); free(dsp); free(rhs_perm);
  }
  return finish_build();
}

int preAlps_OperatorBuild(const char* matrixFilename, MPI_Comm comm) {
  int rank;
  MPI_Comm_rank(comm, &rank);
  CPLM_Mat_CSR_t G = CPLM_MatCSRNULL();
  if (rank == 0) {
    const size_t L = strlen(matrixFilename);
    if (L < 3 || strcmp(matrixFilename + L - 3, "mtx") != 0)
      CPLM_Abort("Please Compile with PETSC to read other matrix file type");  /* ref: operator.c:65 */
    pa_load_mtx(matrixFilename, &G, 1);
    pa_sym_scale(&G);
  }
  return build_from_root(&G, comm, NULL, 0);
}

/* "%"-comment lines, optional "nrows ncols" header, one value per line (ref: cplm_v0_dvector.c:120-218) */
static void load_vector(const char* fn, double** v, int* n) {
  FILE* f = fopen(fn, "r");
  if (!f) CPLM_Abort("Impossible to open the file %s", fn);
  char line[512];
  int cap = 1024, cnt = 0, header_checked = 0;
  double* x = (double*)pa_xmalloc(sizeof(double) * (size_t)cap);
  while (fgets(line, sizeof line, f)) {
    if (line[0] == '%' || line[0] == '\n') continue;
    double a, b;
    const int k = sscanf(line, "%lf %lf", &a, &b);
    if (!header_checked) { header_checked = 1; if (k == 2) continue; }
    if (k < 1) continue;
    if (cnt == cap) { cap *= 2; x = (double*)realloc(x, sizeof(double) * (size_t)cap); }
    x[cnt++] = a;
  }
  fclose(f);
  *v = x; *n = cnt;
}

int preAlps_OperatorRHSBuild(const char* matrixFilename, const char* rhsFilename, double** rhs, MPI_Comm comm) {
  int rank;
  MPI_Comm_rank(comm, &rank);
  CPLM_Mat_CSR_t G = CPLM_MatCSRNULL();
  int n = 0;
  *rhs = NULL;
  if (rank == 0) {
    pa_load_mtx(matrixFilename, &G, 1);
    printf("Load of %s ...\n", rhsFilename);
    load_vector(rhsFilename, rhs, &n);
    if (n != G.info.m) CPLM_Abort("right-hand side has %d entries, matrix has %d rows", n, G.info.m);
    /* ref: operator.c:172-187: a_ij /= sqrt(r_i * r_j) with r = row max; the rhs is NOT scaled */
    double* R = (double*)pa_xcalloc((size_t)G.info.m, sizeof(double));
    for (int i = 0; i < G.info.m; ++i)
      for (int p = G.rowPtr[i]; p < G.rowPtr[i + 1]; ++p) { double a = fabs(G.val[p]); if (a > R[i]) R[i] = a; }
    for (int i = 0; i < G.info.m; ++i)
      for (int p = G.rowPtr[i]; p < G.rowPtr[i + 1]; ++p) G.val[p] /= sqrt(R[i] * R[G.colInd[p]]);
    free(R);
  }
  return build_from_root(&G, comm, rhs, n);
}

int preAlps_OperatorBuildNoPerm(CPLM_Mat_CSR_t* locA, int* idxRowBegin, int nbBlockPerProcs, MPI_Comm comm) {
  pa_state_t* g = &pa_g;
  if (nbBlockPerProcs != 1) CPLM_Abort("[OperatorBuildNoPerm] Each MPI process must have one (and only one) metis");
  if (g->built) preAlps_OperatorFree();
  int size, rank;
  MPI_Comm_size(comm, &size);
  MPI_Comm_rank(comm, &rank);
  g->comm = comm; g->nproc = size; g->rank = rank;
  g->S = size; g->s_lo = rank; g->s_hi = rank + 1;
  g->xport = size > 1 ? PA_XPORT_MPI : PA_XPORT_NONE;
  g->A.info = locA->info;
  const int m = locA->info.m, lnnz = locA->info.lnnz;
  g->A.rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)m + 1));
  g->A.colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)(lnnz > 0 ? lnnz : 1));
  g->A.val = (double*)pa_xmalloc(sizeof(double) * (size_t)(lnnz > 0 ? lnnz : 1));
  memcpy(g->A.rowPtr, locA->rowPtr, sizeof(int) * ((size_t)m + 1));
  memcpy(g->A.colInd, locA->colInd, sizeof(int) * (size_t)lnnz);
  memcpy(g->A.val, locA->val, sizeof(double) * (size_t)lnnz);
  g->rowPos = (int*)pa_xmalloc(sizeof(int) * ((size_t)size + 1));
  g->nrowPos = size + 1;
  memcpy(g->rowPos, idxRowBegin, sizeof(int) * ((size_t)size + 1));
  return finish_build();
}

/* ------------------------------------------------------------------ virtual-subdomain entry points */
static void set_single_or_nccl(int S, int s_lo, int s_hi) {
  pa_state_t* g = &pa_g;
  if (g->built) preAlps_OperatorFree();
  if (g->nccl_ready) g->xport = g->nproc > 1 ? PA_XPORT_NCCL : PA_XPORT_NONE;
  else { g->nproc = 1; g->rank = 0; g->xport = PA_XPORT_NONE; }
  if (S < 1 || s_lo < 0 || s_hi > S || s_lo >= s_hi) CPLM_Abort("bad subdomain range [%d,%d) of %d", s_lo, s_hi, S);
  g->S = S; g->s_lo = s_lo; g->s_hi = s_hi;
}

int preAlps_b200_OperatorBuildCSR(int M, const int* rowPtr, const int* colInd, const double* val, int S, int s_lo,
                                  int s_hi, int scale, const int* parts_in) {
  set_single_or_nccl(S, s_lo, s_hi);
  CPLM_Mat_CSR_t G = CPLM_MatCSRNULL();
  const int nnz = rowPtr[M];
  G.info.M = G.info.m = G.info.N = G.info.n = M;
  G.info.nnz = G.info.lnnz = nnz;
  G.info.blockSize = 1; G.info.format = FORMAT_CSR; G.info.structure = SYMMETRIC;
  G.rowPtr = (int*)pa_xmalloc(sizeof(int) * ((size_t)M + 1));
  G.colInd = (int*)pa_xmalloc(sizeof(int) * (size_t)nnz);
  G.val = (double*)pa_xmalloc(sizeof(double) * (size_t)nnz);
  memcpy(G.rowPtr, rowPtr, sizeof(int) * ((size_t)M + 1));
  memcpy(G.colInd, colInd, sizeof(int) * (size_t)nnz);
  memcpy(G.val, val, sizeof(double) * (size_t)nnz);
  if (pa_check_diag(&G)) CPLM_Abort("Diagonal is not set correctly");
  partition_global(&G, S, s_lo, s_hi, scale, parts_in);
  return finish_build();
}

int preAlps_b200_OperatorBuildFile(const char* mtx, int S, int s_lo, int s_hi) {
  set_single_or_nccl(S, s_lo, s_hi);
  CPLM_Mat_CSR_t G = CPLM_MatCSRNULL();
  pa_load_mtx(mtx, &G, 0);
  partition_global(&G, S, s_lo, s_hi, 1, NULL);
  return finish_build();
}

int preAlps_b200_OperatorBuildStencil(int kind, int N, int S, int s_lo, int s_hi) {
  set_single_or_nccl(S, s_lo, s_hi);
  const int timing = getenv("PREALPS_B200_TIMING") != NULL;
  double t0 = pa_wtime();
  CPLM_Mat_CSR_t G = CPLM_MatCSRNULL();
  pa_stencil_csr(kind, N, &G);
  if (timing) fprintf(stderr, "[prealps_b200] setup: %-28s %.3f s\n", "stencil generation", pa_wtime() - t0);
  partition_global(&G, S, s_lo, s_hi, 1, NULL);
  t0 = pa_wtime();
  const int rc = finish_build();
  if (timing) fprintf(stderr, "[prealps_b200] setup: %-28s %.3f s\n", "halo plan + device upload", pa_wtime() - t0);
  return rc;
}

/* ------------------------------------------------------------------ getters (ref: operator.c:353-393) */
int preAlps_OperatorGetA(CPLM_Mat_CSR_t* A) { if (!A) CPLM_Abort(" wrong test 'A != NULL'"); *A = pa_g.A; return 0; }
int preAlps_OperatorGetSizes(int* M, int* m) { *M = pa_g.M; *m = pa_g.m; return 0; }
int preAlps_OperatorGetRowPosPtr(int** rowPos, int* n) { *n = pa_g.nrowPos; *rowPos = pa_g.rowPos; return rowPos == NULL; }
int preAlps_OperatorGetColPosPtr(int** colPos, int* n) { *n = pa_g.ncolPos; *colPos = pa_g.colPos; return colPos == NULL; }
int preAlps_OperatorGetDepPtr(int** dep, int* n) { *n = pa_g.ndep; *dep = pa_g.dep; return dep == NULL; }
int preAlps_b200_GetPartition(int* S, int* s_lo, int* s_hi) { *S = pa_g.S; *s_lo = pa_g.s_lo; *s_hi = pa_g.s_hi; return 0; }
int preAlps_b200_GetPerm(int** perm, int* n) { *perm = pa_g.perm; *n = pa_g.nperm; return pa_g.perm == NULL; }
int preAlps_b200_GetHalo(int** halo, int* n) { *halo = pa_g.halo; *n = pa_g.nhalo; return 0; }
int preAlps_b200_GetHaloPlan(int* nnbr, int** nbr, int** send_ptr, int** send_idx, int** recv_ptr) {
  *nnbr = pa_g.nnbr; *nbr = pa_g.nbr; *send_ptr = pa_g.send_ptr; *send_idx = pa_g.send_idx; *recv_ptr = pa_g.recv_ptr;
  return 0;
}

void preAlps_OperatorPrint(int rank) {
  if (rank != 0) return;
  printf("rowPos:"); for (int i = 0; i < pa_g.nrowPos; ++i) printf(" %d", pa_g.rowPos[i]);
  printf("\ndep:"); for (int i = 0; i < pa_g.ndep; ++i) printf(" %d", pa_g.dep[i]);
  printf("\nA: M=%d N=%d nnz=%d m=%d n=%d lnnz=%d | halo rows %d, neighbours %d\n", pa_g.A.info.M, pa_g.A.info.N,
         pa_g.A.info.nnz, pa_g.A.info.m, pa_g.A.info.n, pa_g.A.info.lnnz, pa_g.nhalo, pa_g.nnbr);
}

/* ------------------------------------------------------------------ staging of host-resident blocks */
void pa_ensure_stage(size_t doubles) {
  pa_state_t* g = &pa_g;
  if (g->stage_cap >= doubles) return;
  if (g->d_stage_in) pcu_free(g->ctx, g->d_stage_in);
  if (g->d_stage_out) pcu_free(g->ctx, g->d_stage_out);
  g->d_stage_in = (double*)pcu_malloc(g->ctx, sizeof(double) * doubles);
  g->d_stage_out = (double*)pcu_malloc(g->ctx, sizeof(double) * doubles);
  if (!g->d_stage_in || !g->d_stage_out) CPLM_Abort("device allocation failed: %s", pcu_last_error());
  g->stage_cap = doubles;
}

/* host block (either storage) -> device row-major m x n (ld = n) */
double* pa_stage_in(const CPLM_Mat_Dense_t* X) {
  const int m = X->info.m, n = X->info.n;
  pa_ensure_stage((size_t)m * n + 8);
  double* h = (double*)pa_xmalloc(sizeof(double) * (size_t)m * n);
  if (X->info.stor_type == COL_MAJOR)
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) h[(size_t)i * n + j] = X->val[(size_t)j * X->info.lda + i];
  else
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) h[(size_t)i * n + j] = X->val[(size_t)i * X->info.lda + j];
  pa_cuda_check(pcu_h2d(pa_g.ctx, pa_g.d_stage_in, h, sizeof(double) * (size_t)m * n), "pcu_h2d");
  free(h);
  return pa_g.d_stage_in;
}

void pa_stage_out(CPLM_Mat_Dense_t* Y) {
  const int m = Y->info.m, n = Y->info.n;
  double* h = (double*)pa_xmalloc(sizeof(double) * (size_t)m * n);
  pa_cuda_check(pcu_d2h(pa_g.ctx, h, pa_g.d_stage_out, sizeof(double) * (size_t)m * n), "pcu_d2h");
  if (Y->info.stor_type == COL_MAJOR)
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) Y->val[(size_t)j * Y->info.lda + i] = h[(size_t)i * n + j];
  else
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) Y->val[(size_t)i * Y->info.lda + j] = h[(size_t)i * n + j];
  free(h);
}

/* ------------------------------------------------------------------ the product */
static void halo_exchange(const double* X, int ldx, int t) {
  pa_state_t* g = &pa_g;
  if (g->nproc <= 1) return;
  if (g->xport == PA_XPORT_NCCL) {
    pa_cuda_check(pcu_spmm_halo_exchange(g->spmm, X, ldx, t), "pcu_spmm_halo_exchange");
    return;
  }
  /* MPI transport: packed boundary rows go through pinned host buffers */
  const int ns = g->send_ptr[g->nnbr], nr = g->recv_ptr[g->nnbr];
  if (g->h_cap_t < t) {
    if (g->h_send) pcu_host_free(g->h_send);
    if (g->h_recv) pcu_host_free(g->h_recv);
    g->h_send = (double*)pcu_host_alloc(sizeof(double) * (size_t)(ns > 0 ? ns : 1) * t);
    g->h_recv = (double*)pcu_host_alloc(sizeof(double) * (size_t)(nr > 0 ? nr : 1) * t);
    if (!g->h_send || !g->h_recv) CPLM_Abort("pinned host allocation failed");
    g->h_cap_t = t;
  }
  double* packed = NULL;
  int nrows = 0;
  pa_cuda_check(pcu_spmm_halo_pack(g->spmm, X, ldx, t, &packed, &nrows), "pcu_spmm_halo_pack");
  if (ns) pa_cuda_check(pcu_d2h(g->ctx, g->h_send, packed, sizeof(double) * (size_t)ns * t), "pcu_d2h");
  MPI_Request* rq = (MPI_Request*)pa_xmalloc(sizeof(MPI_Request) * (size_t)(g->nnbr > 0 ? g->nnbr : 1));
  for (int k = 0; k < g->nnbr; ++k)
    MPI_Isend(g->h_send + (size_t)g->send_ptr[k] * t, (g->send_ptr[k + 1] - g->send_ptr[k]) * t, MPI_DOUBLE, g->nbr[k], 78,
              g->comm, &rq[k]);
  for (int k = 0; k < g->nnbr; ++k)
    MPI_Recv(g->h_recv + (size_t)g->recv_ptr[k] * t, (g->recv_ptr[k + 1] - g->recv_ptr[k]) * t, MPI_DOUBLE, g->nbr[k], 78,
             g->comm, MPI_STATUS_IGNORE);
  MPI_Waitall(g->nnbr, rq, MPI_STATUSES_IGNORE);
  free(rq);
  double* H = pcu_spmm_halo_buffer(g->spmm, t);
  if (!H) CPLM_Abort("halo buffer: %s", pcu_last_error());
  if (nr) pa_cuda_check(pcu_h2d(g->ctx, H, g->h_recv, sizeof(double) * (size_t)nr * t), "pcu_h2d");
}

int preAlps_BlockOperator(CPLM_Mat_Dense_t* X, CPLM_Mat_Dense_t* AX) {
  pa_state_t* g = &pa_g;
  if (!g->built) CPLM_Abort("preAlps_BlockOperator called before preAlps_OperatorBuild");
  if (!X || !X->val) CPLM_Abort(" wrong test 'X->val != NULL'");
  const int t = X->info.n;
  if (X->info.m != g->m) CPLM_Abort("block has %d rows, operator has %d", X->info.m, g->m);
  const int dev_in = pa_is_device_block(X);
  const double* x; int ldx;
  if (dev_in) {
    if (X->info.stor_type != ROW_MAJOR) CPLM_Abort("device blocks must be ROW_MAJOR");
    x = X->val; ldx = X->info.lda;
  } else {
    x = pa_stage_in(X); ldx = t;
  }
  if (AX->val == NULL) {  /* ref: cplm_v0_matmult_v2.c:160-171 allocates the result on demand (host) */
    CPLM_MatDenseSetInfo(AX, g->M, X->info.N, g->m, t, X->info.stor_type);
    AX->val = (double*)pa_xcalloc((size_t)g->m * t, sizeof(double));
  }
  const int dev_out = pa_is_device_block(AX);
  double* y; int ldy;
  if (dev_out) { y = AX->val; ldy = AX->info.lda; }
  else { pa_ensure_stage((size_t)g->m * t + 8); y = g->d_stage_out; ldy = t; if (!dev_in) x = g->d_stage_in; }
  if (g->nproc > 1 && g->xport == PA_XPORT_NCCL) {
    /* exchange + product in one call: the library overlaps the exchange with the local part of the product */
    pa_cuda_check(pcu_spmm_apply_exchange(g->spmm, x, ldx, y, ldy, t), "pcu_spmm_apply_exchange");
  } else {
    halo_exchange(x, ldx, t);
    pa_cuda_check(pcu_spmm_apply(g->spmm, x, ldx, y, ldy, t), "pcu_spmm_apply");
  }
  if (!dev_out) pa_stage_out(AX);
  return 0;
}

double preAlps_b200_Stat(const char* name) {
  pa_state_t* g = &pa_g;
  if (!strcmp(name, "spmm_bytes_t8")) return g->spmm ? pcu_spmm_bytes(g->spmm, 8) : -1;
  if (!strncmp(name, "spmm_bytes_t", 12)) return g->spmm ? pcu_spmm_bytes(g->spmm, atoi(name + 12)) : -1;
  if (!strncmp(name, "bj_bytes_t", 10)) return g->bj ? pcu_bj_bytes(g->bj, atoi(name + 10)) : -1;
  if (!strncmp(name, "bj_stored_bytes_t", 17)) return g->bj ? pcu_bj_stored_bytes(g->bj, atoi(name + 17)) : -1;
  if (!strcmp(name, "bj_nnz_exact")) return g->bj ? pcu_bj_stat(g->bj, 0) : -1;
  if (!strcmp(name, "bj_nnz_stored")) return g->bj ? pcu_bj_stat(g->bj, 1) : -1;
  if (!strcmp(name, "bj_supernodes")) return g->bj ? pcu_bj_stat(g->bj, 2) : -1;
  if (!strcmp(name, "bj_levels")) return g->bj ? pcu_bj_stat(g->bj, 3) : -1;
  if (!strcmp(name, "bj_factor_flops")) return g->bj ? pcu_bj_stat(g->bj, 5) : -1;
  if (!strcmp(name, "bj_factor_s")) return g->bj ? pcu_bj_stat(g->bj, 6) : -1;
  if (!strcmp(name, "bj_analysis_s")) return g->bj ? pcu_bj_stat(g->bj, 7) : -1;
  if (!strcmp(name, "bj_launches")) return g->bj ? pcu_bj_stat(g->bj, 8) : -1;
  if (!strcmp(name, "nhalo")) return g->nhalo;
  if (!strcmp(name, "nnbr")) return g->nnbr;
  if (!strcmp(name, "launches")) return g->ctx ? (double)pcu_launch_count(g->ctx) : 0;
  return -1;
}
