/*
 * oracle/dump_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference library (built by oracle/Makefile) through
 * the same call sequence as /root/reference/examples/test_ecg_prealps_op.c:
 * 158-223 (OperatorBuild, BlockJacobiCreate, srand(0) rhs with the rhs[0]
 * quirk of :183, the RCI loop), and additionally writes everything the parity
 * tests need as raw little-endian arrays into a dump directory:
 *
 *   r<rank>_rowPos.i32 r<rank>_colPos.i32 r<rank>_dep.i32
 *   r<rank>_A_rowPtr.i32 r<rank>_A_colInd.i32 r<rank>_A_val.f64     (row panel, global columns)
 *   r<rank>_D_rowPtr.i32 r<rank>_D_colInd.i32 r<rank>_D_val.f64     (upper-triangular diag block)
 *   r<rank>_rhs.f64 r<rank>_sol.f64
 *   r<rank>_AP1.f64 r<rank>_P1.f64   (first block-Jacobi apply / first SpMM, column-major m x t)
 *   perm.i32 posB.i32 (rank 0: the METIS k-way permutation, perm[new]=old)
 *   summary.json (rank 0: iterations, residual history, true residual, timings)
 *
 * Usage: MPISHIM_NP=<S> ecg_dump_ref -m A.mtx -e t [-o 0|1|2] [-r 0|1] [-t tol] [-i maxit] -d dumpdir
 *        (-o 2 selects ORTHODIR_FUSED with the loop of test_ecg_bench_fused.c:251-259)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <getopt.h>

#include <mpi.h>
#include <mkl.h>

#include "operator.h"
#include "block_jacobi.h"
#include "ecg.h"
#include <cplm_v0_matcsr.h>

static const char* dumpdir = NULL;
static int g_rank = 0;

static void dump(const char* name, const char* ext, const void* p, size_t bytes, int per_rank) {
  if (!dumpdir) return;
  char path[4096];
  if (per_rank) snprintf(path, sizeof path, "%s/r%d_%s.%s", dumpdir, g_rank, name, ext);
  else snprintf(path, sizeof path, "%s/%s.%s", dumpdir, name, ext);
  FILE* f = fopen(path, "wb");
  if (!f) { perror(path); MPI_Abort(MPI_COMM_WORLD, 9); }
  if (bytes) fwrite(p, 1, bytes, f);
  fclose(f);
}

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  double tol = 1e-5;
  int maxIter = 1000, enlFac = 1, ortho = 0, bs_red = 0, c, nodump_big = 0;
  const char* matrixFilename = NULL;
  while ((c = getopt(argc, argv, "e:i:m:o:r:t:d:q")) != -1) switch (c) {
    case 'e': enlFac = atoi(optarg); break;
    case 'i': maxIter = atoi(optarg); break;
    case 'm': matrixFilename = optarg; break;
    case 'o': ortho = atoi(optarg); break;
    case 'r': bs_red = atoi(optarg); break;
    case 't': tol = atof(optarg); break;
    case 'd': dumpdir = optarg; break;
    case 'q': nodump_big = 1; break; /* only summary.json + small arrays */
    default: MPI_Abort(MPI_COMM_WORLD, 2);
  }
  int rank, size;
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  g_rank = rank;
  MKL_Set_Num_Threads(1);

  /* rank 0: replay load -> scale -> k-way ordering to expose perm/posB (operator.c:56-80) */
  if (rank == 0 && dumpdir) {
    CPLM_Mat_CSR_t M0 = CPLM_MatCSRNULL();
    CPLM_LoadMatrixMarket(matrixFilename, &M0);
    double* R = malloc(M0.info.m * sizeof(double));
    double* C = malloc(M0.info.m * sizeof(double));
    CPLM_MatCSRSymRACScaling(&M0, R, C);
    if (!nodump_big) {
      dump("S_rowPtr", "i32", M0.rowPtr, (size_t)(M0.info.m + 1) * 4, 0);
      dump("S_colInd", "i32", M0.colInd, (size_t)M0.info.lnnz * 4, 0);
      dump("S_val", "f64", M0.val, (size_t)M0.info.lnnz * 8, 0);
    }
    free(R); free(C);
    CPLM_IVector_t posB = CPLM_IVectorNULL(), perm = CPLM_IVectorNULL();
    CPLM_metisKwayOrdering(&M0, &perm, size, &posB);
    if (perm.val) dump("perm", "i32", perm.val, (size_t)perm.nval * 4, 0);
    dump("posB", "i32", posB.val, (size_t)posB.nval * 4, 0);
    CPLM_IVectorFree(&posB); CPLM_IVectorFree(&perm); CPLM_MatCSRFree(&M0);
  }

  double t_setup = MPI_Wtime();
  CPLM_Mat_CSR_t A = CPLM_MatCSRNULL();
  int M, m, *rowPos = NULL, *colPos = NULL, *dep = NULL, sizeRowPos, sizeColPos, sizeDep;
  preAlps_OperatorBuild(matrixFilename, MPI_COMM_WORLD);
  preAlps_OperatorGetA(&A);
  preAlps_OperatorGetSizes(&M, &m);
  preAlps_OperatorGetRowPosPtr(&rowPos, &sizeRowPos);
  preAlps_OperatorGetColPosPtr(&colPos, &sizeColPos);
  preAlps_OperatorGetDepPtr(&dep, &sizeDep);
  double t_build = MPI_Wtime() - t_setup;

  dump("rowPos", "i32", rowPos, (size_t)sizeRowPos * 4, 1);
  dump("dep", "i32", dep, (size_t)sizeDep * 4, 1);
  if (!nodump_big) {
    dump("colPos", "i32", colPos, (size_t)sizeColPos * 4, 1);
    dump("A_rowPtr", "i32", A.rowPtr, (size_t)(A.info.m + 1) * 4, 1);
    dump("A_colInd", "i32", A.colInd, (size_t)A.info.lnnz * 4, 1);
    dump("A_val", "f64", A.val, (size_t)A.info.lnnz * 8, 1);
    /* the diag block exactly as block_jacobi.c:48 extracts it */
    CPLM_Mat_CSR_t D = CPLM_MatCSRNULL();
    CPLM_IVector_t rp = CPLM_IVectorNULL(), cp = CPLM_IVectorNULL();
    CPLM_IVectorCreateFromPtr(&rp, sizeRowPos, rowPos);
    CPLM_IVectorCreateFromPtr(&cp, sizeColPos, colPos);
    CPLM_MatCSRGetDiagBlock(&A, &D, &rp, &cp, SYMMETRIC);
    dump("D_rowPtr", "i32", D.rowPtr, (size_t)(D.info.m + 1) * 4, 1);
    dump("D_colInd", "i32", D.colInd, (size_t)D.info.lnnz * 4, 1);
    dump("D_val", "f64", D.val, (size_t)D.info.lnnz * 8, 1);
    CPLM_MatCSRFree(&D);
  }

  double t_fac = MPI_Wtime();
  preAlps_BlockJacobiCreate(&A, rowPos, sizeRowPos, colPos, sizeColPos);
  t_fac = MPI_Wtime() - t_fac;

  /* rhs exactly as test_ecg_prealps_op.c:172-184 */
  double* rhs = (double*)malloc(m * sizeof(double));
  srand(0);
  double normb = 0.0;
  for (int i = 0; i < m; ++i) {
    rhs[i] = ((double)rand() / (double)RAND_MAX);
    normb += pow(rhs[i], 2);
  }
  MPI_Allreduce(MPI_IN_PLACE, &normb, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  normb = sqrt(normb);
  for (int i = 1; i < m; ++i) rhs[i] /= normb;
  if (!nodump_big) dump("rhs", "f64", rhs, (size_t)m * 8, 1);

  preAlps_ECG_t ecg;
  ecg.comm = MPI_COMM_WORLD;
  ecg.globPbSize = M;
  ecg.locPbSize = m;
  ecg.maxIter = maxIter;
  ecg.enlFac = enlFac;
  ecg.tol = tol;
  ecg.ortho_alg = (ortho == 0 ? ORTHODIR : (ortho == 1 ? ORTHOMIN : ORTHODIR_FUSED));
  ecg.bs_red = (bs_red == 0 ? NO_BS_RED : ADAPT_BS);
  int rci_request = 0, stop = 0;
  double* sol = (double*)malloc(m * sizeof(double));
  double* hist = (double*)malloc((size_t)(maxIter + 2) * sizeof(double));
  int* bshist = (int*)malloc((size_t)(maxIter + 2) * sizeof(int));
  int nhist = 0;
  double t_op = 0, t_prec = 0, t0, t_solve = MPI_Wtime();

  preAlps_ECGInitialize(&ecg, rhs, &rci_request);
  t0 = MPI_Wtime(); preAlps_BlockJacobiApply(ecg.R, ecg.P); t_prec += MPI_Wtime() - t0;
  if (!nodump_big) dump("P1", "f64", ecg.P->val, (size_t)m * enlFac * 8, 1);
  t0 = MPI_Wtime(); preAlps_BlockOperator(ecg.P, ecg.AP); t_op += MPI_Wtime() - t0;
  if (!nodump_big) dump("AP1", "f64", ecg.AP->val, (size_t)m * enlFac * 8, 1);

  if (ecg.ortho_alg != ORTHODIR_FUSED) {
    while (stop != 1) {
      preAlps_ECGIterate(&ecg, &rci_request);
      if (rci_request == 0) {
        t0 = MPI_Wtime(); preAlps_BlockOperator(ecg.P, ecg.AP); t_op += MPI_Wtime() - t0;
      } else if (rci_request == 1) {
        preAlps_ECGStoppingCriterion(&ecg, &stop);
        hist[nhist] = ecg.res; bshist[nhist++] = ecg.bs;
        if (stop == 1) break;
        t0 = MPI_Wtime();
        if (ecg.ortho_alg == ORTHOMIN) preAlps_BlockJacobiApply(ecg.R, ecg.Z);
        else preAlps_BlockJacobiApply(ecg.AP, ecg.Z);
        t_prec += MPI_Wtime() - t0;
      }
    }
  } else {
    /* test_ecg_bench_fused.c:245-259 */
    t0 = MPI_Wtime(); preAlps_BlockJacobiApply(ecg.AP, ecg.Z); t_prec += MPI_Wtime() - t0;
    while (rci_request != 1) {
      preAlps_ECGIterate(&ecg, &rci_request);
      hist[nhist] = ecg.res; bshist[nhist++] = ecg.bs;
      if (rci_request == 1) break;
      t0 = MPI_Wtime(); preAlps_BlockOperator(ecg.P, ecg.AP); t_op += MPI_Wtime() - t0;
      t0 = MPI_Wtime(); preAlps_BlockJacobiApply(ecg.AP, ecg.Z); t_prec += MPI_Wtime() - t0;
    }
  }
  int iters = ecg.iter, bs_final = ecg.bs;
  double res_final = ecg.res, ecg_tot = ecg.tot_t, ecg_comm = ecg.comm_t;
  double normb_ecg = ecg.normb;
  preAlps_ECGFinalize(&ecg, sol);
  t_solve = MPI_Wtime() - t_solve;
  if (!nodump_big) dump("sol", "f64", sol, (size_t)m * 8, 1);

  /* true residual ||b - A x|| / ||b|| with the reference's own distributed SpMM */
  CPLM_Mat_Dense_t xs = CPLM_MatDenseNULL(), ax = CPLM_MatDenseNULL();
  CPLM_MatDenseSetInfo(&xs, M, 1, m, 1, COL_MAJOR);
  CPLM_MatDenseSetInfo(&ax, M, 1, m, 1, COL_MAJOR);
  xs.val = sol;
  ax.val = (double*)calloc(m, sizeof(double));
  preAlps_BlockOperator(&xs, &ax);
  double rr[2] = {0, 0};
  for (int i = 0; i < m; ++i) { double d = rhs[i] - ax.val[i]; rr[0] += d * d; rr[1] += rhs[i] * rhs[i]; }
  MPI_Allreduce(MPI_IN_PLACE, rr, 2, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  double tmax[5] = {t_solve, t_op, t_prec, t_fac, t_build};
  MPI_Allreduce(MPI_IN_PLACE, tmax, 5, MPI_DOUBLE, MPI_MAX, MPI_COMM_WORLD);

  if (rank == 0) {
    char path[4096];
    FILE* f = stdout;
    if (dumpdir) { snprintf(path, sizeof path, "%s/summary.json", dumpdir); f = fopen(path, "w"); }
    fprintf(f, "{\"matrix\": \"%s\", \"np\": %d, \"M\": %d, \"enlFac\": %d, \"ortho_alg\": %d, \"bs_red\": %d,\n",
            matrixFilename, size, M, enlFac, ortho, bs_red);
    fprintf(f, " \"tol\": %.17g, \"maxIter\": %d, \"iter\": %d, \"res\": %.17g, \"bs\": %d, \"normb\": %.17g,\n",
            tol, maxIter, iters, res_final, bs_final, normb_ecg);
    fprintf(f, " \"true_relres\": %.17g,\n", sqrt(rr[0]) / sqrt(rr[1]));
    fprintf(f, " \"t_solve\": %.6e, \"t_op\": %.6e, \"t_prec\": %.6e, \"t_factor\": %.6e, \"t_build\": %.6e, \"ecg_tot_t\": %.6e, \"ecg_comm_t\": %.6e,\n",
            tmax[0], tmax[1], tmax[2], tmax[3], tmax[4], ecg_tot, ecg_comm);
    fprintf(f, " \"res_hist\": [");
    for (int i = 0; i < nhist; ++i) fprintf(f, "%s%.17g", i ? ", " : "", hist[i]);
    fprintf(f, "],\n \"bs_hist\": [");
    for (int i = 0; i < nhist; ++i) fprintf(f, "%s%d", i ? ", " : "", bshist[i]);
    fprintf(f, "]}\n");
    if (dumpdir) fclose(f);
    printf("[ecg_dump_ref] np=%d M=%d t=%d iter=%d res=%.6e true_relres=%.6e t_solve=%.3fs (op %.3f, prec %.3f) factor %.3fs\n",
           size, M, enlFac, iters, res_final, sqrt(rr[0]) / sqrt(rr[1]), tmax[0], tmax[1], tmax[2], tmax[3]);
  }
  free(rhs); free(sol); free(hist); free(bshist); free(ax.val);
  preAlps_OperatorFree();
  MPI_Finalize();
  return 0;
}
