/*
 * bench_kernels.c -- the preAlps half of the reference's two kernel benchmarks, without PETSc:
 *   SpMM          ref: examples/test_bench_spmm.c:194-233    (preAlps_BlockOperator)
 *   block Jacobi  ref: examples/test_bench_bjacobi.c:217-255 (preAlps_BlockJacobiApply)
 * Same protocol: the operator is built from a .mtx file, the input block is m x 28 uniforms drawn after
 * srand(0), the block is presented COL_MAJOR with 1, 2, 4, ..., 28 columns, one untimed call and then
 * nrepet = 10 timed calls separated by MPI_Barrier, time and time per right-hand side printed by rank 0.
 * The reference passes HOST blocks, so those timings contain the host <-> device copies of the whole
 * block; the second table times the same kernels on blocks that live in HBM (preAlps_b200_BenchKernel,
 * CUDA events, L2 flushed between repetitions) and gives the algorithmic bytes and GB/s of DESIGN.md 4.
 *
 *   mpirun -n S ./bench_kernels -m A.mtx [-k spmm|bjacobi|both]       (here: MPISHIM_NP=S ./bench_kernels ...)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mpi.h>

#include "operator.h"
#include "block_jacobi.h"
#include "ecg.h"
#include "prealps_b200.h"

#define MAXCOL 28
#define NREPET 10

static int ncols_of(int k) { return k == 0 ? 1 : 2 * k; } /* 1, 2, 4, ..., 28 as in the reference */

static void host_sweep(int what, int M, int m, double* in, double* out, double* tsec) {
  for (int k = 0; k <= MAXCOL / 2; ++k) {
    const int t = ncols_of(k);
    CPLM_Mat_Dense_t X = CPLM_MatDenseNULL(), Y = CPLM_MatDenseNULL();
    CPLM_MatDenseSetInfo(&X, M, t, m, t, COL_MAJOR);
    CPLM_MatDenseSetInfo(&Y, M, t, m, t, COL_MAJOR);
    X.val = in;
    Y.val = out;
    for (int rep = -1; rep < NREPET; ++rep) {
      if (rep == 0) tsec[k] = MPI_Wtime();
      if (what == 0) preAlps_BlockOperator(&X, &Y);
      else preAlps_BlockJacobiApply(&X, &Y);
      MPI_Barrier(MPI_COMM_WORLD);
    }
    tsec[k] = MPI_Wtime() - tsec[k];
  }
}

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank, size;
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  const char* matrixFilename = NULL;
  const char* which = "both";
  for (int i = 1; i + 1 < argc; ++i) {
    if (!strcmp(argv[i], "-m") || !strcmp(argv[i], "--matrix")) matrixFilename = argv[i + 1];
    if (!strcmp(argv[i], "-k") || !strcmp(argv[i], "--kernel")) which = argv[i + 1];
  }
  if (!matrixFilename) {
    if (rank == 0) printf("USAGE\n\tmpirun -n nb_proc ./bench_kernels -m/--matrix file [-k spmm|bjacobi|both]\n");
    MPI_Finalize();
    return 1;
  }
  const int do_spmm = strcmp(which, "bjacobi") != 0, do_bj = strcmp(which, "spmm") != 0;
  if (rank == 0) printf("=== Parameters ===\n\tmatrix: %s\n\tnrepet: %d\n\tmaxCol: %d\n", matrixFilename, NREPET, MAXCOL);

  CPLM_Mat_CSR_t A = CPLM_MatCSRNULL();
  int M, m, sizeRowPos, sizeColPos;
  int *rowPos = NULL, *colPos = NULL;
  preAlps_OperatorBuild(matrixFilename, MPI_COMM_WORLD);
  preAlps_OperatorGetA(&A);
  preAlps_OperatorGetSizes(&M, &m);
  preAlps_OperatorGetRowPosPtr(&rowPos, &sizeRowPos);
  preAlps_OperatorGetColPosPtr(&colPos, &sizeColPos);
  if (do_bj) preAlps_BlockJacobiCreate(&A, rowPos, sizeRowPos, colPos, sizeColPos);
  if (rank == 0) printf("=== Matrix informations ===\n\tsize: %d\n\tnnz : %d\n", A.info.M, A.info.nnz);

  double* in = (double*)malloc(sizeof(double) * (size_t)m * MAXCOL);
  double* out = (double*)malloc(sizeof(double) * (size_t)m * MAXCOL);
  srand(0);
  for (size_t i = 0; i < (size_t)m * MAXCOL; ++i) in[i] = (double)rand() / (double)RAND_MAX;

  double tsec[MAXCOL / 2 + 1];
  for (int what = 0; what < 2; ++what) {
    if ((what == 0 && !do_spmm) || (what == 1 && !do_bj)) continue;
    host_sweep(what, M, m, in, out, tsec);
    if (rank == 0) {
      printf("=== ECG timings (host blocks, as the reference calls it) ===\n\trhs\ttime\t\ttime/rhs\n");
      for (int k = 0; k <= MAXCOL / 2; ++k)
        printf("%s\t%2d\t%e\t%e\n", k ? "" : (what ? "trsm" : "spmm"), ncols_of(k), tsec[k], tsec[k] / ncols_of(k));
    }
    /* the same kernel on blocks resident in HBM */
    if (rank == 0) printf("=== %s on device-resident blocks ===\n\trhs\tms/call\t\tMB/call\t\tGB/s\n", what ? "block Jacobi" : "SpMM");
    for (int k = 0; k <= MAXCOL / 2; ++k) {
      const int t = ncols_of(k);
      float ms = 0.f;
      char name[32];
      preAlps_b200_BenchKernel(what, t, NREPET, 1, &ms);
      snprintf(name, sizeof name, what ? "bj_bytes_t%d" : "spmm_bytes_t%d", t);
      const double bytes = preAlps_b200_Stat(name);
      if (rank == 0) printf("\t%2d\t%e\t%e\t%.1f\n", t, (double)ms, bytes / 1e6, bytes / ((double)ms * 1e-3) / 1e9);
      MPI_Barrier(MPI_COMM_WORLD);
    }
  }

  free(in);
  free(out);
  if (do_bj) preAlps_BlockJacobiFree();
  preAlps_OperatorFree();
  MPI_Finalize();
  return 0;
}
