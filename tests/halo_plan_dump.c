/* tests/halo_plan_dump.c -- multi-process CPU check of the host layer (no GPU: PREALPS_B200_HOST_ONLY=1).
 * MPISHIM_NP=<S> ./halo_plan_dump A.mtx outdir : every rank builds the operator exactly like the reference
 * driver does (preAlps_OperatorBuild) and dumps its panel, maps and halo plan as raw int32 files. */
#include <stdio.h>
#include <stdlib.h>
#include <mpi.h>
#include "operator.h"
#include "prealps_b200.h"

static void dump(const char* dir, int rank, const char* name, const int* p, int n) {
  char path[4096];
  snprintf(path, sizeof path, "%s/r%d_%s.i32", dir, rank, name);
  FILE* f = fopen(path, "wb");
  if (n > 0) fwrite(p, sizeof(int), (size_t)n, f);
  fclose(f);
}

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  preAlps_OperatorBuild(argv[1], MPI_COMM_WORLD);
  CPLM_Mat_CSR_t A;
  preAlps_OperatorGetA(&A);
  int *rowPos, *colPos, *dep, n1, n2, n3, *halo, nh, nnbr, *nbr, *sp, *si, *rp;
  preAlps_OperatorGetRowPosPtr(&rowPos, &n1);
  preAlps_OperatorGetColPosPtr(&colPos, &n2);
  preAlps_OperatorGetDepPtr(&dep, &n3);
  preAlps_b200_GetHalo(&halo, &nh);
  preAlps_b200_GetHaloPlan(&nnbr, &nbr, &sp, &si, &rp);
  dump(argv[2], rank, "rowPos", rowPos, n1);
  dump(argv[2], rank, "colPos", colPos, n2);
  dump(argv[2], rank, "dep", dep, n3);
  dump(argv[2], rank, "A_rowPtr", A.rowPtr, A.info.m + 1);
  dump(argv[2], rank, "A_colInd", A.colInd, A.info.lnnz);
  dump(argv[2], rank, "halo", halo, nh);
  dump(argv[2], rank, "nbr", nbr, nnbr);
  dump(argv[2], rank, "send_ptr", sp, nnbr + 1);
  dump(argv[2], rank, "send_idx", si, nnbr ? sp[nnbr] : 0);
  dump(argv[2], rank, "recv_ptr", rp, nnbr + 1);
  preAlps_OperatorFree();
  MPI_Finalize();
  return 0;
}
