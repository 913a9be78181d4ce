/* tests/alt_build_dump.c -- the two alternative operator builders of the reference API, as real processes over the MPI
 * shim.  Uses only operator.h, so the same file is compiled against this library (CPU: PREALPS_B200_HOST_ONLY=1) and,
 * by tests/golden/make_golden_altbuild.py, against the unmodified reference (oracle/_ref/libprealps_ref.so).
 *   MPISHIM_NP=<S> ./alt_build_dump rhs    A.mtx rhs.txt outdir   preAlps_OperatorRHSBuild  (ref: operator.c:136-268)
 *   MPISHIM_NP=<S> ./alt_build_dump noperm outdir                 preAlps_OperatorBuildNoPerm on the panels the first
 *                                                                 mode left in outdir    (ref: operator.c:271-308) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "operator.h"

static void dump(const char* dir, int rank, const char* name, const char* ext, const void* p, size_t bytes) {
  char path[4096];
  snprintf(path, sizeof path, "%s/r%d_%s.%s", dir, rank, name, ext);
  FILE* f = fopen(path, "wb");
  if (bytes > 0) fwrite(p, 1, bytes, f);
  fclose(f);
}

static void* slurp(const char* dir, int rank, const char* name, const char* ext, size_t* bytes) {
  char path[4096];
  snprintf(path, sizeof path, "%s/r%d_%s.%s", dir, rank, name, ext);
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(1); }
  fseek(f, 0, SEEK_END);
  *bytes = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  void* p = malloc(*bytes ? *bytes : 1);
  if (fread(p, 1, *bytes, f) != *bytes) exit(1);
  fclose(f);
  return p;
}

static void dump_maps(const char* dir, int rank, const char* tag) {
  int *rowPos, *colPos, *dep, n1, n2, n3;
  char name[64];
  preAlps_OperatorGetRowPosPtr(&rowPos, &n1);
  preAlps_OperatorGetColPosPtr(&colPos, &n2);
  preAlps_OperatorGetDepPtr(&dep, &n3);
  snprintf(name, sizeof name, "%s_rowPos", tag); dump(dir, rank, name, "i32", rowPos, sizeof(int) * (size_t)n1);
  snprintf(name, sizeof name, "%s_colPos", tag); dump(dir, rank, name, "i32", colPos, sizeof(int) * (size_t)n2);
  snprintf(name, sizeof name, "%s_dep", tag); dump(dir, rank, name, "i32", dep, sizeof(int) * (size_t)n3);
}

int main(int argc, char** argv) {
  MPI_Init(&argc, &argv);
  int rank, size;
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  MPI_Comm_size(MPI_COMM_WORLD, &size);
  if (argc >= 5 && !strcmp(argv[1], "rhs")) {
    const char* dir = argv[4];
    double* rhs = NULL;
    preAlps_OperatorRHSBuild(argv[2], argv[3], &rhs, MPI_COMM_WORLD);
    CPLM_Mat_CSR_t A;
    preAlps_OperatorGetA(&A);
    int M, m;
    preAlps_OperatorGetSizes(&M, &m);
    dump(dir, rank, "A_rowPtr", "i32", A.rowPtr, sizeof(int) * (size_t)(m + 1));
    dump(dir, rank, "A_colInd", "i32", A.colInd, sizeof(int) * (size_t)A.rowPtr[m]);
    dump(dir, rank, "A_val", "f64", A.val, sizeof(double) * (size_t)A.rowPtr[m]);
    dump(dir, rank, "rhs", "f64", rhs, sizeof(double) * (size_t)m);
    dump_maps(dir, rank, "rhs");
  } else if (argc >= 3 && !strcmp(argv[1], "noperm")) {
    const char* dir = argv[2];
    size_t b1, b2, b3, b4;
    int* rp = (int*)slurp(dir, rank, "A_rowPtr", "i32", &b1);
    int* ci = (int*)slurp(dir, rank, "A_colInd", "i32", &b2);
    double* v = (double*)slurp(dir, rank, "A_val", "f64", &b3);
    int* rowPos = (int*)slurp(dir, rank, "rhs_rowPos", "i32", &b4);
    const int m = (int)(b1 / sizeof(int)) - 1;
    CPLM_Mat_CSR_t loc = CPLM_MatCSRNULL();
    loc.info.M = rowPos[size]; loc.info.N = rowPos[size]; loc.info.m = m; loc.info.n = rowPos[size];
    loc.info.nnz = rp[m]; loc.info.lnnz = rp[m];
    loc.info.blockSize = 1; loc.info.format = FORMAT_CSR; loc.info.structure = UNSYMMETRIC;
    loc.rowPtr = rp; loc.colInd = ci; loc.val = v;
    preAlps_OperatorBuildNoPerm(&loc, rowPos, 1, MPI_COMM_WORLD);
    CPLM_Mat_CSR_t A;
    preAlps_OperatorGetA(&A);
    int M, mm;
    preAlps_OperatorGetSizes(&M, &mm);
    int sz[2] = {M, mm};
    dump(dir, rank, "noperm_sizes", "i32", sz, sizeof sz);
    dump(dir, rank, "noperm_A_colInd", "i32", A.colInd, sizeof(int) * (size_t)A.rowPtr[mm]);
    dump_maps(dir, rank, "noperm");
  } else {
    if (rank == 0) fprintf(stderr, "usage: alt_build_dump rhs A.mtx rhs.txt outdir | noperm outdir\n");
    MPI_Finalize();
    return 2;
  }
  MPI_Finalize();
  return 0;
}
