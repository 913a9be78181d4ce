/*
 * oracle/ref_glue.c -- TEST INFRASTRUCTURE ONLY.
 * utils/operator.c of the reference links against two helpers that live in
 * utils/preAlps_utils.c (a file that drags in ParMETIS and is otherwise off the
 * ECG path).  Only preAlps_abort is needed; this is its behaviour
 * (/root/reference/utils/preAlps_utils.c:34-50: print the message, exit(1)).
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <mpi.h>

void preAlps_abort(char* s, ...) {
  va_list ap;
  va_start(ap, s);
  fprintf(stderr, "===================\nAborting ... ");
  vfprintf(stderr, s, ap);
  fprintf(stderr, "\n===================\n");
  va_end(ap);
  MPI_Abort(MPI_COMM_WORLD, 1);
  exit(1);
}
