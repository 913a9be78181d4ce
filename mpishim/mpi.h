/*
 * mpishim/mpi.h -- a single-node, fork-based subset of MPI-1/2.
 *
 * The container and the GPU boxes this project is built on ship no MPI.  The
 * preAlps driver (examples/test_ecg_prealps_op.c) is an MPI SPMD program, so
 * both the oracle build of the unmodified reference sources (oracle/Makefile)
 * and the unchanged-driver build of the B200 library need *some* mpi.h.  This
 * shim implements exactly the calls those two builds reference:
 *
 *   MPISHIM_NP=<n> ./prog args...
 *
 * MPI_Init() maps an anonymous shared arena and fork()s n-1 children; rank 0
 * is the original process.  Point-to-point is eager (a send copies into the
 * arena and returns), collectives are built on point-to-point in rank order, so
 * every reduction is deterministic.  If a real MPI is installed, drop this
 * directory from the include path and link with mpicc instead: the library
 * code only uses the standard names below.
 */
#ifndef MPISHIM_MPI_H
#define MPISHIM_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPISHIM 1

typedef int  MPI_Comm;
typedef int  MPI_Datatype;
typedef int  MPI_Op;
typedef int  MPI_Request;
typedef long MPI_Aint;
typedef int  MPI_Fint;

typedef struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  long _bytes;
} MPI_Status;

#define MPI_COMM_WORLD ((MPI_Comm)0)
#define MPI_COMM_SELF  ((MPI_Comm)1)
#define MPI_COMM_NULL  ((MPI_Comm)-1)

/* predefined datatypes (value = table index, see mpishim.c) */
#define MPI_DATATYPE_NULL  ((MPI_Datatype)0)
#define MPI_CHAR           ((MPI_Datatype)1)
#define MPI_BYTE           ((MPI_Datatype)2)
#define MPI_INT            ((MPI_Datatype)3)
#define MPI_DOUBLE         ((MPI_Datatype)4)
#define MPI_LONG           ((MPI_Datatype)5)
#define MPI_FLOAT          ((MPI_Datatype)6)
#define MPI_UNSIGNED       ((MPI_Datatype)7)
#define MPI_LONG_LONG      ((MPI_Datatype)8)
#define MPI_LONG_LONG_INT  MPI_LONG_LONG
#define MPI_UNSIGNED_LONG  ((MPI_Datatype)9)
#define MPI_SHORT          ((MPI_Datatype)10)
#define MPI_INT64_T        MPI_LONG_LONG
#define MPI_INT32_T        MPI_INT

#define MPI_SUM  ((MPI_Op)1)
#define MPI_MAX  ((MPI_Op)2)
#define MPI_MIN  ((MPI_Op)3)
#define MPI_PROD ((MPI_Op)4)

#define MPI_SUCCESS   0
#define MPI_ERR_COMM  5
#define MPI_ERR_COUNT 2
#define MPI_ERR_TYPE  3
#define MPI_ERR_TAG   4
#define MPI_ERR_RANK  6
#define MPI_ERR_OTHER 15

#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG    (-1)
#define MPI_PROC_NULL  (-2)
#define MPI_UNDEFINED  (-32766)

#define MPI_IN_PLACE        ((void*)(-1L))
#define MPI_BOTTOM          ((void*)0)
#define MPI_STATUS_IGNORE   ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL    ((MPI_Request)-1)
#define MPI_MAX_PROCESSOR_NAME 64

int MPI_Init(int* argc, char*** argv);
int MPI_Initialized(int* flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);

int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm);
int MPI_Comm_free(MPI_Comm* comm);
int MPI_Barrier(MPI_Comm comm);

int MPI_Send(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status* st);
int MPI_Isend(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Wait(MPI_Request* req, MPI_Status* st);
int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* sts);
int MPI_Test(MPI_Request* req, int* flag, MPI_Status* st);
int MPI_Iprobe(int src, int tag, MPI_Comm comm, int* flag, MPI_Status* st);
int MPI_Probe(int src, int tag, MPI_Comm comm, MPI_Status* st);
int MPI_Get_count(const MPI_Status* st, MPI_Datatype dt, int* count);

int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Gatherv(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, const int* rcounts, const int* displs, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Allgatherv(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, const int* rcounts, const int* displs, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Scatter(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Scatterv(const void* sbuf, const int* scounts, const int* displs, MPI_Datatype sdt, void* rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm);

int MPI_Get_address(const void* location, MPI_Aint* address);
int MPI_Type_create_struct(int count, const int* blocklens, const MPI_Aint* displs, const MPI_Datatype* types, MPI_Datatype* newtype);
int MPI_Type_struct(int count, int* blocklens, MPI_Aint* displs, MPI_Datatype* types, MPI_Datatype* newtype);
int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* dt);
int MPI_Type_free(MPI_Datatype* dt);
int MPI_Type_size(MPI_Datatype dt, int* size);
int MPI_Get_processor_name(char* name, int* len);
int MPI_Error_string(int errorcode, char* string, int* resultlen);

#ifdef __cplusplus
}
#endif
#endif /* MPISHIM_MPI_H */
