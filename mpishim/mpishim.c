/*
 * mpishim/mpishim.c -- fork + shared-memory implementation of the MPI subset
 * declared in mpishim/mpi.h.  See the header for the launch convention
 * (MPISHIM_NP=<n> ./prog).
 *
 * Layout of the shared arena (one anonymous MAP_SHARED mapping, created before
 * fork so every rank sees it at the same virtual address):
 *
 *   [ hdr_t | heap ... ]
 *
 * hdr_t holds one process-shared mutex + condition variable, the barrier
 * state, the child pids and one FIFO message queue per destination rank.
 * Messages live in the heap (first-fit allocator with coalescing, protected by
 * the same mutex).  Sends are eager: the payload is copied into the arena and
 * the call returns, so no rendezvous deadlock is possible.
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <errno.h>
#include <pthread.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define MAXRANKS 64
#define MAXTYPES 256
#define MAXREQS  4096
#define TAG_BCAST   (-1001)
#define TAG_REDUCE  (-1002)
#define TAG_GATHER  (-1003)
#define TAG_SCATTER (-1004)

typedef struct blk {
  size_t size;      /* payload bytes */
  size_t prev_size; /* payload bytes of the physically previous block (0 if first) */
  int    free_;
  int    last_;
} blk_t;

typedef struct msg {
  int src, tag, comm;
  size_t bytes;
  size_t next; /* arena offset of next message in the queue, 0 = none */
} msg_t;

typedef struct hdr {
  pthread_mutex_t mu;
  pthread_cond_t  cv;
  int np;
  int aborted;
  int abort_code;
  int bar_count;
  int bar_gen;
  pid_t pids[MAXRANKS];
  size_t qhead[MAXRANKS], qtail[MAXRANKS];
  size_t heap_off, heap_size;
} hdr_t;

static hdr_t* H = NULL;
static char*  base_ = NULL;
static size_t arena_bytes = 0;
static int my_rank = 0, n_ranks = 1, inited = 0, finalized = 0;

static size_t type_size[MAXTYPES] = {0, 1, 1, sizeof(int), sizeof(double), sizeof(long),
                                     sizeof(float), sizeof(unsigned), sizeof(long long),
                                     sizeof(unsigned long), sizeof(short)};
static int n_types = 11;

typedef struct req {
  int active, is_recv;
  void* buf; int count; MPI_Datatype dt; int src, tag, comm;
} req_t;
static req_t reqs[MAXREQS];

/* ------------------------------------------------------------------ heap */
#define ALIGN 64
#define HDRSZ ((sizeof(blk_t) + ALIGN - 1) / ALIGN * ALIGN)

static blk_t* blk_at(size_t off) { return (blk_t*)(base_ + off); }
static size_t blk_off(blk_t* b) { return (size_t)((char*)b - base_); }
static blk_t* blk_next(blk_t* b) { return b->last_ ? NULL : (blk_t*)((char*)b + HDRSZ + b->size); }
static blk_t* blk_prev(blk_t* b) {
  if (blk_off(b) == H->heap_off) return NULL;
  return (blk_t*)((char*)b - b->prev_size - HDRSZ);
}

/* caller holds the lock */
static void* heap_alloc(size_t n) {
  n = (n + ALIGN - 1) / ALIGN * ALIGN;
  if (n == 0) n = ALIGN;
  for (blk_t* b = blk_at(H->heap_off); b; b = blk_next(b)) {
    if (!b->free_ || b->size < n) continue;
    if (b->size >= n + HDRSZ + ALIGN) { /* split */
      blk_t* r = (blk_t*)((char*)b + HDRSZ + n);
      r->size = b->size - n - HDRSZ;
      r->prev_size = n;
      r->free_ = 1;
      r->last_ = b->last_;
      b->size = n;
      b->last_ = 0;
      blk_t* rn = blk_next(r);
      if (rn) rn->prev_size = r->size;
    }
    b->free_ = 0;
    return (char*)b + HDRSZ;
  }
  return NULL;
}

static void heap_free(void* p) {
  blk_t* b = (blk_t*)((char*)p - HDRSZ);
  b->free_ = 1;
  blk_t* n = blk_next(b);
  if (n && n->free_) {
    b->size += HDRSZ + n->size;
    b->last_ = n->last_;
    blk_t* nn = blk_next(b);
    if (nn) nn->prev_size = b->size;
  }
  blk_t* pv = blk_prev(b);
  if (pv && pv->free_) {
    pv->size += HDRSZ + b->size;
    pv->last_ = b->last_;
    blk_t* nn = blk_next(pv);
    if (nn) nn->prev_size = pv->size;
  }
}

/* ------------------------------------------------------------- utilities */
static void die_all(int code) {
  if (H) {
    H->aborted = 1;
    H->abort_code = code;
    for (int r = 0; r < n_ranks; ++r)
      if (r != my_rank && H->pids[r] > 0) kill(H->pids[r], SIGKILL);
  }
  _exit(code ? code : 1);
}

static void lock(void) { pthread_mutex_lock(&H->mu); }
static void unlock(void) { pthread_mutex_unlock(&H->mu); }

/* wait on the condition variable with a timeout so that a rank that died
 * (segfault, abort) cannot hang the others forever */
static void cv_wait(void) {
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  ts.tv_nsec += 200 * 1000 * 1000;
  if (ts.tv_nsec >= 1000000000L) { ts.tv_sec += 1; ts.tv_nsec -= 1000000000L; }
  pthread_cond_timedwait(&H->cv, &H->mu, &ts);
  if (H->aborted) { unlock(); _exit(H->abort_code ? H->abort_code : 1); }
  if (my_rank == 0) {
    for (int r = 1; r < n_ranks; ++r) {
      int st;
      if (H->pids[r] > 0 && waitpid(H->pids[r], &st, WNOHANG) == H->pids[r]) {
        if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) {
          fprintf(stderr, "[mpishim] rank %d died (status 0x%x); aborting\n", r, st);
          unlock();
          die_all(1);
        }
        H->pids[r] = -1; /* exited cleanly */
      }
    }
  } else if (getppid() == 1) { /* rank 0 vanished */
    unlock();
    _exit(1);
  }
}

static size_t dt_size(MPI_Datatype dt) {
  if (dt <= 0 || dt >= n_types) {
    fprintf(stderr, "[mpishim] bad datatype %d\n", dt);
    die_all(3);
  }
  return type_size[dt];
}

/* ------------------------------------------------------------ init / exit */
int MPI_Init(int* argc, char*** argv) {
  (void)argc; (void)argv;
  if (inited) return MPI_SUCCESS;
  const char* e = getenv("MPISHIM_NP");
  n_ranks = e ? atoi(e) : 1;
  if (n_ranks < 1) n_ranks = 1;
  if (n_ranks > MAXRANKS) { fprintf(stderr, "[mpishim] MPISHIM_NP > %d\n", MAXRANKS); exit(2); }
  const char* am = getenv("MPISHIM_ARENA_MB");
  size_t mb = am ? (size_t)atol(am) : (n_ranks > 1 ? 16384 : 64);
  arena_bytes = mb << 20;
  base_ = (char*)mmap(NULL, arena_bytes, PROT_READ | PROT_WRITE,
                      MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (base_ == MAP_FAILED) { perror("[mpishim] mmap"); exit(2); }
  H = (hdr_t*)base_;
  memset(H, 0, sizeof(*H));
  pthread_mutexattr_t ma; pthread_mutexattr_init(&ma);
  pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
  pthread_mutex_init(&H->mu, &ma);
  pthread_condattr_t ca; pthread_condattr_init(&ca);
  pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
  pthread_cond_init(&H->cv, &ca);
  H->np = n_ranks;
  H->heap_off = (sizeof(hdr_t) + 4095) / 4096 * 4096;
  H->heap_size = arena_bytes - H->heap_off;
  blk_t* b0 = blk_at(H->heap_off);
  b0->size = H->heap_size - HDRSZ; b0->prev_size = 0; b0->free_ = 1; b0->last_ = 1;
  H->pids[0] = getpid();
  fflush(stdout); fflush(stderr);
  my_rank = 0;
  for (int r = 1; r < n_ranks; ++r) {
    pid_t p = fork();
    if (p < 0) { perror("[mpishim] fork"); die_all(2); }
    if (p == 0) { my_rank = r; break; }
    H->pids[r] = p;
  }
  if (my_rank != 0) H->pids[my_rank] = getpid();
  inited = 1;
  MPI_Barrier(MPI_COMM_WORLD);
  return MPI_SUCCESS;
}

int MPI_Initialized(int* flag) { *flag = inited; return MPI_SUCCESS; }

int MPI_Finalize(void) {
  if (!inited || finalized) return MPI_SUCCESS;
  MPI_Barrier(MPI_COMM_WORLD);
  finalized = 1;
  fflush(stdout); fflush(stderr);
  if (my_rank != 0) _exit(0); /* children never return into main's epilogue twice */
  for (int r = 1; r < n_ranks; ++r)
    if (H->pids[r] > 0) { int st; waitpid(H->pids[r], &st, 0); }
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode) {
  (void)comm;
  fflush(stdout); fflush(stderr);
  if (!inited) _exit(errorcode ? errorcode : 1);
  die_all(errorcode ? errorcode : 1);
  return MPI_SUCCESS;
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Comm_rank(MPI_Comm comm, int* rank) { *rank = (comm == MPI_COMM_SELF) ? 0 : my_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int* size) { *size = (comm == MPI_COMM_SELF) ? 1 : n_ranks; return MPI_SUCCESS; }
static int next_comm = 2;
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm* newcomm) {
  if (comm == MPI_COMM_SELF) { *newcomm = MPI_COMM_SELF; return MPI_SUCCESS; }
  *newcomm = 2 * (next_comm++); /* even ids = world-group duplicates */
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm* comm) { *comm = MPI_COMM_NULL; return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm comm) {
  if (comm == MPI_COMM_SELF || n_ranks == 1) return MPI_SUCCESS;
  lock();
  int gen = H->bar_gen;
  if (++H->bar_count == n_ranks) {
    H->bar_count = 0;
    H->bar_gen++;
    pthread_cond_broadcast(&H->cv);
  } else {
    while (H->bar_gen == gen) cv_wait();
  }
  unlock();
  return MPI_SUCCESS;
}

/* ------------------------------------------------------- point to point */
static void push_msg(int dest, int tag, int comm, const void* buf, size_t bytes) {
  lock();
  msg_t* m;
  while ((m = (msg_t*)heap_alloc(sizeof(msg_t) + bytes)) == NULL) {
    /* arena full: wait for receivers to drain */
    cv_wait();
  }
  m->src = my_rank; m->tag = tag; m->comm = comm; m->bytes = bytes; m->next = 0;
  unlock();
  memcpy((char*)m + sizeof(msg_t), buf, bytes); /* copy outside the lock */
  lock();
  size_t off = (size_t)((char*)m - base_);
  if (H->qtail[dest]) ((msg_t*)(base_ + H->qtail[dest]))->next = off; else H->qhead[dest] = off;
  H->qtail[dest] = off;
  pthread_cond_broadcast(&H->cv);
  unlock();
}

/* caller holds the lock; returns the matching message (unlinked iff take) */
static msg_t* find_msg(int src, int tag, int comm, int take) {
  size_t prev = 0;
  for (size_t off = H->qhead[my_rank]; off; prev = off, off = ((msg_t*)(base_ + off))->next) {
    msg_t* m = (msg_t*)(base_ + off);
    if (m->comm != comm) continue;
    if (src != MPI_ANY_SOURCE && m->src != src) continue;
    if (tag != MPI_ANY_TAG && m->tag != tag) continue;
    if (tag == MPI_ANY_TAG && m->tag < -1000) continue; /* never steal collective traffic */
    if (take) {
      if (prev) ((msg_t*)(base_ + prev))->next = m->next; else H->qhead[my_rank] = m->next;
      if (H->qtail[my_rank] == off) H->qtail[my_rank] = prev;
    }
    return m;
  }
  return NULL;
}

static void recv_bytes(void* buf, size_t maxbytes, int src, int tag, int comm, MPI_Status* st) {
  lock();
  msg_t* m;
  while ((m = find_msg(src, tag, comm, 1)) == NULL) cv_wait();
  unlock();
  if (m->bytes > maxbytes) {
    fprintf(stderr, "[mpishim] rank %d: message truncated (%zu > %zu) src %d tag %d\n",
            my_rank, m->bytes, maxbytes, m->src, m->tag);
    die_all(4);
  }
  memcpy(buf, (char*)m + sizeof(msg_t), m->bytes);
  if (st) { st->MPI_SOURCE = m->src; st->MPI_TAG = m->tag; st->MPI_ERROR = MPI_SUCCESS; st->_bytes = (long)m->bytes; }
  lock();
  heap_free(m);
  pthread_cond_broadcast(&H->cv);
  unlock();
}

int MPI_Send(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm) {
  if (dest == MPI_PROC_NULL) return MPI_SUCCESS;
  if (dest < 0 || dest >= n_ranks) return MPI_ERR_RANK;
  push_msg(dest, tag, comm, buf, (size_t)count * dt_size(dt));
  return MPI_SUCCESS;
}

int MPI_Recv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status* st) {
  if (src == MPI_PROC_NULL) return MPI_SUCCESS;
  recv_bytes(buf, (size_t)count * dt_size(dt), src, tag, comm, st);
  return MPI_SUCCESS;
}

static int new_req(void) {
  for (int i = 0; i < MAXREQS; ++i) if (!reqs[i].active) { reqs[i].active = 1; return i; }
  fprintf(stderr, "[mpishim] out of requests\n");
  die_all(5);
  return -1;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request* req) {
  int rc = MPI_Send(buf, count, dt, dest, tag, comm);
  int i = new_req();
  reqs[i].is_recv = 0;
  *req = i;
  return rc;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request* req) {
  int i = new_req();
  reqs[i].is_recv = 1; reqs[i].buf = buf; reqs[i].count = count; reqs[i].dt = dt;
  reqs[i].src = src; reqs[i].tag = tag; reqs[i].comm = comm;
  *req = i;
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request* req, MPI_Status* st) {
  if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
  req_t* r = &reqs[*req];
  if (r->active && r->is_recv)
    recv_bytes(r->buf, (size_t)r->count * dt_size(r->dt), r->src, r->tag, r->comm, st);
  r->active = 0;
  *req = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}

int MPI_Waitall(int n, MPI_Request* rq, MPI_Status* sts) {
  for (int i = 0; i < n; ++i) MPI_Wait(&rq[i], sts ? &sts[i] : NULL);
  return MPI_SUCCESS;
}

int MPI_Test(MPI_Request* req, int* flag, MPI_Status* st) {
  if (*req == MPI_REQUEST_NULL) { *flag = 1; return MPI_SUCCESS; }
  req_t* r = &reqs[*req];
  if (!r->is_recv) { *flag = 1; r->active = 0; *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }
  lock();
  msg_t* m = find_msg(r->src, r->tag, r->comm, 0);
  unlock();
  if (m) { *flag = 1; return MPI_Wait(req, st); }
  *flag = 0;
  return MPI_SUCCESS;
}

int MPI_Iprobe(int src, int tag, MPI_Comm comm, int* flag, MPI_Status* st) {
  lock();
  msg_t* m = find_msg(src, tag, comm, 0);
  if (m && st) { st->MPI_SOURCE = m->src; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; st->_bytes = (long)m->bytes; }
  *flag = (m != NULL);
  unlock();
  return MPI_SUCCESS;
}

int MPI_Probe(int src, int tag, MPI_Comm comm, MPI_Status* st) {
  lock();
  msg_t* m;
  while ((m = find_msg(src, tag, comm, 0)) == NULL) cv_wait();
  if (st) { st->MPI_SOURCE = m->src; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; st->_bytes = (long)m->bytes; }
  unlock();
  return MPI_SUCCESS;
}

int MPI_Get_count(const MPI_Status* st, MPI_Datatype dt, int* count) {
  *count = (int)((size_t)st->_bytes / dt_size(dt));
  return MPI_SUCCESS;
}

/* ----------------------------------------------------------- collectives */
int MPI_Bcast(void* buf, int count, MPI_Datatype dt, int root, MPI_Comm comm) {
  if (comm == MPI_COMM_SELF || n_ranks == 1) return MPI_SUCCESS;
  size_t bytes = (size_t)count * dt_size(dt);
  if (my_rank == root) {
    for (int r = 0; r < n_ranks; ++r) if (r != root) push_msg(r, TAG_BCAST, comm, buf, bytes);
  } else {
    recv_bytes(buf, bytes, root, TAG_BCAST, comm, NULL);
  }
  return MPI_SUCCESS;
}

static void reduce_into(void* acc, const void* in, int count, MPI_Datatype dt, MPI_Op op) {
#define RED(T) do { T* a = (T*)acc; const T* b = (const T*)in;                   \
    for (int i = 0; i < count; ++i) {                                             \
      switch (op) { case MPI_SUM: a[i] = a[i] + b[i]; break;                      \
                    case MPI_PROD: a[i] = a[i] * b[i]; break;                     \
                    case MPI_MAX: a[i] = a[i] > b[i] ? a[i] : b[i]; break;        \
                    case MPI_MIN: a[i] = a[i] < b[i] ? a[i] : b[i]; break;        \
                    default: die_all(6); } } } while (0)
  switch (dt) {
    case MPI_INT: RED(int); break;
    case MPI_DOUBLE: RED(double); break;
    case MPI_LONG: RED(long); break;
    case MPI_FLOAT: RED(float); break;
    case MPI_UNSIGNED: RED(unsigned); break;
    case MPI_LONG_LONG: RED(long long); break;
    case MPI_UNSIGNED_LONG: RED(unsigned long); break;
    case MPI_CHAR: RED(char); break;
    default: fprintf(stderr, "[mpishim] reduce on datatype %d\n", dt); die_all(6);
  }
#undef RED
}

int MPI_Reduce(const void* sbuf, void* rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm) {
  size_t bytes = (size_t)count * dt_size(dt);
  if (comm == MPI_COMM_SELF || n_ranks == 1) {
    if (sbuf != MPI_IN_PLACE) memcpy(rbuf, sbuf, bytes);
    return MPI_SUCCESS;
  }
  if (my_rank != root) {
    push_msg(root, TAG_REDUCE, comm, sbuf == MPI_IN_PLACE ? rbuf : sbuf, bytes);
    return MPI_SUCCESS;
  }
  /* rank-ordered reduction: acc = x_0 + x_1 + ... (deterministic) */
  void* mine = malloc(bytes ? bytes : 1);
  void* tmp = malloc(bytes ? bytes : 1);
  memcpy(mine, sbuf == MPI_IN_PLACE ? rbuf : sbuf, bytes);
  int first = 1;
  for (int r = 0; r < n_ranks; ++r) {
    const void* contrib = mine;
    if (r != root) { recv_bytes(tmp, bytes, r, TAG_REDUCE, comm, NULL); contrib = tmp; }
    if (first) { memcpy(rbuf, contrib, bytes); first = 0; }
    else reduce_into(rbuf, contrib, count, dt, op);
  }
  free(mine); free(tmp);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void* sbuf, void* rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm) {
  MPI_Reduce(sbuf, rbuf, count, dt, op, 0, comm);
  return MPI_Bcast(rbuf, count, dt, 0, comm);
}

int MPI_Gatherv(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, const int* rcounts,
                const int* displs, MPI_Datatype rdt, int root, MPI_Comm comm) {
  size_t sb = (size_t)scount * dt_size(sdt);
  if (my_rank != root) { push_msg(root, TAG_GATHER, comm, sbuf, sb); return MPI_SUCCESS; }
  size_t rs = dt_size(rdt);
  for (int r = 0; r < n_ranks; ++r) {
    char* dst = (char*)rbuf + (size_t)displs[r] * rs;
    if (r == root) { if (sbuf != MPI_IN_PLACE) memcpy(dst, sbuf, sb); }
    else recv_bytes(dst, (size_t)rcounts[r] * rs, r, TAG_GATHER, comm, NULL);
  }
  return MPI_SUCCESS;
}

int MPI_Gather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount,
               MPI_Datatype rdt, int root, MPI_Comm comm) {
  int rc[MAXRANKS], ds[MAXRANKS];
  for (int r = 0; r < n_ranks; ++r) { rc[r] = rcount; ds[r] = r * rcount; }
  return MPI_Gatherv(sbuf, scount, sdt, rbuf, rc, ds, rdt, root, comm);
}

int MPI_Allgatherv(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, const int* rcounts,
                   const int* displs, MPI_Datatype rdt, MPI_Comm comm) {
  MPI_Gatherv(sbuf, scount, sdt, rbuf, rcounts, displs, rdt, 0, comm);
  int tot = 0;
  for (int r = 0; r < n_ranks; ++r) if (displs[r] + rcounts[r] > tot) tot = displs[r] + rcounts[r];
  return MPI_Bcast(rbuf, tot, rdt, 0, comm);
}

int MPI_Allgather(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount,
                  MPI_Datatype rdt, MPI_Comm comm) {
  MPI_Gather(sbuf, scount, sdt, rbuf, rcount, rdt, 0, comm);
  return MPI_Bcast(rbuf, rcount * n_ranks, rdt, 0, comm);
}

int MPI_Scatterv(const void* sbuf, const int* scounts, const int* displs, MPI_Datatype sdt,
                 void* rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm) {
  if (my_rank == root) {
    size_t ss = dt_size(sdt);
    for (int r = 0; r < n_ranks; ++r) {
      const char* src = (const char*)sbuf + (size_t)displs[r] * ss;
      if (r == root) { if (rbuf != MPI_IN_PLACE) memcpy(rbuf, src, (size_t)scounts[r] * ss); }
      else push_msg(r, TAG_SCATTER, comm, src, (size_t)scounts[r] * ss);
    }
  } else {
    recv_bytes(rbuf, (size_t)rcount * dt_size(rdt), root, TAG_SCATTER, comm, NULL);
  }
  return MPI_SUCCESS;
}

int MPI_Scatter(const void* sbuf, int scount, MPI_Datatype sdt, void* rbuf, int rcount,
                MPI_Datatype rdt, int root, MPI_Comm comm) {
  int sc[MAXRANKS], ds[MAXRANKS];
  for (int r = 0; r < n_ranks; ++r) { sc[r] = scount; ds[r] = r * scount; }
  return MPI_Scatterv(sbuf, sc, ds, sdt, rbuf, rcount, rdt, root, comm);
}

/* ------------------------------------------------------------- datatypes */
int MPI_Get_address(const void* location, MPI_Aint* address) { *address = (MPI_Aint)(intptr_t)location; return MPI_SUCCESS; }

int MPI_Type_create_struct(int count, const int* blocklens, const MPI_Aint* displs,
                           const MPI_Datatype* types, MPI_Datatype* newtype) {
  /* extent = end of the furthest member; good enough for the packed POD
   * structs (all-int headers) the callers describe */
  size_t ext = 0;
  for (int i = 0; i < count; ++i) {
    size_t end = (size_t)displs[i] + (size_t)blocklens[i] * dt_size(types[i]);
    if (end > ext) ext = end;
  }
  if (n_types >= MAXTYPES) die_all(7);
  type_size[n_types] = ext;
  *newtype = n_types++;
  return MPI_SUCCESS;
}

int MPI_Type_struct(int count, int* blocklens, MPI_Aint* displs, MPI_Datatype* types, MPI_Datatype* newtype) {
  return MPI_Type_create_struct(count, blocklens, displs, types, newtype);
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype* newtype) {
  if (n_types >= MAXTYPES) die_all(7);
  type_size[n_types] = (size_t)count * dt_size(oldtype);
  *newtype = n_types++;
  return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype* dt) { (void)dt; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype* dt) { *dt = MPI_DATATYPE_NULL; return MPI_SUCCESS; }
int MPI_Type_size(MPI_Datatype dt, int* size) { *size = (int)dt_size(dt); return MPI_SUCCESS; }

int MPI_Get_processor_name(char* name, int* len) {
  gethostname(name, MPI_MAX_PROCESSOR_NAME - 1);
  name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
  *len = (int)strlen(name);
  return MPI_SUCCESS;
}

int MPI_Error_string(int errorcode, char* string, int* resultlen) {
  *resultlen = snprintf(string, 64, "mpishim error %d", errorcode);
  return MPI_SUCCESS;
}
