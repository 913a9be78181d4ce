// ctx.cu -- context, memory, timers and the NCCL plumbing of libprealps_cuda.
#include "common.cuh"

#include <dlfcn.h>
#include <stdarg.h>

namespace pcu {
static thread_local char g_err[1024] = "no error";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  if (getenv("PREALPS_CUDA_VERBOSE")) fprintf(stderr, "[prealps_cuda] %s\n", g_err);
}

int ensure_partials(pcu_ctx* ctx, size_t doubles) {
  if (ctx->red_partials_doubles >= doubles) return 0;
  if (ctx->red_partials) cudaFree(ctx->red_partials);
  ctx->red_partials = nullptr;
  ctx->red_partials_doubles = 0;
  PCU_CUDA(cudaMalloc(&ctx->red_partials, doubles * sizeof(double)));
  ctx->red_partials_doubles = doubles;
  return 0;
}

// ---- NCCL through dlopen: the library must load (and fail loudly only when
// multi-rank features are requested) on boxes without libnccl.
typedef struct { char internal[PCU_NCCL_ID_BYTES]; } nccl_uid_t;
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(nccl_uid_t*) = nullptr;
  int (*CommInitRank)(void**, int, nccl_uid_t, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0;

static int load_nccl() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.h; ++i) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.h) { set_error("NCCL requested but libnccl.so.2 cannot be loaded: %s", dlerror()); return 1; }
#define SYM(field, name)                                                       \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                            \
  if (!g_nccl.field) { set_error("libnccl lacks symbol %s", name); return 1; }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(Broadcast, "ncclBroadcast")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}
#define PCU_NCCL(call)                                                                       \
  do {                                                                                       \
    int r_ = (call);                                                                         \
    if (r_ != 0) { set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); return 1; } \
  } while (0)

int nccl_send(pcu_ctx* ctx, const void* buf, size_t count, int is_double, int peer, cudaStream_t st) {
  PCU_NCCL(g_nccl.Send(buf, count, is_double ? kNcclFloat64 : kNcclInt32, peer, ctx->nccl_comm, st ? st : ctx->stream));
  return 0;
}
int nccl_recv(pcu_ctx* ctx, void* buf, size_t count, int is_double, int peer, cudaStream_t st) {
  PCU_NCCL(g_nccl.Recv(buf, count, is_double ? kNcclFloat64 : kNcclInt32, peer, ctx->nccl_comm, st ? st : ctx->stream));
  return 0;
}
int nccl_group_start(pcu_ctx*) { PCU_NCCL(g_nccl.GroupStart()); return 0; }
int nccl_group_end(pcu_ctx*) { PCU_NCCL(g_nccl.GroupEnd()); return 0; }
}  // namespace pcu

using namespace pcu;

extern "C" {

const char* pcu_last_error(void) { return g_err; }

int pcu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int pcu_ctx_create(int device, pcu_ctx** out) {
  PCU_CHECK(out != nullptr, "pcu_ctx_create: null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("pcu_ctx_create: no CUDA device available (%s); libprealps_cuda has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  PCU_CHECK(device >= 0 && device < n, "pcu_ctx_create: device %d out of range (have %d)", device, n);
  PCU_CUDA(cudaSetDevice(device));
  pcu_ctx* c = new pcu_ctx();
  c->device = device;
  cudaDeviceProp prop;
  PCU_CUDA(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  PCU_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    cudaMemPool_t pool;
    PCU_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    PCU_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  for (int i = 0; i < 16; ++i) {
    PCU_CUDA(cudaEventCreate(&c->ev_start[i]));
    PCU_CUDA(cudaEventCreate(&c->ev_stop[i]));
  }
  *out = c;
  return 0;
}

int pcu_ctx_destroy(pcu_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
  if (c->red_partials) cudaFree(c->red_partials);
  for (int i = 0; i < 16; ++i) { cudaEventDestroy(c->ev_start[i]); cudaEventDestroy(c->ev_stop[i]); }
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

void* pcu_ctx_stream(pcu_ctx* c) { return (void*)c->stream; }

int pcu_sync(pcu_ctx* c) {
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// Stream-ordered allocations from the device's default pool, which is told to keep freed memory:
// a solver that is created and destroyed per solve (ECGInitialize .. ECGFinalize) then never pays
// for cudaMalloc/cudaFree again (measured: 0.17 s + 0.16 s per solve at 64^3 with the plain calls).
void* pcu_malloc(pcu_ctx* c, size_t bytes) {
  void* p = nullptr;
  cudaSetDevice(c->device);
  cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 8, c->stream);
  if (e != cudaSuccess) { set_error("pcu_malloc(%zu): %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
  return p;
}
int pcu_free(pcu_ctx* c, void* p) {
  if (!p) return 0;
  PCU_CUDA(cudaFreeAsync(p, c->stream));
  return 0;
}
int pcu_memset(pcu_ctx* c, void* p, int byte, size_t bytes) {
  PCU_CUDA(cudaMemsetAsync(p, byte, bytes, c->stream));
  return 0;
}
int pcu_h2d(pcu_ctx* c, void* dst, const void* src, size_t bytes) {
  PCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int pcu_d2h(pcu_ctx* c, void* dst, const void* src, size_t bytes) {
  PCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int pcu_d2d(pcu_ctx* c, void* dst, const void* src, size_t bytes) {
  PCU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
int pcu_copy_cols(pcu_ctx* c, int m, int ncols, double* dst, int ldd, const double* src, int lds) {
  PCU_CHECK(c && dst && src && m >= 0 && ncols >= 0 && ldd >= ncols && lds >= ncols, "pcu_copy_cols: bad arguments");
  if (m == 0 || ncols == 0) return 0;
  PCU_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)ldd, src, sizeof(double) * (size_t)lds,
                             sizeof(double) * (size_t)ncols, (size_t)m, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
int pcu_zero_cols(pcu_ctx* c, int m, int ncols, double* dst, int ldd) {
  PCU_CHECK(c && dst && m >= 0 && ncols >= 0 && ldd >= ncols, "pcu_zero_cols: bad arguments");
  if (m == 0 || ncols == 0) return 0;
  PCU_CUDA(cudaMemset2DAsync(dst, sizeof(double) * (size_t)ldd, 0, sizeof(double) * (size_t)ncols, (size_t)m, c->stream));
  return 0;
}
void* pcu_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
int pcu_host_free(void* p) {
  if (p) PCU_CUDA(cudaFreeHost(p));
  return 0;
}

int pcu_ptr_is_device(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int pcu_flush_l2(pcu_ctx* c) {
  static void* scratch = nullptr;
  const size_t bytes = (size_t)256 << 20;
  if (!scratch) PCU_CUDA(cudaMalloc(&scratch, bytes));
  PCU_CUDA(cudaMemsetAsync(scratch, 1, bytes, c->stream));
  return 0;
}

int pcu_timer_start(pcu_ctx* c, int s) {
  PCU_CHECK(s >= 0 && s < 16, "timer slot");
  PCU_CUDA(cudaEventRecord(c->ev_start[s], c->stream));
  return 0;
}
int pcu_timer_stop(pcu_ctx* c, int s) {
  PCU_CHECK(s >= 0 && s < 16, "timer slot");
  PCU_CUDA(cudaEventRecord(c->ev_stop[s], c->stream));
  return 0;
}
int pcu_timer_elapsed_ms(pcu_ctx* c, int s, float* ms) {
  PCU_CHECK(s >= 0 && s < 16, "timer slot");
  PCU_CUDA(cudaEventSynchronize(c->ev_stop[s]));
  PCU_CUDA(cudaEventElapsedTime(ms, c->ev_start[s], c->ev_stop[s]));
  return 0;
}
int64_t pcu_launch_count(pcu_ctx* c) { return c->launches; }

int pcu_nccl_unique_id(void* out) {
  if (load_nccl()) return 1;
  nccl_uid_t id;
  PCU_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out, &id, PCU_NCCL_ID_BYTES);
  return 0;
}

int pcu_ctx_init_nccl(pcu_ctx* c, int nranks, int rank, const void* id128) {
  PCU_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "pcu_ctx_init_nccl: bad rank %d/%d", rank, nranks);
  c->nranks = nranks;
  c->rank = rank;
  if (nranks == 1) return 0;
  if (load_nccl()) return 1;
  nccl_uid_t id;
  memcpy(&id, id128, PCU_NCCL_ID_BYTES);
  PCU_CUDA(cudaSetDevice(c->device));
  PCU_NCCL(g_nccl.CommInitRank(&c->nccl_comm, nranks, id, rank));
  // NCCL sets its channels up on the first collective (8 ranks: ~8 s, which round 1 had mistaken for eight METIS runs
  // competing for the host): pay for it here, once, where it can be told apart from the operator build
  // ... and so are the broadcast ring (the partition is broadcast from rank 0) and the point-to-point connections (halo
  // plan and halo exchange): one grouped send/recv with every peer connects them all at once instead of one by one
  // inside the first operator build
  double* d = nullptr;
  PCU_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)(2 * nranks + 2)));
  PCU_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)(2 * nranks + 2), c->stream));
  PCU_NCCL(g_nccl.AllReduce(d, d, 1, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream));
  PCU_NCCL(g_nccl.Broadcast(d, d, 1, kNcclFloat64, 0, c->nccl_comm, c->stream));
  if (nccl_group_start(c)) return 1;
  for (int q = 0; q < nranks; ++q) {
    if (q == rank) continue;
    PCU_NCCL(g_nccl.Send(d + 2 + q, 1, kNcclFloat64, q, c->nccl_comm, c->stream));
    PCU_NCCL(g_nccl.Recv(d + 2 + nranks + q, 1, kNcclFloat64, q, c->nccl_comm, c->stream));
  }
  if (nccl_group_end(c)) return 1;
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return 0;
}
int pcu_comm_size(pcu_ctx* c) { return c->nranks; }
int pcu_comm_rank(pcu_ctx* c) { return c->rank; }

int pcu_exchange_ints(pcu_ctx* c, int nnbr, const int* nbr, const int* send_ptr, const int* send_data,
                      const int* recv_ptr, int* recv_data) {
  if (nnbr == 0) return 0;
  PCU_CHECK(c->nccl_comm != nullptr, "pcu_exchange_ints: NCCL communicator not initialised");
  const int ns = send_ptr[nnbr], nr = recv_ptr[nnbr];
  int *ds = nullptr, *dr = nullptr;
  PCU_CUDA(cudaMalloc(&ds, sizeof(int) * (size_t)(ns > 0 ? ns : 1)));
  PCU_CUDA(cudaMalloc(&dr, sizeof(int) * (size_t)(nr > 0 ? nr : 1)));
  if (ns) PCU_CUDA(cudaMemcpyAsync(ds, send_data, sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, c->stream));
  if (nccl_group_start(c)) return 1;
  for (int q = 0; q < nnbr; ++q) {
    const int a = send_ptr[q + 1] - send_ptr[q], b = recv_ptr[q + 1] - recv_ptr[q];
    if (a > 0 && nccl_send(c, ds + send_ptr[q], (size_t)a, 0, nbr[q])) return 1;
    if (b > 0 && nccl_recv(c, dr + recv_ptr[q], (size_t)b, 0, nbr[q])) return 1;
  }
  if (nccl_group_end(c)) return 1;
  if (nr) PCU_CUDA(cudaMemcpyAsync(recv_data, dr, sizeof(int) * (size_t)nr, cudaMemcpyDeviceToHost, c->stream));
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(ds); cudaFree(dr);
  return 0;
}

int pcu_bcast_ints(pcu_ctx* c, int* host_data, long long n, int root) {
  if (c->nranks == 1 || n == 0) return 0;
  PCU_CHECK(c->nccl_comm != nullptr, "pcu_bcast_ints: NCCL communicator not initialised");
  int* d = nullptr;
  PCU_CUDA(cudaMalloc(&d, sizeof(int) * (size_t)n));
  if (c->rank == root) PCU_CUDA(cudaMemcpyAsync(d, host_data, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  PCU_NCCL(g_nccl.Broadcast(d, d, (size_t)n, kNcclInt32, root, c->nccl_comm, c->stream));
  if (c->rank != root) PCU_CUDA(cudaMemcpyAsync(host_data, d, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return 0;
}

int pcu_allreduce_sum(pcu_ctx* c, double* dbuf, int n) {
  if (c->nranks == 1 || n == 0) return 0;
  PCU_CHECK(c->nccl_comm != nullptr, "pcu_allreduce_sum: NCCL communicator not initialised");
  PCU_NCCL(g_nccl.AllReduce(dbuf, dbuf, (size_t)n, kNcclFloat64, kNcclSum, c->nccl_comm, c->stream));
  return 0;
}

}  // extern "C"
