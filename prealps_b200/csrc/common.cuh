// common.cuh -- shared declarations for libprealps_cuda (sm_100a only).
#pragma once
#ifdef PCU_EMUL  // tests/emul: CPU emulation of the kernels (test infrastructure, never defined for the product build)
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/prealps_cuda.h"

namespace pcu {
void set_error(const char* fmt, ...);
}

#define PCU_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      pcu::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

#define PCU_CHECK(cond, ...)                  \
  do {                                        \
    if (!(cond)) {                            \
      pcu::set_error(__VA_ARGS__);            \
      return 1;                               \
    }                                         \
  } while (0)

struct pcu_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start[16] = {};
  cudaEvent_t ev_stop[16] = {};
  int64_t launches = 0;
  // scratch for two-stage reductions: partials[grid][<=4096] and a small result area
  double* red_partials = nullptr;
  size_t red_partials_doubles = 0;
  // NCCL (loaded with dlopen; absent => single rank)
  void* nccl_comm = nullptr;
  int nranks = 1;
  int rank = 0;
};

namespace pcu {

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// grid for streaming kernels: a multiple of the SM count
inline int stream_grid(const pcu_ctx* ctx, int64_t work_items, int per_block, int max_waves = 8) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)ctx->num_sms * max_waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

int ensure_partials(pcu_ctx* ctx, size_t doubles);

// NCCL thin wrappers (ctx.cu)
// st == nullptr: the library stream
int nccl_send(pcu_ctx* ctx, const void* buf, size_t count, int is_double, int peer, cudaStream_t st = nullptr);
int nccl_recv(pcu_ctx* ctx, void* buf, size_t count, int is_double, int peer, cudaStream_t st = nullptr);
int nccl_group_start(pcu_ctx* ctx);
int nccl_group_end(pcu_ctx* ctx);

}  // namespace pcu

#define PCU_LAUNCH_CHECK(ctx)                                                          \
  do {                                                                                 \
    (ctx)->launches++;                                                                 \
    cudaError_t e_ = cudaGetLastError();                                               \
    if (e_ != cudaSuccess) {                                                           \
      pcu::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)
