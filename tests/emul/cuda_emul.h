// cuda_emul.h -- TEST INFRASTRUCTURE ONLY: a miniature CUDA execution model and runtime on the CPU, enough to run
// prealps_b200/csrc/bj_factor.cu and bj_solve.cu (the block-Jacobi factorisation and sweeps) unmodified apart from their
// PCU_EMUL branches (inline PTX -> plain C++).  One pthread per CUDA thread, blocks of a grid one after the other,
// __syncthreads / __syncwarp as barriers that tolerate threads which have already returned, mma.sync.m8n8k4.f64 as a
// warp-wide exchange.  "Device" memory is host memory; streams, events and launch attributes are accepted and ignored.
// It checks index logic and data flow (who reads what after which barrier is NOT checked: cp.async completes at
// once, griddepcontrol is a no-op).  Never linked into the product libraries: tests/test_bj_emul.py builds it.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

// ---------------------------------------------------------------- vector types, qualifiers
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

extern thread_local dim3 threadIdx;
extern dim3 blockIdx, blockDim, gridDim;

void __syncthreads();
void __syncwarp(unsigned mask = 0xffffffffu);
inline void __threadfence() {}
template <class T> inline T __ldg(const T* p) { return *p; }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline long long clock64() { return 0; }
inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T> inline T __ldcg(const T* p) { return *p; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
using std::fma;
using std::max;
using std::min;
void* emul_dyn_smem();
// every lane of the calling warp contributes v; all[0..31] receives the 32 values (lanes that have returned: 0)
void emul_warp_allgather(double v, double (&all)[32]);

// values that fit a double exactly (the kernels shuffle small ints)
template <class T> inline T __shfl_sync(unsigned, T v, int src_lane) {
  double all[32];
  emul_warp_allgather((double)v, all);
  return (T)all[src_lane & 31];
}

template <class T> inline T __shfl_sync(unsigned, T v, int src_lane, int width) {  // sub-warp segments of `width` lanes
  double all[32];
  emul_warp_allgather((double)v, all);
  const int lane = (int)(threadIdx.x & 31), base = lane / width * width;
  return (T)all[base + (src_lane % width)];
}
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) {
  double all[32];
  emul_warp_allgather((double)v, all);
  const int lane = (int)(threadIdx.x & 31);
  return (lane + delta < 32) ? (T)all[lane + delta] : v;
}

template <class T> inline T __shfl_xor_sync(unsigned, T v, int mask) {
  double all[32];
  emul_warp_allgather((double)v, all);
  return (T)all[((int)(threadIdx.x & 31) ^ mask) & 31];
}

inline void pcu_emul_check_aligned(const void* p, size_t a) {
  if (reinterpret_cast<uintptr_t>(p) % a != 0) {
    std::fprintf(stderr, "[emul] misaligned %zu-byte vector access at %p\n", a, p);
    std::abort();
  }
}

// ---------------------------------------------------------------- runtime
typedef int cudaError_t;
typedef void* cudaStream_t;
struct emul_event { double t; };
typedef emul_event* cudaEvent_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamCaptureModeThreadLocal = 1 };
enum { cudaLaunchAttributeProgrammaticStreamSerialization = 6 };
struct cudaLaunchAttribute {
  int id;
  struct { int programmaticStreamSerializationAllowed; } val;
};
struct cudaLaunchConfig_t {
  dim3 gridDim, blockDim;
  size_t dynamicSmemBytes = 0;
  cudaStream_t stream = nullptr;
  cudaLaunchAttribute* attrs = nullptr;
  unsigned numAttrs = 0;
};

void emul_register(const void* p, size_t bytes);   // "device" allocations, for cudaPointerGetAttributes
void emul_unregister(const void* p);
bool emul_is_device(const void* p);
struct cudaDeviceProp { int multiProcessorCount; };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; };
typedef void* cudaMemPool_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaMemPoolAttrReleaseThreshold = 4 };
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 4; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void*) { return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  a->type = emul_is_device(p) ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered;
  return cudaSuccess;
}
inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  void* q = nullptr;
  if (posix_memalign(&q, 256, bytes ? bytes : 256) != 0) return 2;
  std::memset(q, 0xEE, bytes ? bytes : 256);  // fresh device memory is garbage
  emul_register(q, bytes ? bytes : 256);
  *p = static_cast<T*>(q);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { if (p) emul_unregister(p); std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)64 << 30; return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocAsync(T** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t bytes) { *p = static_cast<T*>(std::malloc(bytes ? bytes : 8)); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t i = 0; i < h; ++i) std::memcpy(static_cast<char*>(d) + i * dp, static_cast<const char*>(s) + i * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemset2DAsync(void* d, size_t dp, int v, size_t w, size_t h, cudaStream_t = nullptr) {
  for (size_t i = 0; i < h; ++i) std::memset(static_cast<char*>(d) + i * dp, v, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
inline double emul_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emul_event{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emul_now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((b->t - a->t) * 1e3); return cudaSuccess; }
// CUDA graphs: a capture records the launches (with their arguments) instead of running them, a graph launch replays them
struct emul_graph { std::vector<std::function<void()>> launches; };
typedef emul_graph* cudaGraph_t;
typedef emul_graph* cudaGraphExec_t;
cudaError_t cudaStreamBeginCapture(cudaStream_t, int);
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g);
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, int) { *e = new emul_graph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { for (auto& f : e->launches) f(); return cudaSuccess; }
extern long long emul_graph_replays;  // kernels run from a graph launch (tests)

// run body() for every thread of every block
void emul_launch_impl(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body);
template <class F>
inline void emul_launch(dim3 grid, dim3 block, size_t dyn_smem, F body) { emul_launch_impl(grid, block, dyn_smem, std::function<void()>(body)); }

template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args... args) {
  emul_launch(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, [=] { kernel(KArgs(args)...); });
  return cudaSuccess;
}
