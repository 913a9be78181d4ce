// cuda_shim.h -- TEST INFRASTRUCTURE ONLY: just enough of the CUDA execution model to run the kernels of
// prealps_b200/csrc/spmm_kernels.cuh on the CPU, one pthread per CUDA thread, one block at a time, so that the index
// logic of a kernel can be checked in a container without a GPU.  It is never linked into libprealps_cuda /
// libprealps_b200 (the product has no CPU path); only tests/test_spmm_emul.py builds and loads it.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct emul_dim3 { unsigned x = 1, y = 1, z = 1; };
struct int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }

extern thread_local emul_dim3 threadIdx;
extern emul_dim3 blockIdx, blockDim, gridDim;
void __syncthreads();

#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

template <class T> inline T __ldg(const T* p) { return *p; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
using std::fma;
using std::max;
using std::min;

inline void pcu_emul_check_aligned(const void* p, size_t a) {
  if (reinterpret_cast<uintptr_t>(p) % a != 0) {
    std::fprintf(stderr, "[emul] misaligned %zu-byte vector access at %p\n", a, p);
    std::abort();
  }
}

// run `kernel()` for every thread of every block of a 1-D grid
template <class F>
void emul_launch(int grid, int block, F kernel);
void emul_run_block(int block, void (*thunk)(void*), void* ctx);

template <class F>
void emul_launch(int grid, int block, F kernel) {
  gridDim.x = (unsigned)grid;
  blockDim.x = (unsigned)block;
  for (int b = 0; b < grid; ++b) {
    blockIdx.x = (unsigned)b;
    emul_run_block(block, [](void* k) { (*static_cast<F*>(k))(); }, &kernel);
  }
}
