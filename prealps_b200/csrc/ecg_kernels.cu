// ecg_kernels.cu -- K4/K5/K6: the tall-skinny kernels of the ECG iteration.
//
// Replaces the MKL dgemm/dtrsm/dpotrf/domatcopy calls of _preAlps_ECGIterateOdir
// (ref: src/solvers/ecg.c:402-530) by three fused, HBM-bound streaming passes over
// row-major m x t blocks:
//   gram2         two t x t Gram products in one pass (warp-level 4x4 register
//                 tiles, block reduction in shared memory, fixed-order final sum)
//   ortho_update  t x t Cholesky + triangular inverse in shared memory (every CTA
//                 redundantly, t <= 32), then P,AP <- .U^{-1}, X += P a, R -= AP a
//                 and ||R||_F^2 in ONE pass over the four blocks
//   update_z      Z -= P b1 + Pprev b2
// All reductions are two-stage with a fixed grid => bit-reproducible run to run.
#include "common.cuh"

#include <algorithm>

namespace {

constexpr int kThreads = 256;
constexpr int kMaxT = 32;
constexpr int kWaves = 8;  // CTAs per SM launched by the streaming passes (grid = 148 * kWaves, fixed => deterministic sums)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------- gram2
// Thread = one 4x4 tile of (a,b) pairs for one stripe of rows.  NT = tiles per dimension.
template <int NT, bool VEC>
__global__ void __launch_bounds__(kThreads) gram2_kernel(int m, int t, const double* __restrict__ A1, int lda1,
                                                         const double* __restrict__ B1, int ldb1,
                                                         const double* __restrict__ A2, int lda2,
                                                         const double* __restrict__ B2, int ldb2,
                                                         double* __restrict__ partials) {
  constexpr int TILES = NT * NT;          // threads cooperating on one row
  constexpr int RG = kThreads / TILES;    // row stripes per CTA
  const int tid = threadIdx.x;
  const int tile = tid % TILES, rg = tid / TILES;
  const int ia = (tile / NT) * 4, ib = (tile % NT) * 4;
  const bool two = (A2 != nullptr);
  double acc1[4][4], acc2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc1[i][j] = 0.0; acc2[i][j] = 0.0; }

  for (int64_t r = (int64_t)blockIdx.x * RG + rg; r < m; r += (int64_t)gridDim.x * RG) {
    double a[4], b[4];
    if (VEC) {  // t is a multiple of 4 and every row is 16-byte aligned
      const double2 a01 = __ldg(reinterpret_cast<const double2*>(A1 + r * lda1 + ia));
      const double2 a23 = __ldg(reinterpret_cast<const double2*>(A1 + r * lda1 + ia + 2));
      const double2 b01 = __ldg(reinterpret_cast<const double2*>(B1 + r * ldb1 + ib));
      const double2 b23 = __ldg(reinterpret_cast<const double2*>(B1 + r * ldb1 + ib + 2));
      a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y;
      b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = (ia + i < t) ? __ldg(A1 + r * lda1 + ia + i) : 0.0;
        b[i] = (ib + i < t) ? __ldg(B1 + r * ldb1 + ib + i) : 0.0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc1[i][j] = fma(a[i], b[j], acc1[i][j]);
    if (two) {
      if (VEC) {
        const double2 a01 = __ldg(reinterpret_cast<const double2*>(A2 + r * lda2 + ia));
        const double2 a23 = __ldg(reinterpret_cast<const double2*>(A2 + r * lda2 + ia + 2));
        const double2 b01 = __ldg(reinterpret_cast<const double2*>(B2 + r * ldb2 + ib));
        const double2 b23 = __ldg(reinterpret_cast<const double2*>(B2 + r * ldb2 + ib + 2));
        a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y;
        b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a[i] = (ia + i < t) ? __ldg(A2 + r * lda2 + ia + i) : 0.0;
          b[i] = (ib + i < t) ? __ldg(B2 + r * ldb2 + ib + i) : 0.0;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = fma(a[i], b[j], acc2[i][j]);
    }
  }
  // block reduction over the RG stripes, fixed order
  __shared__ double red[kThreads * 16];
  const int ngram = two ? 2 : 1;
  for (int g = 0; g < ngram; ++g) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(rg * TILES + tile) * 16 + i * 4 + j] = g == 0 ? acc1[i][j] : acc2[i][j];
    __syncthreads();
    for (int e = tid; e < TILES * 16; e += kThreads) {
      double s = 0.0;
      for (int q = 0; q < RG; ++q) s += red[q * TILES * 16 + e];
      const int tl = e / 16, i = (e % 16) / 4, j = e % 4;
      const int aa = (tl / NT) * 4 + i, bb = (tl % NT) * 4 + j;
      if (aa < t && bb < t) partials[((size_t)blockIdx.x * ngram + g) * (size_t)(t * t) + aa + (size_t)bb * t] = s;
    }
  }
}

// out[e] = sum_b partials[b][e]: one warp per element, lane l adds b = l, l+32, ... in ascending order, then a
// fixed shuffle tree -- the order never depends on timing, so the result is bit-reproducible
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblocks, int n,
                                       double* __restrict__ out1, int n1, double* __restrict__ out2) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x & 31;
  if (e >= n) return;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partials[(size_t)b * n + e];
  s = warp_sum(s);
  if (lane == 0) { if (e < n1) out1[e] = s; else out2[e - n1] = s; }
}

// -------------------------------------------------------- small t x t algebra in shared memory
// S (col-major t x t, upper triangle valid) -> U with S = U^T U, in place; returns 0 or j+1 of the
// first non-positive pivot in *fail (like dpotrf's info).
__device__ void smem_chol_upper(double* S, int t, int* fail) {
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int j = 0; j < t; ++j) {
    if (tid == 0) {
      const double d = S[j + j * t];
      if (!(d > 0.0)) { if (*fail == 0) *fail = j + 1; S[j + j * t] = 1.0; }
      else S[j + j * t] = sqrt(d);
    }
    __syncthreads();
    if (tid > j && tid < t) S[j + tid * t] /= S[j + j * t];
    __syncthreads();
    for (int idx = tid; idx < t * t; idx += nth) {
      const int k = idx % t, i = idx / t;  // element (k,i), k <= i
      if (k > j && i >= k) S[k + i * t] -= S[j + k * t] * S[j + i * t];
    }
    __syncthreads();
  }
}

// Ui = U^{-1} (upper triangular, col-major); thread i builds column i by back substitution
__device__ void smem_triu_inverse(const double* U, double* Ui, int t) {
  const int i = threadIdx.x;
  if (i < t) {
    for (int k = 0; k < t; ++k) Ui[k + i * t] = 0.0;
    Ui[i + i * t] = 1.0 / U[i + i * t];
    for (int k = i - 1; k >= 0; --k) {
      double s = 0.0;
      for (int l = k + 1; l <= i; ++l) s += U[k + l * t] * Ui[l + i * t];
      Ui[k + i * t] = -s / U[k + k * t];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------ ortho_update
// Row-per-thread streaming pass.  T = compile-time bound on t (registers), t = run-time width.
template <int T>
__global__ void __launch_bounds__(kThreads) ortho_update_kernel(int m, int t, const double* __restrict__ G,
                                                                const double* __restrict__ Gpr, double* P, int ldp,
                                                                double* AP, int ldap, double* X, int ldx, double* R,
                                                                int ldr, double* U_out, double* alpha_out,
                                                                double* __restrict__ rr_partials, int* status) {
  __shared__ double sU[T * T];
  __shared__ double sUi[T * T];
  __shared__ double sAl[T * T];
  __shared__ double sred[kThreads / 32];
  __shared__ int sfail;
  const int tid = threadIdx.x;
  if (tid == 0) sfail = 0;
  for (int e = tid; e < t * t; e += kThreads) {
    const int a = e % t, b = e / t;
    sU[e] = (a <= b) ? G[e] : 0.0;  // upper triangle of AP^T P (ref: ecg.c:431 'U')
  }
  __syncthreads();
  smem_chol_upper(sU, t, &sfail);
  smem_triu_inverse(sU, sUi, t);
  const bool upd = (X != nullptr);
  if (upd) {
    // alpha = U^{-T} Gpr : alpha[a][c] = sum_{k<=a} Ui[k][a] * Gpr[k][c]
    for (int e = tid; e < t * t; e += kThreads) {
      const int a = e % t, c = e / t;
      double s = 0.0;
      for (int k = 0; k <= a; ++k) s += sUi[k + a * t] * Gpr[k + c * t];
      sAl[e] = s;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int e = tid; e < t * t; e += kThreads) {
      if (U_out) U_out[e] = sU[e];
      if (upd && alpha_out) alpha_out[e] = sAl[e];
    }
    if (tid == 0 && status) status[0] = sfail;
  }

  double rr = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + tid; r < m; r += (int64_t)gridDim.x * kThreads) {
    double p[T], q[T];
    // ---- P row: q = p * Ui ; X row += q * alpha
#pragma unroll
    for (int a = 0; a < T; ++a) p[a] = (a < t) ? P[r * ldp + a] : 0.0;
#pragma unroll
    for (int b = 0; b < T; ++b) {
      double s = 0.0;
      if (b < t) {
#pragma unroll
        for (int a = 0; a < T; ++a) if (a <= b) s = fma(p[a], sUi[a + b * t], s);
      }
      q[b] = s;
    }
#pragma unroll
    for (int b = 0; b < T; ++b) if (b < t) P[r * ldp + b] = q[b];
    if (upd) {
#pragma unroll
      for (int c = 0; c < T; ++c) {
        if (c < t) {
          double s = X[r * ldx + c];
#pragma unroll
          for (int a = 0; a < T; ++a) if (a < t) s = fma(q[a], sAl[a + c * t], s);
          X[r * ldx + c] = s;
        }
      }
    }
    // ---- AP row: q = ap * Ui ; R row -= q * alpha
#pragma unroll
    for (int a = 0; a < T; ++a) p[a] = (a < t) ? AP[r * ldap + a] : 0.0;
#pragma unroll
    for (int b = 0; b < T; ++b) {
      double s = 0.0;
      if (b < t) {
#pragma unroll
        for (int a = 0; a < T; ++a) if (a <= b) s = fma(p[a], sUi[a + b * t], s);
      }
      q[b] = s;
    }
#pragma unroll
    for (int b = 0; b < T; ++b) if (b < t) AP[r * ldap + b] = q[b];
    if (upd) {
#pragma unroll
      for (int c = 0; c < T; ++c) {
        if (c < t) {
          double s = R[r * ldr + c];
#pragma unroll
          for (int a = 0; a < T; ++a) if (a < t) s = fma(-q[a], sAl[a + c * t], s);
          R[r * ldr + c] = s;
          rr = fma(s, s, rr);
        }
      }
    }
  }
  if (upd && rr_partials) {
    rr = warp_sum(rr);
    if ((tid & 31) == 0) sred[tid >> 5] = rr;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) s += sred[w];
      rr_partials[blockIdx.x] = s;
    }
  }
}

// ------------------------------------------------------------------ lanes-per-row variants (t == T, T even)
// G = T/2 lanes own one row, each lane two adjacent columns (one 16-byte access per block and row): a warp
// touches 32/G consecutive rows = 512 contiguous bytes per instruction, the ideal 4 wavefronts.  The small
// t x t products run over the group with warp shuffles; the matrices sit transposed (row-major) in shared
// memory so that a lane's two coefficients are one conflict-free 16-byte read.
template <int T>
__device__ __forceinline__ double2 row_times(const double2 v, const double* __restrict__ Mt, int lig, double2 acc,
                                             double sign) {
  constexpr int G = T / 2;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const double ax = sign * __shfl_sync(0xffffffffu, v.x, g, G);
    const double ay = sign * __shfl_sync(0xffffffffu, v.y, g, G);
    const double2 m0 = *reinterpret_cast<const double2*>(Mt + (2 * g) * T + 2 * lig);
    const double2 m1 = *reinterpret_cast<const double2*>(Mt + (2 * g + 1) * T + 2 * lig);
    acc.x = fma(ax, m0.x, acc.x);
    acc.y = fma(ax, m0.y, acc.y);
    acc.x = fma(ay, m1.x, acc.x);
    acc.y = fma(ay, m1.y, acc.y);
  }
  return acc;
}

template <int T>
__global__ void __launch_bounds__(kThreads) ortho_update_v2_kernel(int m, const double* __restrict__ G_,
                                                                   const double* __restrict__ Gpr, double* P, int ldp,
                                                                   double* AP, int ldap, double* X, int ldx, double* R,
                                                                   int ldr, double* U_out, double* alpha_out,
                                                                   double* __restrict__ rr_partials, int* status) {
  constexpr int G = T / 2, RPW = 32 / G, RPB = kThreads / G;
  __shared__ double sU[T * T];
  __shared__ double sUi[T * T];
  __shared__ __align__(16) double sUiT[T * T];  // Ui transposed: [a*T + b] = Ui(a,b)
  __shared__ __align__(16) double sAlT[T * T];  // alpha transposed: [a*T + c] = alpha(a,c)
  __shared__ double sred[kThreads / 32];
  __shared__ int sfail;
  const int tid = threadIdx.x;
  if (tid == 0) sfail = 0;
  for (int e = tid; e < T * T; e += kThreads) {
    const int a = e % T, b = e / T;
    sU[e] = (a <= b) ? G_[e] : 0.0;
  }
  __syncthreads();
  smem_chol_upper(sU, T, &sfail);
  smem_triu_inverse(sU, sUi, T);
  const bool upd = (X != nullptr);
  for (int e = tid; e < T * T; e += kThreads) {
    const int a = e % T, c = e / T;  // column-major index e = a + c*T
    sUiT[a * T + c] = sUi[e];
    double s = 0.0;
    if (upd) for (int k = 0; k <= a; ++k) s += sUi[k + a * T] * Gpr[k + c * T];
    sAlT[a * T + c] = s;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int e = tid; e < T * T; e += kThreads) {
      if (U_out) U_out[e] = sU[e];
      if (upd && alpha_out) alpha_out[e] = sAlT[(e % T) * T + e / T];
    }
    if (tid == 0 && status) status[0] = sfail;
  }
  const int lane = tid & 31, lig = lane % G;
  double rr = 0.0;
  const double2 zero = make_double2(0.0, 0.0);
  const int64_t stride = (int64_t)gridDim.x * RPB;
  // two row groups per trip, every load issued before the first use: 8 x 16 B in flight per lane
  for (int64_t rb = (int64_t)blockIdx.x * RPB + (tid >> 5) * RPW; rb < m; rb += 2 * stride) {
    int64_t r[2];
    bool act[2];
    double2 p[2], ap[2], x[2], rv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      r[u] = rb + u * stride + lane / G;
      act[u] = r[u] < m;
      p[u] = act[u] ? *reinterpret_cast<const double2*>(P + r[u] * ldp + 2 * lig) : zero;
      ap[u] = act[u] ? *reinterpret_cast<const double2*>(AP + r[u] * ldap + 2 * lig) : zero;
      x[u] = (act[u] && upd) ? *reinterpret_cast<const double2*>(X + r[u] * ldx + 2 * lig) : zero;
      rv[u] = (act[u] && upd) ? *reinterpret_cast<const double2*>(R + r[u] * ldr + 2 * lig) : zero;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const double2 q = row_times<T>(p[u], sUiT, lig, zero, 1.0);
      const double2 q2 = row_times<T>(ap[u], sUiT, lig, zero, 1.0);
      if (upd) {
        x[u] = row_times<T>(q, sAlT, lig, x[u], 1.0);
        rv[u] = row_times<T>(q2, sAlT, lig, rv[u], -1.0);
      }
      if (act[u]) {
        *reinterpret_cast<double2*>(P + r[u] * ldp + 2 * lig) = q;
        *reinterpret_cast<double2*>(AP + r[u] * ldap + 2 * lig) = q2;
        if (upd) {
          *reinterpret_cast<double2*>(X + r[u] * ldx + 2 * lig) = x[u];
          *reinterpret_cast<double2*>(R + r[u] * ldr + 2 * lig) = rv[u];
          rr = fma(rv[u].x, rv[u].x, rr);
          rr = fma(rv[u].y, rv[u].y, rr);
        }
      }
    }
  }
  if (upd && rr_partials) {
    rr = warp_sum(rr);
    if ((tid & 31) == 0) sred[tid >> 5] = rr;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) s += sred[w];
      rr_partials[blockIdx.x] = s;
    }
  }
}

// Z -= P b1 + Pprev b2 (all T wide)
template <int T>
__global__ void __launch_bounds__(kThreads) update_z_v2_kernel(int m, double* Z, int ldz, const double* __restrict__ P,
                                                               int ldp, const double* __restrict__ beta1,
                                                               const double* __restrict__ Pp, int ldpp,
                                                               const double* __restrict__ beta2) {
  constexpr int G = T / 2, RPW = 32 / G, RPB = kThreads / G;
  __shared__ __align__(16) double sB1[T * T];
  __shared__ __align__(16) double sB2[T * T];
  const int tid = threadIdx.x;
  for (int e = tid; e < T * T; e += kThreads) {
    const int a = e % T, c = e / T;
    sB1[a * T + c] = beta1[e];
    sB2[a * T + c] = Pp ? beta2[e] : 0.0;
  }
  __syncthreads();
  const int lane = tid & 31, lig = lane % G;
  const double2 zero = make_double2(0.0, 0.0);
  const int64_t stride = (int64_t)gridDim.x * RPB;
  for (int64_t rb = (int64_t)blockIdx.x * RPB + (tid >> 5) * RPW; rb < m; rb += 2 * stride) {
    int64_t r[2];
    bool act[2];
    double2 z[2], p[2], pp[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      r[u] = rb + u * stride + lane / G;
      act[u] = r[u] < m;
      z[u] = act[u] ? *reinterpret_cast<const double2*>(Z + r[u] * ldz + 2 * lig) : zero;
      p[u] = act[u] ? *reinterpret_cast<const double2*>(P + r[u] * ldp + 2 * lig) : zero;
      pp[u] = (act[u] && Pp) ? *reinterpret_cast<const double2*>(Pp + r[u] * ldpp + 2 * lig) : zero;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      z[u] = row_times<T>(p[u], sB1, lig, z[u], -1.0);
      if (Pp) z[u] = row_times<T>(pp[u], sB2, lig, z[u], -1.0);
      if (act[u]) *reinterpret_cast<double2*>(Z + r[u] * ldz + 2 * lig) = z[u];
    }
  }
}

// ------------------------------------------------------------------ FP64 tensor-core (DMMA) variants, T = 8 or 16
// A rows-times-(t x t) product is a skinny GEMM.  mma.sync.m8n8k4.f64 computes an 8-row x 8-column block per warp:
//   A fragment: lane holds A(row = lane/4, k = lane%4)        -> one 8-byte load per lane and k-step
//   B fragment: lane holds B(k = lane%4, n = lane/4)          -> the small matrix, kept in registers
//   C fragment: lane holds C(row = lane/4, 2*(lane%4) + {0,1}) -> one coalesced 16-byte access
// so the products need NO shuffles and NO shared-memory reads in the loop (the CUDA-core versions above are
// bound by exactly those: ncu showed l1tex at 50-80 % with DRAM at 44-52 %), and only 2 accumulator doubles per
// lane and 8x8 block, which leaves the registers for loads in flight.
#ifndef PCU_EMUL
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#else  // tests/emul: the same product as a warp-wide exchange (lane holds A[lane/4][lane%4], B[lane%4][lane/4], D[lane/4][2*(lane%4)+{0,1}])
inline void dmma884(double& d0, double& d1, double a, double b) {
  double A[32], B[32];
  emul_warp_allgather(a, A);
  emul_warp_allgather(b, B);
  const int lane = (int)(threadIdx.x & 31), row = lane >> 2, n0 = 2 * (lane & 3);
  for (int k = 0; k < 4; ++k) {
    d0 = fma(A[row * 4 + k], B[n0 * 4 + k], d0);
    d1 = fma(A[row * 4 + k], B[(n0 + 1) * 4 + k], d1);
  }
}
#endif

template <int T>
__global__ void __launch_bounds__(kThreads) gram2_mma_kernel(int m, const double* __restrict__ A1, int lda1,
                                                             const double* __restrict__ B1, int ldb1,
                                                             const double* __restrict__ A2, int lda2,
                                                             const double* __restrict__ B2, int ldb2,
                                                             double* __restrict__ partials) {
  constexpr int NB = T / 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane & 3, lc = lane >> 2;  // row within the 4-row group, column within the 8-column block
  const bool two = (A2 != nullptr);
  double c1[NB][NB][2], c2[NB][NB][2];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) { c1[i][j][0] = c1[i][j][1] = 0.0; c2[i][j][0] = c2[i][j][1] = 0.0; }
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t ngroups = (m + 3) / 4;
  constexpr int U = 4;  // 4-row groups per trip: all loads first
  for (int64_t g0 = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * U; g0 < ngroups; g0 += nwarps * U) {
    double a1[U][NB], b1[U][NB], a2[U][NB], b2[U][NB];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = (g0 + u) * 4 + lr;
      const bool act = r < m;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        a1[u][nb] = act ? __ldg(A1 + r * lda1 + 8 * nb + lc) : 0.0;
        b1[u][nb] = act ? __ldg(B1 + r * ldb1 + 8 * nb + lc) : 0.0;
        if (two) {
          a2[u][nb] = act ? __ldg(A2 + r * lda2 + 8 * nb + lc) : 0.0;
          b2[u][nb] = act ? __ldg(B2 + r * ldb2 + 8 * nb + lc) : 0.0;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          dmma884(c1[i][j][0], c1[i][j][1], a1[u][i], b1[u][j]);
          if (two) dmma884(c2[i][j][0], c2[i][j][1], a2[u][i], b2[u][j]);
        }
  }
  // block reduction over the warps in fixed order; element (a, b) with a = 8i + lane/4, b = 8j + 2*(lane%4) + e
  __shared__ double red[(kThreads / 32) * 2 * T * T];
  const int ngram = two ? 2 : 1;
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int a = 8 * i + lc, b = 8 * j + 2 * lr + e;
        red[(warp * 2 + 0) * T * T + a + b * T] = c1[i][j][e];
        red[(warp * 2 + 1) * T * T + a + b * T] = c2[i][j][e];
      }
  __syncthreads();
  for (int e = tid; e < ngram * T * T; e += kThreads) {
    const int g = e / (T * T), idx = e % (T * T);
    double sacc = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) sacc += red[(w * 2 + g) * T * T + idx];
    partials[((size_t)blockIdx.x * ngram + g) * (size_t)(T * T) + idx] = sacc;
  }
}

// P <- P Ui, AP <- AP Ui, X += P_old W, R -= AP_old W with W = Ui alpha; rr = ||R||^2
template <int T>
__global__ void __launch_bounds__(kThreads) ortho_update_mma_kernel(int m, const double* __restrict__ G_,
                                                                    const double* __restrict__ Gpr, double* __restrict__ P,
                                                                    int ldp, double* __restrict__ AP, int ldap,
                                                                    double* __restrict__ X, int ldx, double* __restrict__ R,
                                                                    int ldr, double* U_out, double* alpha_out,
                                                                    double* __restrict__ rr_partials, int* status) {
  constexpr int NB = T / 8, KB = T / 4;
  __shared__ double sU[T * T];
  __shared__ double sUi[T * T];
  __shared__ double sAl[T * T];
  __shared__ double sW[T * T];
  __shared__ double sred[kThreads / 32];
  __shared__ int sfail;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) sfail = 0;
  for (int e = tid; e < T * T; e += kThreads) sU[e] = ((e % T) <= (e / T)) ? G_[e] : 0.0;
  __syncthreads();
  smem_chol_upper(sU, T, &sfail);
  smem_triu_inverse(sU, sUi, T);
  const bool upd = (X != nullptr);
  for (int e = tid; e < T * T; e += kThreads) {
    const int a = e % T, c = e / T;
    double sacc = 0.0;
    if (upd) for (int k = 0; k <= a; ++k) sacc += sUi[k + a * T] * Gpr[k + c * T];  // alpha = Ui^T Gpr
    sAl[e] = sacc;
  }
  __syncthreads();
  for (int e = tid; e < T * T; e += kThreads) {
    const int a = e % T, c = e / T;
    double sacc = 0.0;
    for (int k = a; k < T; ++k) sacc += sUi[a + k * T] * sAl[k + c * T];          // W = Ui alpha (Ui upper triangular)
    sW[e] = sacc;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int e = tid; e < T * T; e += kThreads) {
      if (U_out) U_out[e] = sU[e];
      if (upd && alpha_out) alpha_out[e] = sAl[e];
    }
    if (tid == 0 && status) status[0] = sfail;
  }
  const int lr = lane >> 2, lk = lane & 3;  // row within the 8-row group, k within the k-step / column pair
  double bu[KB][NB], bw[KB][NB];
#pragma unroll
  for (int kb = 0; kb < KB; ++kb)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      bu[kb][nb] = sUi[(4 * kb + lk) + (8 * nb + lr) * T];   // B(k, n) with k = lane%4, n = lane/4
      bw[kb][nb] = sW[(4 * kb + lk) + (8 * nb + lr) * T];
    }
  double rr = 0.0;
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t ngroups = (m + 7) / 8;
  constexpr int U = (T == 8) ? 2 : 1;
  for (int64_t g0 = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * U; g0 < ngroups; g0 += nwarps * U) {
    double ap_[U][KB], aap[U][KB], cx[U][NB][2], cr[U][NB][2];
    bool act[U];
    int64_t r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r[u] = (g0 + u) * 8 + lr;
      act[u] = r[u] < m;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        ap_[u][kb] = act[u] ? P[r[u] * ldp + 4 * kb + lk] : 0.0;
        aap[u][kb] = act[u] ? AP[r[u] * ldap + 4 * kb + lk] : 0.0;
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        double2 vx = make_double2(0.0, 0.0), vr = make_double2(0.0, 0.0);
        if (act[u] && upd) {
          vx = *reinterpret_cast<const double2*>(X + r[u] * ldx + 8 * nb + 2 * lk);
          vr = *reinterpret_cast<const double2*>(R + r[u] * ldr + 8 * nb + 2 * lk);
        }
        cx[u][nb][0] = vx.x; cx[u][nb][1] = vx.y;
        cr[u][nb][0] = -vr.x; cr[u][nb][1] = -vr.y;   // accumulate -R + AP W, negate on store
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double cp[NB][2], cap[NB][2];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        cp[nb][0] = cp[nb][1] = cap[nb][0] = cap[nb][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          dmma884(cp[nb][0], cp[nb][1], ap_[u][kb], bu[kb][nb]);
          dmma884(cap[nb][0], cap[nb][1], aap[u][kb], bu[kb][nb]);
          if (upd) {
            dmma884(cx[u][nb][0], cx[u][nb][1], ap_[u][kb], bw[kb][nb]);
            dmma884(cr[u][nb][0], cr[u][nb][1], aap[u][kb], bw[kb][nb]);
          }
        }
      }
      if (act[u]) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          *reinterpret_cast<double2*>(P + r[u] * ldp + 8 * nb + 2 * lk) = make_double2(cp[nb][0], cp[nb][1]);
          *reinterpret_cast<double2*>(AP + r[u] * ldap + 8 * nb + 2 * lk) = make_double2(cap[nb][0], cap[nb][1]);
          if (upd) {
            *reinterpret_cast<double2*>(X + r[u] * ldx + 8 * nb + 2 * lk) = make_double2(cx[u][nb][0], cx[u][nb][1]);
            const double r0 = -cr[u][nb][0], r1 = -cr[u][nb][1];
            *reinterpret_cast<double2*>(R + r[u] * ldr + 8 * nb + 2 * lk) = make_double2(r0, r1);
            rr = fma(r0, r0, rr);
            rr = fma(r1, r1, rr);
          }
        }
      }
    }
  }
  if (upd && rr_partials) {
    rr = warp_sum(rr);
    if (lane == 0) sred[warp] = rr;
    __syncthreads();
    if (tid == 0) {
      double sacc = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) sacc += sred[w];
      rr_partials[blockIdx.x] = sacc;
    }
  }
}

// Z -= P b1 + Pprev b2
template <int T>
__global__ void __launch_bounds__(kThreads) update_z_mma_kernel(int m, double* __restrict__ Z, int ldz,
                                                                const double* __restrict__ P, int ldp,
                                                                const double* __restrict__ beta1,
                                                                const double* __restrict__ Pp, int ldpp,
                                                                const double* __restrict__ beta2) {
  constexpr int NB = T / 8, KB = T / 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lk = lane & 3;
  double b1[KB][NB], b2[KB][NB];
#pragma unroll
  for (int kb = 0; kb < KB; ++kb)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      b1[kb][nb] = -beta1[(4 * kb + lk) + (8 * nb + lr) * T];
      b2[kb][nb] = Pp ? -beta2[(4 * kb + lk) + (8 * nb + lr) * T] : 0.0;
    }
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t ngroups = (m + 7) / 8;
  constexpr int U = (T == 8) ? 4 : 2;
  for (int64_t g0 = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * U; g0 < ngroups; g0 += nwarps * U) {
    double a1[U][KB], a2[U][KB], cz[U][NB][2];
    bool act[U];
    int64_t r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r[u] = (g0 + u) * 8 + lr;
      act[u] = r[u] < m;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        a1[u][kb] = act[u] ? __ldg(P + r[u] * ldp + 4 * kb + lk) : 0.0;
        a2[u][kb] = (act[u] && Pp) ? __ldg(Pp + r[u] * ldpp + 4 * kb + lk) : 0.0;
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        double2 v = make_double2(0.0, 0.0);
        if (act[u]) v = *reinterpret_cast<const double2*>(Z + r[u] * ldz + 8 * nb + 2 * lk);
        cz[u][nb][0] = v.x; cz[u][nb][1] = v.y;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          dmma884(cz[u][nb][0], cz[u][nb][1], a1[u][kb], b1[kb][nb]);
          if (Pp) dmma884(cz[u][nb][0], cz[u][nb][1], a2[u][kb], b2[kb][nb]);
        }
      if (act[u]) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          *reinterpret_cast<double2*>(Z + r[u] * ldz + 8 * nb + 2 * lk) = make_double2(cz[u][nb][0], cz[u][nb][1]);
      }
    }
  }
}

// X += P alpha, R -= AP alpha, rr = ||R||^2
template <int T>
__global__ void __launch_bounds__(kThreads) update_xr_kernel(int m, int t, const double* __restrict__ P, int ldp,
                                                             const double* __restrict__ AP, int ldap,
                                                             const double* __restrict__ alpha, double* X, int ldx,
                                                             double* R, int ldr, double* __restrict__ rr_partials) {
  __shared__ double sAl[T * T];
  __shared__ double sred[kThreads / 32];
  const int tid = threadIdx.x;
  for (int e = tid; e < t * t; e += kThreads) sAl[e] = alpha[e];
  __syncthreads();
  double rr = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + tid; r < m; r += (int64_t)gridDim.x * kThreads) {
    double p[T];
#pragma unroll
    for (int a = 0; a < T; ++a) p[a] = (a < t) ? P[r * ldp + a] : 0.0;
#pragma unroll
    for (int c = 0; c < T; ++c) if (c < t) {
      double s = X[r * ldx + c];
#pragma unroll
      for (int a = 0; a < T; ++a) if (a < t) s = fma(p[a], sAl[a + c * t], s);
      X[r * ldx + c] = s;
    }
#pragma unroll
    for (int a = 0; a < T; ++a) p[a] = (a < t) ? AP[r * ldap + a] : 0.0;
#pragma unroll
    for (int c = 0; c < T; ++c) if (c < t) {
      double s = R[r * ldr + c];
#pragma unroll
      for (int a = 0; a < T; ++a) if (a < t) s = fma(-p[a], sAl[a + c * t], s);
      R[r * ldr + c] = s;
      rr = fma(s, s, rr);
    }
  }
  rr = warp_sum(rr);
  if ((tid & 31) == 0) sred[tid >> 5] = rr;
  __syncthreads();
  if (tid == 0 && rr_partials) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += sred[w];
    rr_partials[blockIdx.x] = s;
  }
}

// ADAPT_BS after a reduction: P <- P Wm, AP <- AP Wm (in place, row by row), X += P_old Wx, R -= AP_old Wx, rr = ||R||^2,
// with Wm and Wx general t x t matrices built on the host (column-major, tight ld).  AP / X / R may be null.
// One pass over the four blocks instead of memset + product into a scratch block + copy back for each of P and AP and a
// separate X / R update (19 block passes -> 8).
template <int T>
__global__ void __launch_bounds__(kThreads) transform_update_kernel(int m, int t, const double* __restrict__ Wm,
                                                                    const double* __restrict__ Wx, double* P, int ldp,
                                                                    double* AP, int ldap, double* X, int ldx, double* R,
                                                                    int ldr, double* __restrict__ rr_partials) {
  __shared__ double sWm[T * T];
  __shared__ double sWx[T * T];
  __shared__ double sred[kThreads / 32];
  const int tid = threadIdx.x;
  for (int e = tid; e < t * t; e += kThreads) { sWm[e] = Wm[e]; sWx[e] = Wx ? Wx[e] : 0.0; }
  __syncthreads();
  double rr = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + tid; r < m; r += (int64_t)gridDim.x * kThreads) {
    double p[T];
#pragma unroll
    for (int a = 0; a < T; ++a) p[a] = (a < t) ? P[r * ldp + a] : 0.0;
    if (X) {
#pragma unroll
      for (int c = 0; c < T; ++c) if (c < t) {
        double s = X[r * ldx + c];
#pragma unroll
        for (int a = 0; a < T; ++a) if (a < t) s = fma(p[a], sWx[a + c * t], s);
        X[r * ldx + c] = s;
      }
    }
#pragma unroll
    for (int c = 0; c < T; ++c) if (c < t) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < T; ++a) if (a < t) s = fma(p[a], sWm[a + c * t], s);
      P[r * ldp + c] = s;
    }
    if (AP) {
#pragma unroll
      for (int a = 0; a < T; ++a) p[a] = (a < t) ? AP[r * ldap + a] : 0.0;
      if (R) {
#pragma unroll
        for (int c = 0; c < T; ++c) if (c < t) {
          double s = R[r * ldr + c];
#pragma unroll
          for (int a = 0; a < T; ++a) if (a < t) s = fma(-p[a], sWx[a + c * t], s);
          R[r * ldr + c] = s;
          rr = fma(s, s, rr);
        }
      }
#pragma unroll
      for (int c = 0; c < T; ++c) if (c < t) {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < T; ++a) if (a < t) s = fma(p[a], sWm[a + c * t], s);
        AP[r * ldap + c] = s;
      }
    }
  }
  rr = warp_sum(rr);
  if ((tid & 31) == 0) sred[tid >> 5] = rr;
  __syncthreads();
  if (tid == 0 && rr_partials) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += sred[w];
    rr_partials[blockIdx.x] = s;
  }
}

// Z -= P b1 + Pprev b2 ; b1 is t1 x tz, b2 is t2 x tz, both column-major with tight ld
template <int T>
__global__ void __launch_bounds__(kThreads) update_z_kernel(int m, int tz, double* Z, int ldz,
                                                            const double* __restrict__ P, int ldp, int t1,
                                                            const double* __restrict__ beta1,
                                                            const double* __restrict__ Pp, int ldpp, int t2,
                                                            const double* __restrict__ beta2) {
  __shared__ double sB1[T * T];
  __shared__ double sB2[T * T];
  const int tid = threadIdx.x;
  for (int e = tid; e < t1 * tz; e += kThreads) sB1[e] = beta1[e];
  for (int e = tid; e < t2 * tz; e += kThreads) sB2[e] = beta2[e];
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * kThreads + tid; r < m; r += (int64_t)gridDim.x * kThreads) {
    double z[T], p[T];
#pragma unroll
    for (int c = 0; c < T; ++c) z[c] = (c < tz) ? Z[r * ldz + c] : 0.0;
#pragma unroll
    for (int a = 0; a < T; ++a) p[a] = (a < t1) ? P[r * ldp + a] : 0.0;
#pragma unroll
    for (int c = 0; c < T; ++c) if (c < tz) {
      double s = z[c];
#pragma unroll
      for (int a = 0; a < T; ++a) if (a < t1) s = fma(-p[a], sB1[a + c * t1], s);
      z[c] = s;
    }
    if (t2 > 0) {
#pragma unroll
      for (int a = 0; a < T; ++a) p[a] = (a < t2) ? Pp[r * ldpp + a] : 0.0;
#pragma unroll
      for (int c = 0; c < T; ++c) if (c < tz) {
        double s = z[c];
#pragma unroll
        for (int a = 0; a < T; ++a) if (a < t2) s = fma(-p[a], sB2[a + c * t2], s);
        z[c] = s;
      }
    }
#pragma unroll
    for (int c = 0; c < T; ++c) if (c < tz) Z[r * ldz + c] = z[c];
  }
}

__global__ void sum_columns_kernel(int m, int t, const double* __restrict__ X, int ldx, double* __restrict__ sol) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < t; ++c) s += X[r * ldx + c];  // same order as dgemv with a vector of ones
    sol[r] = s;
  }
}

__global__ void split_rhs_kernel(int m, int t, const double* __restrict__ rhs, const int* __restrict__ col_of_row,
                                 double* __restrict__ R, int ldr) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
    const int cc = col_of_row[r];
    for (int c = 0; c < t; ++c) R[r * ldr + c] = (c == cc) ? rhs[r] : 0.0;
  }
}

__global__ void __launch_bounds__(kThreads) fro2_kernel(int m, int t, const double* __restrict__ R, int ldr,
                                                        double* __restrict__ partials) {
  __shared__ double sred[kThreads / 32];
  const int tid = threadIdx.x;
  double s = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * kThreads + tid; r < m; r += (int64_t)gridDim.x * kThreads)
    for (int c = 0; c < t; ++c) { const double v = R[r * ldr + c]; s = fma(v, v, s); }
  s = warp_sum(s);
  if ((tid & 31) == 0) sred[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double q = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) q += sred[w];
    partials[blockIdx.x] = q;
  }
}

// ORTHODIR_FUSED small-matrix step (ref: ecg.c:577-587): U = chol(mu); beta1 <- U^-T beta1 U^-1; beta2 <- beta2 U^-1
__global__ void __launch_bounds__(kThreads) fused_small_kernel(int t, const double* __restrict__ mu, double* beta1,
                                                               double* beta2, double* U_out, int* status) {
  __shared__ double sU[kMaxT * kMaxT], sUi[kMaxT * kMaxT], sT[kMaxT * kMaxT], sT2[kMaxT * kMaxT];
  __shared__ int sfail;
  const int tid = threadIdx.x;
  if (tid == 0) sfail = 0;
  for (int e = tid; e < t * t; e += kThreads) sU[e] = ((e % t) <= (e / t)) ? mu[e] : 0.0;
  __syncthreads();
  smem_chol_upper(sU, t, &sfail);
  smem_triu_inverse(sU, sUi, t);
  // sT = beta1 Ui ; beta1' = Ui^T sT
  for (int e = tid; e < t * t; e += kThreads) {
    const int a = e % t, c = e / t;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k <= c; ++k) { s1 += beta1[a + k * t] * sUi[k + c * t]; s2 += beta2[a + k * t] * sUi[k + c * t]; }
    sT[e] = s1;
    sT2[e] = s2;
  }
  __syncthreads();
  for (int e = tid; e < t * t; e += kThreads) {
    const int a = e % t, c = e / t;
    double s1 = 0.0;
    for (int k = 0; k <= a; ++k) s1 += sUi[k + a * t] * sT[k + c * t];
    beta1[e] = s1;
    beta2[e] = sT2[e];
    if (U_out) U_out[e] = ((e % t) <= (e / t)) ? sU[e] : 0.0;
  }
  if (tid == 0 && status) status[0] = sfail;
}

// Z <- Z U^-1 with U = chol_upper(G): row-per-thread, the t x t factor recomputed per CTA in shared memory
template <int T>
__global__ void __launch_bounds__(kThreads) right_solve_kernel(int m, int t, const double* __restrict__ G, double* Z, int ldz) {
  __shared__ double sU[T * T], sUi[T * T];
  __shared__ int sfail;
  const int tid = threadIdx.x;
  if (tid == 0) sfail = 0;
  for (int e = tid; e < t * t; e += kThreads) sU[e] = ((e % t) <= (e / t)) ? G[e] : 0.0;
  __syncthreads();
  smem_chol_upper(sU, t, &sfail);
  smem_triu_inverse(sU, sUi, t);
  for (int64_t r = (int64_t)blockIdx.x * kThreads + tid; r < m; r += (int64_t)gridDim.x * kThreads) {
    double z[T], q[T];
#pragma unroll
    for (int a = 0; a < T; ++a) z[a] = (a < t) ? Z[r * ldz + a] : 0.0;
#pragma unroll
    for (int b = 0; b < T; ++b) {
      double sacc = 0.0;
      if (b < t) {
#pragma unroll
        for (int a = 0; a < T; ++a) if (a <= b) sacc = fma(z[a], sUi[a + b * t], sacc);
      }
      q[b] = sacc;
    }
#pragma unroll
    for (int b = 0; b < T; ++b) if (b < t) Z[r * ldz + b] = q[b];
  }
}

int pick_T(int t) { return t <= 1 ? 1 : t <= 2 ? 2 : t <= 4 ? 4 : t <= 8 ? 8 : t <= 16 ? 16 : 32; }

}  // namespace

using namespace pcu;

#define DISPATCH_T(T_, ...)                      \
  switch (T_) {                                   \
    case 1: { constexpr int TT = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int TT = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int TT = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int TT = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int TT = 16; __VA_ARGS__; } break; \
    default: { constexpr int TT = 32; __VA_ARGS__; } break; \
  }

extern "C" {

int pcu_gram2(pcu_ctx* c, int m, int t, const double* A1, int lda1, const double* B1, int ldb1, double* G1,
              const double* A2, int lda2, const double* B2, int ldb2, double* G2) {
  PCU_CHECK(c && A1 && B1 && G1 && t >= 1 && t <= kMaxT, "pcu_gram2: bad arguments (t=%d)", t);
  const int ngram = A2 ? 2 : 1;
  PCU_CHECK(!A2 || (B2 && G2), "pcu_gram2: second pair incomplete");
  const int nt = (t + 3) / 4;  // tiles per dimension: 1, 2, 4 or 8
  const int NT = nt <= 1 ? 1 : nt <= 2 ? 2 : nt <= 4 ? 4 : 8;
  const int rows_per_cta = kThreads / (NT * NT);
  const int grid = stream_grid(c, m, rows_per_cta * 8, kWaves);
  if (ensure_partials(c, (size_t)grid * ngram * t * t + 16)) return 1;
  const bool vec = (t % 4 == 0) && lda1 % 2 == 0 && ldb1 % 2 == 0 && (!A2 || (lda2 % 2 == 0 && ldb2 % 2 == 0)) &&
                   ((uintptr_t)A1 % 16 == 0) && ((uintptr_t)B1 % 16 == 0) && ((uintptr_t)A2 % 16 == 0) &&
                   ((uintptr_t)B2 % 16 == 0);
  const bool use_mma = vec && (t == 8 || t == 16) && !getenv("PREALPS_ECG_NOMMA");
  if (use_mma) {
    const int gridm = stream_grid(c, (m + 3) / 4, (kThreads / 32) * 4 * 4, kWaves);
    if (ensure_partials(c, (size_t)gridm * ngram * t * t + 16)) return 1;
    if (t == 8) gram2_mma_kernel<8><<<gridm, kThreads, 0, c->stream>>>(m, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, c->red_partials);
    else gram2_mma_kernel<16><<<gridm, kThreads, 0, c->stream>>>(m, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, c->red_partials);
    PCU_LAUNCH_CHECK(c);
    const int nn = ngram * t * t;
    reduce_partials_kernel<<<ceil_div(nn * 32, 256), 256, 0, c->stream>>>(c->red_partials, gridm, nn, G1, t * t, G2);
    PCU_LAUNCH_CHECK(c);
    return 0;
  }
#define GRAM_LAUNCH(NT_, V_) gram2_kernel<NT_, V_><<<grid, kThreads, 0, c->stream>>>(m, t, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, c->red_partials)
  switch (NT) {
    case 1: if (vec) GRAM_LAUNCH(1, true); else GRAM_LAUNCH(1, false); break;
    case 2: if (vec) GRAM_LAUNCH(2, true); else GRAM_LAUNCH(2, false); break;
    case 4: if (vec) GRAM_LAUNCH(4, true); else GRAM_LAUNCH(4, false); break;
    default: if (vec) GRAM_LAUNCH(8, true); else GRAM_LAUNCH(8, false); break;
  }
#undef GRAM_LAUNCH
  PCU_LAUNCH_CHECK(c);
  const int n = ngram * t * t;
  reduce_partials_kernel<<<ceil_div(n * 32, 256), 256, 0, c->stream>>>(c->red_partials, grid, n, G1, t * t, G2);
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_ortho_update(pcu_ctx* c, int m, int t, const double* G, const double* Gpr, double* P, int ldp, double* AP,
                     int ldap, double* X, int ldx, double* R, int ldr, double* U_out, double* alpha_out, double* rr,
                     int* status_dev) {
  PCU_CHECK(c && G && P && AP && t >= 1 && t <= kMaxT, "pcu_ortho_update: bad arguments (t=%d)", t);
  PCU_CHECK((X == nullptr) == (R == nullptr), "pcu_ortho_update: X and R must both be given or both be NULL");
  PCU_CHECK(!X || Gpr, "pcu_ortho_update: Gpr missing");
  const bool v2 = (t == 2 || t == 4 || t == 8 || t == 16 || t == 32) && ldp % 2 == 0 && ldap % 2 == 0 &&
                  (!X || (ldx % 2 == 0 && ldr % 2 == 0)) && ((uintptr_t)P % 16 == 0) && ((uintptr_t)AP % 16 == 0) &&
                  ((uintptr_t)X % 16 == 0) && ((uintptr_t)R % 16 == 0) && !getenv("PREALPS_ECG_V1");
  const int grid = stream_grid(c, m, v2 ? kThreads / (t / 2) * 4 : kThreads, kWaves);
  if (ensure_partials(c, (size_t)grid + 16)) return 1;
  if (v2 && (t == 8 || t == 16) && !getenv("PREALPS_ECG_NOMMA")) {
    const int gridm = stream_grid(c, (m + 7) / 8, (kThreads / 32) * 4, kWaves);
    if (ensure_partials(c, (size_t)gridm + 16)) return 1;
    if (t == 8) ortho_update_mma_kernel<8><<<gridm, kThreads, 0, c->stream>>>(m, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev);
    else ortho_update_mma_kernel<16><<<gridm, kThreads, 0, c->stream>>>(m, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev);
    PCU_LAUNCH_CHECK(c);
    if (X && rr) {
      reduce_partials_kernel<<<1, 32, 0, c->stream>>>(c->red_partials, gridm, 1, rr, 1, nullptr);
      PCU_LAUNCH_CHECK(c);
    }
    return 0;
  }
  if (v2) {
    switch (t) {
      case 2: ortho_update_v2_kernel<2><<<grid, kThreads, 0, c->stream>>>(m, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev); break;
      case 4: ortho_update_v2_kernel<4><<<grid, kThreads, 0, c->stream>>>(m, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev); break;
      case 8: ortho_update_v2_kernel<8><<<grid, kThreads, 0, c->stream>>>(m, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev); break;
      case 16: ortho_update_v2_kernel<16><<<grid, kThreads, 0, c->stream>>>(m, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev); break;
      default: ortho_update_v2_kernel<32><<<grid, kThreads, 0, c->stream>>>(m, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev); break;
    }
  } else
  DISPATCH_T(pick_T(t), ortho_update_kernel<TT><<<grid, kThreads, 0, c->stream>>>(
                            m, t, G, Gpr, P, ldp, AP, ldap, X, ldx, R, ldr, U_out, alpha_out, c->red_partials, status_dev));
  PCU_LAUNCH_CHECK(c);
  if (X && rr) {
    reduce_partials_kernel<<<1, 32, 0, c->stream>>>(c->red_partials, grid, 1, rr, 1, nullptr);
    PCU_LAUNCH_CHECK(c);
  }
  return 0;
}

int pcu_update_xr(pcu_ctx* c, int m, int t, const double* P, int ldp, const double* AP, int ldap, const double* alpha,
                  double* X, int ldx, double* R, int ldr, double* rr) {
  PCU_CHECK(c && P && AP && alpha && X && R && t >= 1 && t <= kMaxT, "pcu_update_xr: bad arguments (t=%d)", t);
  const int grid = stream_grid(c, m, kThreads, kWaves);
  if (ensure_partials(c, (size_t)grid + 16)) return 1;
  DISPATCH_T(pick_T(t), update_xr_kernel<TT><<<grid, kThreads, 0, c->stream>>>(m, t, P, ldp, AP, ldap, alpha, X, ldx, R,
                                                                              ldr, c->red_partials));
  PCU_LAUNCH_CHECK(c);
  if (rr) {
    reduce_partials_kernel<<<1, 32, 0, c->stream>>>(c->red_partials, grid, 1, rr, 1, nullptr);
    PCU_LAUNCH_CHECK(c);
  }
  return 0;
}

int pcu_transform_update(pcu_ctx* c, int m, int t, const double* Wm, const double* Wx, double* P, int ldp, double* AP, int ldap,
                         double* X, int ldx, double* R, int ldr, double* rr) {
  PCU_CHECK(c && Wm && P && t >= 1 && t <= kMaxT, "pcu_transform_update: bad arguments (t=%d)", t);
  PCU_CHECK((X == nullptr) == (R == nullptr) && (!X || (Wx && AP)), "pcu_transform_update: X, R, Wx and AP go together");
  const int grid = stream_grid(c, m, kThreads, kWaves);
  if (ensure_partials(c, (size_t)grid + 16)) return 1;
  DISPATCH_T(pick_T(t), transform_update_kernel<TT><<<grid, kThreads, 0, c->stream>>>(m, t, Wm, Wx, P, ldp, AP, ldap, X, ldx, R, ldr,
                                                                                     c->red_partials));
  PCU_LAUNCH_CHECK(c);
  if (R && rr) {
    reduce_partials_kernel<<<1, 32, 0, c->stream>>>(c->red_partials, grid, 1, rr, 1, nullptr);
    PCU_LAUNCH_CHECK(c);
  }
  return 0;
}

int pcu_update_z(pcu_ctx* c, int m, int tz, double* Z, int ldz, const double* P, int ldp, int t1, const double* beta1,
                 const double* Pp, int ldpp, int t2, const double* beta2) {
  PCU_CHECK(c && Z && P && beta1 && tz >= 1 && tz <= kMaxT && t1 >= 1 && t1 <= kMaxT && t2 >= 0 && t2 <= kMaxT,
            "pcu_update_z: bad arguments");
  PCU_CHECK(t2 == 0 || (Pp && beta2), "pcu_update_z: Pprev/beta2 missing");
  const int T = pick_T(std::max(tz, std::max(t1, t2)));
  const bool v2 = (tz == T) && T >= 2 && t1 == tz && (t2 == 0 || t2 == tz) && ldz % 2 == 0 && ldp % 2 == 0 &&
                  (t2 == 0 || ldpp % 2 == 0) && ((uintptr_t)Z % 16 == 0) && ((uintptr_t)P % 16 == 0) &&
                  ((uintptr_t)Pp % 16 == 0) && !getenv("PREALPS_ECG_V1");
  const int grid = stream_grid(c, m, v2 ? kThreads / (T / 2) * 4 : kThreads, kWaves);
  if (v2 && (T == 8 || T == 16) && !getenv("PREALPS_ECG_NOMMA")) {
    const double* pp = t2 ? Pp : nullptr;
    const int gridm = stream_grid(c, (m + 7) / 8, (kThreads / 32) * 8, kWaves);
    if (T == 8) update_z_mma_kernel<8><<<gridm, kThreads, 0, c->stream>>>(m, Z, ldz, P, ldp, beta1, pp, ldpp, beta2);
    else update_z_mma_kernel<16><<<gridm, kThreads, 0, c->stream>>>(m, Z, ldz, P, ldp, beta1, pp, ldpp, beta2);
    PCU_LAUNCH_CHECK(c);
    return 0;
  }
  if (v2) {
    const double* pp = t2 ? Pp : nullptr;
    switch (T) {
      case 2: update_z_v2_kernel<2><<<grid, kThreads, 0, c->stream>>>(m, Z, ldz, P, ldp, beta1, pp, ldpp, beta2); break;
      case 4: update_z_v2_kernel<4><<<grid, kThreads, 0, c->stream>>>(m, Z, ldz, P, ldp, beta1, pp, ldpp, beta2); break;
      case 8: update_z_v2_kernel<8><<<grid, kThreads, 0, c->stream>>>(m, Z, ldz, P, ldp, beta1, pp, ldpp, beta2); break;
      case 16: update_z_v2_kernel<16><<<grid, kThreads, 0, c->stream>>>(m, Z, ldz, P, ldp, beta1, pp, ldpp, beta2); break;
      default: update_z_v2_kernel<32><<<grid, kThreads, 0, c->stream>>>(m, Z, ldz, P, ldp, beta1, pp, ldpp, beta2); break;
    }
  } else
  DISPATCH_T(T, update_z_kernel<TT><<<grid, kThreads, 0, c->stream>>>(m, tz, Z, ldz, P, ldp, t1, beta1, Pp, ldpp, t2, beta2));
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_fused_small(pcu_ctx* c, int t, const double* mu, double* beta1, double* beta2, double* U_out, int* status_dev) {
  PCU_CHECK(c && mu && beta1 && beta2 && t >= 1 && t <= kMaxT, "pcu_fused_small: bad arguments");
  fused_small_kernel<<<1, kThreads, 0, c->stream>>>(t, mu, beta1, beta2, U_out, status_dev);
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_right_solve(pcu_ctx* c, int m, int t, const double* G, double* Z, int ldz) {
  PCU_CHECK(c && G && Z && t >= 1 && t <= kMaxT, "pcu_right_solve: bad arguments");
  const int grid = stream_grid(c, m, kThreads, kWaves);
  DISPATCH_T(pick_T(t), right_solve_kernel<TT><<<grid, kThreads, 0, c->stream>>>(m, t, G, Z, ldz));
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_sum_columns(pcu_ctx* c, int m, int t, const double* X, int ldx, double* sol) {
  sum_columns_kernel<<<stream_grid(c, m, 256, 4), 256, 0, c->stream>>>(m, t, X, ldx, sol);
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_split_rhs(pcu_ctx* c, int m, int t, const double* rhs, const int* col_of_row, double* R, int ldr) {
  split_rhs_kernel<<<stream_grid(c, m, 256, 4), 256, 0, c->stream>>>(m, t, rhs, col_of_row, R, ldr);
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_fro2(pcu_ctx* c, int m, int t, const double* R, int ldr, double* out) {
  const int grid = stream_grid(c, m, kThreads, kWaves);
  if (ensure_partials(c, (size_t)grid + 16)) return 1;
  fro2_kernel<<<grid, kThreads, 0, c->stream>>>(m, t, R, ldr, c->red_partials);
  PCU_LAUNCH_CHECK(c);
  reduce_partials_kernel<<<1, 32, 0, c->stream>>>(c->red_partials, grid, 1, out, 1, nullptr);
  PCU_LAUNCH_CHECK(c);
  return 0;
}

}  // extern "C"
