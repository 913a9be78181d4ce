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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------- gram2
// Thread = one 4x4 tile of (a,b) pairs for one stripe of rows.  NT = tiles per dimension.
template <int NT>
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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = (ia + i < t) ? __ldg(A1 + r * lda1 + ia + i) : 0.0;
      b[i] = (ib + i < t) ? __ldg(B1 + r * ldb1 + ib + i) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc1[i][j] = fma(a[i], b[j], acc1[i][j]);
    if (two) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = (ia + i < t) ? __ldg(A2 + r * lda2 + ia + i) : 0.0;
        b[i] = (ib + i < t) ? __ldg(B2 + r * ldb2 + ib + i) : 0.0;
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

// out[e] = sum_b partials[b][e], b ascending (deterministic)
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblocks, int n,
                                       double* __restrict__ out1, int n1, double* __restrict__ out2) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * n + e];
  if (e < n1) out1[e] = s; else out2[e - n1] = s;
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
  const int grid = stream_grid(c, m, rows_per_cta * 8, 2);
  if (ensure_partials(c, (size_t)grid * ngram * t * t + 16)) return 1;
  switch (NT) {
    case 1: gram2_kernel<1><<<grid, kThreads, 0, c->stream>>>(m, t, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, c->red_partials); break;
    case 2: gram2_kernel<2><<<grid, kThreads, 0, c->stream>>>(m, t, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, c->red_partials); break;
    case 4: gram2_kernel<4><<<grid, kThreads, 0, c->stream>>>(m, t, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, c->red_partials); break;
    default: gram2_kernel<8><<<grid, kThreads, 0, c->stream>>>(m, t, A1, lda1, B1, ldb1, A2, lda2, B2, ldb2, c->red_partials); break;
  }
  PCU_LAUNCH_CHECK(c);
  const int n = ngram * t * t;
  reduce_partials_kernel<<<ceil_div(n, 128), 128, 0, c->stream>>>(c->red_partials, grid, n, G1, t * t, G2);
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_ortho_update(pcu_ctx* c, int m, int t, const double* G, const double* Gpr, double* P, int ldp, double* AP,
                     int ldap, double* X, int ldx, double* R, int ldr, double* U_out, double* alpha_out, double* rr,
                     int* status_dev) {
  PCU_CHECK(c && G && P && AP && t >= 1 && t <= kMaxT, "pcu_ortho_update: bad arguments (t=%d)", t);
  PCU_CHECK((X == nullptr) == (R == nullptr), "pcu_ortho_update: X and R must both be given or both be NULL");
  PCU_CHECK(!X || Gpr, "pcu_ortho_update: Gpr missing");
  const int grid = stream_grid(c, m, kThreads, 2);
  if (ensure_partials(c, (size_t)grid + 16)) return 1;
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
  const int grid = stream_grid(c, m, kThreads, 2);
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

int pcu_update_z(pcu_ctx* c, int m, int tz, double* Z, int ldz, const double* P, int ldp, int t1, const double* beta1,
                 const double* Pp, int ldpp, int t2, const double* beta2) {
  PCU_CHECK(c && Z && P && beta1 && tz >= 1 && tz <= kMaxT && t1 >= 1 && t1 <= kMaxT && t2 >= 0 && t2 <= kMaxT,
            "pcu_update_z: bad arguments");
  PCU_CHECK(t2 == 0 || (Pp && beta2), "pcu_update_z: Pprev/beta2 missing");
  const int grid = stream_grid(c, m, kThreads, 2);
  const int T = pick_T(std::max(tz, std::max(t1, t2)));
  DISPATCH_T(T, update_z_kernel<TT><<<grid, kThreads, 0, c->stream>>>(m, tz, Z, ldz, P, ldp, t1, beta1, Pp, ldpp, t2, beta2));
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
  const int grid = stream_grid(c, m, kThreads, 2);
  if (ensure_partials(c, (size_t)grid + 16)) return 1;
  fro2_kernel<<<grid, kThreads, 0, c->stream>>>(m, t, R, ldr, c->red_partials);
  PCU_LAUNCH_CHECK(c);
  reduce_partials_kernel<<<1, 32, 0, c->stream>>>(c->red_partials, grid, 1, out, 1, nullptr);
  PCU_LAUNCH_CHECK(c);
  return 0;
}

}  // extern "C"
