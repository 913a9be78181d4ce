// bj_solve.cu -- K2: block-Jacobi application, X = A_bb^{-1} B for every local diagonal
// block, as level-scheduled multi-RHS sweeps over the supernodal elimination forest.
// Replaces MKL PARDISO phase 33 as called by preAlps_BlockJacobiApply
// (ref: src/preconditioners/block_jacobi.c:93-109, utils/cplm_light/cplm_kernels.c:790-853).
//
// With the factor stored as M_s = [L_ss^{-1}; L_bs L_ss^{-1}] (bj.h) every level is
//   forward : assemble  b_s = B[perm] - sum(update rows of the descendants)   (gather, fixed order)
//             panels    [y_s ; u_s] = M_s b_s                                  (streaming)
//   backward: panels    x_s = M_s^T [y_s ; -x_ancestors]                       (streaming)
// The streaming kernel is HBM-bound: a warp reads a 32-row k-major panel with 512-byte
// coalesced loads (each lane a double2 = two rows of one k), the T-wide input row of step
// k is identical for the 16 lanes of a half-warp (broadcast load), every lane keeps
// 2 x T accumulators, and nothing is reduced across lanes until the very end.  Long
// panels get a whole CTA (split-K over 8 warps, fixed-order combine in shared memory).
// No atomics: results are bit-reproducible.
#include <algorithm>

#include "bj.h"

namespace {

using namespace pcu;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- forward right-hand-side assembly: one group of G lanes per column
template <int T>
__global__ void __launch_bounds__(kThreads) assemble_kernel(const int* __restrict__ cols, int ncols,
                                                            const double* __restrict__ B, int ldb, int t,
                                                            const int* __restrict__ perm,
                                                            const long long* __restrict__ gl_ptr,
                                                            const long long* __restrict__ gl_idx,
                                                            const double* __restrict__ U, double* __restrict__ Wk) {
  constexpr int CPL = (T >= 2) ? 2 : 1;
  constexpr int G = T / CPL;
  const int gid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / G);
  const int lig = threadIdx.x % G;
  const int ngroups = (int)((long long)gridDim.x * blockDim.x / G);
  for (int q = gid; q < ncols; q += ngroups) {
    const int c = cols[q];
    const double* src = B + (size_t)perm[c] * ldb;
    double a0 = 0.0, a1 = 0.0;
    const int c0 = CPL * lig;
    if (c0 < t) a0 = src[c0];
    if (CPL == 2 && c0 + 1 < t) a1 = src[c0 + 1];
    for (long long g = gl_ptr[c]; g < gl_ptr[c + 1]; ++g) {
      const double* u = U + (size_t)gl_idx[g] * T + c0;
      if (CPL == 2) { const double2 v = *reinterpret_cast<const double2*>(u); a0 -= v.x; a1 -= v.y; }
      else a0 -= u[0];
    }
    double* dst = Wk + (size_t)c * T + c0;
    if (CPL == 2) *reinterpret_cast<double2*>(dst) = make_double2(a0, a1);
    else dst[0] = a0;
  }
}

struct SweepArgs {
  const WorkUnit* units;
  const void* panels;
  const double* data;
  const double* Wk;     // fwd input
  double* Y;            // fwd output (own columns) / bwd input (own columns)
  double* U;            // fwd output (update rows)
  double* Xp;           // bwd output + input (ancestors), forest order
  const int* rows;      // bwd gather index
  const int* perm;      // bwd: final scatter into the caller's block
  double* Out; int ldo; int t;
};

template <int T, bool FWD>
__global__ void __launch_bounds__(kThreads) sweep_kernel(SweepArgs a) {
  __shared__ double red[32 * T];
  const WorkUnit u = a.units[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, j = lane & 15;
  if (!u.split && warp >= u.count) return;
  const int pidx = u.split ? u.first : u.first + warp;
  long long off;
  int klen, c0, w, h, row0;
  long long uoff = 0, rows_off = 0;
  if (FWD) {
    const FwdPanel p = reinterpret_cast<const FwdPanel*>(a.panels)[pidx];
    off = p.off; klen = p.klen; c0 = p.c0; w = p.w; h = p.h; row0 = p.row0; uoff = p.uoff;
  } else {
    const BwdPanel p = reinterpret_cast<const BwdPanel*>(a.panels)[pidx];
    off = p.off; klen = p.klen; c0 = p.c0; w = p.w; h = p.h; row0 = p.k0; rows_off = p.rows_off;
  }
  // k range of this warp, in pairs of k (each half-warp takes one k of the pair)
  const int npair = klen >> 1;
  int p0 = 0, p1 = npair;
  if (u.split) {
    const int per = (npair + kWarps - 1) / kWarps;
    p0 = min(npair, warp * per);
    p1 = min(npair, p0 + per);
  }
  double acc0[T], acc1[T];
#pragma unroll
  for (int c = 0; c < T; ++c) { acc0[c] = 0.0; acc1[c] = 0.0; }
  const double* base = a.data + off + 2 * j;
  const int* rows = a.rows + rows_off;
#pragma unroll 4
  for (int kp = p0; kp < p1; ++kp) {
    const int k = 2 * kp + half;
    const double2 m = __ldg(reinterpret_cast<const double2*>(base + (size_t)k * 32));
    const double* bp;
    if (FWD) {
      bp = a.Wk + (size_t)(c0 + k) * T;
    } else {
      const int i = min(row0 + k, h - 1);
      bp = (i < w) ? a.Y + (size_t)(c0 + i) * T : a.Xp + (size_t)__ldg(rows + i) * T;
    }
    if (T >= 2) {
#pragma unroll
      for (int c = 0; c < T; c += 2) {
        const double2 b = *reinterpret_cast<const double2*>(bp + c);
        acc0[c] = fma(m.x, b.x, acc0[c]);
        acc1[c] = fma(m.y, b.x, acc1[c]);
        acc0[c + 1] = fma(m.x, b.y, acc0[c + 1]);
        acc1[c + 1] = fma(m.y, b.y, acc1[c + 1]);
      }
    } else {
      const double b = bp[0];
      acc0[0] = fma(m.x, b, acc0[0]);
      acc1[0] = fma(m.y, b, acc1[0]);
    }
  }
  // combine the two half-warps (even k + odd k); afterwards lane (half, j) owns row 2j + half
  double mine[T];
#pragma unroll
  for (int c = 0; c < T; ++c) {
    const double s0 = acc0[c] + __shfl_xor_sync(0xffffffffu, acc0[c], 16);
    const double s1 = acc1[c] + __shfl_xor_sync(0xffffffffu, acc1[c], 16);
    mine[c] = half ? s1 : s0;
  }
  const int rloc = 2 * j + half;
  if (u.split) {
    // fixed-order combine over the warps
    for (int wv = 0; wv < kWarps; ++wv) {
      if (warp == wv) {
#pragma unroll
        for (int c = 0; c < T; ++c) {
          if (wv == 0) red[rloc * T + c] = mine[c];
          else red[rloc * T + c] += mine[c];
        }
      }
      __syncthreads();
    }
    if (warp != 0) return;
#pragma unroll
    for (int c = 0; c < T; ++c) mine[c] = red[rloc * T + c];
  }
  const int r = row0 + rloc;
  if (FWD) {
    if (r < h) {
      double* dst = (r < w) ? a.Y + (size_t)(c0 + r) * T : a.U + (size_t)(uoff + (r - w)) * T;
#pragma unroll
      for (int c = 0; c < T; ++c) dst[c] = mine[c];
    }
  } else {
    if (r < w) {
      double* dst = a.Xp + (size_t)(c0 + r) * T;
#pragma unroll
      for (int c = 0; c < T; ++c) dst[c] = mine[c];
      double* o = a.Out + (size_t)a.perm[c0 + r] * a.ldo;
#pragma unroll
      for (int c = 0; c < T; ++c) if (c < a.t) o[c] = mine[c];
    }
  }
}

int pick_T(int t) { return t <= 1 ? 1 : t <= 2 ? 2 : t <= 4 ? 4 : t <= 8 ? 8 : t <= 16 ? 16 : 32; }

int ensure_work(pcu_bj* bj, int T) {
  if (bj->cap_t >= T) return 0;
  pcu_ctx* c = bj->ctx;
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(bj->Wk); cudaFree(bj->Y); cudaFree(bj->U); cudaFree(bj->Xp);
  bj->Wk = bj->Y = bj->U = bj->Xp = nullptr;
  const size_t nv = ((size_t)bj->n + 4) * T, nuv = ((size_t)bj->nu + 4) * T;
  PCU_CUDA(cudaMalloc(&bj->Wk, nv * sizeof(double)));
  PCU_CUDA(cudaMalloc(&bj->Y, nv * sizeof(double)));
  PCU_CUDA(cudaMalloc(&bj->Xp, nv * sizeof(double)));
  PCU_CUDA(cudaMalloc(&bj->U, nuv * sizeof(double)));
  PCU_CUDA(cudaMemsetAsync(bj->Wk, 0, nv * sizeof(double), c->stream));
  PCU_CUDA(cudaMemsetAsync(bj->Y, 0, nv * sizeof(double), c->stream));
  PCU_CUDA(cudaMemsetAsync(bj->Xp, 0, nv * sizeof(double), c->stream));
  PCU_CUDA(cudaMemsetAsync(bj->U, 0, nuv * sizeof(double), c->stream));
  bj->cap_t = T;
  return 0;
}

template <int T>
int apply_T(pcu_bj* bj, const double* B, int ldb, double* X, int ldx, int t) {
  pcu_ctx* c = bj->ctx;
  cudaStream_t st = c->stream;
  constexpr int G = (T >= 2) ? T / 2 : 1;
  SweepArgs a{};
  a.Wk = bj->Wk; a.Y = bj->Y; a.U = bj->U; a.Xp = bj->Xp; a.rows = bj->rows; a.perm = bj->perm;
  a.Out = X; a.ldo = ldx; a.t = t;
  for (int l = 0; l < bj->nlevels; ++l) {
    const int ncols = bj->lvl_col_ptr[l + 1] - bj->lvl_col_ptr[l];
    if (ncols > 0) {
      const int grid = stream_grid(c, (long long)ncols * G, kThreads, 8);
      assemble_kernel<T><<<grid, kThreads, 0, st>>>(bj->lvl_cols + bj->lvl_col_ptr[l], ncols, B, ldb, t, bj->perm,
                                                   bj->gl_ptr, bj->gl_idx, bj->U, bj->Wk);
      PCU_LAUNCH_CHECK(c);
    }
    const int nu = bj->fwd_unit_ptr[l + 1] - bj->fwd_unit_ptr[l];
    if (nu > 0) {
      a.units = bj->fwd_units + bj->fwd_unit_ptr[l];
      a.panels = bj->fwd_panels;
      a.data = bj->fwd_data;
      sweep_kernel<T, true><<<nu, kThreads, 0, st>>>(a);
      PCU_LAUNCH_CHECK(c);
    }
  }
  for (int l = bj->nlevels - 1; l >= 0; --l) {
    const int nu = bj->bwd_unit_ptr[l + 1] - bj->bwd_unit_ptr[l];
    if (nu > 0) {
      a.units = bj->bwd_units + bj->bwd_unit_ptr[l];
      a.panels = bj->bwd_panels;
      a.data = bj->bwd_data;
      sweep_kernel<T, false><<<nu, kThreads, 0, st>>>(a);
      PCU_LAUNCH_CHECK(c);
    }
  }
  return 0;
}

}  // namespace

extern "C" int pcu_bj_apply(pcu_bj* bj, const double* B, int ldb, double* X, int ldx, int t) {
  PCU_CHECK(bj && B && X && t >= 1 && t <= 32, "pcu_bj_apply: bad arguments (t=%d, need 1..32)", t);
  PCU_CHECK(ldb >= t && ldx >= t, "pcu_bj_apply: leading dimension smaller than t");
  const int T = pick_T(t);
  if (ensure_work(bj, T)) return 1;
  // the work vectors are laid out for cap_t columns; a narrower solve uses its own T, which is
  // consistent within one apply (every vector is rewritten before it is read)
  switch (T) {
    case 1: return apply_T<1>(bj, B, ldb, X, ldx, t);
    case 2: return apply_T<2>(bj, B, ldb, X, ldx, t);
    case 4: return apply_T<4>(bj, B, ldb, X, ldx, t);
    case 8: return apply_T<8>(bj, B, ldb, X, ldx, t);
    case 16: return apply_T<16>(bj, B, ldb, X, ldx, t);
    default: return apply_T<32>(bj, B, ldb, X, ldx, t);
  }
}
