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
#include <string>

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

__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows of the input block staged per tile and per warp (4 KB per buffer at any T)
template <int T>
struct Tile { static constexpr int KT = (T <= 8) ? 64 : (512 / T); };  // k steps per tile, even, KT*T*8 = 4 KB for T >= 8


// One warp per short panel, or the 8 warps of a CTA on disjoint k ranges of one long panel.
// Per warp: (1) a ring of D panel loads (16 B per lane each, HBM latency) is always in flight,
// (2) the T-wide input rows of the next 64 k steps are copied into a warp-private shared-memory tile
// with cp.async (LDGSTS, no registers) while the current tile is consumed, so the FMAs only ever
// wait on shared-memory broadcasts.  Measured before this pipeline: long-scoreboard bound, 2.1 TB/s.
template <int T, bool FWD, int D, bool NOALLOC>
__global__ void __launch_bounds__(kThreads, 2) sweep_kernel(SweepArgs a) {
  constexpr int KT = Tile<T>::KT;       // k steps per tile
  constexpr int KP = KT / 2;            // k pairs per tile
  constexpr int TILE = KT * T;          // doubles per tile buffer
  extern __shared__ __align__(16) double smem[];
  double* red = smem;                   // 32 * T doubles
  const WorkUnit u = a.units[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, j = lane & 15;
  double* tile0 = smem + 32 * T + (size_t)warp * 2 * TILE;
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
  const int npair = klen >> 1;
  int p0 = 0, p1 = npair;
  if (u.split) {
    int per = (npair + kWarps - 1) / kWarps;
    per = (per + KP - 1) / KP * KP;     // whole tiles per warp
    p0 = min(npair, warp * per);
    p1 = min(npair, p0 + per);
  }
  double acc0[T], acc1[T];
#pragma unroll
  for (int c = 0; c < T; ++c) { acc0[c] = 0.0; acc1[c] = 0.0; }
  const double* base = a.data + off + 2 * j + (size_t)half * 32;
  const int* rows = a.rows + rows_off;

  // copy the input rows of k in [2*tp, 2*tp + KT) into buf (rows past the panel are clamped: their
  // panel entries are zero padding)
  auto stage = [&](int tp, double* buf) {
    constexpr int CHUNKS = (T >= 2) ? TILE / 2 : KT;  // 16-byte pieces per tile (rows when T == 1)
    constexpr int CPR = (T >= 2) ? T / 2 : 1; // pieces per row
#pragma unroll
    for (int q0 = 0; q0 < CHUNKS; q0 += 32) {
      const int q = q0 + lane;
      if (T >= 2) {
        const int r = q / CPR, part = q % CPR;
        const int k = min(2 * tp + r, klen - 1);
        const double* src;
        if (FWD) src = a.Wk + (size_t)(c0 + k) * T;
        else {
          const int i = min(row0 + k, h - 1);
          src = (i < w) ? a.Y + (size_t)(c0 + i) * T : a.Xp + (size_t)__ldg(rows + i) * T;
        }
        cp_async16(buf + (size_t)r * T + 2 * part, src + 2 * part);
      } else {  // T == 1: two rows per 16-byte piece do not share a source row; plain loads
        const int k = min(2 * tp + q, klen - 1);
        if (q < KT) {
          const double* src;
          if (FWD) src = a.Wk + (size_t)(c0 + k);
          else { const int i = min(row0 + k, h - 1); src = (i < w) ? a.Y + (size_t)(c0 + i) : a.Xp + (size_t)__ldg(rows + i); }
          buf[q] = *src;
        }
      }
    }
    cp_async_commit();
  };

  double2 ring[D];
  auto issue = [&](int kp, double2& m) {
    if (kp < p1) m = NOALLOC ? ld_stream2(base + (size_t)kp * 64) : __ldg(reinterpret_cast<const double2*>(base + (size_t)kp * 64));
    else m = make_double2(0.0, 0.0);
  };
  if (p0 < p1) {
    stage(p0, tile0);
#pragma unroll
    for (int u2 = 0; u2 < D; ++u2) issue(p0 + u2, ring[u2]);
    int tix = 0;
    for (int tp = p0; tp < p1; tp += KP, ++tix) {
      double* cur = tile0 + (size_t)(tix & 1) * TILE;
      if (tp + KP < p1) { stage(tp + KP, tile0 + (size_t)((tix + 1) & 1) * TILE); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncwarp();
#pragma unroll 1
      for (int kq = 0; kq < KP; kq += D) {
#pragma unroll
        for (int u2 = 0; u2 < D; ++u2) {
          const double2 m = ring[u2];
          issue(tp + kq + u2 + D, ring[u2]);
          const double* bp = cur + (size_t)(2 * (kq + u2) + half) * T;
          if (T >= 2) {
#pragma unroll
            for (int c = 0; c < T; c += 2) {
              const double2 bb = *reinterpret_cast<const double2*>(bp + c);
              acc0[c] = fma(m.x, bb.x, acc0[c]);
              acc1[c] = fma(m.y, bb.x, acc1[c]);
              acc0[c + 1] = fma(m.x, bb.y, acc0[c + 1]);
              acc1[c + 1] = fma(m.y, bb.y, acc1[c + 1]);
            }
          } else {
            const double bb = bp[0];
            acc0[0] = fma(m.x, bb, acc0[0]);
            acc1[0] = fma(m.y, bb, acc1[0]);
          }
        }
      }
      __syncwarp();
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
  const size_t nv = ((size_t)bj->n + 72) * T, nuv = ((size_t)bj->nu + 4) * T;
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

template <int T, bool FWD, int D, bool NOALLOC>
void launch_one(int nu, cudaStream_t st, const SweepArgs& a) {
  constexpr int bytes = (32 * T + kWarps * 2 * Tile<T>::KT * T) * (int)sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(sweep_kernel<T, FWD, D, NOALLOC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    configured = true;
  }
  sweep_kernel<T, FWD, D, NOALLOC><<<nu, kThreads, bytes, st>>>(a);
}

template <int T, bool FWD>
void launch_sweep(int nu, cudaStream_t st, const SweepArgs& a) {
  const char* e = getenv("PREALPS_BJ_VARIANT");
  const int variant = e ? atoi(e) : 0;
  switch (variant) {
    case 1: launch_one<T, FWD, 4, true>(nu, st, a); break;
    case 2: launch_one<T, FWD, 8, false>(nu, st, a); break;
    case 3: launch_one<T, FWD, 4, false>(nu, st, a); break;
    default: launch_one<T, FWD, 8, true>(nu, st, a); break;
  }
}

// PREALPS_BJ_PROFILE=1: per-launch CUDA-event timings of one apply, printed to stderr
struct LevelProfiler {
  bool on = false;
  cudaStream_t st;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> tag;
  std::vector<double> bytes;
  explicit LevelProfiler(cudaStream_t s) : st(s) { on = getenv("PREALPS_BJ_PROFILE") != nullptr; }
  void mark(const std::string& name, double b) {
    if (!on) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
    ev.push_back(e); tag.push_back(name); bytes.push_back(b);
  }
  void report() {
    if (!on || ev.empty()) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); cudaEventSynchronize(e);
    ev.push_back(e);
    double tot = 0, totb = 0;
    for (size_t i = 0; i + 1 < ev.size(); ++i) {
      float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      fprintf(stderr, "  %-22s %9.1f us %10.1f MB %8.1f GB/s\n", tag[i].c_str(), ms * 1e3, bytes[i] / 1e6,
              bytes[i] / (ms * 1e-3) / 1e9);
      tot += ms; totb += bytes[i];
    }
    fprintf(stderr, "  total %.3f ms, %.1f MB, %.1f GB/s\n", tot, totb / 1e6, totb / (tot * 1e-3) / 1e9);
    for (auto x : ev) cudaEventDestroy(x);
  }
};

template <int T>
int apply_T(pcu_bj* bj, const double* B, int ldb, double* X, int ldx, int t) {
  pcu_ctx* c = bj->ctx;
  cudaStream_t st = c->stream;
  LevelProfiler prof(st);
  constexpr int G = (T >= 2) ? T / 2 : 1;
  SweepArgs a{};
  a.Wk = bj->Wk; a.Y = bj->Y; a.U = bj->U; a.Xp = bj->Xp; a.rows = bj->rows; a.perm = bj->perm;
  a.Out = X; a.ldo = ldx; a.t = t;
  for (int l = 0; l < bj->nlevels; ++l) {
    const int ncols = bj->lvl_col_ptr[l + 1] - bj->lvl_col_ptr[l];
    if (ncols > 0) {
      prof.mark("asm L" + std::to_string(l) + " cols=" + std::to_string(ncols), 3.0 * ncols * T * 8);
      const int grid = stream_grid(c, (long long)ncols * G, kThreads, 8);
      assemble_kernel<T><<<grid, kThreads, 0, st>>>(bj->lvl_cols + bj->lvl_col_ptr[l], ncols, B, ldb, t, bj->perm,
                                                   bj->gl_ptr, bj->gl_idx, bj->U, bj->Wk);
      PCU_LAUNCH_CHECK(c);
    }
    const int nu = bj->fwd_unit_ptr[l + 1] - bj->fwd_unit_ptr[l];
    if (nu > 0) {
      prof.mark("fwd L" + std::to_string(l) + " ctas=" + std::to_string(nu), bj->fwd_lvl_bytes[l]);
      a.units = bj->fwd_units + bj->fwd_unit_ptr[l];
      a.panels = bj->fwd_panels;
      a.data = bj->fwd_data;
      launch_sweep<T, true>(nu, st, a);
      PCU_LAUNCH_CHECK(c);
    }
  }
  for (int l = bj->nlevels - 1; l >= 0; --l) {
    const int nu = bj->bwd_unit_ptr[l + 1] - bj->bwd_unit_ptr[l];
    if (nu > 0) {
      prof.mark("bwd L" + std::to_string(l) + " ctas=" + std::to_string(nu), bj->bwd_lvl_bytes[l]);
      a.units = bj->bwd_units + bj->bwd_unit_ptr[l];
      a.panels = bj->bwd_panels;
      a.data = bj->bwd_data;
      launch_sweep<T, false>(nu, st, a);
      PCU_LAUNCH_CHECK(c);
    }
  }
  prof.report();
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
