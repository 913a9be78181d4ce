// bj_solve.cu -- K2: block-Jacobi application, X = A_bb^{-1} B for every local diagonal
// block, as level-scheduled multi-RHS sweeps over the supernodal elimination forest.
// Replaces MKL PARDISO phase 33 as called by preAlps_BlockJacobiApply
// (ref: src/preconditioners/block_jacobi.c:93-109, utils/cplm_light/cplm_kernels.c:790-853).
//
// With the factor stored as M_s = [L_ss^{-1}; L_bs L_ss^{-1}] (bj.h) every level is
//   forward : assemble  b_s = B[perm] - sum(update rows of the descendants)   (gather, fixed order)
//             panels    [y_s ; u_s] = M_s b_s                                  (streaming)
//   backward: panels    x_s = M_s^T [y_s ; -x_ancestors]                       (streaming, the SAME panels, bj.h)
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

// Programmatic dependent launch: the kernels of one apply form a chain (assemble / forward / backward, level after
// level). Each is launched with programmatic stream serialization, so its CTAs may start while the previous
// kernel drains: they fetch what does not depend on it (work units, panel descriptors, the first panel data), then
// wait here for the previous kernel to complete. Without the launch attribute both instructions are no-ops.
#ifndef PCU_EMUL
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else  // tests/emul runs the kernels of a stream one after the other
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}
#endif

// ---- forward right-hand-side assembly: one group of G lanes per column.
// A group's FIRST column fetches everything static -- its forest column, the row of the caller's block it maps to, its
// gather range and the first gather indices -- BEFORE griddepcontrol.wait, i.e. while the previous level's forward sweep
// drains; only the loads of B and U (data written earlier in the stream) come after the wait (measured: 0.941 -> 0.923 ms
// for the apply of one 64^3 block).
// The list is walked in order (fixed summation order), NB update rows in flight per lane group; the indices of the next
// batch are fetched while the rows of the current one are in flight.  NB = 16 for the levels whose columns gather
// hundreds of update rows (the top of the forest: an assembly launch there is one chain of dependent round trips, its
// length is what the launch costs), NB = 8 (fewer registers, more resident warps) where the lists are short or empty.
// Slots past the end of a list point at row `pad` of U, which is zero and never written: the loads of a batch carry no
// predicate, and with the minimum-CTAs launch bound ptxas issues them back to back instead of sinking each next to its
// addition (the padding adds 0.0).  The update rows arrive NEGATED (the rows of M_s below the diagonal block are stored
// negated for the backward sweep, bj.h), so the assembly adds them: b_s = B - sum(u) as before, bit for bit.
template <int T, int NB>
__global__ void __launch_bounds__(kThreads, NB == 16 ? 2 : 4) assemble_kernel(const int* __restrict__ cols, int ncols,
                                                                              const double* __restrict__ B, int ldb, int t,
                                                                              const int* __restrict__ perm,
                                                                              const long long* __restrict__ gl_ptr,
                                                                              const long long* __restrict__ gl_idx,
                                                                              const double* __restrict__ U, long long pad,
                                                                              double* __restrict__ Wk) {
  constexpr int CPL = (T >= 2) ? 2 : 1;
  constexpr int G = T / CPL;
  const int gid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / G);
  const int lig = threadIdx.x % G;
  const int ngroups = (int)((long long)gridDim.x * blockDim.x / G);
  int c_first = 0, row_first = 0;
  long long g_first = 0, g1_first = 0;
  long long idx[NB];
  if (gid < ncols) {
    c_first = __ldg(cols + gid);
    row_first = __ldg(perm + c_first);
    g_first = __ldg(gl_ptr + c_first);
    g1_first = __ldg(gl_ptr + c_first + 1);
#pragma unroll
    for (int j = 0; j < NB; ++j) idx[j] = (g_first + j < g1_first) ? __ldg(gl_idx + g_first + j) : pad;
  }
  pdl_wait();
  pdl_launch_dependents();
  for (int q = gid; q < ncols; q += ngroups) {
    const bool pre = (q == gid);
    const int c = pre ? c_first : cols[q];
    const double* src = B + (size_t)(pre ? row_first : perm[c]) * ldb;
    double a0 = 0.0, a1 = 0.0;
    const int c0 = CPL * lig;
    if (c0 < t) a0 = src[c0];
    if (CPL == 2 && c0 + 1 < t) a1 = src[c0 + 1];
    const long long g0 = pre ? g_first : gl_ptr[c];
    const long long g1 = pre ? g1_first : gl_ptr[c + 1];
    if (!pre) {
#pragma unroll
      for (int j = 0; j < NB; ++j) idx[j] = (g0 + j < g1) ? __ldg(gl_idx + g0 + j) : pad;
    }
    for (long long g = g0; g < g1; g += NB) {
      double2 v[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const double* u = U + (size_t)idx[j] * T + c0;
        if (CPL == 2) v[j] = *reinterpret_cast<const double2*>(u);
        else v[j] = make_double2(u[0], 0.0);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) idx[j] = (g + NB + j < g1) ? __ldg(gl_idx + g + NB + j) : pad;  // next batch
#pragma unroll
      for (int j = 0; j < NB; ++j) { a0 += v[j].x; a1 += v[j].y; }
    }
    double* dst = Wk + (size_t)c * T + c0;
    if (CPL == 2) *reinterpret_cast<double2*>(dst) = make_double2(a0, a1);
    else dst[0] = a0;
  }
}

// The same assembly for the levels whose columns gather long lists (the top of the forest: hundreds of update rows per column,
// few columns): one WARP per column.  The 32 / G lane groups of the warp take the batches of NB rows round robin (32 / G * NB
// rows of a column in flight instead of 16, the chain of dependent round trips that an assembly launch there consists of is
// 32 / G times shorter), then add their partial sums in a fixed order (xor butterfly): bit-reproducible, not bit-identical to
// the sequential order of assemble_kernel.
template <int T>
__global__ void __launch_bounds__(kThreads, 4) assemble_wide_kernel(const int* __restrict__ cols, int ncols,
                                                                    const double* __restrict__ B, int ldb, int t,
                                                                    const int* __restrict__ perm,
                                                                    const long long* __restrict__ gl_ptr,
                                                                    const long long* __restrict__ gl_idx,
                                                                    const double* __restrict__ U, long long pad,
                                                                    double* __restrict__ Wk) {
  constexpr int CPL = (T >= 2) ? 2 : 1;
  constexpr int G = T / CPL;      // lanes per row
  constexpr int SG = 32 / G;      // lane groups per warp
  constexpr int NB = 8;
  const int lane = threadIdx.x & 31, sub = lane / G, lig = lane % G;
  const int wq = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  int c_first = 0, row_first = 0;
  long long g_first = 0, g1_first = 0;
  long long idx[NB];
  if (wq < ncols) {
    c_first = __ldg(cols + wq);
    row_first = __ldg(perm + c_first);
    g_first = __ldg(gl_ptr + c_first);
    g1_first = __ldg(gl_ptr + c_first + 1);
#pragma unroll
    for (int j = 0; j < NB; ++j) idx[j] = (g_first + sub * NB + j < g1_first) ? __ldg(gl_idx + g_first + sub * NB + j) : pad;
  }
  pdl_wait();
  pdl_launch_dependents();
  for (int q = wq; q < ncols; q += nwarps) {
    const bool pre = (q == wq);
    const int c = pre ? c_first : cols[q];
    const int c0 = CPL * lig;
    double a0 = 0.0, a1 = 0.0;
    if (sub == 0) {
      const double* src = B + (size_t)(pre ? row_first : perm[c]) * ldb;
      if (c0 < t) a0 = src[c0];
      if (CPL == 2 && c0 + 1 < t) a1 = src[c0 + 1];
    }
    const long long g0 = pre ? g_first : gl_ptr[c];
    const long long g1 = pre ? g1_first : gl_ptr[c + 1];
    if (!pre) {
#pragma unroll
      for (int j = 0; j < NB; ++j) idx[j] = (g0 + sub * NB + j < g1) ? __ldg(gl_idx + g0 + sub * NB + j) : pad;
    }
    for (long long g = g0 + sub * NB; g < g1; g += SG * NB) {
      double2 v[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const double* u = U + (size_t)idx[j] * T + c0;
        if (CPL == 2) v[j] = *reinterpret_cast<const double2*>(u);
        else v[j] = make_double2(u[0], 0.0);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) idx[j] = (g + SG * NB + j < g1) ? __ldg(gl_idx + g + SG * NB + j) : pad;  // next batch
#pragma unroll
      for (int j = 0; j < NB; ++j) { a0 += v[j].x; a1 += v[j].y; }
    }
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (sub == 0) {
      double* dst = Wk + (size_t)c * T + c0;
      if (CPL == 2) *reinterpret_cast<double2*>(dst) = make_double2(a0, a1);
      else dst[0] = a0;
    }
  }
}

struct SweepArgs {
  const WorkUnit* units;
  const void* panels;
  const double* data;
  const double* Wk;     // fwd input
  const double* Bsrc;   // fwd, leaves only: read the caller's block directly (row perm[c], ld ldb) instead of Wk -- the
  int ldb;              // columns of a leaf gather nothing, their assembly would be a plain permuted copy
  double* Y;            // fwd output (own columns) / bwd input (own columns)
  double* U;            // fwd output (update rows)
  double* Xp;           // bwd output + input (ancestors), forest order
  const int* rows;      // bwd gather index
  const int* perm;      // bwd: final scatter into the caller's block
  double* Out; int ldo; int t;
  double* scratch; int* counters;   // inter-CTA split-K
};

// one 256-bit load per lane (SASS LDG.E.NA.EFL2.256): the 32 bytes a lane owns of a k-block, not allocated in L1 and
// first in line for eviction from L2 -- the panels are streamed once per apply (19 GB) and must not push the block
// vectors (update rows, ancestor rows gathered by the backward sweep) out of the 126 MB L2
#ifndef PCU_EMUL
__device__ __forceinline__ void ld_stream4(const double* p, double2& a, double2& b) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0, %1, %2, %3}, [%4];"
               : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// warp-private ring of panel data in shared memory, filled by cp.async.bulk (SASS UBLKCP) issued by one lane, completion
// on one mbarrier per stage
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_stream(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar,
                                                unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_addr(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {  // bounded: a lost copy traps, it does not hang
  unsigned ok = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(phase) : "memory");
    if (ok) return;
    if (++spins > (1u << 26)) __trap();
  }
}
#else  // tests/emul: the same data movement without the cache hints; cp.async completes at once
inline void ld_stream4(const double* p, double2& a, double2& b) {
  pcu_emul_check_aligned(p, 32);
  a = make_double2(p[0], p[1]);
  b = make_double2(p[2], p[3]);
}
inline unsigned long long l2_evict_first_policy() { return 0; }
inline void cp_async16(void* smem_dst, const void* gmem_src) {
  pcu_emul_check_aligned(smem_dst, 16);
  pcu_emul_check_aligned(gmem_src, 16);
  std::memcpy(smem_dst, gmem_src, 16);
}
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
inline void mbar_init(unsigned long long*, unsigned) {}
inline void fence_mbar_init() {}
inline void mbar_arrive_expect_tx(unsigned long long*, unsigned) {}
inline void bulk_g2s_stream(void* smem, const void* gmem, unsigned bytes, unsigned long long*, unsigned long long) {
  pcu_emul_check_aligned(smem, 16);
  pcu_emul_check_aligned(gmem, 16);
  std::memcpy(smem, gmem, bytes);
}
inline void mbar_wait(unsigned long long*, unsigned) {}
#endif

// rows of the input block staged per tile and per warp (4 KB per buffer at any T)
template <int T, bool STREAM>
struct Tile {
  // k steps of input rows per tile.  STREAM (contiguous panels, the shared-memory panel ring takes 8 KB per warp): 32 (16 from
  // t = 16 up), KT*T*8 = 2 KB at t = 8 and 16; the tile-wise backward sweep of the single-copy factor keeps its panel data in
  // registers, consumes whole 32-row tiles of M and stages 64 rows (4 KB at t = 8)
  static constexpr int KT0 = STREAM ? ((T <= 8) ? 32 : 16) : ((T <= 8) ? 64 : (512 / T));
  static constexpr int KT = (!STREAM && KT0 < 32) ? 32 : KT0;
};
// Shared-memory ring of panel data per warp: RS stages of RK k-blocks (1 KB each), one cp.async.bulk copy per stage.
// Measured on B200 (128^3, 8 blocks / one 64^3 block, t = 8, apply in ms; profiles/r02_smem_ring.md): register ring of 4
// k-blocks 3.507 / 0.754; shared-memory ring 8 x 1: 3.530 / 0.772; 4 x 2: 3.286 / 0.696; 2 x 4: 3.226 / 0.678 -- the bytes
// in flight per warp double without a register spent on them (80 instead of 128 registers), and few large copies beat
// many small ones; 3 CTAs/SM with 6 KB rings lose (3.242 / 0.737 and 3.624 / 0.828).
template <int T>
struct Ring { static constexpr int RK = (T <= 16) ? 4 : 2, RS = 2, DOUBLES = RK * RS * 128; };

#ifndef PCU_EMUL
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#define PCU_DYN_SMEM(name) extern __shared__ __align__(16) double name[]
#else
// D (8x8) += A (8x4, row-major fragment: lane holds A[lane/4][lane%4]) * B (4x8, lane holds B[lane%4][lane/4]);
// lane holds D[lane/4][2*(lane%4) + {0,1}]
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
#define PCU_DYN_SMEM(name) double* name = static_cast<double*>(emul_dyn_smem())
#endif

// Backward sweep: the A fragments of M_s^T straight out of the forward panels.  In a k-block (32 rows x 4 columns) the
// 32 bytes of "slot" s hold M(8*rg + s/4, 4*kb + s%4), rg = 0..3.  For the 32 rows x 8 columns of two consecutive k-blocks
// (a unit) lane l needs, as the A fragment of step group ib = 0..7, M(4*ib + l%4, 8-column index l/4): k-block l/16, slot
// 16*(ib%2) + 4*(l%4) + (l/4)%4, register ib/2.  So a lane reads the slots 4*(l%4) + (l/4)%4 (ib even) and 16 more (ib
// odd) of k-block l/16 -- two 32-byte loads per unit like the forward sweep, the warp covering 2 x 512 contiguous bytes
// per load: the transposition is in the addressing, there is nothing to exchange between lanes.
__device__ __forceinline__ int bwd_lane_offset(int lane) {   // doubles, relative to the unit's first k-block
  return (lane >> 4) * 128 + (4 * (lane & 3) + ((lane >> 2) & 3)) * 4;
}

// lane owns, for every row group rg and column block nb: row row0 + 8*rg + lane/4, columns 8*nb + 2*(lane%4) + {0,1}
template <int T, bool FWD>
__device__ __forceinline__ void store_outputs(const double (&acc)[4][(T + 7) / 8][2], const SweepArgs& a, int row0, int c0,
                                              int w, int h, long long uoff, int lr, int lk) {
  constexpr int NB = (T + 7) / 8;
#pragma unroll
  for (int rg = 0; rg < 4; ++rg) {
    const int r = row0 + 8 * rg + lr;
    double* dst = nullptr;
    double* out = nullptr;
    if (FWD) {
      if (r < h) dst = (r < w) ? a.Y + (size_t)(c0 + r) * T : a.U + (size_t)(uoff + (r - w)) * T;
    } else if (r < w) {
      dst = a.Xp + (size_t)(c0 + r) * T;
      out = a.Out + (size_t)a.perm[c0 + r] * a.ldo;
    }
    if (!dst) continue;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int col = 8 * nb + 2 * lk;
      if (T >= 2) {
        if (col < T) {
          *reinterpret_cast<double2*>(dst + col) = make_double2(acc[rg][nb][0], acc[rg][nb][1]);
          if (out) {
            if (col < a.t) out[col] = acc[rg][nb][0];
            if (col + 1 < a.t) out[col + 1] = acc[rg][nb][1];
          }
        }
      } else if (col == 0) {
        dst[0] = acc[rg][nb][0];
        if (out) out[0] = acc[rg][nb][0];
      }
    }
  }
}

// One warp per short panel (and on levels with many short panels several panels in a row), or the 8 warps of a CTA on
// disjoint k ranges of one long panel.  Per warp and k-block (4 steps): 1 KB of the panel (32 B per lane, already in
// A-fragment order), one shared-memory double per lane (B fragment of the T-wide input rows) and 4 DMMA per 8 output columns.
//   (1) panel data: a warp-private shared-memory ring of 2 stages x 4 k-blocks filled by cp.async.bulk (struct Ring), always
//       8 KB in flight per warp (HBM latency) without a register spent on it; the tile-wise backward sweep of the
//       single-copy factor keeps a register ring of two 2 KB units,
//   (2) the input rows of the next tile of k steps are copied into a warp-private shared-memory buffer with
//       cp.async while the current tile is consumed,
//   (3) the accumulators are 2 doubles per 8x8 block (8 doubles at T = 8).
// History (128^3, t=8, whole apply): plain FMA loop 2.1 TB/s (long-scoreboard bound), + register ring 2.6,
// + cp.async tiles 3.5, DMMA + fragment-order panels 4.3 ... 5.2, shared-memory ring 5.7: see profiles/.
// TCOPY (backward only): the factor also holds the transposed panels (bj.h), the backward sweep streams them exactly like
// the forward sweep streams M; without them it reads M tile by tile.
template <int T, bool FWD, int D, bool NOALLOC, int OCC, bool TCOPY>
__global__ void __launch_bounds__(kThreads, (T <= 8) ? OCC : ((T == 32 && !FWD && !TCOPY) ? 1 : 2)) sweep_kernel(SweepArgs a) {
  constexpr bool STREAM = FWD || TCOPY;   // the panel is one contiguous run of k-blocks
  constexpr int NB = (T + 7) / 8;       // 8-column blocks of the output
  constexpr int KT = Tile<T, STREAM>::KT;  // k steps per input tile
  constexpr int KB = KT / 4;            // k-blocks per tile
  constexpr int TILE = KT * T;          // doubles per tile buffer
  PCU_DYN_SMEM(smem);
  double* red = smem;                   // 32 * T doubles
  const WorkUnit u = a.units[blockIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = lane >> 2, lk = lane & 3;   // fragment row / k (or column pair)
  double* tile0 = smem + 32 * T + (size_t)warp * 2 * TILE;
  if (!u.split && warp >= u.count) return;
  const int pidx = u.split ? u.first : u.first + warp;
  long long off;
  int klen, c0, w, h, row0;
  long long uoff = 0;
  const int* rows = a.rows;
  auto load_desc = [&](int idx) {
    if (FWD) {
      const FwdPanel p = reinterpret_cast<const FwdPanel*>(a.panels)[idx];
      off = p.off; klen = p.klen; c0 = p.c0; w = p.w; h = p.h; row0 = p.row0; uoff = p.uoff;
    } else {
      const BwdPanel p = reinterpret_cast<const BwdPanel*>(a.panels)[idx];
      off = p.off; klen = p.klen; c0 = p.c0; w = p.w; h = p.h; row0 = p.k0; rows = a.rows + p.rows_off;
    }
  };
  load_desc(pidx);
  const int nkb = klen >> 2;
  int q0 = 0, q1 = nkb;
  if (u.split) {
    const int s0 = (u.split == 2) ? u.kb0 : 0, s1 = (u.split == 2) ? u.kb1 : nkb;
    int per = (s1 - s0 + kWarps - 1) / kWarps;
    per = (per + KB - 1) / KB * KB;     // whole tiles per warp
    q0 = min(s1, s0 + warp * per);
    q1 = min(s1, q0 + per);
  }
  double acc[4][NB][2];
#pragma unroll
  for (int rg = 0; rg < 4; ++rg)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) acc[rg][nb][0] = acc[rg][nb][1] = 0.0;

  // copy the input rows of steps [4*tq, 4*tq + KT) into buf (steps past the panel are clamped: their
  // panel entries are zero padding)
  auto stage = [&](int tq, double* buf) {
    constexpr int CHUNKS = (T >= 2) ? TILE / 2 : KT;  // 16-byte pieces per tile (rows when T == 1)
    constexpr int CPR = (T >= 2) ? T / 2 : 1;         // pieces per row
#pragma unroll
    for (int c0q = 0; c0q < CHUNKS; c0q += 32) {
      const int q = c0q + lane;
      if (T >= 2) {
        const int r = q / CPR, part = q % CPR;
        const int k = min(4 * tq + r, klen - 1);
        const double* src;
        if (FWD) src = a.Bsrc ? a.Bsrc + (size_t)__ldg(a.perm + c0 + k) * a.ldb : a.Wk + (size_t)(c0 + k) * T;
        else {
          const int i = min(row0 + k, h - 1);
          src = (i < w) ? a.Y + (size_t)(c0 + i) * T : a.Xp + (size_t)__ldg(rows + i) * T;
        }
        cp_async16(buf + (size_t)r * T + 2 * part, src + 2 * part);
      } else if (q < KT) {
        const int k = min(4 * tq + q, klen - 1);
        const double* src;
        if (FWD) src = a.Bsrc ? a.Bsrc + (size_t)__ldg(a.perm + c0 + k) * a.ldb : a.Wk + (size_t)(c0 + k);
        else { const int i = min(row0 + k, h - 1); src = (i < w) ? a.Y + (size_t)(c0 + i) : a.Xp + (size_t)__ldg(rows + i); }
        buf[q] = *src;
      }
    }
    cp_async_commit();
  };

  if constexpr (STREAM) {
    // panel data through a warp-private shared-memory ring: RS stages of RK k-blocks, each stage one cp.async.bulk copy
    // (SASS UBLKCP, L2 evict-first) issued by lane 0 and completed on its own mbarrier; the lanes read their A fragments
    // from it with two conflict-free 16-byte loads per k-block (struct Ring)
    constexpr int RK = Ring<T>::RK, RS = Ring<T>::RS;
    static_assert(KB % RK == 0, "a tile is a whole number of ring stages");
    double* ringbuf = smem + 32 * T + (size_t)kWarps * 2 * TILE + (size_t)warp * Ring<T>::DOUBLES;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + 32 * T + (size_t)kWarps * 2 * TILE +
                                                                     (size_t)kWarps * Ring<T>::DOUBLES) + warp * RS;
    const unsigned long long pol = l2_evict_first_policy();
    if (lane == 0) {
#pragma unroll
      for (int sg = 0; sg < RS; ++sg) mbar_init(bars + sg, 1);
      fence_mbar_init();
    }
    __syncwarp();
    // A warp that owns a whole panel goes on to the warp's next panels (u.first + warp + 8, + 16, ...): lane 0 keeps
    // the ring filled ACROSS panels (a stage never mixes two panels) and the first input tile of the next panel is staged
    // during the last tile of the current one, so the chain "descriptor -> first panel data -> gather indices -> input
    // rows" is paid once per warp, not once per panel (the short panels of the lower half of the forest are a few
    // microseconds of streaming each).
    const int pidx_end = u.split ? 0 : u.first + u.count;
    const double* pdata = a.data + off;
    int pk = q0, pend = q1, pnext = pidx + kWarps;   // producer (lane 0): next k-block to fetch, end, the panel after
    auto refill = [&](int sg) {
      if (pk >= pend) {
        if (pnext >= pidx_end) return;
        long long noff; int nklen;
        if (FWD) { const FwdPanel* np = reinterpret_cast<const FwdPanel*>(a.panels) + pnext; noff = __ldg(&np->off); nklen = __ldg(&np->klen); }
        else { const BwdPanel* np = reinterpret_cast<const BwdPanel*>(a.panels) + pnext; noff = __ldg(&np->off); nklen = __ldg(&np->klen); }
        pdata = a.data + noff; pk = 0; pend = nklen >> 2; pnext += kWarps;
      }
      const int nk = min(RK, pend - pk);
      mbar_arrive_expect_tx(bars + sg, (unsigned)nk * 1024u);
      bulk_g2s_stream(ringbuf + (size_t)sg * RK * 128, pdata + (size_t)pk * 128, (unsigned)nk * 1024u, bars + sg, pol);
      pk += nk;
    };
    // panel data does not depend on the previous kernel: in flight before the wait
    if (lane == 0) {
#pragma unroll
      for (int sg = 0; sg < RS; ++sg) refill(sg);
    }
    pdl_wait();
    pdl_launch_dependents();
    if (q0 < q1) {
      stage(q0, tile0);
      int tix = 0, it = 0, cidx = pidx;
      for (;;) {
      int s_row0 = row0, s_c0 = c0, s_w = w, s_h = h;   // what storing this panel's results needs
      long long s_uoff = uoff;
      bool more = false;
      for (int tq = q0; tq < q1; tq += KB, ++tix) {
        const double* cur = tile0 + (size_t)(tix & 1) * TILE;
        if (tq + KB < q1) { stage(tq + KB, tile0 + (size_t)((tix + 1) & 1) * TILE); cp_async_wait<1>(); }
        else if (cidx + kWarps < pidx_end) {   // last tile: the first input tile of the warp's next panel
          cidx += kWarps;
          load_desc(cidx);
          stage(0, tile0 + (size_t)((tix + 1) & 1) * TILE);
          more = true;
          cp_async_wait<1>();
        } else cp_async_wait<0>();
        __syncwarp();
#pragma unroll 1
        for (int kq = 0; kq < KB && tq + kq < q1; kq += RK, ++it) {
          const int sg = it % RS;
          mbar_wait(bars + sg, (unsigned)(it / RS) & 1u);
#pragma unroll
          for (int u2 = 0; u2 < RK; ++u2) {
            if (tq + kq + u2 < q1) {
              const double* mp = ringbuf + (size_t)(sg * RK + u2) * 128 + lane * 4;
              const double2 m0 = *reinterpret_cast<const double2*>(mp), m1 = *reinterpret_cast<const double2*>(mp + 2);
              const double* brow = cur + (size_t)(4 * (kq + u2) + lk) * T;
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                const double bf = (8 * nb + lr < T) ? brow[8 * nb + lr] : 0.0;
                dmma884(acc[0][nb][0], acc[0][nb][1], m0.x, bf);
                dmma884(acc[1][nb][0], acc[1][nb][1], m0.y, bf);
                dmma884(acc[2][nb][0], acc[2][nb][1], m1.x, bf);
                dmma884(acc[3][nb][0], acc[3][nb][1], m1.y, bf);
              }
            }
          }
          __syncwarp();
          if (lane == 0) refill(sg);
        }
        __syncwarp();
      }
      if (u.split) break;
      store_outputs<T, FWD>(acc, a, s_row0, s_c0, s_w, s_h, s_uoff, lr, lk);
      if (!more) return;
      q0 = 0; q1 = klen >> 2;
#pragma unroll
      for (int rg = 0; rg < 4; ++rg)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[rg][nb][0] = acc[rg][nb][1] = 0.0;
      }
    }
  } else {
    // Backward: the slice is the 32 columns [row0, row0 + 32) of the supernode's forward panels.  Tile by tile (32 rows
    // of forward slice p = row0/32 + tile), the k-blocks [8*qcol, 8*qcol + nkq) of that slice are read in units of two
    // (2 KB; the ring holds two units = the 4 KB in flight per warp of the forward sweep) in the lane order of
    // bwd_lane_offset: unit jg feeds the output rows [8*jg, 8*jg + 8) with the tile's 32 input rows (8 DMMA steps).
    static_assert(KB % 8 == 0, "the backward sweep consumes whole 32-row tiles");
    const int W4 = (w + 3) >> 2;
    const int qcol = row0 >> 5;
    const int nkq = min(8, W4 - 8 * qcol);
    const int tl1 = q1 >> 3;
    const int khalf = lane >> 4;   // which k-block of a unit this lane reads
    int pcur = qcol + (q0 >> 3);
    const double* cur_base = a.data + off + panel_cum(w, pcur) + (long long)qcol * 1024 + bwd_lane_offset(lane);
    const double* nxt_base = cur_base + 128ll * min(W4, 8 * pcur + 8);
    double2 ue0[2], ue1[2], uo0[2], uo1[2];   // per ring slot: even steps (ib = 0, 2, 4, 6), odd steps
    auto issue_unit = [&](const double* tb, int jg, bool valid, double2& e0, double2& e1, double2& o0, double2& o1) {
      if (valid && 2 * jg + khalf < nkq) {
        ld_stream4(tb + (size_t)(2 * jg) * 128, e0, e1);
        ld_stream4(tb + (size_t)(2 * jg) * 128 + 64, o0, o1);
      } else {
        e0 = make_double2(0.0, 0.0); e1 = e0; o0 = e0; o1 = e0;
      }
    };
    issue_unit(cur_base, 0, q0 < q1, ue0[0], ue1[0], uo0[0], uo1[0]);
    issue_unit(cur_base, 1, q0 < q1, ue0[1], ue1[1], uo0[1], uo1[1]);
    pdl_wait();
    pdl_launch_dependents();
    if (q0 < q1) {
      stage(q0, tile0);
      int tix = 0;
      for (int tq = q0; tq < q1; tq += KB, ++tix) {
        const double* cur = tile0 + (size_t)(tix & 1) * TILE;
        if (tq + KB < q1) { stage(tq + KB, tile0 + (size_t)((tix + 1) & 1) * TILE); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
#pragma unroll 1
        for (int mt = 0; mt < KB / 8; ++mt) {
          const int ti = (tq >> 3) + mt;   // tile of the slice
          if (ti >= tl1) break;
          const bool more = ti + 1 < tl1;
#pragma unroll
          for (int jg = 0; jg < 4; ++jg) {
            const bool live = 2 * jg < nkq;
            const double av[8] = {ue0[jg & 1].x, uo0[jg & 1].x, ue0[jg & 1].y, uo0[jg & 1].y,
                                  ue1[jg & 1].x, uo1[jg & 1].x, ue1[jg & 1].y, uo1[jg & 1].y};
            if (jg < 2) issue_unit(cur_base, jg + 2, true, ue0[jg & 1], ue1[jg & 1], uo0[jg & 1], uo1[jg & 1]);
            else issue_unit(nxt_base, jg - 2, more, ue0[jg & 1], ue1[jg & 1], uo0[jg & 1], uo1[jg & 1]);
            if (live) {
              const double* brow = cur + (size_t)(32 * mt + lk) * T;
#pragma unroll
              for (int ib = 0; ib < 8; ++ib)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                  const double bf = (8 * nb + lr < T) ? brow[(size_t)(4 * ib) * T + 8 * nb + lr] : 0.0;
                  dmma884(acc[jg][nb][0], acc[jg][nb][1], av[ib], bf);
                }
            }
          }
          ++pcur;
          cur_base = nxt_base;
          nxt_base += 128ll * min(W4, 8 * pcur + 8);
        }
        __syncwarp();
      }
    }
  }
  // lane owns, for every row group rg and column block nb: row 8*rg + lane/4, columns 8*nb + 2*(lane%4) + {0,1}
  if (u.split) {
    // fixed-order combine over the warps: every warp parks its partial sums in its own (now idle) tile buffer,
    // then each thread adds the 8 partials of one or more entries in warp order
#pragma unroll
    for (int rg = 0; rg < 4; ++rg)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * nb + 2 * lk + e;
          if (col < T) tile0[(8 * rg + lr) * T + col] = acc[rg][nb][e];
        }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * T; e += kThreads) {
      double sum = smem[32 * T + e];
#pragma unroll
      for (int wv = 1; wv < kWarps; ++wv) sum += smem[32 * T + (size_t)wv * 2 * TILE + e];
      red[e] = sum;
    }
    __syncthreads();
    if (warp != 0) return;
    if (u.split == 2) {
      // long panel cut across CTAs: publish this slice, the CTA that arrives last adds all slices in slice order
      double* mine = a.scratch + (size_t)(u.slot + u.chunk) * 32 * T;
      for (int e = lane; e < 32 * T; e += 32) mine[e] = red[e];
      __threadfence();
      int last = 0;
      if (lane == 0) last = (atomicAdd(a.counters + u.cidx, 1) == u.nchunks - 1);
      last = __shfl_sync(0xffffffffu, last, 0);
      if (!last) return;
      __threadfence();
      if (lane == 0) a.counters[u.cidx] = 0;  // ready for the next apply
      for (int e = lane; e < 32 * T; e += 32) {
        double sacc = 0.0;
        for (int ch = 0; ch < u.nchunks; ++ch) sacc += __ldcg(a.scratch + (size_t)(u.slot + ch) * 32 * T + e);
        red[e] = sacc;
      }
      __syncwarp();
    }
#pragma unroll
    for (int rg = 0; rg < 4; ++rg)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * nb + 2 * lk + e;
          if (col < T) acc[rg][nb][e] = red[(8 * rg + lr) * T + col];
        }
  }
  store_outputs<T, FWD>(acc, a, row0, c0, w, h, uoff, lr, lk);
}

int pick_T(int t) { return t <= 1 ? 1 : t <= 2 ? 2 : t <= 4 ? 4 : t <= 8 ? 8 : t <= 16 ? 16 : 32; }

int ensure_work(pcu_bj* bj, int T) {
  if (bj->cap_t >= T) return 0;
  pcu_ctx* c = bj->ctx;
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(bj->Wk); cudaFree(bj->Y); cudaFree(bj->U); cudaFree(bj->Xp);
  bj->Wk = bj->Y = bj->U = bj->Xp = nullptr;
  cudaFree(bj->scratch); cudaFree(bj->counters);
  bj->scratch = nullptr; bj->counters = nullptr;
  PCU_CUDA(cudaMalloc(&bj->scratch, sizeof(double) * (size_t)(bj->scratch_slots + 1) * 32 * T));
  PCU_CUDA(cudaMalloc(&bj->counters, sizeof(int) * (size_t)(bj->ncounters + 1)));
  PCU_CUDA(cudaMemsetAsync(bj->counters, 0, sizeof(int) * (size_t)(bj->ncounters + 1), c->stream));
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

// launch with programmatic stream serialization (see pdl_wait); PREALPS_BJ_NOPDL=1 falls back to plain launches
template <typename... KArgs, typename... Args>
void launch_chain(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
  static const bool pdl = getenv("PREALPS_BJ_NOPDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <int T, bool FWD, int D, bool NOALLOC, int OCC, bool TCOPY>
void launch_one(int nu, cudaStream_t st, const SweepArgs& a) {
  constexpr int bytes = (32 * T + kWarps * 2 * Tile<T, FWD || TCOPY>::KT * T + ((FWD || TCOPY) ? kWarps * (Ring<T>::DOUBLES + Ring<T>::RS) : 0)) * (int)sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(sweep_kernel<T, FWD, D, NOALLOC, OCC, TCOPY>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    configured = true;
  }
  launch_chain(sweep_kernel<T, FWD, D, NOALLOC, OCC, TCOPY>, nu, kThreads, bytes, st, a);
}

template <int T, bool FWD>
void launch_sweep(int nu, cudaStream_t st, const SweepArgs& a, bool tcopy) {
  // ring of 4 k-blocks per warp, streaming (no L1 allocation) panel loads, 2 CTAs per SM: the best of the
  // combinations measured on B200 (deeper rings and 3 CTAs/SM spill; allocating loads are 2 % slower)
  if (FWD) launch_one<T, FWD, 4, true, 2, false>(nu, st, a);
  else if (tcopy) launch_one<T, false, 4, true, 2, true>(nu, st, a);
  else launch_one<T, false, 4, true, 2, false>(nu, st, a);
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
  // the zero row behind the update rows: U holds (nu + 4) * cap_t doubles and a narrower solve uses the front of it, so the
  // tail starts at row nu * cap_t / T of the T-wide layout whatever ran before
  const long long pad = bj->nu * (long long)(bj->cap_t / T);
  SweepArgs a{};
  a.Wk = bj->Wk; a.Y = bj->Y; a.U = bj->U; a.Xp = bj->Xp; a.rows = bj->rows; a.perm = bj->perm;
  a.Out = X; a.ldo = ldx; a.t = t;
  a.scratch = bj->scratch; a.counters = bj->counters;
  // level 0 = the leaves: nothing to gather, so the forward panels read B[perm] themselves when the caller's rows can
  // be copied T wide in 16-byte pieces (t == T, even leading dimension, aligned base)
  const char* asm_wide_env = getenv("PREALPS_BJ_ASM_WIDE");
  const int asm_wide = asm_wide_env ? atoi(asm_wide_env) : 1;
  static const bool leaf_direct_on = getenv("PREALPS_BJ_NO_LEAF_DIRECT") == nullptr;
  const bool leaf_direct = leaf_direct_on && t == T && (T == 1 || (ldb % 2 == 0 && (reinterpret_cast<size_t>(B) & 15) == 0));
  for (int l = 0; l < bj->nlevels; ++l) {
    const int ncols = bj->lvl_col_ptr[l + 1] - bj->lvl_col_ptr[l];
    a.Bsrc = (l == 0 && leaf_direct) ? B : nullptr;
    a.ldb = ldb;
    if (ncols > 0 && a.Bsrc == nullptr) {
      prof.mark("asm L" + std::to_string(l) + " cols=" + std::to_string(ncols), 3.0 * ncols * T * 8);
      const int grid = stream_grid(c, (long long)ncols * G, kThreads, 8);
      // one warp per column where the lists are long AND the columns few (one subdomain per GPU: 0.683 -> 0.671 ms per apply);
      // with tens of thousands of columns the 4-lane kernel already fills the machine and is 0.7 % faster (8 blocks per GPU)
      if ((bj->lvl_long_lists[l] && ncols <= 8192 && asm_wide) || asm_wide == 2)   // 2: every level (tests)
        launch_chain(assemble_wide_kernel<T>, stream_grid(c, (long long)ncols * 32, kThreads, 8), kThreads, 0, st,
                     bj->lvl_cols + bj->lvl_col_ptr[l], ncols, B, ldb, t, bj->perm, bj->gl_ptr, bj->gl_idx, bj->U, pad, bj->Wk);
      else if (bj->lvl_long_lists[l])
        launch_chain(assemble_kernel<T, 16>, grid, kThreads, 0, st, bj->lvl_cols + bj->lvl_col_ptr[l], ncols, B, ldb, t,
                     bj->perm, bj->gl_ptr, bj->gl_idx, bj->U, pad, bj->Wk);
      else
        launch_chain(assemble_kernel<T, 8>, grid, kThreads, 0, st, bj->lvl_cols + bj->lvl_col_ptr[l], ncols, B, ldb, t,
                     bj->perm, bj->gl_ptr, bj->gl_idx, bj->U, pad, bj->Wk);
      PCU_LAUNCH_CHECK(c);
    }
    const int nu = bj->fwd_unit_ptr[l + 1] - bj->fwd_unit_ptr[l];
    if (nu > 0) {
      prof.mark("fwd L" + std::to_string(l) + " ctas=" + std::to_string(nu), bj->fwd_lvl_bytes[l]);
      a.units = bj->fwd_units + bj->fwd_unit_ptr[l];
      a.panels = bj->fwd_panels;
      a.data = bj->fwd_data;
      launch_sweep<T, true>(nu, st, a, false);
      PCU_LAUNCH_CHECK(c);
    }
  }
  for (int l = bj->nlevels - 1; l >= 0; --l) {
    const int nu = bj->bwd_unit_ptr[l + 1] - bj->bwd_unit_ptr[l];
    if (nu > 0) {
      prof.mark("bwd L" + std::to_string(l) + " ctas=" + std::to_string(nu), bj->bwd_lvl_bytes[l]);
      a.units = bj->bwd_units + bj->bwd_unit_ptr[l];
      a.panels = bj->bwd_panels;
      a.data = bj->bwd_data ? bj->bwd_data : bj->fwd_data;
      launch_sweep<T, false>(nu, st, a, bj->bwd_data != nullptr);
      PCU_LAUNCH_CHECK(c);
    }
  }
  prof.report();
  return 0;
}

int dispatch_apply(pcu_bj* bj, const double* B, int ldb, double* X, int ldx, int t) {
  // the work vectors are laid out for cap_t columns; a narrower solve uses its own T, which is
  // consistent within one apply (every vector is rewritten before it is read)
  switch (pick_T(t)) {
    case 1: return apply_T<1>(bj, B, ldb, X, ldx, t);
    case 2: return apply_T<2>(bj, B, ldb, X, ldx, t);
    case 4: return apply_T<4>(bj, B, ldb, X, ldx, t);
    case 8: return apply_T<8>(bj, B, ldb, X, ldx, t);
    case 16: return apply_T<16>(bj, B, ldb, X, ldx, t);
    default: return apply_T<32>(bj, B, ldb, X, ldx, t);
  }
}

}  // namespace

extern "C" int pcu_bj_apply(pcu_bj* bj, const double* B, int ldb, double* X, int ldx, int t) {
  PCU_CHECK(bj && B && X && t >= 1 && t <= 32, "pcu_bj_apply: bad arguments (t=%d, need 1..32)", t);
  PCU_CHECK(ldb >= t && ldx >= t, "pcu_bj_apply: leading dimension smaller than t");
  if (ensure_work(bj, pick_T(t))) return 1;
  return dispatch_apply(bj, B, ldb, X, ldx, t);
}
