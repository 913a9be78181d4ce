// spmm_kernels.cuh -- device code of K1 (CSR SpMM over the m x t enlarged block) and the pure host helpers that define
// its data layout (row blocks, local/halo split).  Included by spmm.cu; tests/emul/ compiles the same text with a CPU
// shim (PCU_EMUL: one pthread per CUDA thread, test infrastructure only) to check the index logic without a GPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace {


constexpr int kThreads = 256;
constexpr int kNnzCap = 1792;   // staged entries per CTA (21 KB of shared memory) and rows per CTA: the static arrays
constexpr int kRowCap = 256;    // are sized for the larger of the two row-block shapes below
// Row blocks come in two shapes. With few lanes per row (t <= 4) a CTA covers many rows per pass and larger blocks
// amortise the descriptor + staging latency (7-point, t = 1: 78 -> 61 us); from t = 8 up the smaller blocks win
// (t = 32: 289 vs 316 us): more CTAs in different phases per SM.
constexpr int kShapeRows[2] = {128, 256};
constexpr int kShapeNnz[2] = {1536, 1792};

struct SpmmArgs {
  const int* rowPtr;
  const int* colInd;
  const double* val;
  const int4* blk;  // per row block: {first row, end row, first entry, end entry} -- one load instead of a chain of three
  int m;
  const double* X;
  int ldx;
  const double* H;  // halo rows, ld = t
  double* Y;
  int ldy;
  int t;
};

#ifndef PCU_EMUL
__device__ __forceinline__ double2 ldg2(const double* p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): 4 columns of a row per lane
__device__ __forceinline__ void ldg4(const double* p, double (&x)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
}
__device__ __forceinline__ void stg4(double* p, const double (&x)[4]) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
}
#else  // CPU shim of tests/emul: same accesses, and the alignment the vector instructions need is asserted
inline double2 ldg2(const double* p) { pcu_emul_check_aligned(p, 16); return *reinterpret_cast<const double2*>(p); }
inline void ldg4(const double* p, double (&x)[4]) { pcu_emul_check_aligned(p, 32); for (int j = 0; j < 4; ++j) x[j] = p[j]; }
inline void stg4(double* p, const double (&x)[4]) { pcu_emul_check_aligned(p, 32); for (int j = 0; j < 4; ++j) p[j] = x[j]; }
#endif

// T columns, G = T / CPL lanes per row, each lane owns CPL adjacent columns: 4 with 256-bit accesses (T >= 8, rows
// 32-byte aligned: half the lanes, so twice the rows per warp and half the instructions per row), else 2, or 1 (T == 1).
template <int T, int CPL>
__global__ void __launch_bounds__(kThreads) spmm_kernel(SpmmArgs a) {
  constexpr int G = T / CPL;
  constexpr int NG = kThreads / G;
  __shared__ int s_col[kNnzCap];
  __shared__ double s_val[kNnzCap];
  __shared__ int s_rp[kRowCap + 1];

  const int4 d = __ldg(a.blk + blockIdx.x);
  const int r0 = d.x, r1 = d.y, p0 = d.z, p1 = d.w;
  const int n = p1 - p0;
  const int tid = threadIdx.x;

  if (n <= kNnzCap) {
    for (int i = tid; i < n; i += kThreads) {
      s_col[i] = __ldg(a.colInd + p0 + i);
      s_val[i] = __ldg(a.val + p0 + i);
    }
    for (int i = tid; i <= r1 - r0; i += kThreads) s_rp[i] = __ldg(a.rowPtr + r0 + i) - p0;
    __syncthreads();
    const int grp = tid / G, lig = tid % G;
    for (int r = r0 + grp; r < r1; r += NG) {
      const int b = s_rp[r - r0], e = s_rp[r - r0 + 1];
      if (CPL == 4) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
        for (int p = b; p < e; ++p) {
          const int c = s_col[p];
          const double v = s_val[p];
          const double* src = (c < a.m) ? a.X + (size_t)c * a.ldx : a.H + (size_t)(c - a.m) * a.t;
          double x[4];
          ldg4(src + 4 * lig, x);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fma(v, x[j], acc[j]);
        }
        stg4(a.Y + (size_t)r * a.ldy + 4 * lig, acc);
        continue;
      }
      double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
      for (int p = b; p < e; ++p) {
        const int c = s_col[p];
        const double v = s_val[p];
        const double* src = (c < a.m) ? a.X + (size_t)c * a.ldx : a.H + (size_t)(c - a.m) * a.t;
        if (CPL == 2) {
          const double2 x = ldg2(src + 2 * lig);
          acc0 = fma(v, x.x, acc0);
          acc1 = fma(v, x.y, acc1);
        } else {
          acc0 = fma(v, __ldg(src), acc0);
        }
      }
      double* dst = a.Y + (size_t)r * a.ldy;
      if (CPL == 2) *reinterpret_cast<double2*>(dst + 2 * lig) = make_double2(acc0, acc1);
      else dst[0] = acc0;
    }
  } else {
    // a single very long row (the host never puts two rows in an oversized block):
    // all threads stride over it, then a fixed-order tree reduction in shared memory.
    double* red = s_val;  // kNnzCap >= kThreads * 2
    const int r = r0;
    for (int c0 = 0; c0 < T; c0 += 2) {
      double acc0 = 0.0, acc1 = 0.0;
      for (int p = p0 + tid; p < p1; p += kThreads) {
        const int c = a.colInd[p];
        const double v = a.val[p];
        const double* src = (c < a.m) ? a.X + (size_t)c * a.ldx : a.H + (size_t)(c - a.m) * a.t;
        acc0 = fma(v, src[c0], acc0);
        if (c0 + 1 < T) acc1 = fma(v, src[c0 + 1], acc1);
      }
      red[tid] = acc0;
      red[kThreads + tid] = acc1;
      __syncthreads();
      for (int s = kThreads / 2; s > 0; s >>= 1) {
        if (tid < s) { red[tid] += red[tid + s]; red[kThreads + tid] += red[kThreads + tid + s]; }
        __syncthreads();
      }
      if (tid == 0) {
        a.Y[(size_t)r * a.ldy + c0] = red[0];
        if (c0 + 1 < T) a.Y[(size_t)r * a.ldy + c0 + 1] = red[kThreads];
      }
      __syncthreads();
    }
  }
}

// ---- bulk-copy staging (the default from t = 8 up; measured on B200 against the LDG + STS staging of spmm_kernel:
// 7-point 128^3 t = 8: 109 -> 102 us, 27-point: 304 -> 256 us, profiles/r02_candidates_ab.md)
#ifndef PCU_EMUL
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// cp.async.bulk global -> shared (SASS UBLKCP): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
               "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned phase) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(phase)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void pcu_trap() { __trap(); }
#else
inline void mbar_init(unsigned long long*, unsigned) {}
inline void fence_mbar_init() {}
inline void mbar_arrive_expect_tx(unsigned long long*, unsigned) {}
inline void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long*) {
  pcu_emul_check_aligned(smem, 16); pcu_emul_check_aligned(gmem, 16);
  if (bytes % 16) std::abort();
  std::memcpy(smem, gmem, bytes);
}
inline bool mbar_try_wait(unsigned long long*, unsigned) { return true; }
inline void pcu_trap() { std::abort(); }
#endif

// Same mapping and summation order as spmm_kernel<T, CPL>.  The col/val chunk of the row block lands in shared memory
// through two cp.async.bulk copies issued by one thread (no register round trip, no 6 rounds of LDG + STS by the whole
// CTA), completion on an mbarrier; the chunk starts are aligned down to 16 bytes (oc / ov entries of slack in front)
// and the sizes rounded up (the device arrays carry 16 bytes of slack at the end).  HALO = false (no column >= m: one
// process, or the local part of the overlapped product) drops the "block row or halo row" select, so the row phase is
// LDS + LDS + IMAD.WIDE + LDG + CPL DFMA per entry.  Needs ldx == T and every row block within the staging capacity.
// gathers a lane issues back to back (row-loop unroll) against resident CTAs, measured on B200 (128^3, t = 8, us; profiles/
// r02_spmm_occupancy.md): 4 columns per lane (7-point): unroll 4 at 40 registers / 6 CTAs per SM 101.9, unroll 8 at 48 / 5: 109.6,
// unroll 2 at 32 / 8: 95.3 -- a full SM of threads beats more loads in flight per thread; 2 columns per lane keeps unroll 4
#ifndef PCU_SPMM_MINB4
#define PCU_SPMM_MINB4 8
#endif
#ifndef PCU_SPMM_UNROLL4
#define PCU_SPMM_UNROLL4 2
#endif
#ifndef PCU_SPMM_UNROLL2
#define PCU_SPMM_UNROLL2 4
#endif
constexpr int kRowUnroll4 = PCU_SPMM_UNROLL4, kRowUnroll2 = PCU_SPMM_UNROLL2;
template <int T, int CPL, bool HALO>
__global__ void __launch_bounds__(kThreads, CPL == 4 ? PCU_SPMM_MINB4 : 8) spmm_bulk_kernel(SpmmArgs a) {  // 32 registers, no spills

  static_assert(CPL == 2 || CPL == 4, "lanes own 2 or 4 adjacent columns");
  constexpr int G = T / CPL;
  constexpr int NG = kThreads / G;
  constexpr int kCap = kShapeNnz[0];
  __shared__ __align__(16) int s_col[kCap + 8];
  __shared__ __align__(16) double s_val[kCap + 4];
  __shared__ int s_rp[kShapeRows[0] + 1];
  __shared__ __align__(8) unsigned long long s_bar;

  const int4 d = __ldg(a.blk + blockIdx.x);
  const int r0 = d.x, r1 = d.y, p0 = d.z, p1 = d.w;
  const int tid = threadIdx.x;
  const int pc = p0 & ~3, pv = p0 & ~1;  // 16-byte aligned starts of the two chunks
  const unsigned bc = (unsigned)(((p1 - pc) * 4 + 15) & ~15), bv = (unsigned)(((p1 - pv) * 8 + 15) & ~15);
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0 && p1 > p0) {
    mbar_arrive_expect_tx(&s_bar, bc + bv);
    bulk_g2s(s_col, a.colInd + pc, bc, &s_bar);
    bulk_g2s(s_val, a.val + pv, bv, &s_bar);
  }
  for (int i = tid; i <= r1 - r0; i += kThreads) s_rp[i] = __ldg(a.rowPtr + r0 + i) - p0;
  __syncthreads();
  if (p1 > p0) {  // bounded wait: a lost transaction traps instead of hanging the GPU
    unsigned spins = 0;
    while (!mbar_try_wait(&s_bar, 0))
      if (++spins > (1u << 26)) pcu_trap();
  }
  const int oc = p0 - pc, ov = p0 - pv;
  const int grp = tid / G, lig = tid % G;
  const double* xl = a.X + CPL * lig;
  const double* hl = HALO ? a.H + CPL * lig - (size_t)a.m * T : nullptr;  // halo row c - m at hl + c * T
  for (int r = r0 + grp; r < r1; r += NG) {
    const int b = s_rp[r - r0], e = s_rp[r - r0 + 1];
    double acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = 0.0;
    auto entry = [&](int p) {
      const int c = s_col[oc + p];
      const double v = s_val[ov + p];
      const double* src = ((HALO && c >= a.m) ? hl : xl) + (size_t)c * T;
      if constexpr (CPL == 4) {
        double x[4];
        ldg4(src, x);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fma(v, x[j], acc[j]);
      } else {
        const double2 x = ldg2(src);
        acc[0] = fma(v, x.x, acc[0]);
        acc[1] = fma(v, x.y, acc[1]);
      }
    };
    if constexpr (CPL == 4) {
#pragma unroll (kRowUnroll4)
      for (int p = b; p < e; ++p) entry(p);
    } else {
#pragma unroll (kRowUnroll2)
      for (int p = b; p < e; ++p) entry(p);
    }
    double* dst = a.Y + (size_t)r * a.ldy + CPL * lig;
    if constexpr (CPL == 4) {
      stg4(dst, acc);
    } else {
      *reinterpret_cast<double2*>(dst) = make_double2(acc[0], acc[1]);
    }
  }
}

// any 1 <= t <= 32: 16 lanes per row, lane owns columns lig and lig+16
__global__ void __launch_bounds__(kThreads) spmm_kernel_generic(SpmmArgs a) {
  constexpr int G = 16, NG = kThreads / G;
  __shared__ int s_col[kNnzCap];
  __shared__ double s_val[kNnzCap];
  const int4 d = __ldg(a.blk + blockIdx.x);
  const int r0 = d.x, r1 = d.y;
  const int tid = threadIdx.x, grp = tid / G, lig = tid % G;
  const int t = a.t;
  // rows are processed in sub-blocks that fit the staging buffers
  for (int rs = r0; rs < r1;) {
    int re = rs;
    const int p0 = a.rowPtr[rs];
    while (re < r1 && a.rowPtr[re + 1] - p0 <= kNnzCap) ++re;
    if (re == rs) {  // one row longer than the buffer: direct global reads
      const int p1 = a.rowPtr[rs + 1];
      if (grp == 0) {
        double acc0 = 0.0, acc1 = 0.0;
        for (int p = p0; p < p1; ++p) {
          const int c = a.colInd[p];
          const double v = a.val[p];
          const double* src = (c < a.m) ? a.X + (size_t)c * a.ldx : a.H + (size_t)(c - a.m) * t;
          if (lig < t) acc0 = fma(v, src[lig], acc0);
          if (lig + 16 < t) acc1 = fma(v, src[lig + 16], acc1);
        }
        if (lig < t) a.Y[(size_t)rs * a.ldy + lig] = acc0;
        if (lig + 16 < t) a.Y[(size_t)rs * a.ldy + lig + 16] = acc1;
      }
      rs += 1;
      continue;
    }
    const int n = a.rowPtr[re] - p0;
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) { s_col[i] = a.colInd[p0 + i]; s_val[i] = a.val[p0 + i]; }
    __syncthreads();
    for (int r = rs + grp; r < re; r += NG) {
      const int b = a.rowPtr[r] - p0, e = a.rowPtr[r + 1] - p0;
      double acc0 = 0.0, acc1 = 0.0;
      for (int p = b; p < e; ++p) {
        const int c = s_col[p];
        const double v = s_val[p];
        const double* src = (c < a.m) ? a.X + (size_t)c * a.ldx : a.H + (size_t)(c - a.m) * t;
        if (lig < t) acc0 = fma(v, __ldg(src + lig), acc0);
        if (lig + 16 < t) acc1 = fma(v, __ldg(src + lig + 16), acc1);
      }
      if (lig < t) a.Y[(size_t)r * a.ldy + lig] = acc0;
      if (lig + 16 < t) a.Y[(size_t)r * a.ldy + lig + 16] = acc1;
    }
    rs = re;
  }
}

__global__ void halo_pack_kernel(const double* __restrict__ X, int ldx, int t, const int* __restrict__ idx,
                                 int nrows, double* __restrict__ out) {
  const int64_t total = (int64_t)nrows * t;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / t), c = (int)(i % t);
    out[i] = X[(size_t)idx[r] * ldx + c];
  }
}

// Overlapped product (the default under NCCL; PREALPS_SPMM_OVERLAP=0 serialises): Y[r, :] += sum_k hval[k] * H[hcol[k], :] for the rows that read
// halo rows, after the local kernel has written the partial sums of the entries with column < m.  Halo columns sort
// after the local ones, so continuing each row's FMA chain from the stored partial sum reproduces the merged kernel's
// summation order bit for bit.  16 lanes per row, lane owns columns lig and lig + 16 (any t <= 32).
__global__ void __launch_bounds__(kThreads) halo_add_kernel(int nb, const int* __restrict__ brow, const int* __restrict__ hptr,
                                                            const int* __restrict__ hcol, const double* __restrict__ hval,
                                                            const double* __restrict__ H, int t, double* Y, int ldy) {
  constexpr int G = 16;
  const int gid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / G), lig = threadIdx.x % G;
  const int ngroups = (int)((long long)gridDim.x * blockDim.x / G);
  for (int q = gid; q < nb; q += ngroups) {
    double* y = Y + (size_t)brow[q] * ldy;
    const int b = hptr[q], e = hptr[q + 1];
    double acc0 = (lig < t) ? y[lig] : 0.0, acc1 = (lig + 16 < t) ? y[lig + 16] : 0.0;
    for (int p = b; p < e; ++p) {
      const double v = hval[p];
      const double* src = H + (size_t)hcol[p] * t;
      if (lig < t) acc0 = fma(v, __ldg(src + lig), acc0);
      if (lig + 16 < t) acc1 = fma(v, __ldg(src + lig + 16), acc1);
    }
    if (lig < t) y[lig] = acc0;
    if (lig + 16 < t) y[lig + 16] = acc1;
  }
}

// row blocks of shape sh: <= kShapeRows[sh] rows and <= kShapeNnz[sh] entries; an over-long row gets a block of its own
inline void build_row_blocks(int m, const int* rowPtr, int sh, std::vector<int4>* blk) {
  blk->clear();
  for (int r = 0; r < m;) {
    int e = r;
    while (e < m && e - r < kShapeRows[sh] && rowPtr[e + 1] - rowPtr[r] <= kShapeNnz[sh]) ++e;
    if (e == r) e = r + 1;
    blk->push_back(make_int4(r, e, rowPtr[r], rowPtr[e]));
    r = e;
  }
}

// the panel split for the overlapped product: (lrp, lci, lv) keeps the entries with column < m of every row,
// (brow, hptr, hcol, hv) the halo entries (column - m) of the rows that have any, rows ascending
inline void split_local_halo(int m, const int* rowPtr, const int* colInd, const double* val, std::vector<int>* lrp,
                             std::vector<int>* lci, std::vector<double>* lv, std::vector<int>* brow, std::vector<int>* hptr,
                             std::vector<int>* hcol, std::vector<double>* hv) {
  lrp->assign(m + 1, 0);
  lci->clear(); lv->clear(); brow->clear(); hcol->clear(); hv->clear();
  hptr->assign(1, 0);
  for (int r = 0; r < m; ++r) {
    bool boundary = false;
    for (int p = rowPtr[r]; p < rowPtr[r + 1]; ++p) {
      if (colInd[p] < m) { lci->push_back(colInd[p]); lv->push_back(val[p]); }
      else { hcol->push_back(colInd[p] - m); hv->push_back(val[p]); boundary = true; }
    }
    (*lrp)[r + 1] = (int)lci->size();
    if (boundary) { brow->push_back(r); hptr->push_back((int)hcol->size()); }
  }
}

}  // namespace
