// bj_factor.cu -- K3: numeric supernodal (multifrontal) Cholesky of the block-Jacobi
// diagonal blocks on the device, and packing of the factor into the streaming panel
// layout of bj.h.  Replaces MKL PARDISO phase 12 as called by
// preAlps_BlockJacobiCreate (ref: src/preconditioners/block_jacobi.c:26-63,
// utils/cplm_light/cplm_kernels.c:741-783).
//
// Host: per-block symbolic analysis (bj_symbolic.cpp, integer work), forest assembly,
// task lists.  Device, level by level over the supernodal elimination forest:
//   front assembly (scatter of A, extend-add of the children's update matrices),
//   blocked left-looking Cholesky of the h x w panel (64-wide tiles: FP64 GEMM, tile
//   potrf with explicit tile inverse, row-tile multiply),
//   Schur complement update S = F22 - L21 L21^T,
//   M = [I; L21] L11^{-1} by a right-to-left block recursion (same two kernels),
//   packing of M and M^T into 32-row k-major panels.
// No atomics anywhere: every output element has exactly one writer per launch, and the
// children of one parent are extend-added in separate rounds => bit-reproducible.
#include <algorithm>
#include <chrono>
#include <map>
#include <numeric>
#include <thread>

#include "bj.h"
#include "bj_symbolic.h"

namespace {

using namespace pcu;

constexpr int NB = 64;        // tile width of the dense kernels
constexpr int kThreads = 256;

struct SnDev {
  long long zoff;  // h*w, column-major, ld = h        (level-local buffer)
  long long moff;  // h*w, column-major, ld = h        (level-local buffer)
  long long soff;  // (h-w)^2, column-major, ld = h-w  (global buffer)
  long long doff;  // ceil(w/NB) tiles of NB*NB: inverses of the diagonal tiles (level-local)
  int h, w;
};

struct AEntry { long long dst; double val; };

struct EaTask {    // extend-add of one child's update matrix into its parent's front
  long long src;   // child S offset
  long long rel;   // offset into rel[]: position of each child update row inside the parent's row list
  long long pz;    // parent zoff
  long long ps;    // parent soff
  int uc;          // child update size
  int hp, wp;
  int pad_;
};

struct PackTask {
  long long moff;
  long long dst;
  int h, w;
  int row0;   // fwd: first row of the slice; bwd: first column (= first contributing row)
  int klen;
};

// ------------------------------------------------------------------ kernels
__global__ void scatter_a_kernel(const AEntry* __restrict__ e, long long n, double* __restrict__ Z) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    Z[e[i].dst] += e[i].val;
}

__global__ void extend_add_kernel(const EaTask* __restrict__ tasks, const int* __restrict__ rel,
                                  const double* Sbuf_src, double* Zbuf, double* Sbuf) {
  const EaTask t = tasks[blockIdx.x];
  const int* r = rel + t.rel;
  const double* S = Sbuf_src + t.src;
  const int uc = t.uc, up = t.hp - t.wp;
  for (int j = blockIdx.y; j < uc; j += gridDim.y) {
    const int rj = r[j];
    for (int i = j + threadIdx.x; i < uc; i += blockDim.x) {
      const int ri = r[i];
      const double v = S[i + (long long)j * uc];
      if (rj < t.wp) Zbuf[t.pz + ri + (long long)rj * t.hp] += v;
      else Sbuf[t.ps + (ri - t.wp) + (long long)(rj - t.wp) * up] += v;
    }
  }
}

struct GemmOp {
  const double* A; long long sai, sak;
  const double* B; long long sbj, sbk;
  double* C; long long ldc;
  int m, n, k, lower;
};

// mode 0: left-looking panel update at column block kb
// mode 1: Schur complement S -= Z21 Z21^T (lower triangle)
// mode 2: M recursion at column block jb (= kb)
__device__ __forceinline__ bool make_op(int mode, const SnDev& s, int kb, double* Zbuf, double* Mbuf,
                                        double* Sbuf, GemmOp& op) {
  const long long h = s.h;
  const int w = s.w;
  if (mode == 0) {
    if (kb >= w || kb == 0) return false;
    const int nb = min(NB, w - kb);
    double* Z = Zbuf + s.zoff;
    op.A = Z + kb; op.sai = 1; op.sak = h;
    op.B = Z + kb; op.sbj = 1; op.sbk = h;
    op.C = Z + kb + (long long)kb * h; op.ldc = h;
    op.m = s.h - kb; op.n = nb; op.k = kb; op.lower = 0;
    return true;
  } else if (mode == 1) {
    const int u = s.h - w;
    if (u <= 0) return false;
    double* Z = Zbuf + s.zoff;
    op.A = Z + w; op.sai = 1; op.sak = h;
    op.B = Z + w; op.sbj = 1; op.sbk = h;
    op.C = Sbuf + s.soff; op.ldc = u;
    op.m = u; op.n = u; op.k = w; op.lower = 1;
    return true;
  } else {
    if (kb >= w) return false;
    const int nb = min(NB, w - kb);
    const int kk = w - kb - nb;
    if (kk <= 0) return false;
    double* M = Mbuf + s.moff;
    const double* Z = Zbuf + s.zoff;
    op.A = M + kb + (long long)(kb + nb) * h; op.sai = 1; op.sak = h;
    op.B = Z + (kb + nb) + (long long)kb * h; op.sbj = h; op.sbk = 1;
    op.C = M + kb + (long long)kb * h; op.ldc = h;
    op.m = s.h - kb; op.n = nb; op.k = kk; op.lower = 0;
    return true;
  }
}

// C(i,j) -= sum_k A(i,k) B(j,k); 64x64 tile per CTA, 4x4 per thread, K in slabs of 16.
__global__ void __launch_bounds__(kThreads) gemm_kernel(const SnDev* __restrict__ sns, int mode, int kb,
                                                        double* Zbuf, double* Mbuf, double* Sbuf) {
  const SnDev s = sns[blockIdx.x];
  GemmOp op;
  if (!make_op(mode, s, kb, Zbuf, Mbuf, Sbuf, op)) return;
  __shared__ double As[16][64 + 4];
  __shared__ double Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int ti = tid % 16, tj = tid / 16;
  const int tiles_m = (op.m + 63) / 64, tiles_n = (op.n + 63) / 64;
  for (int bm = blockIdx.y; bm < tiles_m; bm += gridDim.y)
    for (int bn = blockIdx.z; bn < tiles_n; bn += gridDim.z) {
      const int i0 = bm * 64, j0 = bn * 64;
      if (op.lower && i0 + 63 < j0) continue;  // tile strictly above the diagonal
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
      for (int k0 = 0; k0 < op.k; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int idx = tid + q * kThreads;
          int i, kk;
          if (op.sai == 1) { i = idx % 64; kk = idx / 64; } else { kk = idx % 16; i = idx / 16; }
          const int gi = i0 + i, gk = k0 + kk;
          As[kk][i] = (gi < op.m && gk < op.k) ? op.A[gi * op.sai + gk * op.sak] : 0.0;
          int j, kj;
          if (op.sbj == 1) { j = idx % 64; kj = idx / 64; } else { kj = idx % 16; j = idx / 16; }
          const int gj = j0 + j, gk2 = k0 + kj;
          Bs[kj][j] = (gj < op.n && gk2 < op.k) ? op.B[gj * op.sbj + gk2 * op.sbk] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          double a[4], b[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { a[q] = As[kk][ti * 4 + q]; b[q] = Bs[kk][tj * 4 + q]; }
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int y = 0; y < 4; ++y) {
        const int gj = j0 + tj * 4 + y;
        if (gj >= op.n) continue;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int gi = i0 + ti * 4 + x;
          if (gi >= op.m) continue;
          if (op.lower && gi < gj) continue;
          op.C[gi + gj * op.ldc] -= acc[x][y];
        }
      }
    }
}

// Cholesky of the diagonal tile Z[kb:kb+nb, kb:kb+nb] (lower, in place) and its explicit
// inverse (lower) into the supernode's Dinv slot.  One CTA per supernode.
__global__ void __launch_bounds__(kThreads) potrf_tile_kernel(const SnDev* __restrict__ sns, int kb, double* Zbuf,
                                                              double* Dbuf, int* fail) {
  const SnDev s = sns[blockIdx.x];
  if (kb >= s.w) return;
  const int nb = min(NB, s.w - kb);
  __shared__ double L[NB][NB + 1];
  const int tid = threadIdx.x;
  double* Z = Zbuf + s.zoff + kb + (long long)kb * s.h;
  for (int e = tid; e < NB * NB; e += kThreads) {
    const int i = e % NB, j = e / NB;
    L[i][j] = (i < nb && j < nb && i >= j) ? Z[i + (long long)j * s.h] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (tid == 0) {
      const double d = L[j][j];
      if (!(d > 0.0)) { atomicMax(fail, 1); L[j][j] = 1.0; }
      else L[j][j] = sqrt(d);
    }
    __syncthreads();
    if (tid > j && tid < nb) L[tid][j] /= L[j][j];
    __syncthreads();
    // trailing update of the lower triangle: L[i][k] -= L[i][j] * L[k][j], j < k <= i
    for (int e = tid; e < nb * nb; e += kThreads) {
      const int i = e % nb, k = e / nb;
      if (k > j && i >= k) L[i][k] -= L[i][j] * L[k][j];
    }
    __syncthreads();
  }
  double* D = Dbuf + s.doff + (long long)(kb / NB) * NB * NB;  // column-major NB x NB
  for (int e = tid; e < NB * NB; e += kThreads) {
    const int i = e % NB, j = e / NB;
    if (i < nb && j < nb && i >= j) Z[i + (long long)j * s.h] = L[i][j];
  }
  // inverse by forward substitution, one column per thread: L x = e_c
  if (tid < NB) {
    const int c = tid;
    double x[NB];
#pragma unroll 1
    for (int i = 0; i < NB; ++i) x[i] = 0.0;
    if (c < nb) {
      x[c] = 1.0 / L[c][c];
      for (int i = c + 1; i < nb; ++i) {
        double sum = 0.0;
        for (int k = c; k < i; ++k) sum += L[i][k] * x[k];
        x[i] = -sum / L[i][i];
      }
    }
    for (int i = 0; i < NB; ++i) D[i + c * NB] = x[i];
  }
}

// Row-tile multiply with the inverse of a diagonal tile, in place:
//   trans = 1 : X <- X * Dinv^T  on Z[kb+nb:h, kb:kb+nb]   (panel trsm,  X L^T = B)
//   trans = 0 : X <- X * Dinv    on M[kb:h,    kb:kb+nb]   (M recursion, X L   = B)
__global__ void __launch_bounds__(kThreads) tilemul_kernel(const SnDev* __restrict__ sns, int kb, int trans,
                                                           double* Zbuf, double* Mbuf, const double* __restrict__ Dbuf) {
  const SnDev s = sns[blockIdx.x];
  if (kb >= s.w) return;
  const int nb = min(NB, s.w - kb);
  const int rbeg = trans ? kb + nb : kb;
  const int nrows = s.h - rbeg;
  if (nrows <= 0) return;
  __shared__ double D[NB * NB];
  const int tid = threadIdx.x;
  const double* Dg = Dbuf + s.doff + (long long)(kb / NB) * NB * NB;
  for (int e = tid; e < NB * NB; e += kThreads) D[e] = Dg[e];
  __syncthreads();
  double* X = (trans ? Zbuf + s.zoff : Mbuf + s.moff) + rbeg + (long long)kb * s.h;
  const int r = tid % 64, cg = tid / 64;  // 64 rows x 4 column groups of 16
  const int tiles = (nrows + 63) / 64;
  for (int bt = blockIdx.y; bt < tiles; bt += gridDim.y) {
    const int gi = bt * 64 + r;
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0;
    if (gi < nrows) {
      for (int k = 0; k < nb; ++k) {
        const double b = X[gi + (long long)k * s.h];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int jj = cg * 16 + j;
          // trans: X(i,jj) = sum_k B(i,k) Dinv(jj,k) ; else sum_k B(i,k) Dinv(k,jj)
          const double d = trans ? D[jj + k * NB] : D[k + jj * NB];
          acc[j] = fma(b, d, acc[j]);
        }
      }
    }
    __syncthreads();  // every read of this row tile is done before anyone overwrites it
    if (gi < nrows) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int jj = cg * 16 + j;
        if (jj < nb) X[gi + (long long)jj * s.h] = acc[j];
      }
    }
    __syncthreads();
  }
}

// M <- [I ; L21]
__global__ void minit_kernel(const SnDev* __restrict__ sns, const double* __restrict__ Zbuf, double* Mbuf) {
  const SnDev s = sns[blockIdx.x];
  double* M = Mbuf + s.moff;
  const double* Z = Zbuf + s.zoff;
  const long long tot = (long long)s.h * s.w;
  for (long long e = blockIdx.y * (long long)blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.y * blockDim.x) {
    const int i = (int)(e % s.h), j = (int)(e / s.h);
    M[e] = (i < s.w) ? (i == j ? 1.0 : 0.0) : Z[e];
  }
}

// Panel layout (bj.h): k-blocks of 4 steps; inside a k-block the 128 values are stored in DMMA A-fragment
// order, element e = lane*4 + rg holding M(row0 + 8*rg + lane/4, 4*kb + lane%4), so that a lane reads its four
// row-group values of one k-block as 32 contiguous bytes and a warp reads 1 KB contiguous per k-block.
// ONE copy serves both sweeps (bj.h): the rows below the diagonal block are stored negated, which is what the backward
// sweep needs (x_s = L_ss^{-T} (y_s - L_bs^T x_anc)); the forward sweep then produces -u_s and the assembly adds it.
__global__ void pack_kernel(const PackTask* __restrict__ tasks, const double* __restrict__ Mbuf, double* __restrict__ out) {
  const PackTask t = tasks[blockIdx.x];
  const double* M = Mbuf + t.moff;
  double* dst = out + t.dst;
  const long long tot = (long long)t.klen * 32;
  for (long long e = threadIdx.x; e < tot; e += blockDim.x) {
    const int kb = (int)(e / 128), rem = (int)(e % 128), lane = rem / 4, rg = rem % 4;
    const int i = t.row0 + 8 * rg + lane / 4, k = 4 * kb + lane % 4;
    double v = 0.0;
    if (i < t.h && k < t.w && (i >= t.w || k <= i)) { v = M[i + (long long)k * t.h]; if (i >= t.w) v = -v; }
    dst[e] = v;
  }
}

// The optional transposed copy (bj.h): slice q of M_s^T as one contiguous run of k-blocks, steps = supernode rows 32q..
__global__ void pack_bwd_kernel(const PackTask* __restrict__ tasks, const double* __restrict__ Mbuf, double* __restrict__ out) {
  const PackTask t = tasks[blockIdx.x];
  const double* M = Mbuf + t.moff;
  double* dst = out + t.dst;
  const long long tot = (long long)t.klen * 32;
  // output row (a column c of the supernode) = row0 + 8*rg + lane/4 ; step k <-> supernode row i = row0 + 4*kb + lane%4
  for (long long e = threadIdx.x; e < tot; e += blockDim.x) {
    const int kb = (int)(e / 128), rem = (int)(e % 128), lane = rem / 4, rg = rem % 4;
    const int c = t.row0 + 8 * rg + lane / 4, i = t.row0 + 4 * kb + lane % 4;
    double v = 0.0;
    if (i < t.h && c < t.w && i >= c) { v = M[i + (long long)c * t.h]; if (i >= t.w) v = -v; }
    dst[e] = v;
  }
}

template <typename T>
int upload(T** d, const std::vector<T>& h) {
  *d = nullptr;
  PCU_CUDA(cudaMalloc(d, sizeof(T) * std::max<size_t>(h.size(), 1)));
  if (!h.empty()) PCU_CUDA(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return 0;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

int pcu_bj_destroy(pcu_bj* bj) {
  if (!bj) return 0;
  cudaSetDevice(bj->ctx->device);
  cudaStreamSynchronize(bj->ctx->stream);
  cudaFree(bj->fwd_data); cudaFree(bj->bwd_data); cudaFree(bj->fwd_panels); cudaFree(bj->bwd_panels);
  cudaFree(bj->fwd_units); cudaFree(bj->bwd_units); cudaFree(bj->perm); cudaFree(bj->rows);
  cudaFree(bj->lvl_cols); cudaFree(bj->gl_ptr); cudaFree(bj->gl_idx);
  cudaFree(bj->Wk); cudaFree(bj->Y); cudaFree(bj->U); cudaFree(bj->Xp); cudaFree(bj->scratch); cudaFree(bj->counters);
  delete bj;
  return 0;
}

int pcu_bj_create(pcu_ctx* ctx, int nblk, const int* blk_ptr, const int* const* rowPtr, const int* const* colInd,
                  const double* const* val, pcu_bj** out) {
  PCU_CHECK(ctx && nblk >= 1 && blk_ptr && rowPtr && colInd && val && out, "pcu_bj_create: bad arguments");
  PCU_CUDA(cudaSetDevice(ctx->device));
  const double t_an0 = now_s();
  // ---------------------------------------------------------------- symbolic, one thread per block
  std::vector<Symbolic> sym(nblk);
  std::vector<int> rc(nblk, 0);
  SymbolicOptions opt;
  if (const char* e = getenv("PREALPS_BJ_LEAF")) opt.leaf_cols = atoi(e);
  if (const char* e = getenv("PREALPS_BJ_RELAX")) opt.relax_zero = atof(e);
  if (const char* e = getenv("PREALPS_BJ_RELAX_BIG")) opt.relax_big = atof(e);
  if (const char* e = getenv("PREALPS_BJ_RELAX_BIG_COLS")) opt.relax_big_cols = atoi(e);
  {
    unsigned hw = std::thread::hardware_concurrency();
    int nthr = (int)std::min<unsigned>(hw ? hw : 1, (unsigned)nblk);
    if (const char* e = getenv("PREALPS_BJ_THREADS")) nthr = std::max(1, atoi(e));
    std::vector<std::thread> pool;
    std::vector<int> next(1, 0);
    auto work = [&](int tid) {
      for (int b = tid; b < nblk; b += nthr)
        rc[b] = analyze(blk_ptr[b + 1] - blk_ptr[b], rowPtr[b], colInd[b], opt, &sym[b]);
    };
    if (nthr <= 1) work(0);
    else { for (int i = 0; i < nthr; ++i) pool.emplace_back(work, i); for (auto& th : pool) th.join(); }
  }
  for (int b = 0; b < nblk; ++b) PCU_CHECK(rc[b] == 0, "pcu_bj_create: symbolic analysis of block %d failed (%d)", b, rc[b]);
  if (getenv("PREALPS_BJ_PROFILE")) {
    double tmax = 0.0, tsum = 0.0;
    for (auto& S : sym) { tmax = std::max(tmax, S.ordering_seconds); tsum += S.ordering_seconds; }
    fprintf(stderr, "  analysis: METIS_NodeND max %.2f s, sum %.2f s over %d blocks\n", tmax, tsum, nblk);
  }

  // ---------------------------------------------------------------- forest
  pcu_bj* bj = new pcu_bj();
  bj->ctx = ctx;
  bj->nblk = nblk;
  const int n = blk_ptr[nblk] - blk_ptr[0];
  bj->n = n;
  int ns = 0;
  for (auto& S : sym) { ns += S.nsuper; bj->nlevels = std::max(bj->nlevels, S.nlevels); }
  bj->nsuper = ns;
  const int nlev = bj->nlevels;
  std::vector<int> sn_c0(ns), sn_w(ns), sn_h(ns), sn_lev(ns), sn_par(ns), sn_blk(ns);
  std::vector<long long> sn_rp(ns + 1, 0);
  std::vector<int> perm(n);
  {
    int s0 = 0;
    for (int b = 0; b < nblk; ++b) {
      const Symbolic& S = sym[b];
      const int ro = blk_ptr[b] - blk_ptr[0];
      for (int k = 0; k < S.n; ++k) perm[ro + k] = ro + S.perm[k];
      for (int s = 0; s < S.nsuper; ++s) {
        const int g = s0 + s;
        sn_c0[g] = ro + S.sn_col[s];
        sn_w[g] = S.sn_col[s + 1] - S.sn_col[s];
        sn_h[g] = (int)(S.sn_rowptr[s + 1] - S.sn_rowptr[s]);
        sn_lev[g] = S.sn_level[s];
        sn_par[g] = S.sn_parent[s] < 0 ? -1 : s0 + S.sn_parent[s];
        sn_blk[g] = b;
        sn_rp[g + 1] = sn_rp[g] + sn_h[g];
      }
      s0 += S.nsuper;
      bj->stat[0] += (double)S.nnzL_exact;
      bj->stat[1] += (double)S.nnzL_stored;
      bj->stat[5] += S.flops;
    }
  }
  std::vector<int> rows(sn_rp[ns]);  // forest row indices of every supernode
  {
    int s0 = 0;
    for (int b = 0; b < nblk; ++b) {
      const Symbolic& S = sym[b];
      const int ro = blk_ptr[b] - blk_ptr[0];
      for (int s = 0; s < S.nsuper; ++s)
        for (long long p = S.sn_rowptr[s]; p < S.sn_rowptr[s + 1]; ++p)
          rows[sn_rp[s0 + s] + (p - S.sn_rowptr[s])] = ro + S.sn_rows[p];
      s0 += S.nsuper;
    }
  }
  bj->stat[2] = ns;
  bj->stat[3] = nlev;
  // levels: supernodes sorted by (level, w descending)
  std::vector<int> order(ns);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (sn_lev[a] != sn_lev[b]) return sn_lev[a] < sn_lev[b];
    return sn_w[a] > sn_w[b];
  });
  std::vector<int> lev_ptr(nlev + 1, 0);
  for (int s = 0; s < ns; ++s) lev_ptr[sn_lev[s] + 1]++;
  for (int l = 0; l < nlev; ++l) lev_ptr[l + 1] += lev_ptr[l];
  // update-buffer slots, global S offsets, level-local Z/M/D offsets
  std::vector<long long> uoff(ns), soff(ns), zoff(ns), doff(ns);
  long long nu = 0, stot = 0, zmax = 0, dmax = 0;
  for (int s = 0; s < ns; ++s) { uoff[s] = nu; nu += sn_h[s] - sn_w[s]; }
  bj->nu = nu;
  // Update matrices ((h-w)^2 doubles each) live from their supernode's level to their PARENT's level only: offsets come
  // from a first-fit allocator replayed level by level (allocate the level's own, free the children it consumed).  All
  // of them at once would be 4x the factor itself (64^3 block: 4.2 GB against 1.2 GB live at the worst level; 128^3:
  // ~76 GB against ~19 GB), which is what limited the blocks a GPU can factor, not the factor.
  std::vector<std::vector<std::pair<long long, long long>>> s_zero(nlev);   // ranges born at a level: zeroed before use
  {
    std::map<long long, long long> freemap;   // offset -> length, coalesced
    freemap[0] = (long long)1 << 60;
    std::vector<std::vector<int>> kids_of(ns);
    for (int s = 0; s < ns; ++s) if (sn_par[s] >= 0) kids_of[sn_par[s]].push_back(s);
    auto release = [&](long long off, long long len) {
      auto it = freemap.emplace(off, len).first;
      auto nx = std::next(it);
      if (nx != freemap.end() && it->first + it->second == nx->first) { it->second += nx->second; freemap.erase(nx); }
      if (it != freemap.begin()) {
        auto pv = std::prev(it);
        if (pv->first + pv->second == it->first) { pv->second += it->second; freemap.erase(it); }
      }
    };
    for (int l = 0; l < nlev; ++l) {
      for (int q = lev_ptr[l]; q < lev_ptr[l + 1]; ++q) {
        const int s = order[q];
        const long long need = (long long)(sn_h[s] - sn_w[s]) * (sn_h[s] - sn_w[s]);
        soff[s] = 0;
        if (need == 0) continue;
        auto it = freemap.begin();
        while (it->second < need) ++it;   // the last hole is unbounded
        soff[s] = it->first;
        const long long rest = it->second - need, at = it->first + need;
        freemap.erase(it);
        if (rest > 0) freemap[at] = rest;
        stot = std::max(stot, soff[s] + need);
        if (!s_zero[l].empty() && s_zero[l].back().first + s_zero[l].back().second == soff[s]) s_zero[l].back().second += need;
        else s_zero[l].push_back({soff[s], need});
      }
      for (int q = lev_ptr[l]; q < lev_ptr[l + 1]; ++q)
        for (int c : kids_of[order[q]]) {
          const long long len = (long long)(sn_h[c] - sn_w[c]) * (sn_h[c] - sn_w[c]);
          if (len > 0) release(soff[c], len);
        }
    }
    for (int l = 0; l < nlev; ++l) {   // fewer, larger memsets
      auto& z = s_zero[l];
      std::sort(z.begin(), z.end());
      size_t o = 0;
      for (size_t i = 1; i < z.size(); ++i) {
        if (z[o].first + z[o].second == z[i].first) z[o].second += z[i].second;
        else z[++o] = z[i];
      }
      if (!z.empty()) z.resize(o + 1);
    }
  }
  bj->stat[10] = 8.0 * (double)stot;
  for (int l = 0; l < nlev; ++l) {
    long long z = 0, d = 0;
    for (int q = lev_ptr[l]; q < lev_ptr[l + 1]; ++q) {
      const int s = order[q];
      zoff[s] = z; z += (long long)sn_h[s] * sn_w[s];
      doff[s] = d; d += (long long)((sn_w[s] + NB - 1) / NB) * NB * NB;
    }
    zmax = std::max(zmax, z);
    dmax = std::max(dmax, d);
  }
  // children lists and relative indices (position of each child update row in the parent's rows)
  std::vector<int> nchild(ns, 0), child_ord(ns, 0);
  for (int s = 0; s < ns; ++s) if (sn_par[s] >= 0) child_ord[s] = nchild[sn_par[s]]++;
  std::vector<int> rel(nu > 0 ? nu : 1);
  {
    std::vector<int> pos(n, -1);
    // process parents one at a time: mark positions, then fill rel for each child
    std::vector<std::vector<int>> kids(ns);
    for (int s = 0; s < ns; ++s) if (sn_par[s] >= 0) kids[sn_par[s]].push_back(s);
    for (int p = 0; p < ns; ++p) {
      if (kids[p].empty()) continue;
      for (int i = 0; i < sn_h[p]; ++i) pos[rows[sn_rp[p] + i]] = i;
      for (int c : kids[p])
        for (int i = sn_w[c]; i < sn_h[c]; ++i) {
          const int r = pos[rows[sn_rp[c] + i]];
          PCU_CHECK(r >= 0, "pcu_bj_create: internal error, child row missing from parent structure");
          rel[uoff[c] + (i - sn_w[c])] = r;
        }
      for (int i = 0; i < sn_h[p]; ++i) pos[rows[sn_rp[p] + i]] = -1;
    }
  }
  // A entries per level: (permuted row >= permuted col) -> front position
  std::vector<int> col2sn(n);
  for (int s = 0; s < ns; ++s) for (int c = 0; c < sn_w[s]; ++c) col2sn[sn_c0[s] + c] = s;
  std::vector<long long> aent_ptr(nlev + 1, 0);
  std::vector<AEntry> aent;
  {
    long long tot = 0;
    for (int b = 0; b < nblk; ++b) tot += rowPtr[b][blk_ptr[b + 1] - blk_ptr[b]];
    std::vector<int> elev(tot);
    std::vector<AEntry> tmp(tot);
    long long q = 0;
    for (int b = 0; b < nblk; ++b) {
      const Symbolic& S = sym[b];
      const int ro = blk_ptr[b] - blk_ptr[0];
      const int nb_ = blk_ptr[b + 1] - blk_ptr[b];
      for (int i = 0; i < nb_; ++i)
        for (int p = rowPtr[b][i]; p < rowPtr[b][i + 1]; ++p) {
          const int j = colInd[b][p];
          const int a = ro + S.iperm[i], c = ro + S.iperm[j];
          const int col = std::min(a, c), row = std::max(a, c);
          const int s = col2sn[col];
          const int* rs = rows.data() + sn_rp[s];
          const int* it = std::lower_bound(rs, rs + sn_h[s], row);
          PCU_CHECK(it != rs + sn_h[s] && *it == row, "pcu_bj_create: internal error, entry outside the symbolic structure");
          tmp[q].dst = zoff[s] + (it - rs) + (long long)(col - sn_c0[s]) * sn_h[s];
          tmp[q].val = val[b][p];
          elev[q] = sn_lev[s];
          aent_ptr[sn_lev[s] + 1]++;
          ++q;
        }
    }
    for (int l = 0; l < nlev; ++l) aent_ptr[l + 1] += aent_ptr[l];
    aent.resize(tot);
    std::vector<long long> fill(aent_ptr.begin(), aent_ptr.end() - 1);
    for (long long i = 0; i < tot; ++i) aent[fill[elev[i]]++] = tmp[i];
  }
  // ---------------------------------------------------------------- solve-side structures
  // One copy of the panels or two (bj.h)?  The backward sweep is ~10 % faster on the transposed copy (B200, 128^3: apply
  // 3.48 against 3.60 ms); it is kept when both copies take less than a third of the free device memory, else the
  // backward sweep reads M tile by tile (half the factor memory: 256^3 then fits 2 GPUs).  PREALPS_BJ_COPIES=1|2 decides.
  bool tcopy = true;
  {
    double est = 0.0;
    for (int s = 0; s < ns; ++s) {
      est += (double)panel_cum(sn_w[s], (sn_h[s] + 31) / 32);
      for (int p = 0; p * 32 < sn_w[s]; ++p) est += 32.0 * ((sn_h[s] - 32 * p + 3) & ~3);
    }
    size_t mfree = 0, mtot = 0;
    PCU_CUDA(cudaMemGetInfo(&mfree, &mtot));
    tcopy = 8.0 * est <= (double)mfree / 3.0;
    if (const char* e = getenv("PREALPS_BJ_COPIES")) tcopy = atoi(e) >= 2;
  }
  bj->stat[9] = tcopy ? 2.0 : 1.0;
  // panels (sorted by level, long panels first inside a level), work units, gather lists
  std::vector<FwdPanel> fp;
  std::vector<BwdPanel> bp;
  std::vector<int> fp_sn, bp_sn;  // supernode of every panel
  std::vector<WorkUnit> fu, bu;
  bj->fwd_unit_ptr.assign(nlev + 1, 0);
  bj->fwd_lvl_bytes.assign(nlev, 0.0);
  bj->bwd_lvl_bytes.assign(nlev, 0.0);
  bj->bwd_unit_ptr.assign(nlev + 1, 0);
  std::vector<std::vector<PackTask>> pk_f(nlev), pk_b(nlev);
  long long fdoubles = 0, bdoubles = 0;
  std::vector<long long> sn_doff(ns, 0);  // the panels of a supernode are contiguous, slice after slice (bj.h: panel_cum)
  const int kSplitK = 512;  // without level-adaptive cuts (PREALPS_BJ_NOCHUNK): panels at least this long get a whole CTA
  const bool use_chunks = getenv("PREALPS_BJ_NOCHUNK") == nullptr;
  // klen_of: steps of a panel rounded to whole k-blocks (the backward kernel walks whole 32-row tiles: its slices are
  // cut at multiples of 8 k-blocks)
  // Every panel of a level goes to the one sweep launch of that level and direction.  (Until round 2 the panels of <= 32
  // steps -- the leaves of the forest -- had a kernel of their own that staged a whole panel in shared memory; once the sweep
  // kernel streamed its panel data through a shared-memory ring and across the panels of a warp, folding them into the regular
  // launch was faster: 0.670 -> 0.641 ms for one 64^3 block, 3.190 -> 3.154 ms for eight; profiles/r02_smem_ring.md.)
  auto make_units = [&](std::vector<int>& klen_of, bool tiles, int first, int count, std::vector<WorkUnit>& units) {
    auto cost_kb = [&](int i) { return klen_of[i] / 4; };
    const size_t u_begin = units.size();
    // panels [first, first+count) are already sorted by cost descending
    int i = 0;
    int slots = 0, ctrs = 0;
    // Work per level is spread over the 148 x 2 resident CTAs x 8 warps: q = k-blocks per warp when every warp slot
    // is busy. A panel longer than 4q gets a whole CTA (its 8 warps on disjoint k ranges), one longer than ~12q is
    // cut across several CTAs. With many subdomains per GPU q is large and nothing is cut; with one subdomain per
    // GPU (strong scaling) the few long panels of a level are spread over the machine instead of being streamed
    // by one warp or one CTA each.
    long long level_kb = 0;
    for (int j = 0; j < count; ++j) level_kb += cost_kb(first + j);
    int split_kb = kSplitK / 4, chunk_kb = 1 << 30;
    if (use_chunks) {
      const int q = (int)std::max<long long>(16, (level_kb + kWarpSlots - 1) / kWarpSlots);
      // a panel gets a whole CTA from 3/4 of a warp's share of the level on (32 .. 256 k-blocks).  Measured on B200 (apply, ms,
      // one 64^3 block / eight): fixed 128 k-blocks 0.785 / 3.510; 64: 0.744 / 3.598; 256: 0.849 / 3.484; this rule 0.747 / 3.489
      static const int split_a = getenv("PREALPS_BJ_SPLITA") ? atoi(getenv("PREALPS_BJ_SPLITA")) : 6;
      split_kb = std::min(256, std::max(32, (q * split_a / 8 + 15) & ~15));
      // a slice = 10 q: a level's long panels then make ONE wave of CTAs (2 x 148 slots).  Measured, apply of one 64^3
      // block: 0.958 / 0.917 / 0.883 / 0.810 / 0.820 / 0.849 ms with 2 / 4 / 8 / 10 / 12 / 16 q -- smaller slices mean more
      // partial sums and a second, partly filled wave that costs as much as a full one; 8 blocks per GPU: no difference
      const int per_unit = getenv("PREALPS_BJ_CHUNKQ") ? atoi(getenv("PREALPS_BJ_CHUNKQ")) : 10;
      chunk_kb = std::max(kChunkMinKB, (per_unit * q + 7) & ~7);
    }
    int nlong = 0;
    while (nlong < count && cost_kb(first + nlong) >= split_kb) ++nlong;
    while (i < nlong) {
      const int nkb = tiles ? ((klen_of[first + i] / 4 + 7) & ~7) : klen_of[first + i] / 4;
      const int nch = (2 * nkb >= 3 * chunk_kb) ? (nkb + chunk_kb - 1) / chunk_kb : 1;
      if (nch <= 1) units.push_back({first + i, 1, 1, 0, nkb, 0, 1, 0, 0, 0});
      else {
        const int per = (((nkb + nch - 1) / nch) + 7) & ~7;  // even slices, whole 8-k-block groups (one per warp)
        for (int ch = 0; ch < nch; ++ch)
          units.push_back({first + i, 1, 2, std::min(nkb, ch * per), std::min(nkb, (ch + 1) * per), ch, nch, slots, ctrs, 0});
        slots += nch;
        ++ctrs;
      }
      ++i;
    }
    bj->scratch_slots = std::max(bj->scratch_slots, slots);
    bj->ncounters = std::max(bj->ncounters, ctrs);
    // the other panels: one warp per panel, and when a level has many more panels than the machine has warp slots a warp
    // takes several in a row (first + warp, + 8, + 16, ...: the sweep kernel streams across them, bj_solve.cu) -- as many as
    // leave every CTA slot several units.  Not for the tile-wise backward sweep of the single-copy factor.
    const int pw_max = getenv("PREALPS_BJ_PW") ? atoi(getenv("PREALPS_BJ_PW")) : 4;
    const int pw_force = getenv("PREALPS_BJ_PW_FORCE") ? atoi(getenv("PREALPS_BJ_PW_FORCE")) : 0;   // tests
    const int pw = tiles ? 1 : pw_force > 0 ? pw_force : std::max(1, std::min(pw_max, (count - i) / (8 * 148 * 2 * 2)));
    while (i < count) { const int c = std::min(8 * pw, count - i); units.push_back({first + i, c, 0, 0, 0, 0, 1, 0, 0, 0}); i += c; }
    // CTAs are dispatched in unit order: longest first by TIME, not by panel length.  A CTA whose 8 warps each stream a
    // whole panel of just under split_kb k-blocks runs as long as a CTA that shares a panel 8 times as long; left at
    // the end of the list (they were, the list being sorted by panel length) those groups were the tail of every level.
    if (getenv("PREALPS_BJ_NOLPT") == nullptr) {
      auto cost = [&](const WorkUnit& u) {
        if (u.split == 2) return (u.kb1 - u.kb0 + 7) / 8;
        if (u.split == 1) return (cost_kb(u.first) + 7) / 8;
        return cost_kb(u.first) * ((u.count + 7) / 8);  // the first panel of a group is its longest
      };
      std::stable_sort(units.begin() + u_begin, units.end(), [&](const WorkUnit& a, const WorkUnit& b) { return cost(a) > cost(b); });
    }
  };
  for (int l = 0; l < nlev; ++l) {
    // forward
    std::vector<std::pair<int, std::pair<int, int>>> lst;  // (klen, (sn, p))
    for (int q = lev_ptr[l]; q < lev_ptr[l + 1]; ++q) {
      const int s = order[q];
      sn_doff[s] = fdoubles;
      fdoubles += panel_cum(sn_w[s], (sn_h[s] + 31) / 32);
      for (int p = 0; p * 32 < sn_h[s]; ++p) {
        int klen = std::min(sn_w[s], 32 * p + 32);
        klen = (klen + 3) & ~3;  // whole k-blocks of 4 (one DMMA step)
        lst.push_back({klen, {s, p}});
      }
    }
    std::stable_sort(lst.begin(), lst.end(), [](auto& a, auto& b) { return a.first > b.first; });
    const int f0 = (int)fp.size();
    std::vector<int> kl;
    for (auto& e : lst) {
      const int s = e.second.first, p = e.second.second;
      const long long poff = sn_doff[s] + panel_cum(sn_w[s], p);
      FwdPanel P{poff, uoff[s], e.first, sn_c0[s], 32 * p, sn_w[s], sn_h[s], 0};
      pk_f[l].push_back({zoff[s], poff, sn_h[s], sn_w[s], 32 * p, e.first});
      bj->fwd_lvl_bytes[l] += 8.0 * e.first * 32;
      fp.push_back(P);
      fp_sn.push_back(s);
    }
    kl.resize(fp.size());
    for (size_t i = f0; i < fp.size(); ++i) kl[i] = fp[i].klen;
    make_units(kl, false, f0, (int)fp.size() - f0, fu);
    bj->fwd_unit_ptr[l + 1] = (int)fu.size();
    // backward: slice q of M_s^T = the 32 columns [32q, 32q + 32) of the SAME panels, walked tile by tile (32 rows of
    // forward slice p = q, q+1, ...); klen counts rows, whole tiles
    lst.clear();
    for (int q = lev_ptr[l]; q < lev_ptr[l + 1]; ++q) {
      const int s = order[q];
      for (int p = 0; p * 32 < sn_w[s]; ++p) lst.push_back({(sn_h[s] - 32 * p + 3) & ~3, {s, p}});
    }
    std::stable_sort(lst.begin(), lst.end(), [](auto& a, auto& b) { return a.first > b.first; });
    const int b0 = (int)bp.size();
    std::vector<double> bbytes;
    for (auto& e : lst) {
      const int s = e.second.first, p = e.second.second;
      if (tcopy) {
        BwdPanel P{bdoubles, sn_rp[s], e.first, 32 * p, sn_c0[s], sn_w[s], sn_h[s], 0};
        pk_b[l].push_back({zoff[s], bdoubles, sn_h[s], sn_w[s], 32 * p, e.first});
        bdoubles += (long long)e.first * 32;
        bbytes.push_back(8.0 * e.first * 32);
        bp.push_back(P);
      } else {
        BwdPanel P{sn_doff[s], sn_rp[s], (e.first + 31) & ~31, 32 * p, sn_c0[s], sn_w[s], sn_h[s], 0};
        const int nkq = std::min(8, (sn_w[s] + 3) / 4 - 8 * p);
        bbytes.push_back(8.0 * 128 * nkq * (P.klen / 32));
        bp.push_back(P);
      }
      bj->bwd_lvl_bytes[l] += bbytes.back();
      bp_sn.push_back(s);
    }
    kl.assign(bp.size(), 0);
    {
      size_t i = b0;
      for (auto& e : lst) kl[i++] = e.first;   // the exact length decides who shares a CTA and what is cut
    }
    make_units(kl, !tcopy, b0, (int)bp.size() - b0, bu);
    bj->bwd_unit_ptr[l + 1] = (int)bu.size();
  }
  bj->fwd_doubles = fdoubles;
  bj->bwd_doubles = bdoubles;
  // columns by level + gather lists
  std::vector<int> lvl_cols;
  lvl_cols.reserve(n);
  bj->lvl_col_ptr.assign(nlev + 1, 0);
  for (int l = 0; l < nlev; ++l) {
    for (int q = lev_ptr[l]; q < lev_ptr[l + 1]; ++q) {
      const int s = order[q];
      for (int c = 0; c < sn_w[s]; ++c) lvl_cols.push_back(sn_c0[s] + c);
    }
    bj->lvl_col_ptr[l + 1] = (int)lvl_cols.size();
  }
  std::vector<long long> gl_ptr(n + 1, 0), gl_idx(nu > 0 ? nu : 1);
  for (int s = 0; s < ns; ++s)
    for (int i = sn_w[s]; i < sn_h[s]; ++i) gl_ptr[rows[sn_rp[s] + i] + 1]++;
  for (int c = 0; c < n; ++c) gl_ptr[c + 1] += gl_ptr[c];
  {
    bj->lvl_long_lists.assign(nlev, 0);
    for (int l = 0; l < nlev; ++l) {
      long long ent = 0;
      for (int q = bj->lvl_col_ptr[l]; q < bj->lvl_col_ptr[l + 1]; ++q) ent += gl_ptr[lvl_cols[q] + 1] - gl_ptr[lvl_cols[q]];
      bj->lvl_long_lists[l] = ent > 16ll * std::max(1, bj->lvl_col_ptr[l + 1] - bj->lvl_col_ptr[l]);
    }
    std::vector<long long> fill(gl_ptr.begin(), gl_ptr.end() - 1);
    for (int s = 0; s < ns; ++s)  // ascending supernode order => fixed summation order
      for (int i = sn_w[s]; i < sn_h[s]; ++i) gl_idx[fill[rows[sn_rp[s] + i]]++] = uoff[s] + (i - sn_w[s]);
  }
  bj->stat[7] = now_s() - t_an0;

  // ---------------------------------------------------------------- device buffers
  double *Zbuf = nullptr, *Mbuf = nullptr, *Sbuf = nullptr, *Dbuf = nullptr;
  SnDev* d_sn = nullptr;
  AEntry* d_aent = nullptr;
  int* d_rel = nullptr;
  int* d_fail = nullptr;
  std::vector<SnDev> h_sn(ns);
  for (int q = 0; q < ns; ++q) {
    const int s = order[q];
    h_sn[q] = {zoff[s], zoff[s], soff[s], doff[s], sn_h[s], sn_w[s]};
  }
  PCU_CUDA(cudaMalloc(&Zbuf, sizeof(double) * std::max<long long>(zmax, 1)));
  PCU_CUDA(cudaMalloc(&Mbuf, sizeof(double) * std::max<long long>(zmax, 1)));
  PCU_CUDA(cudaMalloc(&Sbuf, sizeof(double) * std::max<long long>(stot, 1)));
  PCU_CUDA(cudaMalloc(&Dbuf, sizeof(double) * std::max<long long>(dmax, 1)));
  PCU_CUDA(cudaMalloc(&d_fail, sizeof(int)));
  PCU_CUDA(cudaMemset(d_fail, 0, sizeof(int)));
  if (upload(&d_sn, h_sn) || upload(&d_aent, aent) || upload(&d_rel, rel)) return 1;
  PCU_CUDA(cudaMalloc(&bj->fwd_data, sizeof(double) * std::max<long long>(fdoubles, 1)));
  if (tcopy) PCU_CUDA(cudaMalloc(&bj->bwd_data, sizeof(double) * std::max<long long>(bdoubles, 1)));
  if (upload(&bj->fwd_panels, fp) || upload(&bj->bwd_panels, bp) || upload(&bj->fwd_units, fu) ||
      upload(&bj->bwd_units, bu) || upload(&bj->perm, perm) || upload(&bj->rows, rows) ||
      upload(&bj->lvl_cols, lvl_cols) || upload(&bj->gl_ptr, gl_ptr) || upload(&bj->gl_idx, gl_idx))
    return 1;

  // ---------------------------------------------------------------- numeric factorisation
  cudaStream_t st = ctx->stream;
  cudaEvent_t e0, e1;
  PCU_CUDA(cudaEventCreate(&e0));
  PCU_CUDA(cudaEventCreate(&e1));
  PCU_CUDA(cudaEventRecord(e0, st));
  std::vector<EaTask> tasks;
  EaTask* d_tasks = nullptr;
  PackTask* d_pack = nullptr;
  size_t tasks_cap = 0, pack_cap = 0;
  std::vector<std::vector<int>> kids(ns);
  for (int s = 0; s < ns; ++s) if (sn_par[s] >= 0) kids[sn_par[s]].push_back(s);
  for (int l = 0; l < nlev; ++l) {
    const int q0 = lev_ptr[l], cnt = lev_ptr[l + 1] - lev_ptr[l];
    if (cnt == 0) continue;
    const SnDev* lsn = d_sn + q0;
    long long zl = 0;
    int maxw = 0, maxh = 0, maxu = 0;
    for (int q = q0; q < q0 + cnt; ++q) {
      const int s = order[q];
      zl += (long long)sn_h[s] * sn_w[s];
      maxw = std::max(maxw, sn_w[s]);
      maxh = std::max(maxh, sn_h[s]);
      maxu = std::max(maxu, sn_h[s] - sn_w[s]);
    }
    PCU_CUDA(cudaMemsetAsync(Zbuf, 0, sizeof(double) * zl, st));
    for (auto& z : s_zero[l]) PCU_CUDA(cudaMemsetAsync(Sbuf + z.first, 0, sizeof(double) * z.second, st));
    // 1. scatter the entries of A that live in this level's supernodes
    const long long na = aent_ptr[l + 1] - aent_ptr[l];
    if (na > 0) {
      scatter_a_kernel<<<stream_grid(ctx, na, 256, 8), 256, 0, st>>>(d_aent + aent_ptr[l], na, Zbuf);
      PCU_LAUNCH_CHECK(ctx);
    }
    // 2. extend-add, one round per child ordinal (children of one parent never share a launch)
    int maxkids = 0;
    for (int q = q0; q < q0 + cnt; ++q) maxkids = std::max(maxkids, (int)kids[order[q]].size());
    for (int r = 0; r < maxkids; ++r) {
      tasks.clear();
      int maxuc = 0;
      for (int q = q0; q < q0 + cnt; ++q) {
        const int p = order[q];
        if ((int)kids[p].size() <= r) continue;
        const int c = kids[p][r];
        const int uc = sn_h[c] - sn_w[c];
        if (uc == 0) continue;
        tasks.push_back({soff[c], uoff[c], zoff[p], soff[p], uc, sn_h[p], sn_w[p], 0});
        maxuc = std::max(maxuc, uc);
      }
      if (tasks.empty()) continue;
      if (tasks.size() > tasks_cap) {
        if (d_tasks) { PCU_CUDA(cudaStreamSynchronize(st)); cudaFree(d_tasks); }
        tasks_cap = tasks.size() * 2;
        PCU_CUDA(cudaMalloc(&d_tasks, sizeof(EaTask) * tasks_cap));
      }
      PCU_CUDA(cudaStreamSynchronize(st));  // the previous round may still read d_tasks
      PCU_CUDA(cudaMemcpyAsync(d_tasks, tasks.data(), sizeof(EaTask) * tasks.size(), cudaMemcpyHostToDevice, st));
      dim3 grid((unsigned)tasks.size(), (unsigned)std::min(maxuc, tasks.size() > 2048 ? 4 : 64));
      extend_add_kernel<<<grid, 128, 0, st>>>(d_tasks, d_rel, Sbuf, Zbuf, Sbuf);
      PCU_LAUNCH_CHECK(ctx);
    }
    // 3. blocked left-looking Cholesky of every panel [F11; F21]
    auto nactive = [&](int kb) {  // supernodes are sorted by w descending inside the level
      int lo = 0, hi = cnt;
      while (lo < hi) { const int mid = (lo + hi) / 2; if (sn_w[order[q0 + mid]] > kb) lo = mid + 1; else hi = mid; }
      return lo;
    };
    auto ytiles = [&](int rows_) { return (unsigned)std::max(1, std::min((rows_ + 63) / 64, 1024)); };
    for (int kb = 0; kb < maxw; kb += NB) {
      const int na_ = nactive(kb);
      if (na_ == 0) break;
      if (kb > 0) {
        gemm_kernel<<<dim3(na_, ytiles(maxh - kb), 1), kThreads, 0, st>>>(lsn, 0, kb, Zbuf, Mbuf, Sbuf);
        PCU_LAUNCH_CHECK(ctx);
      }
      potrf_tile_kernel<<<na_, kThreads, 0, st>>>(lsn, kb, Zbuf, Dbuf, d_fail);
      PCU_LAUNCH_CHECK(ctx);
      if (maxh - kb > 0) {
        tilemul_kernel<<<dim3(na_, ytiles(maxh - kb)), kThreads, 0, st>>>(lsn, kb, 1, Zbuf, Mbuf, Dbuf);
        PCU_LAUNCH_CHECK(ctx);
      }
    }
    // 4. Schur complement (the update matrix handed to the parent)
    if (maxu > 0) {
      const unsigned tu = ytiles(maxu);
      gemm_kernel<<<dim3(cnt, tu, std::min(tu, 64u)), kThreads, 0, st>>>(lsn, 1, 0, Zbuf, Mbuf, Sbuf);
      PCU_LAUNCH_CHECK(ctx);
    }
    // 5. M = [I; L21] L11^{-1}, right-to-left over the column blocks
    minit_kernel<<<dim3(cnt, (unsigned)std::max(1, std::min(256, (int)(((long long)maxh * maxw + 255) / 256)))), 256, 0, st>>>(lsn, Zbuf, Mbuf);
    PCU_LAUNCH_CHECK(ctx);
    for (int jb = ((maxw - 1) / NB) * NB; jb >= 0; jb -= NB) {
      const int na_ = nactive(jb);
      if (na_ == 0) continue;
      gemm_kernel<<<dim3(na_, ytiles(maxh - jb), 1), kThreads, 0, st>>>(lsn, 2, jb, Zbuf, Mbuf, Sbuf);
      PCU_LAUNCH_CHECK(ctx);
      tilemul_kernel<<<dim3(na_, ytiles(maxh - jb)), kThreads, 0, st>>>(lsn, jb, 0, Zbuf, Mbuf, Dbuf);
      PCU_LAUNCH_CHECK(ctx);
    }
    // 6. pack M into panels, and M^T when the factor keeps the transposed copy
    for (int dir = 0; dir < (tcopy ? 2 : 1); ++dir) {
      const std::vector<PackTask>& pk = dir == 0 ? pk_f[l] : pk_b[l];
      if (pk.empty()) continue;
      if (pk.size() > pack_cap) {
        if (d_pack) { PCU_CUDA(cudaStreamSynchronize(st)); cudaFree(d_pack); }
        pack_cap = pk.size() * 2;
        PCU_CUDA(cudaMalloc(&d_pack, sizeof(PackTask) * pack_cap));
      }
      PCU_CUDA(cudaStreamSynchronize(st));
      PCU_CUDA(cudaMemcpyAsync(d_pack, pk.data(), sizeof(PackTask) * pk.size(), cudaMemcpyHostToDevice, st));
      if (dir == 0) pack_kernel<<<(unsigned)pk.size(), 256, 0, st>>>(d_pack, Mbuf, bj->fwd_data);
      else pack_bwd_kernel<<<(unsigned)pk.size(), 256, 0, st>>>(d_pack, Mbuf, bj->bwd_data);
      PCU_LAUNCH_CHECK(ctx);
    }
  }
  PCU_CUDA(cudaEventRecord(e1, st));
  PCU_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  PCU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  bj->stat[6] = ms * 1e-3;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  int fail = 0;
  PCU_CUDA(cudaMemcpy(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(Zbuf); cudaFree(Mbuf); cudaFree(Sbuf); cudaFree(Dbuf); cudaFree(d_sn); cudaFree(d_aent);
  cudaFree(d_rel); cudaFree(d_fail); cudaFree(d_tasks); cudaFree(d_pack);
  if (fail) {
    pcu_bj_destroy(bj);
    set_error("pcu_bj_create: a diagonal block is not positive definite (Cholesky pivot <= 0)");
    return 2;
  }
  bj->stat[8] = 3.0 * nlev;
  *out = bj;
  return 0;
}

int pcu_bj_analyze(int n, const int* rowPtr, const int* colInd, int use_metis, int* perm, int* nsuper, int* sn_col,
                   long long* sn_rowptr, int* sn_rows, long long rows_cap, int* sn_parent, int* sn_level,
                   double* stats4) {
  Symbolic S;
  SymbolicOptions opt;
  opt.use_metis = use_metis != 0;
  if (const char* e = getenv("PREALPS_BJ_LEAF")) opt.leaf_cols = atoi(e);
  if (const char* e = getenv("PREALPS_BJ_RELAX")) opt.relax_zero = atof(e);
  if (const char* e = getenv("PREALPS_BJ_RELAX_BIG")) opt.relax_big = atof(e);
  if (const char* e = getenv("PREALPS_BJ_RELAX_BIG_COLS")) opt.relax_big_cols = atoi(e);
  const int rc = analyze(n, rowPtr, colInd, opt, &S);
  if (rc) return rc;
  if ((long long)S.sn_rows.size() > rows_cap) return -9;
  for (int i = 0; i < n; ++i) perm[i] = S.perm[i];
  *nsuper = S.nsuper;
  for (int s = 0; s <= S.nsuper; ++s) { sn_col[s] = S.sn_col[s]; sn_rowptr[s] = S.sn_rowptr[s]; }
  for (size_t i = 0; i < S.sn_rows.size(); ++i) sn_rows[i] = S.sn_rows[i];
  for (int s = 0; s < S.nsuper; ++s) { sn_parent[s] = S.sn_parent[s]; sn_level[s] = S.sn_level[s]; }
  stats4[0] = (double)S.nnzL_exact; stats4[1] = (double)S.nnzL_stored; stats4[2] = S.nlevels; stats4[3] = S.flops;
  return 0;
}

double pcu_bj_stat(pcu_bj* bj, int which) {
  if (!bj || which < 0 || which >= 16) return -1.0;
  if (which == 4) return pcu_bj_stored_bytes(bj, 8);
  return bj->stat[which];
}

double pcu_bj_bytes(pcu_bj* bj, int t) {
  // SURVEY.md 8(d), dense-supernode form (8 B per non-zero of L, no index per entry): L streamed once forward and once
  // backward -- the EXACT non-zeros of the factor, relaxation zeros and panel padding not counted -- plus one pointer
  // per row and sweep, plus read + write of the m x t block in each sweep
  const double vec = (double)bj->n * t * 8.0;
  return 2.0 * (8.0 * bj->stat[0] + 8.0 * ((double)bj->n + bj->nblk)) + 4.0 * vec;
}

double pcu_bj_stored_bytes(pcu_bj* bj, int t) {
  // what the kernels have to move: the stored panels once per sweep (explicit zeros of the relaxed supernodes and the
  // padding of the 32-row panels included; with one copy in memory both sweeps read the same panels), the block vectors (read B, write/read Wk and Y once each, write X) and
  // the update rows (written by the forward sweep, gathered by the assembly)
  const double vec = (double)bj->n * t * 8.0;
  return 8.0 * ((double)bj->fwd_doubles + (bj->bwd_data ? (double)bj->bwd_doubles : (double)bj->fwd_doubles)) + 6.0 * vec +
         2.0 * (double)bj->nu * t * 8.0 + (double)bj->nu * 8.0;
}

}  // extern "C"
