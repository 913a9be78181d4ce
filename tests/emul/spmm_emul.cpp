// spmm_emul.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_shim.h): the SpMM kernels of prealps_b200/csrc/spmm_kernels.cuh
// executed on the CPU with the launch logic of spmm.cu restated per variant, behind a C entry point for ctypes.
#define PCU_EMUL 1
#include "cuda_shim.h"

thread_local emul_dim3 threadIdx;
emul_dim3 blockIdx, blockDim, gridDim;
static pthread_barrier_t g_barrier;
void __syncthreads() { pthread_barrier_wait(&g_barrier); }

namespace {
struct ThreadArg { int tid; void (*thunk)(void*); void* ctx; };
void* thread_main(void* p) {
  ThreadArg* a = static_cast<ThreadArg*>(p);
  threadIdx.x = (unsigned)a->tid;
  a->thunk(a->ctx);
  return nullptr;
}
}  // namespace

void emul_run_block(int block, void (*thunk)(void*), void* ctx) {
  pthread_barrier_init(&g_barrier, nullptr, (unsigned)block);
  std::vector<pthread_t> th(block);
  std::vector<ThreadArg> args(block);
  pthread_attr_t at;
  pthread_attr_init(&at);
  pthread_attr_setstacksize(&at, 256 * 1024);
  for (int i = 0; i < block; ++i) {
    args[i] = ThreadArg{i, thunk, ctx};
    if (pthread_create(&th[i], &at, thread_main, &args[i]) != 0) { std::perror("pthread_create"); std::abort(); }
  }
  for (int i = 0; i < block; ++i) pthread_join(th[i], nullptr);
  pthread_attr_destroy(&at);
  pthread_barrier_destroy(&g_barrier);
}

#include "../../prealps_b200/csrc/spmm_kernels.cuh"

namespace {

// variant codes of the tests: 0 LDG + STS staging (spmm_kernel), 8 / 9 bulk-copy staging (spmm_bulk_kernel; 9: HALO = false,
// for operators without a column >= m)
template <int T, int CPL>
void run_variant(int lean, const SpmmArgs& a, int nblk) {
  if (lean == 8) { emul_launch(nblk, kThreads, [&] { spmm_bulk_kernel<T, CPL, true>(a); }); return; }
  if (lean == 9) { emul_launch(nblk, kThreads, [&] { spmm_bulk_kernel<T, CPL, false>(a); }); return; }
  emul_launch(nblk, kThreads, [&] { spmm_kernel<T, CPL>(a); });
}

// Y = A [X ; H] with the kernel spmm.cu: launch_spmm would pick for (t, cpl, lean); cpl = 0: the generic kernel
int run_spmm(int lean, int cpl, int m, const int* rowPtr, const int* colInd, const double* val, const double* X, int ldx,
             const double* H, double* Y, int ldy, int t) {
  const int sh = (t <= 4) ? 1 : 0;
  std::vector<int4> blk;
  build_row_blocks(m, rowPtr, sh, &blk);
  if (lean > 0) {
    if (t < 8 || ldx != t) return 2;
    for (const int4& b : blk) if (b.w - b.z > kShapeNnz[0]) return 3;  // the host would not choose the bulk kernel
  }
  // like upload_csr: 16 bytes of slack behind the entries (the bulk copies round their size up), 16-byte aligned starts
  const int nnz = rowPtr[m];
  std::vector<double> vbuf((size_t)nnz + 6);
  std::vector<int> cbuf((size_t)nnz + 12);
  double* vpad = vbuf.data() + ((16 - reinterpret_cast<uintptr_t>(vbuf.data()) % 16) % 16) / 8;
  int* cpad = cbuf.data() + ((16 - reinterpret_cast<uintptr_t>(cbuf.data()) % 16) % 16) / 4;
  std::memcpy(vpad, val, sizeof(double) * (size_t)nnz);
  std::memcpy(cpad, colInd, sizeof(int) * (size_t)nnz);
  for (int k = 0; k < 2; ++k) vpad[nnz + k] = std::nan("");
  for (int k = 0; k < 4; ++k) cpad[nnz + k] = -123456789;
  SpmmArgs a{rowPtr, cpad, vpad, blk.data(), m, X, ldx, H, Y, ldy, t};
  const int nblk = (int)blk.size();
  if (cpl == 0) { emul_launch(nblk, kThreads, [&] { spmm_kernel_generic(a); }); return 0; }
  if (t == 1 && cpl == 1 && lean == 0) { emul_launch(nblk, kThreads, [&] { spmm_kernel<1, 1>(a); }); return 0; }
  if (cpl == 2) {
    switch (t) {
      case 2: if (lean) return 2; emul_launch(nblk, kThreads, [&] { spmm_kernel<2, 2>(a); }); return 0;
      case 4: if (lean) return 2; emul_launch(nblk, kThreads, [&] { spmm_kernel<4, 2>(a); }); return 0;
      case 8: run_variant<8, 2>(lean, a, nblk); return 0;
      case 16: run_variant<16, 2>(lean, a, nblk); return 0;
      case 32: run_variant<32, 2>(lean, a, nblk); return 0;
    }
  }
  if (cpl == 4) {
    switch (t) {
      case 8: run_variant<8, 4>(lean, a, nblk); return 0;
      case 16: run_variant<16, 4>(lean, a, nblk); return 0;
      case 32: run_variant<32, 4>(lean, a, nblk); return 0;
    }
  }
  return 1;
}

}  // namespace

extern "C" {

int emul_spmm(int lean, int cpl, int m, const int* rowPtr, const int* colInd, const double* val, const double* X, int ldx,
              const double* H, double* Y, int ldy, int t) {
  return run_spmm(lean, cpl, m, rowPtr, colInd, val, X, ldx, H, Y, ldy, t);
}

// the overlapped product of pcu_spmm_apply_exchange: local part, then the halo entries of the boundary rows
int emul_spmm_split(int lean, int cpl, int m, const int* rowPtr, const int* colInd, const double* val, const double* X,
                    int ldx, const double* H, double* Y, int ldy, int t, int* nbrow_out) {
  std::vector<int> lrp, lci, brow, hptr, hcol;
  std::vector<double> lv, hv;
  split_local_halo(m, rowPtr, colInd, val, &lrp, &lci, &lv, &brow, &hptr, &hcol, &hv);
  if (lci.empty()) { lci.push_back(0); lv.push_back(0.0); }
  const int rc = run_spmm(lean, cpl, m, lrp.data(), lci.data(), lv.data(), X, ldx, H, Y, ldy, t);
  if (rc) return rc;
  const int nb = (int)brow.size();
  if (nbrow_out) *nbrow_out = nb;
  if (nb > 0) {
    const int grid = std::max(1, std::min((nb * 16 + kThreads - 1) / kThreads, 7));
    emul_launch(grid, kThreads, [&] { halo_add_kernel(nb, brow.data(), hptr.data(), hcol.data(), hv.data(), H, t, Y, ldy); });
  }
  return 0;
}

// halo_pack_kernel: out[r, :] = X[idx[r], :t]
int emul_halo_pack(const double* X, int ldx, int t, const int* idx, int nrows, double* out) {
  emul_launch(3, 256, [&] { halo_pack_kernel(X, ldx, t, idx, nrows, out); });
  return 0;
}

}  // extern "C"
