// spmm.cu -- K1: CSR SpMM over the m x t enlarged block, plus the halo pack/exchange.
//
// Y = A_loc * [X ; H].  Replaces mkl_dcsrmm as driven by CPLM_MatCSRMatMult_v2
// (ref: utils/cplm_v0/cplm_v0_matmult_v2.c:108-343, utils/cplm_light/cplm_kernels.c:620-671).
//
// Kernel shape (HBM-bound, see DESIGN.md "SpMM"):
//   * a CTA owns a contiguous row block whose col/val stream (<= kNnzCap entries)
//     is one contiguous chunk of the CSR arrays: it is staged into shared memory
//     with fully coalesced loads, so the 12 B/nnz stream is read from HBM exactly
//     once and never at sub-sector granularity;
//   * a group of G = T/2 lanes owns one row (vector-per-lane: each lane keeps two
//     adjacent columns in a double2), walks that row's entries out of shared
//     memory (broadcast reads) and gathers the T-wide row-major X rows with
//     128-bit loads, i.e. one 16*G-byte contiguous segment per non-zero;
//   * no atomics, no cross-lane reduction: fixed summation order per row
//     (ascending column), bit-reproducible.
#include "common.cuh"

#include <algorithm>
#include "spmm_kernels.cuh"

namespace {

// Row-block shapes (entries / rows per CTA) of 768/64, 1024/96, 2048/192 and 3072/256 and the 4-columns-per-lane mapping for
// long rows were measured against this one on B200 and lost or tied (profiles/r02_spmm_shapes_rejected.md)
template <int T>
void launch_bulk(const SpmmArgs& a, int nblk, bool wide, bool halo, cudaStream_t st) {
  if (wide) {
    if (halo) spmm_bulk_kernel<T, 4, true><<<nblk, kThreads, 0, st>>>(a);
    else spmm_bulk_kernel<T, 4, false><<<nblk, kThreads, 0, st>>>(a);
  } else {
    if (halo) spmm_bulk_kernel<T, 2, true><<<nblk, kThreads, 0, st>>>(a);
    else spmm_bulk_kernel<T, 2, false><<<nblk, kThreads, 0, st>>>(a);
  }
}

// one CSR on the device with its row blocks
struct CsrDev {
  int* rowPtr = nullptr;
  int* colInd = nullptr;
  double* val = nullptr;
  int4* blk[2] = {nullptr, nullptr};  // row blocks of shape kShapeRows/kShapeNnz[i]
  int nblk[2] = {0, 0};
  int64_t nnz = 0;
  bool fits0 = true;  // every shape-0 row block fits the staging buffer (spmm_bulk_kernel needs that)
};

int upload_csr(int m, const int* rowPtr, const int* colInd, const double* val, CsrDev* d) {
  d->nnz = rowPtr[m];
  std::vector<int4> blk[2];
  for (int sh = 0; sh < 2; ++sh) {
    build_row_blocks(m, rowPtr, sh, &blk[sh]);
    d->nblk[sh] = (int)blk[sh].size();
  }
  for (const int4& b : blk[0])
    if (b.w - b.z > kShapeNnz[0]) d->fits0 = false;
  PCU_CUDA(cudaMalloc(&d->rowPtr, sizeof(int) * (size_t)(m + 1)));
  // 16 bytes of slack: spmm_bulk_kernel rounds the size of its bulk copies up to a multiple of 16
  PCU_CUDA(cudaMalloc(&d->colInd, sizeof(int) * (size_t)(d->nnz + 4)));
  PCU_CUDA(cudaMalloc(&d->val, sizeof(double) * (size_t)(d->nnz + 2)));
  for (int sh = 0; sh < 2; ++sh) PCU_CUDA(cudaMalloc(&d->blk[sh], sizeof(int4) * std::max<size_t>(blk[sh].size(), 1)));
  PCU_CUDA(cudaMemcpy(d->rowPtr, rowPtr, sizeof(int) * (size_t)(m + 1), cudaMemcpyHostToDevice));
  PCU_CUDA(cudaMemcpy(d->colInd, colInd, sizeof(int) * (size_t)d->nnz, cudaMemcpyHostToDevice));
  PCU_CUDA(cudaMemcpy(d->val, val, sizeof(double) * (size_t)d->nnz, cudaMemcpyHostToDevice));
  for (int sh = 0; sh < 2; ++sh)
    if (!blk[sh].empty())
      PCU_CUDA(cudaMemcpy(d->blk[sh], blk[sh].data(), sizeof(int4) * blk[sh].size(), cudaMemcpyHostToDevice));
  return 0;
}

void free_csr(CsrDev* d) {
  cudaFree(d->rowPtr); cudaFree(d->colInd); cudaFree(d->val); cudaFree(d->blk[0]); cudaFree(d->blk[1]);
  *d = CsrDev();
}

}  // namespace

struct pcu_spmm {
  pcu_ctx* ctx = nullptr;
  int m = 0, nhalo = 0;
  int64_t nnz = 0;
  CsrDev A;           // the local row panel, columns >= m read the halo buffer
  bool bulk = true;   // spmm_bulk_kernel (cp.async.bulk staging) from t = 8 up; PREALPS_SPMM_BULK=0 keeps spmm_kernel (A/B runs)
  // nhalo > 0 (unless PREALPS_SPMM_OVERLAP=0): the panel split into its entries with column < m (Aloc, same rows) and the halo
  // entries of the boundary rows; the halo exchange then runs on comm_stream next to the local kernel
  bool overlap = false;
  CsrDev Aloc;
  int nbrow = 0;
  int* d_brow = nullptr;   // boundary rows, ascending
  int* d_hptr = nullptr;   // nbrow + 1
  int* d_hcol = nullptr;   // halo row (column - m)
  double* d_hval = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_x = nullptr, ev_h = nullptr;
  // halo
  int nnbr = 0;
  std::vector<int> nbr_rank, send_ptr, recv_ptr;
  int* d_send_idx = nullptr;
  int nsend = 0;
  double* d_sendbuf = nullptr;
  double* d_halo = nullptr;
  int buf_t = 0;  // width the halo/send buffers are currently sized for
};

using namespace pcu;

static int ensure_halo_buffers(pcu_spmm* op, int t) {
  if (op->buf_t >= t) return 0;
  pcu_ctx* c = op->ctx;
  PCU_CUDA(cudaStreamSynchronize(c->stream));
  if (op->comm_stream) PCU_CUDA(cudaStreamSynchronize(op->comm_stream));
  if (op->d_sendbuf) cudaFree(op->d_sendbuf);
  if (op->d_halo) cudaFree(op->d_halo);
  op->d_sendbuf = op->d_halo = nullptr;
  const int tt = std::max(t, 8);
  PCU_CUDA(cudaMalloc(&op->d_sendbuf, sizeof(double) * (size_t)std::max(op->nsend, 1) * tt));
  PCU_CUDA(cudaMalloc(&op->d_halo, sizeof(double) * (size_t)std::max(op->nhalo, 1) * tt));
  op->buf_t = tt;
  return 0;
}

extern "C" {

int pcu_spmm_create(pcu_ctx* ctx, int m, int nhalo, const int* rowPtr, const int* colInd,
                    const double* val, pcu_spmm** out) {
  PCU_CHECK(ctx && rowPtr && colInd && val && out && m >= 0 && nhalo >= 0, "pcu_spmm_create: bad arguments");
  PCU_CUDA(cudaSetDevice(ctx->device));
  pcu_spmm* op = new pcu_spmm();
  op->ctx = ctx;
  op->m = m;
  op->nhalo = nhalo;
  op->nnz = rowPtr[m];
  for (int64_t p = 0; p < op->nnz; ++p)
    PCU_CHECK(colInd[p] >= 0 && colInd[p] < m + nhalo, "pcu_spmm_create: column index %d out of range at %lld",
              colInd[p], (long long)p);
  if (upload_csr(m, rowPtr, colInd, val, &op->A)) return 1;
  if (const char* e = getenv("PREALPS_SPMM_BULK")) op->bulk = atoi(e) != 0;
  // the halo exchange overlapped with the local part of the product, like the reference's MPI_Isend / diagonal block /
  // MPI_Irecv (ref: cplm_v0_matmult_v2.c:182-276): default whenever there is a halo; PREALPS_SPMM_OVERLAP=0 serialises
  // pack -> exchange -> product (measured on 8 B200: SpMM 58.3 -> 54.5 us, the iteration unchanged at 1.11 ms)
  const char* ov = getenv("PREALPS_SPMM_OVERLAP");
  if (nhalo > 0 && !(ov && atoi(ov) == 0)) {
    // split: Aloc keeps the entries with column < m of every row; (brow, hptr, hcol, hval) the others
    std::vector<int> lrp, lci, brow, hptr, hcol;
    std::vector<double> lv, hv;
    split_local_halo(m, rowPtr, colInd, val, &lrp, &lci, &lv, &brow, &hptr, &hcol, &hv);
    if (lci.empty()) { lci.push_back(0); lv.push_back(0.0); }
    if (upload_csr(m, lrp.data(), lci.data(), lv.data(), &op->Aloc)) return 1;
    op->nbrow = (int)brow.size();
    PCU_CUDA(cudaMalloc(&op->d_brow, sizeof(int) * std::max<size_t>(brow.size(), 1)));
    PCU_CUDA(cudaMalloc(&op->d_hptr, sizeof(int) * hptr.size()));
    PCU_CUDA(cudaMalloc(&op->d_hcol, sizeof(int) * std::max<size_t>(hcol.size(), 1)));
    PCU_CUDA(cudaMalloc(&op->d_hval, sizeof(double) * std::max<size_t>(hv.size(), 1)));
    PCU_CUDA(cudaMemcpy(op->d_brow, brow.data(), sizeof(int) * brow.size(), cudaMemcpyHostToDevice));
    PCU_CUDA(cudaMemcpy(op->d_hptr, hptr.data(), sizeof(int) * hptr.size(), cudaMemcpyHostToDevice));
    PCU_CUDA(cudaMemcpy(op->d_hcol, hcol.data(), sizeof(int) * hcol.size(), cudaMemcpyHostToDevice));
    PCU_CUDA(cudaMemcpy(op->d_hval, hv.data(), sizeof(double) * hv.size(), cudaMemcpyHostToDevice));
    PCU_CUDA(cudaStreamCreateWithFlags(&op->comm_stream, cudaStreamNonBlocking));
    PCU_CUDA(cudaEventCreateWithFlags(&op->ev_x, cudaEventDisableTiming));
    PCU_CUDA(cudaEventCreateWithFlags(&op->ev_h, cudaEventDisableTiming));
    op->overlap = true;
  }
  *out = op;
  return 0;
}

int pcu_spmm_destroy(pcu_spmm* op) {
  if (!op) return 0;
  cudaSetDevice(op->ctx->device);
  cudaStreamSynchronize(op->ctx->stream);
  free_csr(&op->A);
  if (op->overlap) {
    cudaStreamSynchronize(op->comm_stream);
    free_csr(&op->Aloc);
    cudaFree(op->d_brow); cudaFree(op->d_hptr); cudaFree(op->d_hcol); cudaFree(op->d_hval);
    cudaEventDestroy(op->ev_x); cudaEventDestroy(op->ev_h);
    cudaStreamDestroy(op->comm_stream);
  }
  if (op->d_send_idx) cudaFree(op->d_send_idx);
  if (op->d_sendbuf) cudaFree(op->d_sendbuf);
  if (op->d_halo) cudaFree(op->d_halo);
  delete op;
  return 0;
}

int pcu_spmm_set_halo(pcu_spmm* op, int nnbr, const int* nbr_rank, const int* send_ptr,
                      const int* send_idx, const int* recv_ptr) {
  PCU_CHECK(op && nnbr >= 0, "pcu_spmm_set_halo: bad arguments");
  op->nnbr = nnbr;
  op->nbr_rank.assign(nbr_rank, nbr_rank + nnbr);
  op->send_ptr.assign(send_ptr, send_ptr + nnbr + 1);
  op->recv_ptr.assign(recv_ptr, recv_ptr + nnbr + 1);
  op->nsend = nnbr ? send_ptr[nnbr] : 0;
  PCU_CHECK((nnbr ? recv_ptr[nnbr] : 0) == op->nhalo, "pcu_spmm_set_halo: recv_ptr covers %d rows, halo has %d",
            nnbr ? recv_ptr[nnbr] : 0, op->nhalo);
  for (int i = 0; i < op->nsend; ++i)
    PCU_CHECK(send_idx[i] >= 0 && send_idx[i] < op->m, "pcu_spmm_set_halo: send index out of range");
  if (op->d_send_idx) { cudaFree(op->d_send_idx); op->d_send_idx = nullptr; }
  PCU_CUDA(cudaMalloc(&op->d_send_idx, sizeof(int) * (size_t)std::max(op->nsend, 1)));
  if (op->nsend) PCU_CUDA(cudaMemcpy(op->d_send_idx, send_idx, sizeof(int) * (size_t)op->nsend, cudaMemcpyHostToDevice));
  op->buf_t = 0;
  return 0;
}

double* pcu_spmm_halo_buffer(pcu_spmm* op, int t) {
  if (ensure_halo_buffers(op, t)) return nullptr;
  return op->d_halo;
}

static int halo_pack_on(pcu_spmm* op, const double* X, int ldx, int t, cudaStream_t st) {
  pcu_ctx* c = op->ctx;
  if (op->nsend > 0) {
    const int grid = stream_grid(c, (int64_t)op->nsend * t, 256, 4);
    halo_pack_kernel<<<grid, 256, 0, st>>>(X, ldx, t, op->d_send_idx, op->nsend, op->d_sendbuf);
    PCU_LAUNCH_CHECK(c);
  }
  return 0;
}

int pcu_spmm_halo_pack(pcu_spmm* op, const double* X, int ldx, int t, double** packed_dev, int* nrows) {
  if (ensure_halo_buffers(op, t)) return 1;
  if (halo_pack_on(op, X, ldx, t, op->ctx->stream)) return 1;
  if (packed_dev) *packed_dev = op->d_sendbuf;
  if (nrows) *nrows = op->nsend;
  return 0;
}

// pack + grouped ncclSend/ncclRecv of the boundary rows on stream st
static int halo_exchange_on(pcu_spmm* op, const double* X, int ldx, int t, cudaStream_t st) {
  pcu_ctx* c = op->ctx;
  PCU_CHECK(c->nccl_comm != nullptr, "pcu_spmm_halo_exchange: %d neighbours but no NCCL communicator", op->nnbr);
  if (halo_pack_on(op, X, ldx, t, st)) return 1;
  if (nccl_group_start(c)) return 1;
  for (int q = 0; q < op->nnbr; ++q) {
    const int ns = op->send_ptr[q + 1] - op->send_ptr[q], nr = op->recv_ptr[q + 1] - op->recv_ptr[q];
    if (ns > 0 && nccl_send(c, op->d_sendbuf + (size_t)op->send_ptr[q] * t, (size_t)ns * t, 1, op->nbr_rank[q], st)) return 1;
    if (nr > 0 && nccl_recv(c, op->d_halo + (size_t)op->recv_ptr[q] * t, (size_t)nr * t, 1, op->nbr_rank[q], st)) return 1;
  }
  if (nccl_group_end(c)) return 1;
  return 0;
}

int pcu_spmm_halo_exchange(pcu_spmm* op, const double* X, int ldx, int t) {
  if (op->nnbr == 0) return 0;
  if (ensure_halo_buffers(op, t)) return 1;
  return halo_exchange_on(op, X, ldx, t, op->ctx->stream);
}

// Y = A [X ; H] for one CSR of the operator (the whole panel, or its local part with H unused)
static int launch_spmm(pcu_spmm* op, const CsrDev& A, const double* X, int ldx, double* Y, int ldy, int t) {
  pcu_ctx* c = op->ctx;
  const int sh = (t <= 4) ? 1 : 0;
  const int nblk = A.nblk[sh];
  SpmmArgs a{A.rowPtr, A.colInd, A.val, A.blk[sh], op->m, X, ldx, op->d_halo, Y, ldy, t};
  const bool aligned = (ldx % 2 == 0) && (ldy % 2 == 0) && ((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0);
  const bool pow2 = (t == 2 || t == 4 || t == 8 || t == 16 || t == 32);
  // 256-bit accesses need 32-byte aligned rows (the halo buffer has ld = t). They pay for short rows, where the
  // per-row instructions dominate and twice the rows per warp halves them (7-point: 126 -> 112 us at t = 8); with 27
  // entries per row a block holds fewer rows than the CTA has lane groups and the narrow mapping is faster.
  const bool wide = (t % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && ((uintptr_t)X % 32 == 0) && ((uintptr_t)Y % 32 == 0) &&
                    ((uintptr_t)op->d_halo % 32 == 0) && A.nnz <= 12 * (int64_t)op->m;
  const bool bulk = op->bulk && A.fits0 && aligned && pow2 && t >= 8 && ldx == t;
  if (bulk) {
    const bool halo = (&A == &op->A) && op->nhalo > 0;  // the local part of the overlapped product has no column >= m
    if (t == 8) launch_bulk<8>(a, nblk, wide, halo, c->stream);
    else if (t == 16) launch_bulk<16>(a, nblk, wide, halo, c->stream);
    else launch_bulk<32>(a, nblk, wide, halo, c->stream);
  } else if (t == 1) spmm_kernel<1, 1><<<nblk, kThreads, 0, c->stream>>>(a);
  else if (aligned && pow2) {
    switch (t) {
      case 2: spmm_kernel<2, 2><<<nblk, kThreads, 0, c->stream>>>(a); break;
      case 4: spmm_kernel<4, 2><<<nblk, kThreads, 0, c->stream>>>(a); break;
      case 8:
        if (wide) spmm_kernel<8, 4><<<nblk, kThreads, 0, c->stream>>>(a);
        else spmm_kernel<8, 2><<<nblk, kThreads, 0, c->stream>>>(a);
        break;
      case 16:
        if (wide) spmm_kernel<16, 4><<<nblk, kThreads, 0, c->stream>>>(a);
        else spmm_kernel<16, 2><<<nblk, kThreads, 0, c->stream>>>(a);
        break;
      default:
        if (wide) spmm_kernel<32, 4><<<nblk, kThreads, 0, c->stream>>>(a);
        else spmm_kernel<32, 2><<<nblk, kThreads, 0, c->stream>>>(a);
        break;
    }
  } else {
    spmm_kernel_generic<<<nblk, kThreads, 0, c->stream>>>(a);
  }
  PCU_LAUNCH_CHECK(c);
  return 0;
}

int pcu_spmm_apply(pcu_spmm* op, const double* X, int ldx, double* Y, int ldy, int t) {
  PCU_CHECK(op && X && Y && t >= 1, "pcu_spmm_apply: bad arguments");
  PCU_CHECK(X != Y, "pcu_spmm_apply: X and Y must not alias");
  if (op->m == 0) return 0;
  if (op->nhalo > 0 && ensure_halo_buffers(op, t)) return 1;
  PCU_CHECK(t <= 32, "pcu_spmm_apply: t=%d > 32 is not supported", t);
  return launch_spmm(op, op->A, X, ldx, Y, ldy, t);
}

// halo exchange over NCCL + product: the exchange (pack + grouped ncclSend / ncclRecv) runs on a second stream while the
// local part of the product (entries with column < m: > 99 % of them) runs on the library stream, then halo_add_kernel
// adds the halo entries of the boundary rows -- what the reference does with MPI_Isend / diagonal block / MPI_Irecv
// (ref: utils/cplm_v0/cplm_v0_matmult_v2.c:182-276).  PREALPS_SPMM_OVERLAP=0 at creation: one after the other.
int pcu_spmm_apply_exchange(pcu_spmm* op, const double* X, int ldx, double* Y, int ldy, int t) {
  PCU_CHECK(op && X && Y && t >= 1 && t <= 32, "pcu_spmm_apply_exchange: bad arguments");
  PCU_CHECK(X != Y, "pcu_spmm_apply_exchange: X and Y must not alias");
  pcu_ctx* c = op->ctx;
  if ((op->nhalo > 0 || op->nsend > 0) && ensure_halo_buffers(op, t)) return 1;
  if (!op->overlap || op->nnbr == 0) {
    if (op->nnbr > 0 && halo_exchange_on(op, X, ldx, t, c->stream)) return 1;
    if (op->m == 0) return 0;
    return launch_spmm(op, op->A, X, ldx, Y, ldy, t);
  }
  // X is ready once everything queued on the library stream so far has run (that also orders this exchange after
  // the previous product's reads of the halo buffer)
  PCU_CUDA(cudaEventRecord(op->ev_x, c->stream));
  PCU_CUDA(cudaStreamWaitEvent(op->comm_stream, op->ev_x, 0));
  if (halo_exchange_on(op, X, ldx, t, op->comm_stream)) return 1;
  PCU_CUDA(cudaEventRecord(op->ev_h, op->comm_stream));
  if (op->m > 0 && launch_spmm(op, op->Aloc, X, ldx, Y, ldy, t)) return 1;
  PCU_CUDA(cudaStreamWaitEvent(c->stream, op->ev_h, 0));
  if (op->nbrow > 0) {
    const int grid = stream_grid(c, (int64_t)op->nbrow * 16, kThreads, 8);
    halo_add_kernel<<<grid, kThreads, 0, c->stream>>>(op->nbrow, op->d_brow, op->d_hptr, op->d_hcol, op->d_hval, op->d_halo, t, Y,
                                                      ldy);
    PCU_LAUNCH_CHECK(c);
  }
  return 0;
}

double pcu_spmm_bytes(pcu_spmm* op, int t) {
  // nnz*(8+4) + (m+1)*4 + read X once + write Y once (+ halo rows read once)
  return (double)op->nnz * 12.0 + (double)(op->m + 1) * 4.0 + 2.0 * (double)op->m * t * 8.0 +
         (double)op->nhalo * t * 8.0;
}

}  // extern "C"
