/*
 * pa_ecg.c -- preAlps_ECG* reverse-communication solver (ref: src/solvers/ecg.c:41-728).
 *
 * Control flow, stopping rule and protocol are the reference's; the numerics run as three
 * fused streaming passes per iteration on device-resident ROW_MAJOR blocks
 * (include/prealps_cuda.h):
 *   Iterate(rci 0), ref ecg.c:421-507:
 *     pass 1  G = AP^T P and Gpr = P^T R in one sweep            (pcu_gram2)
 *     one all-reduce of 2 t^2 doubles (the reference does two: ecg.c:427 and :441)
 *     pass 2  U = chol(G); P,AP <- .U^{-1}; alpha = U^{-T} Gpr; X += P alpha; R -= AP alpha;
 *             ||R||_F^2                                         (pcu_ortho_update)
 *   StoppingCriterion, ref ecg.c:223-271: res = sqrt(trace(R^T R)) -- the squared Frobenius norm
 *     was produced by pass 2, so only its all-reduce and an 8-byte read-back remain
 *   Iterate(rci 1), ref ecg.c:508-527:
 *     pass 3  beta = [AP^T Z ; APprev^T Z]                       (pcu_gram2), all-reduce 2 t^2
 *     pass 4  Z -= P beta1 + Pprev beta2                         (pcu_update_z)
 *     the three block copies of ecg.c:521-523 become a pointer rotation.
 * alpha = U^{-T}(P_old^T R) equals (P U^{-1})^T R of the reference in exact arithmetic; the
 * rounding differs at the 1e-16 level ("parity").
 *
 * ADAPT_BS with Orthodir (ref: ecg.c:445-497), see adapt_half_step(): until the first reduction the
 * iteration is the one above plus a t x t SVD on the host; from the first reduction on the blocks keep
 * the reference's slot layout (V = [slot0 | slot1], live directions first, discarded ones behind them)
 * and the copies of ecg.c:521-523 are real column-range copies.
 */
#include "pa_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  preAlps_ECG_t* owner;
  int m, t, ld;
  double* buf[7];       /* P, Pprev, AP, APprev, Z, R, X (roles rotate) */
  int nbuf;
  double *P, *Pp, *AP, *APp, *Z, *R, *X;
  double* small;        /* G | Gpr | U | alpha | beta1 | beta2 | rr_local | rr_glob */
  double* rhs_dev;
  int* col_of_row;
  int* status_dev;
  int have_rr;
  int iter_since_reset;
  /* ADAPT_BS: after the first reduction the buffers stop rotating (slot0 = P/AP, slot1 = Pp/APp) */
  int adapt, fixed;
  int tprev;            /* columns of slot1 that still take part in the A-orthogonalisation (kbs - t) */
  int zcols_clean;      /* ADAPT_BS: the columns of Z beyond bs are known to be zero */
} ecg_priv_t;

#define MAX_SOLVERS 16
static ecg_priv_t* g_priv[MAX_SOLVERS];

static ecg_priv_t* priv_of(preAlps_ECG_t* ecg) {
  for (int i = 0; i < MAX_SOLVERS; ++i) if (g_priv[i] && g_priv[i]->owner == ecg) return g_priv[i];
  return NULL;
}

static double* sm_G(ecg_priv_t* p) { return p->small; }
static double* sm_Gpr(ecg_priv_t* p) { return p->small + (size_t)p->t * p->t; }
static double* sm_U(ecg_priv_t* p) { return p->small + 2 * (size_t)p->t * p->t; }
static double* sm_alpha(ecg_priv_t* p) { return p->small + 3 * (size_t)p->t * p->t; }
static double* sm_beta1(ecg_priv_t* p) { return p->small + 4 * (size_t)p->t * p->t; }
static double* sm_beta2(ecg_priv_t* p) { return p->small + 5 * (size_t)p->t * p->t; }
static double* sm_rr(ecg_priv_t* p) { return p->small + 6 * (size_t)p->t * p->t; }
/* ORTHODIR_FUSED keeps its five reduced products adjacent: alpha | beta1 | beta2 | mu | rr (one all-reduce, ref: ecg.c:563) */
static double* fu_alpha(ecg_priv_t* p) { return p->small; }
static double* fu_beta1(ecg_priv_t* p) { return p->small + (size_t)p->t * p->t; }
static double* fu_beta2(ecg_priv_t* p) { return p->small + 2 * (size_t)p->t * p->t; }
static double* fu_mu(ecg_priv_t* p) { return p->small + 3 * (size_t)p->t * p->t; }
static double* fu_rr(ecg_priv_t* p) { return p->small + 4 * (size_t)p->t * p->t; }
/* behind sm_rr / rr_glob (6 t^2, 6 t^2 + 1): no view of the pool overlaps another one that is live at the same time */
static double* fu_aout(ecg_priv_t* p) { return p->small + 6 * (size_t)p->t * p->t + 8; }

static void set_shell(CPLM_Mat_Dense_t* s, double* val, int M, int m, int n, int ld) {
  CPLM_MatDenseSetInfo(s, M, n, m, n, ROW_MAJOR);
  s->info.lda = ld;
  s->val = val;
}

static void refresh_shells(preAlps_ECG_t* ecg, ecg_priv_t* p) {
  const int M = ecg->globPbSize, m = p->m, bs = ecg->bs > 0 ? ecg->bs : p->t;
  set_shell(ecg->X, p->X, M, m, p->t, p->ld);
  set_shell(ecg->R, p->R, M, m, p->t, p->ld);
  set_shell(ecg->P, p->P, M, m, bs, p->ld);
  set_shell(ecg->AP, p->AP, M, m, bs, p->ld);
  set_shell(ecg->V, p->P, M, m, bs, p->ld);
  set_shell(ecg->AV, p->AP, M, m, bs, p->ld);
  set_shell(ecg->Z, p->Z, M, m, bs, p->ld);
  ecg->P_p = p->P; ecg->AP_p = p->AP; ecg->R_p = p->R; ecg->Z_p = p->Z;
}

int _preAlps_ECGMalloc(preAlps_ECG_t* ecg) {
  pcu_ctx* c = pa_ctx();
  const int m = ecg->locPbSize, t = ecg->enlFac;
  if (t < 1 || t > 32) CPLM_Abort("enlarging factor %d is outside the supported range 1..32", t);
  ecg_priv_t* p = (ecg_priv_t*)pa_xcalloc(1, sizeof(ecg_priv_t));
  int slot = -1;
  for (int i = 0; i < MAX_SOLVERS; ++i) if (!g_priv[i]) { slot = i; break; }
  if (slot < 0) CPLM_Abort("too many live ECG solvers");
  g_priv[slot] = p;
  p->owner = ecg; p->m = m; p->t = t;
  p->ld = (t % 2 == 0 || t == 1) ? t : t + 1;  /* even row stride keeps 16-byte vector access legal */
  const size_t blk = (size_t)m * p->ld;
  const size_t smalls = 8 * (size_t)t * t + 16;
  /* one pool, like the reference's mkl_calloc(7mt + 3t^2) (ref: ecg.c:58-62), but in HBM */
  p->nbuf = 7;
  ecg->work = (double*)pcu_malloc(c, sizeof(double) * (p->nbuf * blk + smalls + 8));
  if (!ecg->work) CPLM_Abort("device allocation of the ECG pool failed: %s", pcu_last_error());
  pa_cuda_check(pcu_memset(c, ecg->work, 0, sizeof(double) * (p->nbuf * blk + smalls + 8)), "pcu_memset");
  for (int i = 0; i < p->nbuf; ++i) p->buf[i] = ecg->work + (size_t)i * blk;
  p->small = ecg->work + p->nbuf * blk;
  p->rhs_dev = (double*)pcu_malloc(c, sizeof(double) * (size_t)(m > 0 ? m : 1));
  p->col_of_row = (int*)pcu_malloc(c, sizeof(int) * (size_t)(m > 0 ? m : 1));
  p->status_dev = (int*)pcu_malloc(c, sizeof(int) * 4);
  if (!p->rhs_dev || !p->col_of_row || !p->status_dev) CPLM_Abort("device allocation failed: %s", pcu_last_error());
  ecg->iwork = (int*)pa_xcalloc((size_t)t, sizeof(int));
  ecg->X = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->R = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->V = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->AV = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->Z = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->alpha = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->beta = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->P = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  ecg->AP = (CPLM_Mat_Dense_t*)pa_xcalloc(1, sizeof(CPLM_Mat_Dense_t));
  return 0;
}

int _preAlps_ECGReset(preAlps_ECG_t* ecg, double* rhs, int* rci_request) {
  ecg_priv_t* p = priv_of(ecg);
  if (!p) CPLM_Abort("_preAlps_ECGReset on a solver that was not allocated");
  pcu_ctx* c = pa_ctx();
  pa_state_t* g = &pa_g;
  ecg->tot_t = ecg->comm_t = ecg->trsm_t = ecg->gemm_t = ecg->potrf_t = ecg->pstrf_t = 0.0;
  ecg->lapmt_t = ecg->gesvd_t = ecg->geqrf_t = ecg->ormqr_t = ecg->copy_t = 0.0;
  const int m = p->m, t = p->t;
  const size_t blk = (size_t)m * p->ld;
  p->P = p->buf[0]; p->Pp = p->buf[1]; p->AP = p->buf[2]; p->APp = p->buf[3];
  p->Z = p->buf[4]; p->R = p->buf[5]; p->X = p->buf[6];
  pa_cuda_check(pcu_memset(c, ecg->work, 0, sizeof(double) * p->nbuf * blk), "pcu_memset");
  /* ||b||: per-subdomain sums in row order, then summed over subdomains/processes (ref: ecg.c:143-155) */
  double nb = 0.0;
  int* cor = (int*)pa_xmalloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  const int nsub = g->built ? g->s_hi - g->s_lo : 1;
  double* nb_parts = (double*)pa_xcalloc((size_t)nsub, sizeof(double));
  for (int s = 0; s < nsub; ++s) {
    const int r0 = g->built ? g->rowPos[g->s_lo + s] - g->g0 : 0;
    const int r1 = g->built ? g->rowPos[g->s_lo + s + 1] - g->g0 : m;
    double part = 0.0;
    for (int i = r0; i < r1; ++i) part += rhs[i] * rhs[i];  /* == pow(x, 2) bit for bit */
    nb_parts[s] = part;
    nb += part;
    /* R0 = T(b): the rows of subdomain s feed column s % t (ref: ecg.c:162 with rank -> subdomain id) */
    const int col = ((g->built ? g->s_lo : g->rank) + s) % t;
    for (int i = r0; i < r1; ++i) cor[i] = col;
  }
  if (g->nproc > 1 && g->xport == PA_XPORT_MPI) {
    const double t0 = pa_wtime();
    MPI_Allreduce(MPI_IN_PLACE, &nb, 1, MPI_DOUBLE, MPI_SUM, ecg->comm);
    ecg->comm_t += pa_wtime() - t0;
  } else if (g->nproc > 1 && g->xport == PA_XPORT_NCCL && g->built) {
    const double t0 = pa_wtime();
    nb = pa_sum_over_subdomains(nb_parts);  /* rank order of the reference's reduction, whatever the number of GPUs */
    ecg->comm_t += pa_wtime() - t0;
  }
  free(nb_parts);
  ecg->normb = sqrt(nb);
  ecg->res = 1.0;
  ecg->iter = 0;
  ecg->bs = t;
  ecg->kbs = (ecg->ortho_alg == ORTHOMIN) ? t : 2 * t;
  pa_cuda_check(pcu_h2d(c, p->rhs_dev, rhs, sizeof(double) * (size_t)m), "pcu_h2d");
  pa_cuda_check(pcu_h2d(c, p->col_of_row, cor, sizeof(int) * (size_t)m), "pcu_h2d");
  free(cor);
  pa_cuda_check(pcu_split_rhs(c, m, t, p->rhs_dev, p->col_of_row, p->R, p->ld), "pcu_split_rhs");
  refresh_shells(ecg, p);
  CPLM_MatDenseSetInfo(ecg->alpha, t, t, t, t, COL_MAJOR);
  ecg->alpha->val = sm_alpha(p);
  CPLM_MatDenseSetInfo(ecg->beta, ecg->kbs, t, ecg->kbs, t, COL_MAJOR);
  ecg->beta->val = sm_beta1(p);
  p->have_rr = 0;
  p->iter_since_reset = 0;
  p->adapt = (ecg->bs_red == ADAPT_BS);
  p->fixed = 0;
  p->zcols_clean = 0;
  p->tprev = t;
  *rci_request = 0;
  return 0;
}

/* ref: ecg.c:201-221 -- column colIndex of XSplit <- x, everything else untouched (the caller zeroed it).  XSplit is a
 * library block in HBM or a caller's host block of either storage order. */
int _preAlps_ECGSplit(double* x, CPLM_Mat_Dense_t* XSplit, int colIndex) {
  if (!XSplit || !XSplit->val) CPLM_Abort(" wrong test 'XSplit->val != NULL'");
  if (!x) CPLM_Abort(" wrong test 'x != NULL'");
  const int m = XSplit->info.m;
  if (pa_is_device_block(XSplit)) {
    pcu_ctx* c = pa_ctx();
    if (XSplit->info.stor_type != ROW_MAJOR) CPLM_Abort("device blocks must be ROW_MAJOR");
    double* tmp = (double*)pcu_malloc(c, sizeof(double) * (size_t)(m > 0 ? m : 1));
    if (!tmp) CPLM_Abort("device allocation failed: %s", pcu_last_error());
    pa_cuda_check(pcu_h2d(c, tmp, x, sizeof(double) * (size_t)m), "pcu_h2d");
    pa_cuda_check(pcu_copy_cols(c, m, 1, XSplit->val + colIndex, XSplit->info.lda, tmp, 1), "pcu_copy_cols");
    pcu_free(c, tmp);
    return 0;
  }
  const int row_major = XSplit->info.stor_type == ROW_MAJOR;
  const size_t s1 = row_major ? (size_t)XSplit->info.n : 1, s2 = row_major ? 1 : (size_t)XSplit->info.m;
  for (int i = 0; i < m; ++i) XSplit->val[i * s1 + colIndex * s2] = x[i];
  return 0;
}

int preAlps_ECGInitialize(preAlps_ECG_t* ecg, double* rhs, int* rci_request) {
  /* the reference requires #ranks >= enlFac (ref: ecg.c:178-183); with virtual subdomains the
   * number that matters is the number of METIS subdomains */
  int size = pa_g.built ? pa_g.S : 1;
  if (!pa_g.built) MPI_Comm_size(ecg->comm, &size);
  if (size < ecg->enlFac)
    CPLM_Abort("Enlarging factor must be lower than the number of processors in the MPI communicator! size: %d ; enlarging factor: %d",
               size, ecg->enlFac);
  _preAlps_ECGMalloc(ecg);
  return _preAlps_ECGReset(ecg, rhs, rci_request);
}

static void check_status(preAlps_ECG_t* ecg, ecg_priv_t* p) {
  if (ecg->ortho_alg != ORTHOMIN) return;  /* Orthodir ignores dpotrf's return code (ref: ecg.c:431) */
  int st = 0;
  pa_cuda_check(pcu_d2h(pa_g.ctx, &st, p->status_dev, sizeof(int)), "pcu_d2h");
  if (st != 0) CPLM_Abort("ACHQR: dpotrf:\n ERROR: P^tAP is not spd!");  /* ref: ecg.c:320-322 */
}

/* shared by Orthodir and Orthomin: ref ecg.c:421-443,499-506 == ecg.c:307-343 */
static void descent_half_step(preAlps_ECG_t* ecg, ecg_priv_t* p) {
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, t = ecg->bs, ld = p->ld;
  double t0 = pa_wtime();
  pa_cuda_check(pcu_gram2(c, m, t, p->AP, ld, p->P, ld, sm_G(p), p->P, ld, p->R, ld, sm_Gpr(p)), "pcu_gram2");
  ecg->gemm_t += pa_wtime() - t0;
  pa_allreduce_dev(sm_G(p), 2 * p->t * p->t, &ecg->comm_t);  /* G and Gpr are adjacent */
  t0 = pa_wtime();
  pa_cuda_check(pcu_ortho_update(c, m, t, sm_G(p), sm_Gpr(p), p->P, ld, p->AP, ld, p->X, ld, p->R, ld, sm_U(p),
                                 sm_alpha(p), sm_rr(p), p->status_dev),
                "pcu_ortho_update");
  ecg->trsm_t += pa_wtime() - t0;
  check_status(ecg, p);
  p->have_rr = 1;
  ecg->iter++;
  p->iter_since_reset++;
}

/* ------------------------------------------------------------------------------------------------
 * ADAPT_BS: small dense algebra on the host (t <= 32, column-major) -- stands in for LAPACKE_dpotrf,
 * cblas_dtrsm, LAPACKE_dgesvd('O','N'), dgeqrf and dormqr on t x t data (ref: ecg.c:431-479).
 * ------------------------------------------------------------------------------------------------ */
int pa_h_chol_upper(int n, double* A, int lda) {  /* A = U^T U, upper triangle in place */
  for (int j = 0; j < n; ++j) {
    double d = A[j + (size_t)lda * j];
    for (int k = 0; k < j; ++k) d -= A[k + (size_t)lda * j] * A[k + (size_t)lda * j];
    if (!(d > 0.0)) return j + 1;
    d = sqrt(d);
    A[j + (size_t)lda * j] = d;
    for (int c = j + 1; c < n; ++c) {
      double v = A[j + (size_t)lda * c];
      for (int k = 0; k < j; ++k) v -= A[k + (size_t)lda * j] * A[k + (size_t)lda * c];
      A[j + (size_t)lda * c] = v / d;
    }
  }
  return 0;
}

void pa_h_triu_inv(int n, const double* U, int ldu, double* Ui, int ldi) {
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) Ui[i + (size_t)ldi * j] = 0.0;
    Ui[j + (size_t)ldi * j] = 1.0 / U[j + (size_t)ldu * j];
    for (int i = j - 1; i >= 0; --i) {
      double v = 0.0;
      for (int k = i + 1; k <= j; ++k) v += U[i + (size_t)ldu * k] * Ui[k + (size_t)ldi * j];
      Ui[i + (size_t)ldi * j] = -v / U[i + (size_t)ldu * i];
    }
  }
}

/* Left singular vectors and singular values of the t x n matrix A (column-major, lda): one-sided Jacobi on
 * the rows of A.  On return rows[i*n .. i*n+n) = i-th row of Q^T A (row-major), sv descending, Q t x t
 * column-major (ldq = t). */
void pa_h_left_svd(int t, int n, const double* A, int lda, double* sv, double* Q, double* rows) {
  for (int i = 0; i < t; ++i) for (int c = 0; c < n; ++c) rows[(size_t)i * n + c] = A[i + (size_t)lda * c];
  for (int i = 0; i < t * t; ++i) Q[i] = 0.0;
  for (int i = 0; i < t; ++i) Q[i + (size_t)t * i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int i = 0; i < t - 1; ++i)
      for (int j = i + 1; j < t; ++j) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int k = 0; k < n; ++k) {
          const double x = rows[(size_t)i * n + k], y = rows[(size_t)j * n + k];
          a += x * x; b += y * y; c += x * y;
        }
        if (fabs(c) <= 1e-15 * sqrt(a * b) || c == 0.0) continue;
        rotated = 1;
        const double zeta = (b - a) / (2.0 * c);
        const double tg = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + tg * tg), sn = cs * tg;
        for (int k = 0; k < n; ++k) {
          const double x = rows[(size_t)i * n + k], y = rows[(size_t)j * n + k];
          rows[(size_t)i * n + k] = cs * x - sn * y;
          rows[(size_t)j * n + k] = sn * x + cs * y;
        }
        for (int k = 0; k < t; ++k) {
          const double x = Q[k + (size_t)t * i], y = Q[k + (size_t)t * j];
          Q[k + (size_t)t * i] = cs * x - sn * y;
          Q[k + (size_t)t * j] = sn * x + cs * y;
        }
      }
    if (!rotated) break;
  }
  for (int i = 0; i < t; ++i) {
    double a = 0.0;
    for (int k = 0; k < n; ++k) a += rows[(size_t)i * n + k] * rows[(size_t)i * n + k];
    sv[i] = sqrt(a);
  }
  for (int i = 0; i < t - 1; ++i) {  /* selection sort, descending; rows and columns of Q follow */
    int best = i;
    for (int j = i + 1; j < t; ++j) if (sv[j] > sv[best]) best = j;
    if (best == i) continue;
    double tmp = sv[i]; sv[i] = sv[best]; sv[best] = tmp;
    for (int k = 0; k < n; ++k) { tmp = rows[(size_t)i * n + k]; rows[(size_t)i * n + k] = rows[(size_t)best * n + k]; rows[(size_t)best * n + k] = tmp; }
    for (int k = 0; k < t; ++k) { tmp = Q[k + (size_t)t * i]; Q[k + (size_t)t * i] = Q[k + (size_t)t * best]; Q[k + (size_t)t * best] = tmp; }
  }
}

/* The "rci_request == 0" half of an Orthodir iteration with ADAPT_BS (ref: ecg.c:421-507).
 *
 * Layout: slot0 = (p->P, p->AP) holds [live directions (bs) | every direction discarded so far], slot1 =
 * (p->Pp, p->APp) the previous directions; both T = enlFac columns wide. All block kernels run at the full
 * width T on zero/identity padded small matrices, which is exact: discarded columns are multiplied by the
 * identity, padded rows of alpha are zero. SpMM and block-Jacobi -- where the time goes -- run on bs columns
 * because the shells handed to the caller say n = bs.
 *
 *   G = AP^T P, Gpr = P^T R at width T; all-reduce; read back
 *   host: U = chol(G[:bs,:bs]); alpha = U^-T Gpr[:bs,:]; SVD(alpha) -> t1 = #{sigma > tol*normb/sqrt(T)}
 *   no reduction and no reduction so far: the NO_BS_RED pass (pcu_ortho_update)
 *   otherwise: W = U^-1 Q (Q = left singular vectors; the reference applies the same Q through
 *     dgeqrf/dormqr, ecg.c:470-479 -- equal up to the signs of the columns, which cancel in P alpha)
 *     slot0 <- slot0 * diag(W, I);  X += slot0 * [Q^T alpha (first t1 rows); 0];  R -= A slot0 * [..]
 *     bs = t1, kbs = T + (bs before) (ecg.c:491-492): the columns of slot1 beyond the old bs leave the
 *     A-orthogonalisation for good and are zeroed. */
static void small_matmul(int T, const double* A, const double* B, double* C) {  /* C = A B, T x T column-major */
  for (int j = 0; j < T; ++j)
    for (int i = 0; i < T; ++i) {
      double v = 0.0;
      for (int k = 0; k < T; ++k) v += A[i + (size_t)T * k] * B[k + (size_t)T * j];
      C[i + (size_t)T * j] = v;
    }
}

static void adapt_half_step(preAlps_ECG_t* ecg, ecg_priv_t* p) {
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, T = p->t, ld = p->ld, bs = ecg->bs;
  double t0 = pa_wtime();
  pa_cuda_check(pcu_gram2(c, m, T, p->AP, ld, p->P, ld, sm_G(p), p->P, ld, p->R, ld, sm_Gpr(p)), "pcu_gram2");
  ecg->gemm_t += pa_wtime() - t0;
  pa_allreduce_dev(sm_G(p), 2 * T * T, &ecg->comm_t);
  double G[2 * 32 * 32], Ui[32 * 32], alpha[32 * 32], Q[32 * 32], rows[32 * 32], sv[32];
  pa_cuda_check(pcu_d2h(c, G, sm_G(p), sizeof(double) * 2 * (size_t)T * T), "pcu_d2h");
  const double* Gpr = G + (size_t)T * T;
  t0 = pa_wtime();
  pa_h_chol_upper(bs, G, T);  /* Orthodir ignores dpotrf's return code (ref: ecg.c:431) */
  ecg->potrf_t += pa_wtime() - t0;
  t0 = pa_wtime();
  pa_h_triu_inv(bs, G, T, Ui, bs);
  for (int j = 0; j < T; ++j)      /* alpha = U^-T Gpr[:bs, :] (bs x T, ld bs) */
    for (int i = 0; i < bs; ++i) {
      double v = 0.0;
      for (int k = 0; k <= i; ++k) v += Ui[k + (size_t)bs * i] * Gpr[k + (size_t)T * j];
      alpha[i + (size_t)bs * j] = v;
    }
  ecg->trsm_t += pa_wtime() - t0;
  t0 = pa_wtime();
  pa_h_left_svd(bs, T, alpha, bs, sv, Q, rows);
  ecg->gesvd_t += pa_wtime() - t0;
  const double cut = ecg->tol * ecg->normb / sqrt((double)T);  /* ref: ecg.c:420 */
  int t1 = 0;
  for (int i = 0; i < bs; ++i) { if (sv[i] > cut) t1++; else break; }  /* ref: ecg.c:460-464 */
  const int reduce = (t1 > 0 && t1 < T && t1 < bs);                     /* ref: ecg.c:467 */
  if (!reduce && !p->fixed) {
    t0 = pa_wtime();
    pa_cuda_check(pcu_ortho_update(c, m, T, sm_G(p), sm_Gpr(p), p->P, ld, p->AP, ld, p->X, ld, p->R, ld, sm_U(p),
                                   sm_alpha(p), sm_rr(p), p->status_dev), "pcu_ortho_update");
    ecg->trsm_t += pa_wtime() - t0;
  } else {
    /* W = diag(W, I), Af = [alpha'; 0] and Wx = W Af, T x T column-major: slot0 <- slot0 W, X += slot0_old Wx, ... */
    double W[32 * 32], Af[32 * 32], Wx[32 * 32];
    for (int i = 0; i < T * T; ++i) { W[i] = 0.0; Af[i] = 0.0; }
    for (int j = 0; j < T; ++j) W[j + (size_t)T * j] = 1.0;
    const int keep = reduce ? t1 : bs;
    for (int j = 0; j < bs; ++j)
      for (int i = 0; i < bs; ++i) {
        double v = 0.0;
        if (reduce) { for (int k = i; k < bs; ++k) v += Ui[i + (size_t)bs * k] * Q[k + (size_t)bs * j]; }
        else v = Ui[i + (size_t)bs * j];
        W[i + (size_t)T * j] = v;
      }
    for (int j = 0; j < T; ++j)
      for (int i = 0; i < keep; ++i) Af[i + (size_t)T * j] = reduce ? rows[(size_t)i * T + j] : alpha[i + (size_t)bs * j];
    small_matmul(T, W, Af, Wx);
    pa_cuda_check(pcu_h2d(c, sm_U(p), W, sizeof(double) * (size_t)T * T), "pcu_h2d");
    pa_cuda_check(pcu_h2d(c, sm_alpha(p), Wx, sizeof(double) * (size_t)T * T), "pcu_h2d");
    t0 = pa_wtime();
    /* one pass: slot0 <- slot0 W, A slot0 likewise, X += .., R -= .., ||R||^2 (pcu_transform_update) */
    pa_cuda_check(pcu_transform_update(c, m, T, sm_U(p), sm_alpha(p), p->P, ld, p->AP, ld, p->X, ld, p->R, ld, sm_rr(p)),
                  "pcu_transform_update");
    ecg->ormqr_t += pa_wtime() - t0;
    /* the columns of Z beyond the live directions must read as zero in the full-width passes of the other half step:
     * block-Jacobi only writes bs columns from now on */
    if (reduce || !p->zcols_clean) {
      pa_cuda_check(pcu_zero_cols(c, m, T - keep, p->Z + keep, ld), "pcu_zero_cols");
      p->zcols_clean = 1;
    }
    if (reduce) {
      ecg->bs = t1;
      ecg->kbs = bs + T;
      p->tprev = bs;
      p->fixed = 1;
      pa_cuda_check(pcu_zero_cols(c, m, T - bs, p->Pp + bs, ld), "pcu_zero_cols");
      pa_cuda_check(pcu_zero_cols(c, m, T - bs, p->APp + bs, ld), "pcu_zero_cols");
      refresh_shells(ecg, p);
    }
  }
  p->have_rr = 1;
  ecg->iter++;
  p->iter_since_reset++;
}

int _preAlps_ECGIterateOdir(preAlps_ECG_t* ecg, int* rci_request) {
  ecg_priv_t* p = priv_of(ecg);
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, t = ecg->bs, ld = p->ld;
  if (*rci_request == 0) {
    if (p->adapt) adapt_half_step(ecg, p);
    else descent_half_step(ecg, p);
    *rci_request = 1;
  } else if (*rci_request == 1 && p->fixed) {
    /* after a reduction: beta = AV[:, :kbs]^T Z, Z -= V[:, :kbs] beta at the full width T -- the columns of Z
     * beyond bs and the columns of slot1 beyond kbs - T are zero -- then the copies of ref: ecg.c:521-523 */
    const int T = p->t, bs = ecg->bs;
    double t0 = pa_wtime();
    pa_cuda_check(pcu_gram2(c, m, T, p->AP, ld, p->Z, ld, sm_beta1(p), p->APp, ld, p->Z, ld, sm_beta2(p)), "pcu_gram2");
    ecg->gemm_t += pa_wtime() - t0;
    pa_allreduce_dev(sm_beta1(p), 2 * T * T, &ecg->comm_t);
    t0 = pa_wtime();
    pa_cuda_check(pcu_update_z(c, m, T, p->Z, ld, p->P, ld, T, sm_beta1(p), p->Pp, ld, T, sm_beta2(p)), "pcu_update_z");
    ecg->gemm_t += pa_wtime() - t0;
    t0 = pa_wtime();
    pa_cuda_check(pcu_copy_cols(c, m, bs, p->Pp, ld, p->P, ld), "pcu_copy_cols");
    pa_cuda_check(pcu_copy_cols(c, m, bs, p->APp, ld, p->AP, ld), "pcu_copy_cols");
    pa_cuda_check(pcu_copy_cols(c, m, bs, p->P, ld, p->Z, ld), "pcu_copy_cols");
    ecg->copy_t += pa_wtime() - t0;
    *rci_request = 0;
  } else if (*rci_request == 1) {
    double t0 = pa_wtime();
    pa_cuda_check(pcu_gram2(c, m, t, p->AP, ld, p->Z, ld, sm_beta1(p), p->APp, ld, p->Z, ld, sm_beta2(p)), "pcu_gram2");
    ecg->gemm_t += pa_wtime() - t0;
    pa_allreduce_dev(sm_beta1(p), 2 * p->t * p->t, &ecg->comm_t);
    t0 = pa_wtime();
    pa_cuda_check(pcu_update_z(c, m, t, p->Z, ld, p->P, ld, t, sm_beta1(p), p->Pp, ld, t, sm_beta2(p)), "pcu_update_z");
    ecg->gemm_t += pa_wtime() - t0;
    /* ref: ecg.c:521-523 copies P -> P_prev, AP -> AP_prev, Z -> P; here the buffers trade roles */
    t0 = pa_wtime();
    double* oldPp = p->Pp; double* oldAPp = p->APp;
    p->Pp = p->P; p->P = p->Z; p->Z = oldPp;
    p->APp = p->AP; p->AP = oldAPp;
    refresh_shells(ecg, p);
    ecg->copy_t += pa_wtime() - t0;
    *rci_request = 0;
  }
  return 0;
}

/* Cholesky with complete (diagonal) pivoting of the symmetric n x n matrix A (column-major, full storage):
 * A[piv, piv] = U^T U, stops at the first pivot <= tol; tol < 0 selects LAPACK's default n * eps * max(diag)
 * (stands in for LAPACKE_dpstrf('U'), ref: ecg.c:375).  U is returned in the upper triangle of A. */
int pa_h_pivoted_chol(int n, double* A, int lda, int* piv, double tol) {
  double dmax = 0.0;
  for (int i = 0; i < n; ++i) { piv[i] = i; if (A[i + (size_t)lda * i] > dmax) dmax = A[i + (size_t)lda * i]; }
  if (tol < 0.0) tol = n * 1.1102230246251565e-16 * dmax;
  for (int j = 0; j < n; ++j) {
    int p = j;
    for (int i = j + 1; i < n; ++i) if (A[i + (size_t)lda * i] > A[p + (size_t)lda * p]) p = i;
    if (!(A[p + (size_t)lda * p] > tol)) return j;
    if (p != j) {  /* symmetric swap of rows/columns j and p */
      for (int k = 0; k < n; ++k) { double tmp = A[j + (size_t)lda * k]; A[j + (size_t)lda * k] = A[p + (size_t)lda * k]; A[p + (size_t)lda * k] = tmp; }
      for (int k = 0; k < n; ++k) { double tmp = A[k + (size_t)lda * j]; A[k + (size_t)lda * j] = A[k + (size_t)lda * p]; A[k + (size_t)lda * p] = tmp; }
      int ti = piv[j]; piv[j] = piv[p]; piv[p] = ti;
    }
    const double d = sqrt(A[j + (size_t)lda * j]);
    A[j + (size_t)lda * j] = d;
    for (int c = j + 1; c < n; ++c) A[j + (size_t)lda * c] /= d;
    for (int c = j + 1; c < n; ++c)      /* trailing update (full symmetric storage) */
      for (int r = j + 1; r < n; ++r) A[r + (size_t)lda * c] -= A[j + (size_t)lda * r] * A[j + (size_t)lda * c];
  }
  return n;
}

/* Orthomin with ADAPT_BS (ref: ecg.c:360-393): after P <- Z the T new directions are orthonormalised by a rank-revealing
 * Cholesky QR: C = P^T P (all-reduced), C[piv, piv] = U^T U by dpstrf, which stops at the first pivot <= T eps max(diag):
 * `rank` columns.  P <- P[:, piv[:rank]] U^-1 (dlapmt + dtrsm, ecg.c:380-385), block size <- rank (ecg.c:391).
 * Here W = the T x T matrix with W[piv[k], j] = (U^-1)[k, j] (k <= j < rank) and zero columns from `rank` on, P <- P W in one
 * pass: the dropped directions become zero columns, every block kernel keeps running at the full width T (exact), SpMM
 * and block-Jacobi run on `rank` columns because the shells say so.
 * The reference is only consistent up to its first rank drop: afterwards it keeps forming P^T P with the stale column
 * count of P's info (ecg.c:366 after the copy of nrhs columns at :357) and factors an nrhs x nrhs array that is partly
 * stale.  This code continues with the algorithm the reduction stands for: every iteration orthonormalises all T new
 * directions Z - P beta again, the rank may shrink further (or recover), bs = rank. */
static void omin_rrqr(preAlps_ECG_t* ecg, ecg_priv_t* p) {
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, T = p->t, ld = p->ld;
  double t0 = pa_wtime();
  pa_cuda_check(pcu_gram2(c, m, T, p->P, ld, p->P, ld, sm_G(p), NULL, 0, NULL, 0, NULL), "pcu_gram2");
  ecg->gemm_t += pa_wtime() - t0;
  pa_allreduce_dev(sm_G(p), T * T, &ecg->comm_t);
  double C[32 * 32], Ui[32 * 32], W[32 * 32];
  int piv[32];
  pa_cuda_check(pcu_d2h(c, C, sm_G(p), sizeof(double) * (size_t)T * T), "pcu_d2h");
  for (int j = 0; j < T; ++j) for (int i = j + 1; i < T; ++i) C[i + (size_t)T * j] = C[j + (size_t)T * i];  /* 'U' triangle */
  t0 = pa_wtime();
  const int rank = pa_h_pivoted_chol(T, C, T, piv, -1.0);
  ecg->pstrf_t += pa_wtime() - t0;
  for (int k = 0; k < T; ++k) ecg->iwork[k] = piv[k] + 1;  /* dpstrf's 1-based pivots, where the reference leaves them */
  t0 = pa_wtime();
  for (int i = 0; i < T * T; ++i) W[i] = 0.0;
  if (rank > 0) {
    pa_h_triu_inv(rank, C, T, Ui, T);
    for (int j = 0; j < rank; ++j) for (int k = 0; k <= j; ++k) W[piv[k] + (size_t)T * j] = Ui[k + (size_t)T * j];
  }
  pa_cuda_check(pcu_h2d(c, sm_U(p), W, sizeof(double) * (size_t)T * T), "pcu_h2d");
  pa_cuda_check(pcu_transform_update(c, m, T, sm_U(p), NULL, p->P, ld, NULL, 0, NULL, 0, NULL, 0, NULL), "pcu_transform_update");
  ecg->lapmt_t += pa_wtime() - t0;
  ecg->bs = rank;
  if (rank < T) p->fixed = 1;  /* from here on the descent step takes the padded path below */
}

/* The "rci_request == 0" half of an Orthomin iteration once directions have been dropped (bs < T): P has zero columns
 * from bs on, AP = A P was formed on bs columns only (the rest of its buffer is stale).  Everything at the full width T:
 *   G = AP^T P, Gpr = P^T R (only G[:bs, :bs] and Gpr[:bs, :] are read); host: U = chol(G[:bs, :bs]) (ref: ecg.c:318-322),
 *   W = diag(U^-1, 0), alpha = U^-T Gpr[:bs, :];  P <- P W, AP <- AP W (which also clears the stale columns);
 *   X += P [alpha; 0], R -= AP [alpha; 0]   (ref: ecg.c:324-341). */
static void omin_reduced_half_step(preAlps_ECG_t* ecg, ecg_priv_t* p) {
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, T = p->t, ld = p->ld, bs = ecg->bs;
  double t0 = pa_wtime();
  pa_cuda_check(pcu_gram2(c, m, T, p->AP, ld, p->P, ld, sm_G(p), p->P, ld, p->R, ld, sm_Gpr(p)), "pcu_gram2");
  ecg->gemm_t += pa_wtime() - t0;
  pa_allreduce_dev(sm_G(p), 2 * T * T, &ecg->comm_t);
  double G[2 * 32 * 32], Ui[32 * 32], W[32 * 32], Af[32 * 32], Wx[32 * 32];
  pa_cuda_check(pcu_d2h(c, G, sm_G(p), sizeof(double) * 2 * (size_t)T * T), "pcu_d2h");
  const double* Gpr = G + (size_t)T * T;
  t0 = pa_wtime();
  if (pa_h_chol_upper(bs, G, T) != 0) CPLM_Abort("ACHQR: dpotrf:\n ERROR: P^tAP is not spd!");  /* ref: ecg.c:320-322 */
  ecg->potrf_t += pa_wtime() - t0;
  t0 = pa_wtime();
  pa_h_triu_inv(bs, G, T, Ui, bs);
  for (int i = 0; i < T * T; ++i) { W[i] = 0.0; Af[i] = 0.0; }
  for (int j = 0; j < bs; ++j) for (int i = 0; i <= j; ++i) W[i + (size_t)T * j] = Ui[i + (size_t)bs * j];
  for (int j = 0; j < T; ++j)
    for (int i = 0; i < bs; ++i) {
      double v = 0.0;
      for (int k = 0; k <= i; ++k) v += Ui[k + (size_t)bs * i] * Gpr[k + (size_t)T * j];
      Af[i + (size_t)T * j] = v;
    }
  small_matmul(T, W, Af, Wx);
  pa_cuda_check(pcu_h2d(c, sm_U(p), W, sizeof(double) * (size_t)T * T), "pcu_h2d");
  pa_cuda_check(pcu_h2d(c, sm_alpha(p), Wx, sizeof(double) * (size_t)T * T), "pcu_h2d");
  /* P <- P W, AP <- AP W (which also clears the stale columns of AP), X += P_old W Af, R -= AP_old W Af: one pass */
  pa_cuda_check(pcu_transform_update(c, m, T, sm_U(p), sm_alpha(p), p->P, ld, p->AP, ld, p->X, ld, p->R, ld, sm_rr(p)),
                "pcu_transform_update");
  ecg->trsm_t += pa_wtime() - t0;
  refresh_shells(ecg, p);
  p->have_rr = 1;
  ecg->iter++;
  p->iter_since_reset++;
}

int _preAlps_ECGIterateOmin(preAlps_ECG_t* ecg, int* rci_request) {
  ecg_priv_t* p = priv_of(ecg);
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, ld = p->ld;
  if (*rci_request == 0) {
    if (p->fixed) omin_reduced_half_step(ecg, p);
    else descent_half_step(ecg, p);
    *rci_request = 1;
  } else if (*rci_request == 1) {
    /* beta = AP^T Z ; Z -= P beta ; P <- Z   (ref: ecg.c:345-359); after a reduction at the full width: the dropped
     * columns of P and AP are zero, Z = M^-1 R always has T columns */
    const int t = p->fixed ? p->t : ecg->bs;
    double t0 = pa_wtime();
    pa_cuda_check(pcu_gram2(c, m, t, p->AP, ld, p->Z, ld, sm_beta1(p), NULL, 0, NULL, 0, NULL), "pcu_gram2");
    ecg->gemm_t += pa_wtime() - t0;
    pa_allreduce_dev(sm_beta1(p), p->t * p->t, &ecg->comm_t);
    t0 = pa_wtime();
    pa_cuda_check(pcu_update_z(c, m, t, p->Z, ld, p->P, ld, t, sm_beta1(p), NULL, 0, 0, NULL), "pcu_update_z");
    ecg->gemm_t += pa_wtime() - t0;
    double* oldP = p->P;
    p->P = p->Z; p->Z = oldP;
    if (p->adapt) omin_rrqr(ecg, p);
    refresh_shells(ecg, p);
    *rci_request = 0;
  }
  return 0;
}

/* ORTHODIR_FUSED with ADAPT_BS (ref: ecg.c:532-658, reduction at :593-641). Same slot layout and zero/identity
 * padding as adapt_half_step. The five Gram products are taken on the un-normalised blocks at the full width T and
 * reduced together; the t x t algebra (Cholesky, the scalings of ecg.c:580-587, SVD) runs on the host; with
 * W = U^-1 Q (Q = I when nothing is dropped):
 *   slot0 <- slot0 diag(W, I);  A slot0 likewise;
 *   Z <- Z W - slot0 [Q^T b1p Q; b1h Q] - slot1 [b2 Q]      (b1p/b1h: rows of beta for the live / the discarded
 *                                                            directions, b2: previous directions, all scaled as ecg.c:582,586)
 *   X += slot0 [Q^T alpha (first t1 rows); 0];  R -= A slot0 [..];  copies of ecg.c:651-653 on bs columns.
 * Until the first reduction an iteration that drops nothing takes the NO_BS_RED path. Returns 1 if that path
 * must be taken by the caller. */
static int fused_adapt_step(preAlps_ECG_t* ecg, ecg_priv_t* p, const double* H) {
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, T = p->t, ld = p->ld, bs = ecg->bs;
  const double* alpha_f = H;
  const double* beta1_f = H + (size_t)T * T;
  const double* beta2_f = H + 2 * (size_t)T * T;
  double mu[32 * 32], Ui[32 * 32], alpha[32 * 32], Q[32 * 32], rows[32 * 32], sv[32];
  for (int i = 0; i < T * T; ++i) mu[i] = H[3 * (size_t)T * T + i];
  double t0 = pa_wtime();
  pa_h_chol_upper(bs, mu, T);
  ecg->potrf_t += pa_wtime() - t0;
  pa_h_triu_inv(bs, mu, T, Ui, bs);
  for (int j = 0; j < T; ++j)
    for (int i = 0; i < bs; ++i) {
      double v = 0.0;
      for (int k = 0; k <= i; ++k) v += Ui[k + (size_t)bs * i] * alpha_f[k + (size_t)T * j];
      alpha[i + (size_t)bs * j] = v;
    }
  t0 = pa_wtime();
  pa_h_left_svd(bs, T, alpha, bs, sv, Q, rows);
  ecg->gesvd_t += pa_wtime() - t0;
  const double cut = ecg->tol * ecg->normb / sqrt((double)T);
  int t1 = 0;
  for (int i = 0; i < bs; ++i) { if (sv[i] > cut) t1++; else break; }
  const int reduce = (t1 > 0 && t1 < T && t1 < bs);
  if (!reduce && !p->fixed) return 1;
  if (!reduce) { for (int i = 0; i < bs * bs; ++i) Q[i] = 0.0; for (int i = 0; i < bs; ++i) Q[i + (size_t)bs * i] = 1.0; }
  /* B1 = beta1[:, :bs] U^-1 (all T rows), first bs rows also U^-T from the left; B2 = beta2[:, :bs] U^-1 */
  double B1[32 * 32], B2[32 * 32], tmp[32 * 32];
  for (int j = 0; j < bs; ++j)
    for (int i = 0; i < T; ++i) {
      double v1 = 0.0, v2 = 0.0;
      for (int k = 0; k <= j; ++k) { v1 += beta1_f[i + (size_t)T * k] * Ui[k + (size_t)bs * j]; v2 += beta2_f[i + (size_t)T * k] * Ui[k + (size_t)bs * j]; }
      B1[i + (size_t)T * j] = v1; B2[i + (size_t)T * j] = v2;
    }
  for (int j = 0; j < bs; ++j) {
    for (int i = 0; i < bs; ++i) {
      double v = 0.0;
      for (int k = 0; k <= i; ++k) v += Ui[k + (size_t)bs * i] * B1[k + (size_t)T * j];
      tmp[i] = v;
    }
    for (int i = 0; i < bs; ++i) B1[i + (size_t)T * j] = tmp[i];
  }
  /* W = diag(U^-1 Q, I); Bf1 = [Q^T B1p Q; B1h Q]; Bf2 = B2 Q; Af = [Q^T alpha (keep rows); 0]; Wx = W Af -- T x T */
  double W[32 * 32], Bf1[32 * 32], Bf2[32 * 32], Af[32 * 32], Wx[32 * 32];
  for (int i = 0; i < T * T; ++i) { W[i] = 0.0; Bf1[i] = 0.0; Bf2[i] = 0.0; Af[i] = 0.0; }
  for (int j = 0; j < T; ++j) W[j + (size_t)T * j] = 1.0;
  for (int j = 0; j < bs; ++j)
    for (int i = 0; i < bs; ++i) {
      double v = 0.0;
      for (int k = i; k < bs; ++k) v += Ui[i + (size_t)bs * k] * Q[k + (size_t)bs * j];
      W[i + (size_t)T * j] = v;
    }
  for (int j = 0; j < bs; ++j)       /* tmp2 = B1 Q and B2 Q (T x bs) */
    for (int i = 0; i < T; ++i) {
      double v1 = 0.0, v2 = 0.0;
      for (int k = 0; k < bs; ++k) { v1 += B1[i + (size_t)T * k] * Q[k + (size_t)bs * j]; v2 += B2[i + (size_t)T * k] * Q[k + (size_t)bs * j]; }
      Bf1[i + (size_t)T * j] = v1; Bf2[i + (size_t)T * j] = v2;
    }
  for (int j = 0; j < bs; ++j) {     /* first bs rows of Bf1: Q^T from the left */
    for (int i = 0; i < bs; ++i) {
      double v = 0.0;
      for (int k = 0; k < bs; ++k) v += Q[k + (size_t)bs * i] * Bf1[k + (size_t)T * j];
      tmp[i] = v;
    }
    for (int i = 0; i < bs; ++i) Bf1[i + (size_t)T * j] = tmp[i];
  }
  const int keep = reduce ? t1 : bs;
  for (int j = 0; j < T; ++j)
    for (int i = 0; i < keep; ++i) Af[i + (size_t)T * j] = reduce ? rows[(size_t)i * T + j] : alpha[i + (size_t)bs * j];
  small_matmul(T, W, Af, Wx);
  pa_cuda_check(pcu_h2d(c, fu_mu(p), W, sizeof(double) * (size_t)T * T), "pcu_h2d");
  pa_cuda_check(pcu_h2d(c, fu_beta1(p), Bf1, sizeof(double) * (size_t)T * T), "pcu_h2d");
  pa_cuda_check(pcu_h2d(c, fu_beta2(p), Bf2, sizeof(double) * (size_t)T * T), "pcu_h2d");
  pa_cuda_check(pcu_h2d(c, fu_alpha(p), Wx, sizeof(double) * (size_t)T * T), "pcu_h2d");
  t0 = pa_wtime();
  /* slot0 <- slot0 W, A slot0 likewise, X += slot0_old W Af, R -= A slot0_old W Af, ||R||^2: one pass; then Z <- Z W in place */
  pa_cuda_check(pcu_transform_update(c, m, T, fu_mu(p), fu_alpha(p), p->P, ld, p->AP, ld, p->X, ld, p->R, ld, sm_rr(p)),
                "pcu_transform_update");
  pa_cuda_check(pcu_transform_update(c, m, T, fu_mu(p), NULL, p->Z, ld, NULL, 0, NULL, 0, NULL, 0, NULL), "pcu_transform_update");
  ecg->ormqr_t += pa_wtime() - t0;
  t0 = pa_wtime();
  pa_cuda_check(pcu_update_z(c, m, T, p->Z, ld, p->P, ld, T, fu_beta1(p), p->Pp, ld, T, fu_beta2(p)), "pcu_update_z");
  if (keep < T) pa_cuda_check(pcu_zero_cols(c, m, T - keep, p->Z + keep, ld), "pcu_zero_cols");
  ecg->gemm_t += pa_wtime() - t0;
  if (reduce) {
    ecg->bs = t1;
    ecg->kbs = bs + T;
    p->tprev = bs;
    p->fixed = 1;
    pa_cuda_check(pcu_zero_cols(c, m, T - bs, p->Pp + bs, ld), "pcu_zero_cols");
    pa_cuda_check(pcu_zero_cols(c, m, T - bs, p->APp + bs, ld), "pcu_zero_cols");
  }
  ecg->iter++;
  p->have_rr = 1;
  t0 = pa_wtime();
  pa_cuda_check(pcu_copy_cols(c, m, ecg->bs, p->Pp, ld, p->P, ld), "pcu_copy_cols");
  pa_cuda_check(pcu_copy_cols(c, m, ecg->bs, p->APp, ld, p->AP, ld), "pcu_copy_cols");
  pa_cuda_check(pcu_copy_cols(c, m, ecg->bs, p->P, ld, p->Z, ld), "pcu_copy_cols");
  ecg->copy_t += pa_wtime() - t0;
  refresh_shells(ecg, p);
  return 0;
}

/* ref: ecg.c:532-658.  One all-reduce per iteration: alpha = P^T R, beta = [AP^T Z; APprev^T Z], mu = AP^T P and
 * R^T R are formed from the un-normalised blocks, reduced together, and the normalisation by U = chol(mu) is
 * applied afterwards (P, AP, Z <- . U^-1; alpha <- U^-T alpha; beta1 <- U^-T beta1 U^-1; beta2 <- beta2 U^-1).
 * The residual test therefore lags one iteration (ecg.c:566-574); rci_request = 1 means converged. */
int _preAlps_ECGIterateOdirFused(preAlps_ECG_t* ecg, int* rci_request) {
  ecg_priv_t* p = priv_of(ecg);
  pcu_ctx* c = pa_g.ctx;
  const int m = p->m, t = p->adapt ? p->t : ecg->bs, ld = p->ld;  /* ADAPT_BS: full width, zero-padded blocks */
  double t0 = pa_wtime();
  pa_cuda_check(pcu_gram2(c, m, t, p->P, ld, p->R, ld, fu_alpha(p), p->AP, ld, p->P, ld, fu_mu(p)), "pcu_gram2");
  pa_cuda_check(pcu_gram2(c, m, t, p->AP, ld, p->Z, ld, fu_beta1(p), p->APp, ld, p->Z, ld, fu_beta2(p)), "pcu_gram2");
  pa_cuda_check(pcu_fro2(c, m, t, p->R, ld, fu_rr(p)), "pcu_fro2");
  ecg->gemm_t += pa_wtime() - t0;
  pa_allreduce_dev(fu_alpha(p), 4 * p->t * p->t + 1, &ecg->comm_t);
  double rr = 0.0;
  double Hs[4 * 32 * 32 + 1];
  if (p->adapt) {
    pa_cuda_check(pcu_d2h(c, Hs, fu_alpha(p), sizeof(double) * (4 * (size_t)p->t * p->t + 1)), "pcu_d2h");
    rr = Hs[4 * (size_t)p->t * p->t];
  } else {
    pa_cuda_check(pcu_d2h(c, &rr, fu_rr(p), sizeof(double)), "pcu_d2h");
  }
  ecg->res = sqrt(rr);
  if (ecg->res < ecg->tol * ecg->normb || ecg->iter > ecg->maxIter) *rci_request = 1;
  else *rci_request = 0;
  if (p->adapt && !fused_adapt_step(ecg, p, Hs)) return 0;
  t0 = pa_wtime();
  pa_cuda_check(pcu_fused_small(c, t, fu_mu(p), fu_beta1(p), fu_beta2(p), NULL, p->status_dev), "pcu_fused_small");
  pa_cuda_check(pcu_right_solve(c, m, t, fu_mu(p), p->Z, ld), "pcu_right_solve");
  /* P,AP <- .U^-1 ; alpha = U^-T alpha ; X += P alpha ; R -= AP alpha */
  pa_cuda_check(pcu_ortho_update(c, m, t, fu_mu(p), fu_alpha(p), p->P, ld, p->AP, ld, p->X, ld, p->R, ld, NULL, fu_aout(p),
                                 sm_rr(p), p->status_dev), "pcu_ortho_update");
  ecg->trsm_t += pa_wtime() - t0;
  t0 = pa_wtime();
  pa_cuda_check(pcu_update_z(c, m, t, p->Z, ld, p->P, ld, t, fu_beta1(p), p->Pp, ld, t, fu_beta2(p)), "pcu_update_z");
  ecg->gemm_t += pa_wtime() - t0;
  ecg->iter++;
  p->have_rr = 1;
  double* oldPp = p->Pp; double* oldAPp = p->APp;
  p->Pp = p->P; p->P = p->Z; p->Z = oldPp;
  p->APp = p->AP; p->AP = oldAPp;
  refresh_shells(ecg, p);
  return 0;
}

int preAlps_ECGIterate(preAlps_ECG_t* ecg, int* rci_request) {
  const double t0 = pa_wtime();
  if (ecg->ortho_alg == ORTHOMIN) _preAlps_ECGIterateOmin(ecg, rci_request);
  else if (ecg->ortho_alg == ORTHODIR) _preAlps_ECGIterateOdir(ecg, rci_request);
  else _preAlps_ECGIterateOdirFused(ecg, rci_request);
  ecg->tot_t += pa_wtime() - t0;
  return 0;
}

int preAlps_ECGStoppingCriterion(preAlps_ECG_t* ecg, int* stop) {
  const double t0 = pa_wtime();
  ecg_priv_t* p = priv_of(ecg);
  pcu_ctx* c = pa_g.ctx;
  if (!stop) CPLM_Abort(" wrong test 'stop != NULL'");
  if (!p->have_rr) {  /* R changed outside Iterate (or first call): one pass over R */
    const double tg = pa_wtime();
    pa_cuda_check(pcu_fro2(c, p->m, p->t, p->R, p->ld, sm_rr(p)), "pcu_fro2");
    ecg->gemm_t += pa_wtime() - tg;
    p->have_rr = 1;
  }
  double* rr_glob = sm_rr(p) + 1;
  pa_cuda_check(pcu_d2d(c, rr_glob, sm_rr(p), sizeof(double)), "pcu_d2d");
  pa_allreduce_dev(rr_glob, 1, &ecg->comm_t);
  double rr = 0.0;
  pa_cuda_check(pcu_d2h(c, &rr, rr_glob, sizeof(double)), "pcu_d2h");
  ecg->res = sqrt(rr);
  /* ref: ecg.c:264 */
  if (ecg->res > ecg->normb * ecg->tol && ecg->iter < ecg->maxIter && ecg->bs > 0) *stop = 0;
  else *stop = 1;
  ecg->tot_t += pa_wtime() - t0;
  return 0;
}

int _preAlps_ECGWrapUp(preAlps_ECG_t* ecg, double* solution) {
  ecg_priv_t* p = priv_of(ecg);
  pcu_ctx* c = pa_g.ctx;
  pa_cuda_check(pcu_sum_columns(c, p->m, p->t, p->X, p->ld, p->rhs_dev), "pcu_sum_columns");
  pa_cuda_check(pcu_d2h(c, solution, p->rhs_dev, sizeof(double) * (size_t)p->m), "pcu_d2h");
  return 0;
}

void _preAlps_ECGFree(preAlps_ECG_t* ecg) {
  ecg_priv_t* p = priv_of(ecg);
  pcu_ctx* c = pa_g.ctx;
  free(ecg->X); free(ecg->R); free(ecg->V); free(ecg->AV); free(ecg->alpha); free(ecg->beta); free(ecg->Z);
  free(ecg->P); free(ecg->AP);
  ecg->X = ecg->R = ecg->V = ecg->AV = ecg->alpha = ecg->beta = ecg->Z = ecg->P = ecg->AP = NULL;
  if (c && ecg->work) pcu_free(c, ecg->work);
  ecg->work = NULL;
  free(ecg->iwork); ecg->iwork = NULL;
  if (p) {
    if (c) { pcu_free(c, p->rhs_dev); pcu_free(c, p->col_of_row); pcu_free(c, p->status_dev); }
    for (int i = 0; i < MAX_SOLVERS; ++i) if (g_priv[i] == p) g_priv[i] = NULL;
    free(p);
  }
}

int preAlps_ECGFinalize(preAlps_ECG_t* ecg, double* solution) {
  const int ierr = _preAlps_ECGWrapUp(ecg, solution);
  _preAlps_ECGFree(ecg);
  return ierr;
}

static void print_info(const char* name, const CPLM_Mat_Dense_t* A) {  /* ref: cplm_matdense.c:315-325 */
  if (!A) { printf("%s\n(released)\n", name); return; }
  printf("%s\nDense Matrix %dx%d\tLocal Data: %dx%d\t%s\tLeading Dimension Array %d\tNumber of elements of array %d\n", name,
         A->info.M, A->info.N, A->info.m, A->info.n, (A->info.stor_type == ROW_MAJOR) ? "Storage: Row major\n" : "Storage: Col major\n",
         A->info.lda, A->info.nval);
}

void preAlps_ECGPrint(preAlps_ECG_t* ecg, int verbosity) {
  int rank = pa_g.rank;
  printf("[%d] prints ECG_t...\n", rank);
  printf("=== Summary ===\n");
  printf("\titer: %d\n\tres : %e\n\tbs  : %1d\n", ecg->iter, ecg->res, ecg->bs);
  printf("=== Timings ===\n");
  printf("\ttot_t  : %e s\n", ecg->tot_t);
  printf("\tcomm_t : %e s\n", ecg->comm_t);
  printf("\ttrsm_t : %e s\n", ecg->trsm_t);
  printf("\tgemm_t : %e s\n", ecg->gemm_t);
  printf("\tpotrf_t: %e s\n", ecg->potrf_t);
  printf("\tpstrf_t: %e s\n", ecg->pstrf_t);
  printf("\tlapmt_t: %e s\n", ecg->lapmt_t);
  printf("\tgesvd_t: %e s\n", ecg->gesvd_t);
  printf("\tgeqrf_t: %e s\n", ecg->geqrf_t);
  printf("\tormqr_t: %e s\n", ecg->ormqr_t);
  printf("\tcopy_t : %e s\n", ecg->copy_t);
  if (verbosity > 1) {  /* ref: ecg.c:713-726; after Finalize the blocks are released (the reference reads freed memory there) */
    printf("=== Memory consumption ===\n");
    print_info("X", ecg->X); print_info("R", ecg->R); print_info("V", ecg->V); print_info("AV", ecg->AV);
    print_info("P", ecg->P); print_info("AP", ecg->AP); print_info("Z", ecg->Z);
    print_info("alpha", ecg->alpha); print_info("beta", ecg->beta);
    printf("\n");
  }
  printf("[%d] ends printing ECG_t!\n", rank);
}
