/*
 * pa_driver.c -- the driver-side pieces of examples/test_ecg_prealps_op.c as library calls
 * (so that tests and bench.py run exactly the reference's sequence per virtual subdomain):
 * the srand(0) right-hand side (ref: test_ecg_prealps_op.c:172-184), the RCI loop
 * (ref: :203-223), plus the timing regions bench.py reports.
 */
#include "pa_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int preAlps_b200_DriverRhs(double* rhs) {
  pa_state_t* g = &pa_g;
  if (!g->built) CPLM_Abort("preAlps_b200_DriverRhs called before the operator was built");
  /* every reference rank draws its m values after srand(0): identical streams per subdomain */
  double nb = 0.0;
  double* parts = (double*)pa_xcalloc((size_t)(g->s_hi - g->s_lo), sizeof(double));
  for (int s = g->s_lo; s < g->s_hi; ++s) {
    const int r0 = g->rowPos[s] - g->g0, r1 = g->rowPos[s + 1] - g->g0;
    srand(0);
    double part = 0.0;
    for (int i = r0; i < r1; ++i) {
      rhs[i] = ((double)rand() / (double)RAND_MAX);
      part += pow(rhs[i], 2);
    }
    parts[s - g->s_lo] = part;
    nb += part;
  }
  if (g->nproc > 1) {
    if (g->xport == PA_XPORT_MPI) MPI_Allreduce(MPI_IN_PLACE, &nb, 1, MPI_DOUBLE, MPI_SUM, g->comm);
    else nb = pa_sum_over_subdomains(parts);  /* subdomain order: the same bits on 1, 2, 4 or 8 GPUs */
  }
  free(parts);
  nb = sqrt(nb);
  /* the reference's loop starts at i = 1: the first entry of every rank stays unscaled (:183) */
  for (int s = g->s_lo; s < g->s_hi; ++s) {
    const int r0 = g->rowPos[s] - g->g0, r1 = g->rowPos[s + 1] - g->g0;
    for (int i = r0 + 1; i < r1; ++i) rhs[i] /= nb;
  }
  return 0;
}

static void fill_ecg(preAlps_ECG_t* ecg, int enlFac, double tol, int maxIter, int ortho_alg, int bs_red) {
  memset(ecg, 0, sizeof *ecg);
  ecg->comm = pa_g.comm;
  ecg->globPbSize = pa_g.M;
  ecg->locPbSize = pa_g.m;
  ecg->maxIter = maxIter;
  ecg->enlFac = enlFac;
  ecg->tol = tol;
  ecg->ortho_alg = (ortho_alg == 0 ? ORTHODIR : (ortho_alg == 1 ? ORTHOMIN : ORTHODIR_FUSED));
  ecg->bs_red = (bs_red == 0 ? NO_BS_RED : ADAPT_BS);
}

static int g_bs_hist[4096];
static int g_bs_nhist = 0;

int preAlps_b200_LastBlockSizes(int* out, int max) {
  const int n = g_bs_nhist < max ? g_bs_nhist : max;
  for (int i = 0; i < n; ++i) out[i] = g_bs_hist[i];
  return g_bs_nhist;
}

int preAlps_b200_Solve(int enlFac, double tol, int maxIter, int ortho_alg, int bs_red, double* rhs, double* sol,
                       double* res_hist, int max_hist, preAlps_b200_SolveInfo* info) {
  pa_state_t* g = &pa_g;
  if (!g->built || !g->bj) CPLM_Abort("preAlps_b200_Solve needs an operator and a block-Jacobi preconditioner");
  pcu_ctx* c = pa_ctx();
  preAlps_ECG_t ecg;
  fill_ecg(&ecg, enlFac, tol, maxIter, ortho_alg, bs_red);
  int rci = 0, stop = 0, nh = 0;
  g_bs_nhist = 0;
  pa_cuda_check(pcu_sync(c), "pcu_sync");
  const double t0 = pa_wtime();
  pcu_timer_start(c, 0);
  preAlps_ECGInitialize(&ecg, rhs, &rci);
  const double t_init = pa_wtime();
  preAlps_BlockJacobiApply(ecg.R, ecg.P);
  if (ecg.ortho_alg == ORTHODIR_FUSED) {
    /* loop of the reference's examples/test_ecg_bench_fused.c:245-259 */
    while (rci != 1) {
      preAlps_BlockOperator(ecg.P, ecg.AP);
      preAlps_BlockJacobiApply(ecg.AP, ecg.Z);
      preAlps_ECGIterate(&ecg, &rci);
      if (res_hist && nh < max_hist) res_hist[nh] = ecg.res;
      if (g_bs_nhist < 4096) g_bs_hist[g_bs_nhist++] = ecg.bs;
      ++nh;
    }
    stop = 1;
  } else
  preAlps_BlockOperator(ecg.P, ecg.AP);
  while (stop != 1) {
    preAlps_ECGIterate(&ecg, &rci);
    if (rci == 0) {
      preAlps_BlockOperator(ecg.P, ecg.AP);
    } else if (rci == 1) {
      preAlps_ECGStoppingCriterion(&ecg, &stop);
      if (res_hist && nh < max_hist) res_hist[nh] = ecg.res;
      if (g_bs_nhist < 4096) g_bs_hist[g_bs_nhist++] = ecg.bs;
      ++nh;
      if (stop == 1) break;
      if (ecg.ortho_alg == ORTHOMIN) preAlps_BlockJacobiApply(ecg.R, ecg.Z);
      else preAlps_BlockJacobiApply(ecg.AP, ecg.Z);
    }
  }
  const int iter = ecg.iter;
  const double res = ecg.res, normb = ecg.normb;
  const double t_loop = pa_wtime();
  preAlps_ECGFinalize(&ecg, sol);
  pcu_timer_stop(c, 0);
  const double t1 = pa_wtime();
  if (getenv("PREALPS_B200_TIMING"))
    fprintf(stderr, "[prealps_b200] solve: init %.4f s, loop %.4f s (%d iterations), finalize %.4f s\n", t_init - t0,
            t_loop - t_init, iter, t1 - t_loop);
  float ms = 0.f;
  pcu_timer_elapsed_ms(c, 0, &ms);
  if (info) {
    info->iter = iter; info->res = res; info->normb = normb; info->t_solve = t1 - t0; info->t_dev_ms = ms;
    info->nhist = nh; info->stopped = stop;
    /* true residual with the library's own operator (host vectors are staged through HBM) */
    CPLM_Mat_Dense_t xs = CPLM_MatDenseNULL(), ax = CPLM_MatDenseNULL();
    CPLM_MatDenseSetInfo(&xs, g->M, 1, g->m, 1, COL_MAJOR);
    CPLM_MatDenseSetInfo(&ax, g->M, 1, g->m, 1, COL_MAJOR);
    xs.val = sol;
    ax.val = (double*)pa_xcalloc((size_t)(g->m > 0 ? g->m : 1), sizeof(double));
    preAlps_BlockOperator(&xs, &ax);
    double rr[2] = {0.0, 0.0};
    for (int i = 0; i < g->m; ++i) { const double d = rhs[i] - ax.val[i]; rr[0] += d * d; rr[1] += rhs[i] * rhs[i]; }
    free(ax.val);
    if (g->nproc > 1) {
      if (g->xport == PA_XPORT_MPI) MPI_Allreduce(MPI_IN_PLACE, rr, 2, MPI_DOUBLE, MPI_SUM, g->comm);
      else {
        double* d = (double*)pcu_malloc(c, 2 * sizeof(double));
        pa_cuda_check(pcu_h2d(c, d, rr, 2 * sizeof(double)), "pcu_h2d");
        pa_cuda_check(pcu_allreduce_sum(c, d, 2), "pcu_allreduce_sum");
        pa_cuda_check(pcu_d2h(c, rr, d, 2 * sizeof(double)), "pcu_d2h");
        pcu_free(c, d);
      }
    }
    info->true_relres = sqrt(rr[0]) / sqrt(rr[1]);
  }
  return 0;
}

/* `warmup` then `steps` ECG iterations; a solve that converges (or reaches 200 iterations) is wrapped
 * up and restarted with the same right-hand side, its re-initialisation being part of the region. */
int preAlps_b200_BenchIterations(int enlFac, double tol, int ortho_alg, double* rhs, int warmup, int steps,
                                 float* ms_out, long long* launches_out) {
  pa_state_t* g = &pa_g;
  if (!g->built || !g->bj) CPLM_Abort("preAlps_b200_BenchIterations needs an operator and a block-Jacobi preconditioner");
  pcu_ctx* c = pa_ctx();
  preAlps_ECG_t ecg;
  double* sol = (double*)pa_xmalloc(sizeof(double) * (size_t)(g->m > 0 ? g->m : 1));
  int done = 0, live = 0, rci = 0, stop = 0;
  long long l0 = 0;
  const int total = warmup + steps;
  while (done < total) {
    if (!live) {
      fill_ecg(&ecg, enlFac, tol, 200, ortho_alg, 0);
      preAlps_ECGInitialize(&ecg, rhs, &rci);
      preAlps_BlockJacobiApply(ecg.R, ecg.P);
      preAlps_BlockOperator(ecg.P, ecg.AP);
      live = 1; stop = 0;
    }
    if (done == warmup) {
      pa_cuda_check(pcu_sync(c), "pcu_sync");
      l0 = pcu_launch_count(c);
      pcu_timer_start(c, 1);
    }
    /* one iteration = Iterate(0) + stopping test + block-Jacobi + Iterate(1) + SpMM */
    preAlps_ECGIterate(&ecg, &rci);
    preAlps_ECGStoppingCriterion(&ecg, &stop);
    ++done;
    if (stop == 1) {
      preAlps_ECGFinalize(&ecg, sol);
      live = 0;
      continue;
    }
    if (ecg.ortho_alg == ORTHOMIN) preAlps_BlockJacobiApply(ecg.R, ecg.Z);
    else preAlps_BlockJacobiApply(ecg.AP, ecg.Z);
    preAlps_ECGIterate(&ecg, &rci);
    preAlps_BlockOperator(ecg.P, ecg.AP);
  }
  pcu_timer_stop(c, 1);
  pa_cuda_check(pcu_timer_elapsed_ms(c, 1, ms_out), "pcu_timer_elapsed_ms");
  if (launches_out) *launches_out = pcu_launch_count(c) - l0;
  if (live) preAlps_ECGFinalize(&ecg, sol);
  free(sol);
  return 0;
}

int preAlps_b200_BenchKernel(int what, int t, int reps, int flush_l2, float* ms_out) {
  pa_state_t* g = &pa_g;
  pcu_ctx* c = pa_ctx();
  const int m = g->m;
  const int ld = (t % 2 == 0 || t == 1) ? t : t + 1;
  const size_t blk = (size_t)m * ld;
  double* pool = (double*)pcu_malloc(c, sizeof(double) * (7 * blk + 8 * (size_t)t * t + 64));
  if (!pool) CPLM_Abort("device allocation failed: %s", pcu_last_error());
  /* deterministic non-trivial content: srand(0) uniforms, like test_bench_spmm.c */
  double* h = (double*)pa_xmalloc(sizeof(double) * blk);
  srand(0);
  for (size_t i = 0; i < blk; ++i) h[i] = (double)rand() / (double)RAND_MAX;
  for (int k = 0; k < 7; ++k) pa_cuda_check(pcu_h2d(c, pool + k * blk, h, sizeof(double) * blk), "pcu_h2d");
  free(h);
  double *P = pool, *AP = pool + blk, *Z = pool + 2 * blk, *R = pool + 3 * blk, *X = pool + 4 * blk, *Pp = pool + 5 * blk,
         *APp = pool + 6 * blk, *sm = pool + 7 * blk;
  int* st = (int*)pcu_malloc(c, 16);
  /* median of the timed repetitions */
  float* samples = (float*)pa_xmalloc(sizeof(float) * (size_t)(reps > 0 ? reps : 1));
  for (int r = -2; r < reps; ++r) {  /* two untimed warm-up calls */
    if (flush_l2) pa_cuda_check(pcu_flush_l2(c), "pcu_flush_l2");
    pcu_timer_start(c, 2);
    if (what == 0) {
      if (g->nproc > 1 && g->xport == PA_XPORT_NCCL) {
        pa_cuda_check(pcu_spmm_apply_exchange(g->spmm, P, ld, AP, ld, t), "pcu_spmm_apply_exchange");
      } else {
        if (g->nproc > 1 && g->xport == PA_XPORT_NCCL) pa_cuda_check(pcu_spmm_halo_exchange(g->spmm, P, ld, t), "halo");
        pa_cuda_check(pcu_spmm_apply(g->spmm, P, ld, AP, ld, t), "pcu_spmm_apply");
      }
    } else if (what == 1) {
      pa_cuda_check(pcu_bj_apply(g->bj, AP, ld, Z, ld, t), "pcu_bj_apply");
    } else {
      /* G must stay SPD for the in-kernel Cholesky: use P^T P */
      pa_cuda_check(pcu_gram2(c, m, t, P, ld, P, ld, sm, P, ld, R, ld, sm + t * t), "pcu_gram2");
      pa_cuda_check(pcu_ortho_update(c, m, t, sm, sm + t * t, Pp, ld, APp, ld, X, ld, R, ld, sm + 2 * t * t, sm + 3 * t * t,
                                     sm + 6 * t * t, st), "pcu_ortho_update");
      pa_cuda_check(pcu_gram2(c, m, t, AP, ld, Z, ld, sm + 4 * t * t, APp, ld, Z, ld, sm + 5 * t * t), "pcu_gram2");
      pa_cuda_check(pcu_update_z(c, m, t, Z, ld, P, ld, t, sm + 4 * t * t, Pp, ld, t, sm + 5 * t * t), "pcu_update_z");
    }
    pcu_timer_stop(c, 2);
    float ms = 0.f;
    pa_cuda_check(pcu_timer_elapsed_ms(c, 2, &ms), "pcu_timer_elapsed_ms");
    if (r >= 0) samples[r] = ms;
  }
  for (int i = 1; i < reps; ++i) {  /* insertion sort */
    const float v = samples[i];
    int j = i - 1;
    for (; j >= 0 && samples[j] > v; --j) samples[j + 1] = samples[j];
    samples[j + 1] = v;
  }
  *ms_out = reps > 0 ? (reps % 2 ? samples[reps / 2] : 0.5f * (samples[reps / 2 - 1] + samples[reps / 2])) : 0.f;
  free(samples);
  pcu_free(c, pool);
  pcu_free(c, st);
  return 0;
}
