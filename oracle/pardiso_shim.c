/*
 * oracle/pardiso_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A plain-C sparse Cholesky standing in for MKL PARDISO exactly as the
 * preAlps reference drives it (/root/reference/utils/cplm_light/cplm_kernels.c:
 * 677-694 parameters, :741-783 phase 12, :790-853 phase 33, :700-735 phase -1):
 *   mtype = 2 (real SPD), upper-triangular CSR input, iparm[34] = 1 (0-based),
 *   iparm[1] = 2 (METIS nested-dissection ordering), column-major multi-RHS,
 *   iparm[5] = 1 => solution overwrites b.
 * Algorithm: METIS_NodeND fill-reducing ordering, elimination tree, up-looking
 * (row-by-row) numeric factorisation P A P^T = L L^T, then forward/backward
 * substitution over all right-hand sides at once.  Any exact Cholesky yields
 * the same M^{-1} up to rounding, so PARDISO's internal ordering need not be
 * matched.
 */
#include "shim/mkl.h"
#include "shim/metis.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int n;
  int64_t* Lp; /* column pointers, n+1 */
  int* Li;     /* row indices (diagonal first in each column) */
  double* Lx;
  int* perm;   /* perm[new] = old */
  int64_t lnz;
} chol_t;

static void chol_free(chol_t* f) {
  if (!f) return;
  free(f->Lp); free(f->Li); free(f->Lx); free(f->perm); free(f);
}

/* nonzero pattern of row k of L: walk the elimination tree from every entry
 * of column k of the (upper) matrix until a marked node is met */
static int ereach(const int64_t* Cp, const int* Ci, int k, const int* parent, int* s, int* w, int n) {
  int top = n;
  w[k] = k; /* mark k */
  for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
    int i = Ci[p];
    if (i > k) continue;
    int len = 0;
    for (; w[i] != k; i = parent[i]) { s[len++] = i; w[i] = k; }
    while (len > 0) s[--top] = s[--len];
  }
  return top;
}

static chol_t* chol_factor(int n, const int* ia, const int* ja, const double* a, int base, int* err) {
  *err = 0;
  chol_t* f = (chol_t*)calloc(1, sizeof(chol_t));
  f->n = n;
  f->perm = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int* iperm = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  const int64_t nnzU = (int64_t)ia[n] - base;

  /* ---- ordering: symmetric graph without the diagonal -> METIS_NodeND */
  {
    int64_t* deg = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    int64_t nedge = 0;
    for (int i = 0; i < n; ++i)
      for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
        int j = ja[p] - base;
        if (j != i) { deg[i + 1]++; deg[j + 1]++; nedge += 2; }
      }
    if (nedge == 0 || n < 8) {
      for (int i = 0; i < n; ++i) { f->perm[i] = i; iperm[i] = i; }
    } else {
      idx_t* xadj = (idx_t*)malloc(sizeof(idx_t) * ((size_t)n + 1));
      idx_t* adj = (idx_t*)malloc(sizeof(idx_t) * (size_t)nedge);
      xadj[0] = 0;
      for (int i = 0; i < n; ++i) xadj[i + 1] = xadj[i] + deg[i + 1];
      int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
      for (int i = 0; i < n; ++i) pos[i] = xadj[i];
      for (int i = 0; i < n; ++i)
        for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
          int j = ja[p] - base;
          if (j != i) { adj[pos[i]++] = j; adj[pos[j]++] = i; }
        }
      idx_t nv = n;
      idx_t* mp = (idx_t*)malloc(sizeof(idx_t) * (size_t)n);
      idx_t* mip = (idx_t*)malloc(sizeof(idx_t) * (size_t)n);
      int rc = METIS_NodeND(&nv, xadj, adj, NULL, NULL, mp, mip);
      if (rc != METIS_OK) { *err = -3; }
      for (int i = 0; i < n; ++i) { f->perm[i] = (int)mp[i]; iperm[i] = (int)mip[i]; }
      free(xadj); free(adj); free(pos); free(mp); free(mip);
    }
    free(deg);
    if (*err) { free(iperm); chol_free(f); return NULL; }
  }

  /* ---- C = upper triangle of P A P^T, compressed by columns */
  int64_t* Cp = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  int* Ci = (int*)malloc(sizeof(int) * (size_t)(nnzU > 0 ? nnzU : 1));
  double* Cx = (double*)malloc(sizeof(double) * (size_t)(nnzU > 0 ? nnzU : 1));
  for (int i = 0; i < n; ++i)
    for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
      int ni = iperm[i], nj = iperm[ja[p] - base];
      Cp[(ni > nj ? ni : nj) + 1]++;
    }
  for (int i = 0; i < n; ++i) Cp[i + 1] += Cp[i];
  {
    int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) pos[i] = Cp[i];
    for (int i = 0; i < n; ++i)
      for (int p = ia[i] - base; p < ia[i + 1] - base; ++p) {
        int ni = iperm[i], nj = iperm[ja[p] - base];
        int c = ni > nj ? ni : nj, r = ni > nj ? nj : ni;
        Ci[pos[c]] = r; Cx[pos[c]] = a[p]; pos[c]++;
      }
    free(pos);
  }

  /* ---- elimination tree (Liu, with ancestor path compression) */
  int* parent = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int* anc = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int k = 0; k < n; ++k) {
    parent[k] = -1; anc[k] = -1;
    for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
      int i = Ci[p];
      while (i != -1 && i < k) {
        int nx = anc[i];
        anc[i] = k;
        if (nx == -1) parent[i] = k;
        i = nx;
      }
    }
  }
  free(anc);

  /* ---- column counts by symbolic row reaches */
  int* s = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int* w = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int64_t* cnt = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  for (int i = 0; i < n; ++i) w[i] = -1;
  for (int k = 0; k < n; ++k) {
    int top = ereach(Cp, Ci, k, parent, s, w, n);
    for (int t = top; t < n; ++t) cnt[s[t]]++;
    cnt[k]++; /* diagonal */
  }
  f->Lp = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n + 1));
  f->Lp[0] = 0;
  for (int i = 0; i < n; ++i) f->Lp[i + 1] = f->Lp[i] + cnt[i];
  f->lnz = f->Lp[n];
  f->Li = (int*)malloc(sizeof(int) * (size_t)(f->lnz > 0 ? f->lnz : 1));
  f->Lx = (double*)malloc(sizeof(double) * (size_t)(f->lnz > 0 ? f->lnz : 1));
  if (!f->Li || !f->Lx) { *err = -2; goto done; }

  /* ---- numeric up-looking factorisation */
  {
    int64_t* c = cnt; /* reuse as "next free slot" per column */
    double* x = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    for (int i = 0; i < n; ++i) { c[i] = f->Lp[i]; w[i] = -1; }
    for (int k = 0; k < n && !*err; ++k) {
      int top = ereach(Cp, Ci, k, parent, s, w, n);
      x[k] = 0.0;
      for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p)
        if (Ci[p] <= k) x[Ci[p]] += Cx[p];
      double d = x[k];
      x[k] = 0.0;
      for (; top < n; ++top) {
        int i = s[top];
        double lki = x[i] / f->Lx[f->Lp[i]];
        x[i] = 0.0;
        for (int64_t p = f->Lp[i] + 1; p < c[i]; ++p) x[f->Li[p]] -= f->Lx[p] * lki;
        d -= lki * lki;
        int64_t q = c[i]++;
        f->Li[q] = k; f->Lx[q] = lki;
      }
      if (!(d > 0.0)) { *err = -4; break; } /* not positive definite */
      int64_t q = c[k]++;
      f->Li[q] = k; f->Lx[q] = sqrt(d);
    }
    free(x);
  }
done:
  free(Cp); free(Ci); free(Cx); free(parent); free(s); free(w); free(cnt); free(iperm);
  if (*err) { chol_free(f); return NULL; }
  return f;
}

static void chol_solve(const chol_t* f, int nrhs, const double* b, double* x) {
  const int n = f->n;
  double* y = (double*)malloc(sizeof(double) * (size_t)n * (size_t)(nrhs > 0 ? nrhs : 1));
  /* y (row-major n x nrhs) = P b */
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < nrhs; ++r) y[(size_t)i * nrhs + r] = b[(size_t)r * n + f->perm[i]];
  /* L y = y */
  for (int j = 0; j < n; ++j) {
    double* yj = y + (size_t)j * nrhs;
    const double dinv = 1.0 / f->Lx[f->Lp[j]];
    for (int r = 0; r < nrhs; ++r) yj[r] *= dinv;
    for (int64_t p = f->Lp[j] + 1; p < f->Lp[j + 1]; ++p) {
      double* yi = y + (size_t)f->Li[p] * nrhs;
      const double l = f->Lx[p];
      for (int r = 0; r < nrhs; ++r) yi[r] -= l * yj[r];
    }
  }
  /* L^T y = y */
  for (int j = n - 1; j >= 0; --j) {
    double* yj = y + (size_t)j * nrhs;
    for (int64_t p = f->Lp[j] + 1; p < f->Lp[j + 1]; ++p) {
      const double* yi = y + (size_t)f->Li[p] * nrhs;
      const double l = f->Lx[p];
      for (int r = 0; r < nrhs; ++r) yj[r] -= l * yi[r];
    }
    const double dinv = 1.0 / f->Lx[f->Lp[j]];
    for (int r = 0; r < nrhs; ++r) yj[r] *= dinv;
  }
  /* x = P^T y */
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < nrhs; ++r) x[(size_t)r * n + f->perm[i]] = y[(size_t)i * nrhs + r];
  free(y);
}

void pardisoinit(_MKL_DSS_HANDLE_t pt, const MKL_INT* mtype, MKL_INT* iparm) {
  (void)mtype;
  memset(pt, 0, 64 * sizeof(void*));
  memset(iparm, 0, 64 * sizeof(MKL_INT));
}

void pardiso(_MKL_DSS_HANDLE_t pt, const MKL_INT* maxfct, const MKL_INT* mnum, const MKL_INT* mtype,
             const MKL_INT* phase, const MKL_INT* n, const void* a, const MKL_INT* ia,
             const MKL_INT* ja, MKL_INT* perm, const MKL_INT* nrhs, MKL_INT* iparm,
             const MKL_INT* msglvl, void* b, void* x, MKL_INT* error) {
  (void)maxfct; (void)mnum; (void)perm; (void)msglvl;
  chol_t** slot = (chol_t**)pt; /* the caller's handle is >= 64 ints: room for one pointer */
  *error = 0;
  if (*mtype != 2) {
    fprintf(stderr, "[pardiso_shim] only mtype=2 (real SPD) is implemented, got %d\n", *mtype);
    *error = -1;
    return;
  }
  const int base = iparm[34] ? 0 : 1;
  switch (*phase) {
    case 11: /* analysis only: folded into the numeric phase */
      break;
    case 12:
    case 22:
    case 13: {
      if (*slot) { chol_free(*slot); *slot = NULL; }
      int err = 0;
      *slot = chol_factor(*n, ia, ja, (const double*)a, base, &err);
      if (err) { *error = err; return; }
      iparm[17] = (MKL_INT)((*slot)->lnz > 2147483647LL ? -1 : (*slot)->lnz);
      if (*phase != 13) break;
    } /* fall through for 13 */
    case 33: {
      if (!*slot) { *error = -1; return; }
      if (iparm[5] == 1) { /* in-place */
        double* tmp = (double*)malloc(sizeof(double) * (size_t)(*n) * (size_t)(*nrhs > 0 ? *nrhs : 1));
        chol_solve(*slot, *nrhs, (const double*)b, tmp);
        memcpy(b, tmp, sizeof(double) * (size_t)(*n) * (size_t)(*nrhs));
        free(tmp);
      } else {
        chol_solve(*slot, *nrhs, (const double*)b, (double*)x);
      }
      break;
    }
    case -1:
    case 0:
      if (*slot) { chol_free(*slot); *slot = NULL; }
      break;
    default:
      fprintf(stderr, "[pardiso_shim] phase %d not implemented\n", *phase);
      *error = -1;
  }
}
