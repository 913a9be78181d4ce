/* oracle/metis_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * Thin int32 front-end to METIS_PartGraphKway / METIS_NodeND (CUDA toolkit's libmetis_static.a,
 * idx_t = int64) for the numpy restatement in oracle/restate.py; call sequence of the reference's
 * callKway (/root/reference/utils/cplm_core/cplm_matcsr_core.c:394-457): ncon = 1, no weights,
 * options = NULL. */
#include <stdint.h>
#include <stdlib.h>
int METIS_PartGraphKway(int64_t*, int64_t*, int64_t*, int64_t*, int64_t*, int64_t*, int64_t*, int64_t*, float*, float*,
                        int64_t*, int64_t*, int64_t*);
int oracle_kway(int n, const int* xadj, const int* adj, int nparts, int* parts) {
  int64_t nv = n, ncon = 1, np = nparts, obj = 0;
  int64_t* xa = malloc(sizeof(int64_t) * (n + 1));
  int64_t* ad = malloc(sizeof(int64_t) * (xadj[n] > 0 ? xadj[n] : 1));
  int64_t* p = malloc(sizeof(int64_t) * n);
  for (int i = 0; i <= n; ++i) xa[i] = xadj[i];
  for (int i = 0; i < xadj[n]; ++i) ad[i] = adj[i];
  int rc = METIS_PartGraphKway(&nv, &ncon, xa, ad, NULL, NULL, NULL, &np, NULL, NULL, NULL, &obj, p);
  for (int i = 0; i < n; ++i) parts[i] = (int)p[i];
  free(xa); free(ad); free(p);
  return rc;
}
