/*
 * oracle/shim/metis.h -- TEST INFRASTRUCTURE ONLY.
 * Prototype subset of METIS 5 matching the binary that is actually present in
 * this image: /usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a, which
 * was built with 64-bit idx_t and 32-bit real_t (probed).
 */
#ifndef ORACLE_SHIM_METIS_H
#define ORACLE_SHIM_METIS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int64_t idx_t;
typedef float real_t;
#define IDXTYPEWIDTH 64
#define REALTYPEWIDTH 32
#define METIS_NOPTIONS 40
#define METIS_OK 1
#define METIS_ERROR_INPUT (-2)
#define METIS_ERROR_MEMORY (-3)
#define METIS_ERROR (-4)
int METIS_PartGraphKway(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt,
                        idx_t* vsize, idx_t* adjwgt, idx_t* nparts, real_t* tpwgts, real_t* ubvec,
                        idx_t* options, idx_t* edgecut, idx_t* part);
int METIS_PartGraphRecursive(idx_t* nvtxs, idx_t* ncon, idx_t* xadj, idx_t* adjncy, idx_t* vwgt,
                             idx_t* vsize, idx_t* adjwgt, idx_t* nparts, real_t* tpwgts,
                             real_t* ubvec, idx_t* options, idx_t* edgecut, idx_t* part);
int METIS_NodeND(idx_t* nvtxs, idx_t* xadj, idx_t* adjncy, idx_t* vwgt, idx_t* options, idx_t* perm,
                 idx_t* iperm);
int METIS_SetDefaultOptions(idx_t* options);
#ifdef __cplusplus
}
#endif
#endif
