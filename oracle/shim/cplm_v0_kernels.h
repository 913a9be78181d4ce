/* oracle/shim/cplm_v0_kernels.h -- TEST INFRASTRUCTURE ONLY.
 * The reference's src/preconditioners/block_jacobi.h:25 includes this header,
 * which does not exist in the shipped tree; these two includes provide what
 * block_jacobi.{c,h} actually use. */
#include <cplm_kernels.h>
#include <cplm_v0_dvector.h>
