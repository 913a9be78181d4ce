/* compat/mkl.h -- the reference driver includes <mkl.h> only to call
 * MKL_Set_Num_Threads(1) (ref: examples/test_ecg_prealps_op.c:19,149).  This library does
 * not use MKL; the call is a no-op here.  Put include/compat on the include path only
 * when no real MKL is installed. */
#ifndef PREALPS_B200_COMPAT_MKL_H
#define PREALPS_B200_COMPAT_MKL_H
static inline void MKL_Set_Num_Threads(int n) { (void)n; }
#define mkl_set_num_threads MKL_Set_Num_Threads
#endif
