/* oracle/shim/mkl_pardiso.h -- TEST INFRASTRUCTURE ONLY; pardiso() is declared in the shim mkl.h */
#include "mkl.h"
