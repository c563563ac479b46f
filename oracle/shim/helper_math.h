// Stand-in for CUDA-samples <helper_math.h>; the reference includes it but uses
// none of its operators.  TEST INFRASTRUCTURE ONLY (see helper_cuda.h here).
#ifndef NM_ORACLE_SHIM_HELPER_MATH_H
#define NM_ORACLE_SHIM_HELPER_MATH_H
#include <cuda_runtime.h>
#endif
