// Minimal stand-in for the CUDA-samples header the reference includes
// (<helper_cuda.h>, normally found under $CUDA/samples/common/inc; not shipped
// with CUDA 12.9).  TEST INFRASTRUCTURE ONLY: used when oracle/build_ref.sh
// compiles the unmodified reference sources from /root/reference into
// oracle/_ref/.  Only the three names the reference actually uses are provided.
#ifndef NM_ORACLE_SHIM_HELPER_CUDA_H
#define NM_ORACLE_SHIM_HELPER_CUDA_H
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

static inline void nm_shim_check(cudaError_t e, const char* what, const char* file, int line)
{
    if (e != cudaSuccess) {
        std::fprintf(stderr, "[nmref] CUDA error %d (%s) at %s:%d: %s\n", (int)e,
                     cudaGetErrorString(e), file, line, what);
        std::exit(EXIT_FAILURE);
    }
}
#define checkCudaErrors(expr) nm_shim_check((expr), #expr, __FILE__, __LINE__)
#define getLastCudaError(msg) nm_shim_check(cudaGetLastError(), (msg), __FILE__, __LINE__)
static inline int gpuGetMaxGflopsDeviceId() { return 0; }
#endif
