// compat_utils.cu -- libgpuutils.a of the drop-in layer: CudaTex2D and CudaTimer
// (reference src/gpu/utils/cudatex2D.cu:4-39, cudatimer.cu:3-22).
#include "nm_compat.hpp"

namespace {
inline void cuda_check(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) RUNTIME_EXCEPTION(std::string(what) + ": " + cudaGetErrorString(e));
}
} // namespace

CudaTex2D::CudaTex2D(cudaArray* array) : _tex(0) { set(array); }

CudaTex2D::~CudaTex2D() { release(); }

void CudaTex2D::set(cudaArray* array, cudaTextureReadMode read_mode)
{
    release();
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = array;
    cudaTextureDesc tex{};
    tex.addressMode[0] = tex.addressMode[1] = cudaAddressModeBorder;   // zero outside the image
    tex.filterMode = cudaFilterModeLinear;                             // exact texels at centres
    tex.readMode = read_mode;
    tex.normalizedCoords = 0;
    cuda_check(cudaCreateTextureObject(&_tex, &res, &tex, nullptr), "cudaCreateTextureObject");
}

void CudaTex2D::release()
{
    if (_tex) cudaDestroyTextureObject(_tex);
    _tex = 0;
}

CudaTimer::CudaTimer(cudaStream_t stream) : _stream(stream)
{
    cuda_check(cudaEventCreate(&_start), "cudaEventCreate");
    cuda_check(cudaEventCreate(&_stop), "cudaEventCreate");
}

CudaTimer::~CudaTimer()
{
    cudaEventDestroy(_start);
    cudaEventDestroy(_stop);
}

void CudaTimer::start() { cuda_check(cudaEventRecord(_start, _stream), "cudaEventRecord"); }

float CudaTimer::stop()
{
    float ms = 0.f;
    cuda_check(cudaEventRecord(_stop, _stream), "cudaEventRecord");
    cuda_check(cudaEventSynchronize(_stop), "cudaEventSynchronize");
    cuda_check(cudaEventElapsedTime(&ms, _start, _stop), "cudaEventElapsedTime");
    return ms;
}
