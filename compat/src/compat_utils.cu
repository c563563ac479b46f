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

// cudautils.h (gpu/utils/cudautils.cpp:8-31)
int CudaUtils::_max_gflops_device_id = -1;
int CudaUtils::get_max_flops_device_id()
{
    if (_max_gflops_device_id != -1) return _max_gflops_device_id;
    int n = 0, best = 0;
    double best_score = -1.0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) RUNTIME_EXCEPTION("no CUDA device");
    for (int d = 0; d < n; ++d) {
        int sms = 0, khz = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d);
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, d);
        const double score = (double)sms * khz;
        if (score > best_score) { best_score = score; best = d; }
    }
    _max_gflops_device_id = best;
    return best;
}
void CudaUtils::setup_CUDA(int device_id)
{
    if (cudaSetDevice(device_id) != cudaSuccess) RUNTIME_EXCEPTION("Could not set the CUDA device");
}
