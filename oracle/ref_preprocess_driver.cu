// oracle/ref_preprocess_driver.cu -- TEST INFRASTRUCTURE ONLY (never part of the product path).
//
// Host-array entry points around the reference's input-preprocessing functions (gpu/kernels/bgra_2_gray.h,
// cast.h, undistort.h, resample.h: resample_undistort), public API only.  oracle/build_ref.sh compiles it
// with the reference's own sources into oracle/_ref/libnmref.so (entry points nmref_*); `make compat-client`
// compiles the same file against the drop-in headers of this repository (-DNM_COMPAT_BUILD, nmcompat_*).
#include "bgra_2_gray.h"
#include "cast.h"
#include "undistort.h"
#include "resample.h"
#include "cudatex2D.h"
#include "downsample.h"
#include <cuda_runtime.h>

#ifndef NM_COMPAT_BUILD
#define NMREF(name) nmref_##name
#else
#define NMREF(name) nmcompat_##name
#endif

namespace {
template <typename T>
struct Dev {
    T* p = nullptr;
    size_t n;
    explicit Dev(size_t count, const T* host = nullptr) : n(count)
    {
        cudaMalloc(&p, sizeof(T) * (n ? n : 1));
        if (host) cudaMemcpy(p, host, sizeof(T) * n, cudaMemcpyHostToDevice);
    }
    void get(T* host) const { cudaMemcpy(host, p, sizeof(T) * n, cudaMemcpyDeviceToHost); }
    ~Dev() { cudaFree(p); }
};
int status() { cudaError_t e = cudaDeviceSynchronize(); return e == cudaSuccess ? 0 : 1000 + (int)e; }
} // namespace

extern "C" {

int NMREF(grayscale)(const unsigned char* bgra, int w, int h, float* out)
{
    Dev<uchar4> in((size_t)w * h, reinterpret_cast<const uchar4*>(bgra));
    Dev<float> o((size_t)w * h);
    cuda_grayscale<float>(in.p, o.p, w, h, 0);
    const int rc = status();
    o.get(out);
    return rc;
}

// extract every channel, write them back rotated (b <- g, g <- r, r <- b, a <- put(3) = 255), then
// cuda_set_alpha_to_const(alpha): chan_out = 4 float planes, bgra_out = the rewritten frame
int NMREF(channels)(const unsigned char* bgra, int w, int h, int alpha, float* chan_out, unsigned char* bgra_out)
{
    const size_t n = (size_t)w * h;
    Dev<uchar4> px(n, reinterpret_cast<const uchar4*>(bgra));
    Dev<float> ch(4 * n);
    for (int c = 0; c < 4; ++c) cuda_extract_channel<float>(px.p, ch.p + c * n, w, h, c, 0);
    int rc = status();
    ch.get(chan_out);
    cuda_put_channel<float>(px.p, ch.p + 1 * n, w, h, 0, 0);
    cuda_put_channel<float>(px.p, ch.p + 2 * n, w, h, 1, 0);
    cuda_put_channel<float>(px.p, ch.p + 0 * n, w, h, 2, 0);
    cuda_put_channel<float>(px.p, ch.p + 0 * n, w, h, 3, 0);
    if (alpha >= 0) cuda_set_alpha_to_const(px.p, w, h, (unsigned char)alpha, 0);
    rc = rc ? rc : status();
    px.get(reinterpret_cast<uchar4*>(bgra_out));
    return rc;
}

int NMREF(downsample_bgra)(const unsigned char* bgra, int w, int h, unsigned char* out)
{
    const int rw = w / 2, rh = h / 2;
    Dev<uchar4> px((size_t)w * h, reinterpret_cast<const uchar4*>(bgra)), o((size_t)rw * rh);
    downsample_by_2<uchar4>(o.p, rw, rh, px.p, w, h, 0);
    const int rc = status();
    o.get(reinterpret_cast<uchar4*>(out));
    return rc;
}

int NMREF(cast)(const float* src, int cols, int rows, unsigned char* dst, int max_val)
{
    Dev<float> in((size_t)cols * rows, src);
    Dev<unsigned char> o((size_t)cols * rows);
    cuda_cast<float, unsigned char>(in.p, (size_t)cols, (size_t)rows, o.p, (unsigned char)max_val, 0);
    const int rc = status();
    o.get(dst);
    return rc;
}

int NMREF(undistort)(const float* x, const float* y, int cols, int rows, const float* camera4, const float* dist3,
                     float* u, float* v)
{
    const size_t n = (size_t)cols * rows;
    Dev<float> dx(n, x), dy(n, y), cam(4, camera4), dist(3, dist3), du(n), dv(n);
    cuda_undistort(dx.p, dy.p, (size_t)cols, (size_t)rows, cam.p, dist.p, du.p, dv.p, 0);
    const int rc = status();
    du.get(u);
    dv.get(v);
    return rc;
}

// 8-bit image -> cudaArray -> CudaTex2D (normalised float reads, linear filtering, border addressing) ->
// resample_undistort at the coordinates (x, y)
int NMREF(resample_undistort)(const unsigned char* image, int w, int h, const float* x, const float* y, int cols, int rows,
                              float* out)
{
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<unsigned char>();
    cudaArray* arr = nullptr;
    if (cudaMallocArray(&arr, &desc, w, h) != cudaSuccess) return 999;
    cudaMemcpy2DToArray(arr, 0, 0, image, w, w, h, cudaMemcpyHostToDevice);
    int rc;
    {
        CudaTex2D tex;
        tex.set(arr);
        const size_t n = (size_t)cols * rows;
        Dev<float> dx(n, x), dy(n, y), o(n);
        resample_undistort(tex, dx.p, dy.p, (size_t)cols, (size_t)rows, o.p, 0);
        rc = status();
        o.get(out);
    }
    cudaFreeArray(arr);
    return rc;
}

} // extern "C"
