// oracle/ref_mosaic_driver.cu -- TEST INFRASTRUCTURE ONLY (never part of the product path).
//
// Host-array entry points around the reference's mosaic functions (gpu/kernels/resample.h:7-23), public API
// only: the textures are set up the way a reference client does (cudaArray + CudaTex2D: linear filtering,
// border addressing; uchar4 / uchar arrays read as normalised floats, the weight map as a float array).
// Built into oracle/_ref/libnmref.so (nmref_*) by oracle/build_ref.sh and, with -DNM_COMPAT_BUILD, against the
// drop-in headers of this repository (nmcompat_*).
#include "resample.h"
#include "cudatex2D.h"
#include <cuda_runtime.h>
#include <cstring>

#ifndef NM_COMPAT_BUILD
#define NMREF(name) nmref_##name
#else
#define NMREF(name) nmcompat_##name
#endif

namespace {
template <typename T>
struct Arr {
    cudaArray* a = nullptr;
    Arr(const T* host, int w, int h)
    {
        cudaChannelFormatDesc d = cudaCreateChannelDesc<T>();
        cudaMallocArray(&a, &d, w, h);
        cudaMemcpy2DToArray(a, 0, 0, host, sizeof(T) * w, sizeof(T) * w, h, cudaMemcpyHostToDevice);
    }
    ~Arr() { cudaFreeArray(a); }
};
template <typename T>
struct Dev {
    T* p = nullptr;
    size_t n;
    explicit Dev(size_t count, const T* host = nullptr) : n(count)
    {
        cudaMalloc(&p, sizeof(T) * (n ? n : 1));
        if (host) cudaMemcpy(p, host, sizeof(T) * n, cudaMemcpyHostToDevice);
        else cudaMemset(p, 0, sizeof(T) * (n ? n : 1));
    }
    void get(T* host) const { cudaMemcpy(host, p, sizeof(T) * n, cudaMemcpyDeviceToHost); }
    ~Dev() { cudaFree(p); }
};
int status() { cudaError_t e = cudaDeviceSynchronize(); return e == cudaSuccess ? 0 : 1000 + (int)e; }
} // namespace

extern "C" {

// frame: fw x fh BGRA bytes.  result: cols x rows BGRA bytes, x_pos / y_pos: cols x rows floats.
int NMREF(resample_perspective)(const unsigned char* frame, int fw, int fh, const float* mat9, int inverse, int cols, int rows,
                                unsigned char* result, float* x_pos, float* y_pos)
{
    Arr<uchar4> fa(reinterpret_cast<const uchar4*>(frame), fw, fh);
    int rc;
    {
        CudaTex2D tex;
        tex.set(fa.a);
        const size_t n = (size_t)cols * rows;
        Dev<uchar4> out(n);
        Dev<float> xp(n), yp(n), m(9, mat9);
        resample_perspective_transform(out.p, tex, cols, rows, xp.p, yp.p, m.p, inverse != 0, 0);
        rc = status();
        out.get(reinterpret_cast<uchar4*>(result));
        xp.get(x_pos);
        yp.get(y_pos);
    }
    return rc;
}

// mask: mw x mh bytes, sampled at (x_pos, y_pos)
int NMREF(resample_mask)(const unsigned char* mask, int mw, int mh, const float* x_pos, const float* y_pos, int cols, int rows,
                         float threshold, unsigned char* result)
{
    Arr<unsigned char> ma(mask, mw, mh);
    int rc;
    {
        CudaTex2D tex;
        tex.set(ma.a);
        const size_t n = (size_t)cols * rows;
        Dev<unsigned char> out(n);
        Dev<float> xp(n, x_pos), yp(n, y_pos);
        resample_mask(out.p, tex, cols, rows, xp.p, yp.p, threshold, 0);
        rc = status();
        out.get(result);
    }
    return rc;
}

// Blend `n_frames` warps of the same frame (mat9[k], offsets (tx[k], ty[k])) into a cw x ch canvas that starts
// empty (weights 0): the first warp takes the "empty pixel" branch, the later ones the weighted one.
int NMREF(transform_blend)(const unsigned char* frame, const unsigned char* mask, const float* wts, int fw, int fh, int n_frames,
                           const float* mat9, const int* tx, const int* ty, int nw, int nh, int cw, int ch, unsigned char* canvas,
                           float* canvas_wts)
{
    Arr<uchar4> fa(reinterpret_cast<const uchar4*>(frame), fw, fh);
    Arr<unsigned char> ma(mask, fw, fh);
    Arr<float> wa(wts, fw, fh);
    int rc;
    {
        CudaTex2D tf, tm, tw;
        tf.set(fa.a);
        tm.set(ma.a);
        tw.set(wa.a, cudaReadModeElementType);
        Dev<uchar4> cv((size_t)cw * ch);
        Dev<float> cwts((size_t)cw * ch), m((size_t)9 * n_frames, mat9);
        for (int k = 0; k < n_frames; ++k)
            transform_blend(cv.p, cw, ch, tf, fw, fh, nw, nh, m.p + 9 * k, tx[k], ty[k], tm, cwts.p, tw, 0);
        rc = status();
        cv.get(reinterpret_cast<uchar4*>(canvas));
        cwts.get(canvas_wts);
    }
    return rc;
}

} // extern "C"
