// nm_pyramid.cuh -- internal interface of the Gaussian scale-space kernels.
#pragma once
#include "nm_common.cuh"
#include <cuda.h>      // CUtensorMap (type only; the encoder is fetched at run time)

struct NmBlurArgs {
    const float* src;       // [batch][h][src_pitch]
    float*       dst;       // [batch][h][dst_pitch]
    const float* taps;      // device, 2R+1 floats
    const float* taps_host; // the same values in host memory, or null (the streaming kernel takes them as parameters)
    float*       dst2;      // optional decimated copy dst2[y/2][x/2] (next octave base), or null
    float*       scratch;   // only for the generic (R > 16) path: [batch][h][w]
    long long    src_fstride, dst_fstride, dst2_fstride;   // floats between frames
    int          w, h, src_pitch, dst_pitch, dst2_pitch, batch, radius;
    int          src_bgra;  // src holds BGRA words (uchar4), converted to grey while staged (strip kernel only)
};

// TMA descriptor of a blur source: 3-D tensor (x: w, y: h, frame: batch) with the row pitch
// and frame stride of NmBlurArgs, box = the kernel's staged window.  Encoding costs a few
// microseconds on the host, so the batched driver caches one per (octave, level).
struct NmBlurTma {
    CUtensorMap map;
    bool        valid;
    CUtensorMap map_strip;      // same tensor, box of 64 rows: the strip-walking kernel's chunk
    bool        valid_strip;
    CUtensorMap map_stream;     // same tensor, box of 8 rows: one stage of the streaming kernel
    bool        valid_stream;
};
// Returns false when the source cannot be described to TMA (pointer / pitch not 16-byte
// aligned, radius outside the tiled kernel's range): the caller then uses the plain-load
// kernel, which computes the same values.
bool nm_blur_make_tma(NmBlurTma* t, const float* src, int w, int h, int pitch, long long fstride,
                      int batch, int radius, bool words_u32 = false);
// True when nm_blur_launch would take the strip-walking kernel for these arguments (the only kernel that
// converts a BGRA source on the fly).
bool nm_blur_uses_strip(const NmBlurArgs& a, const NmBlurTma* tma);
// BGRA words -> grey floats, n pixels (nm_preprocess.cu).
int nm_grayscale_launch(const void* bgra, float* out, long long n, cudaStream_t st);

// Generic 3-D fp32 tiled descriptor (no swizzle, zero fill outside the tensor); false if it cannot be encoded.
bool nm_tma_encode_3d(NmBlurTma* t, const float* base, const unsigned long long dims[3],
                      const unsigned long long strides_bytes[2], const unsigned box[3]);

// Fused separable blur (rows then columns, zero padding, reference order of operations).
int nm_blur_launch(const NmBlurArgs& a, cudaStream_t stream, const NmBlurTma* tma = nullptr);

// dst[y][x] = src[2y][2x]
int nm_downsample_launch(float* dst, int dw, int dh, int dpitch, long long dfstride,
                         const float* src, int spitch, long long sfstride, int batch,
                         cudaStream_t stream);
