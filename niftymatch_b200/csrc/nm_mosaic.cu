// nm_mosaic.cu -- the step AFTER registration: perspective coordinate maps, resampling
// through the caller's textures and the weighted blend of a warped frame into the mosaic canvas.
//
// Replaces resample_perspective_transform (gpu/kernels/resample.cu:119-219, :221-233), resample_mask
// (:67-81, :236-244) and transform_blend (:7-65, :246-258).  The texture objects are the caller's (CudaTex2D:
// linear filtering, border addressing, un-normalised coordinates), so the filtered samples are produced by
// the same texture unit as in the reference; the arithmetic around them is written in the reference's
// order.  Flat 1-D grids: every kernel here is one independent HBM-bound pass over the output pixels.
#include "nm_common.cuh"

namespace {

struct Mat3 { float m[9]; };

// The reference's a*x + b*y + c and a*b - c*d as its build contracts them (established against the vectors
// the reference produced, tests/golden/mosaic_128x90.npz): the first product is fused into the sum, the second
// one is rounded on its own.  Written with explicit intrinsics so that this build cannot choose differently.
__device__ __forceinline__ float lin3(float a, float x, float b, float y, float c)
{
    return __fadd_rn(__fmaf_rn(a, x, __fmul_rn(b, y)), c);
}
__device__ __forceinline__ float det2(float a, float b, float c, float d)     // a*b - c*d
{
    return __fmaf_rn(a, b, -__fmul_rn(c, d));
}

// the inverse of apply_perspective_inverse (resample.cu:133-150), evaluated by every thread from the same
// nine floats (the reference lets thread 0 of each block do it: same values)
__device__ __forceinline__ Mat3 invert3(const float* __restrict__ t)
{
    Mat3 r;
    const float det = __fmaf_rn(t[2], det2(t[3], t[7], t[4], t[6]),
                                __fmaf_rn(t[0], det2(t[4], t[8], t[7], t[5]), -__fmul_rn(t[1], det2(t[3], t[8], t[5], t[6]))));
    const float invdet = 1 / det;
    r.m[0] = __fmul_rn(det2(t[4], t[8], t[7], t[5]), invdet);
    r.m[1] = __fmul_rn(det2(t[2], t[7], t[1], t[8]), invdet);
    r.m[2] = __fmul_rn(det2(t[1], t[5], t[2], t[4]), invdet);
    r.m[3] = __fmul_rn(det2(t[5], t[6], t[3], t[8]), invdet);
    r.m[4] = __fmul_rn(det2(t[0], t[8], t[2], t[6]), invdet);
    r.m[5] = __fmul_rn(det2(t[3], t[2], t[0], t[5]), invdet);
    r.m[6] = __fmul_rn(det2(t[3], t[7], t[6], t[4]), invdet);
    r.m[7] = __fmul_rn(det2(t[6], t[1], t[0], t[7]), invdet);
    r.m[8] = __fmul_rn(det2(t[0], t[4], t[3], t[1]), invdet);
    return r;
}

// apply_perspective / apply_perspective_inverse (resample.cu:119-191) fused with resample_2D<uchar4> (:83-102)
template <bool INVERSE>
__global__ void __launch_bounds__(256) perspective_resample_kernel(uchar4* __restrict__ result, cudaTextureObject_t tex, int width,
                                                                   int height, float* __restrict__ x_pos, float* __restrict__ y_pos,
                                                                   const float* __restrict__ mat3x3)
{
    __shared__ float t[9];
    if (threadIdx.x == 0) {
        if (INVERSE) {
            const Mat3 inv = invert3(mat3x3);
            for (int i = 0; i < 9; ++i) t[i] = inv.m[i];
        } else {
            for (int i = 0; i < 9; ++i) t[i] = mat3x3[i];
        }
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)width * height) return;
    const int y = (int)(i / width), x = (int)(i - (long long)y * width);
    const float xf = (float)x, yf = (float)y;
    const float x_p = lin3(t[0], xf, t[1], yf, t[2]);
    const float y_p = lin3(t[3], xf, t[4], yf, t[5]);
    const float s_p = lin3(t[6], xf, t[7], yf, t[8]);
    const float xs = __fdiv_rn(x_p, s_p), ys = __fdiv_rn(y_p, s_p);
    x_pos[i] = xs;
    y_pos[i] = ys;
    const float4 res = tex2D<float4>(tex, xs + 0.5f, ys + 0.5f);
    uchar4 o;
    o.x = (unsigned char)(res.x * 255.9999f);
    o.y = (unsigned char)(res.y * 255.9999f);
    o.z = (unsigned char)(res.z * 255.9999f);
    o.w = (unsigned char)(res.w * 255.9999f);
    result[i] = o;
}

// resample_mask_2D (resample.cu:67-81)
__global__ void __launch_bounds__(256) resample_mask_kernel(unsigned char* __restrict__ result, cudaTextureObject_t tex, long long n,
                                                            const float* __restrict__ x, const float* __restrict__ y, float lower_limit)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float res = tex2D<float>(tex, x[i] + 0.5f, y[i] + 0.5f);
    if (res <= lower_limit) result[i] = 0;
    else result[i] = res * 255.999f;
}

// transform_and_blend (resample.cu:7-65)
__global__ void __launch_bounds__(256) blend_kernel(uchar4* __restrict__ canvas, int cw, int ch, cudaTextureObject_t frame, int fw, int fh,
                                                    int nw, int nh, const float* __restrict__ mat3x3, int tx, int ty,
                                                    cudaTextureObject_t mask, float* __restrict__ canvas_wts, cudaTextureObject_t frame_wts)
{
    __shared__ float t[9];
    if (threadIdx.x < 9) t[threadIdx.x] = mat3x3[threadIdx.x];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)nw * nh) return;
    const int y = (int)(i / nw), x = (int)(i - (long long)y * nw);
    const int pos_x = x + tx, pos_y = y + ty;
    if (pos_x < 0 || pos_x >= cw || pos_y < 0 || pos_y >= ch) return;
    const float xf = (float)x, yf = (float)y;
    const float s_p = lin3(t[6], xf, t[7], yf, t[8]);
    const float x_p = __fdiv_rn(lin3(t[0], xf, t[1], yf, t[2]), s_p);
    const float y_p = __fdiv_rn(lin3(t[3], xf, t[4], yf, t[5]), s_p);
    if (x_p >= fw || y_p >= fh) return;
    const float4 res = tex2D<float4>(frame, x_p + 0.5f, y_p + 0.5f);
    const float in_mask = tex2D<float>(mask, x_p + 0.5f, y_p + 0.5f);
    if (in_mask <= 0.5) return;
    const float new_weight = tex2D<float>(frame_wts, x_p + 0.5f, y_p + 0.5f);
    const long long index = (long long)pos_y * cw + pos_x;
    const float current_weight = canvas_wts[index];
    uchar4 o;
    if (current_weight == 0) {
        o.x = (unsigned char)(res.x * 255.9999f);
        o.y = (unsigned char)(res.y * 255.9999f);
        o.z = (unsigned char)(res.z * 255.9999f);
        o.w = 255;
        canvas[index] = o;
        canvas_wts[index] = new_weight;
    } else {
        const uchar4 cur = canvas[index];
        const float sum_wts = current_weight + new_weight;
        o.x = (unsigned char)__fdiv_rn(__fmaf_rn(__fmul_rn(res.x, new_weight), 255.9999f, __fmul_rn((float)cur.x, current_weight)), sum_wts);
        o.y = (unsigned char)__fdiv_rn(__fmaf_rn(__fmul_rn(res.y, new_weight), 255.9999f, __fmul_rn((float)cur.y, current_weight)), sum_wts);
        o.z = (unsigned char)__fdiv_rn(__fmaf_rn(__fmul_rn(res.z, new_weight), 255.9999f, __fmul_rn((float)cur.z, current_weight)), sum_wts);
        o.w = 255;
        canvas[index] = o;
        canvas_wts[index] = current_weight + new_weight;
    }
}

} // namespace

extern "C" int nm_resample_perspective_bgra(void* result, unsigned long long tex, int cols, int rows, float* x_pos, float* y_pos,
                                            const float* mat3x3, int inverse, nm_stream_t stream)
{
    if (cols < 0 || rows < 0) return NM_ERR_INVALID;
    const long long n = (long long)cols * rows;
    if (n == 0) return NM_OK;
    if (!result || !tex || !x_pos || !y_pos || !mat3x3) return NM_ERR_INVALID;
    const unsigned grid = (unsigned)nm_div_up64(n, 256);
    if (inverse)
        perspective_resample_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<uchar4*>(result), (cudaTextureObject_t)tex,
                                                                                  cols, rows, x_pos, y_pos, mat3x3);
    else
        perspective_resample_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<uchar4*>(result), (cudaTextureObject_t)tex,
                                                                                   cols, rows, x_pos, y_pos, mat3x3);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

extern "C" int nm_resample_mask_tex_u8(unsigned char* result, unsigned long long tex, int cols, int rows, const float* x_pos,
                                       const float* y_pos, float threshold, nm_stream_t stream)
{
    if (cols < 0 || rows < 0) return NM_ERR_INVALID;
    const long long n = (long long)cols * rows;
    if (n == 0) return NM_OK;
    if (!result || !tex || !x_pos || !y_pos) return NM_ERR_INVALID;
    resample_mask_kernel<<<(unsigned)nm_div_up64(n, 256), 256, 0, (cudaStream_t)stream>>>(result, (cudaTextureObject_t)tex, n, x_pos, y_pos,
                                                                                         threshold);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

extern "C" int nm_transform_blend_bgra(void* canvas, int cw, int ch, unsigned long long frame_tex, int fw, int fh, int nw, int nh,
                                       const float* mat3x3, int tx, int ty, unsigned long long mask_tex, float* canvas_wts,
                                       unsigned long long frame_wts_tex, nm_stream_t stream)
{
    if (cw < 0 || ch < 0 || nw < 0 || nh < 0) return NM_ERR_INVALID;
    const long long n = (long long)nw * nh;
    if (n == 0 || (long long)cw * ch == 0) return NM_OK;
    if (!canvas || !frame_tex || !mat3x3 || !mask_tex || !canvas_wts || !frame_wts_tex) return NM_ERR_INVALID;
    blend_kernel<<<(unsigned)nm_div_up64(n, 256), 256, 0, (cudaStream_t)stream>>>(static_cast<uchar4*>(canvas), cw, ch,
                                                                                 (cudaTextureObject_t)frame_tex, fw, fh, nw, nh, mat3x3, tx, ty,
                                                                                 (cudaTextureObject_t)mask_tex, canvas_wts,
                                                                                 (cudaTextureObject_t)frame_wts_tex);
    NM_LAUNCH_CHECK();
    return NM_OK;
}
