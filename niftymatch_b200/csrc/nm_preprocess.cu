// nm_preprocess.cu -- the step BEFORE the pyramid: BGRA -> grey, cast to bytes, the
// lens-undistortion coordinate map and the resampling through it.
//
// Replaces cuda_grayscale<float> (gpu/kernels/bgra_2_gray.cu:8-29), cuda_cast<float, unsigned char>
// (cast.cu:7-39), cuda_undistort (undistort.cu:6-64) and resample_undistort (resample.cu:104-117, :235-248).
//
// Grey value.  The reference evaluates 0.07*b + 0.72*g + 0.21*r in DOUBLE (the literals are doubles) and
// rounds the sum to float.  In real arithmetic that is N/100 with N = 7b + 72g + 21r <= 25 500, an integer.
// N/100 is either a float itself (N a multiple of 25) or at least 1/(100 * 2^17) away from every midpoint
// between two floats below 256, which is ~10^8 times the rounding error of the three double operations, so
// the reference's result is RN_float(N/100) for every input -- computed here as one 4-way byte dot product
// (dp4a with the packed weights 7, 72, 21, 0), an int -> float conversion (exact) and ONE correctly rounded
// fp32 division.  tests/test_gpu_preprocess.py checks all 2^24 (b, g, r) against the double formula and
// against the reference's kernel.
//
// In the batched SIFT path the conversion is fused into the base blur (nm_pyramid.cu: the BGRA words are
// staged by TMA and converted in shared memory), so no grey frame is ever written to HBM.
#include "nm_common.cuh"
#include "nm_pyramid.cuh"

namespace {

// 4 pixels per thread: one 16-byte load, one 16-byte store
__global__ void __launch_bounds__(256) gray4_kernel(const uint4* __restrict__ bgra, float4* __restrict__ out, long long n4)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const uint4 p = __ldg(bgra + i);
    out[i] = make_float4(nm_gray_from_bgra(p.x), nm_gray_from_bgra(p.y), nm_gray_from_bgra(p.z), nm_gray_from_bgra(p.w));
}
__global__ void __launch_bounds__(256) gray1_kernel(const unsigned* __restrict__ bgra, float* __restrict__ out, long long first,
                                                    long long n)
{
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = nm_gray_from_bgra(__ldg(bgra + i));
}

// The element-wise kernels below process FOUR consecutive elements per thread with 16-byte loads / stores when
// the pointers allow it (first = 0, count = n / 4 quads), and the remainder (or everything, for unaligned
// pointers) one element per thread: a one-element-per-thread cast reached 0.31 of the HBM roofline.
__device__ __forceinline__ unsigned char cast_one(float v, unsigned char max_val)
{
    return (max_val != 0 && v >= max_val) ? max_val : (unsigned char)(v);
}
// cast.cu:7-21.  The reference indexes pos = j*cols + i with i up to the padded grid width and only tests
// pos < cols*rows, so every element is written (some twice, with the same value): a flat loop is equivalent.
__global__ void __launch_bounds__(256) cast4_kernel(const float4* __restrict__ src, uchar4* __restrict__ dst, long long n4,
                                                    unsigned char max_val)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(src + i);
    dst[i] = make_uchar4(cast_one(v.x, max_val), cast_one(v.y, max_val), cast_one(v.z, max_val), cast_one(v.w, max_val));
}
__global__ void __launch_bounds__(256) cast1_kernel(const float* __restrict__ src, unsigned char* __restrict__ dst, long long first,
                                                    long long n, unsigned char max_val)
{
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = cast_one(src[i], max_val);
}

// undistort.cu:6-47 (the stores to u/v between the steps round to fp32, as the local floats here do)
struct Camera { float k1, k2, k3, fx, fy, cx, cy; };
__device__ __forceinline__ Camera load_camera(const float* __restrict__ distortion_coeffs, const float* __restrict__ camera_matrix)
{
    return Camera{distortion_coeffs[0], distortion_coeffs[1], distortion_coeffs[2], camera_matrix[0], camera_matrix[1],
                  camera_matrix[2], camera_matrix[3]};
}
__device__ __forceinline__ void undistort_one(const Camera& c, float x, float y, float& u, float& v)
{
    float uu = x;
    uu -= c.cx;
    uu /= c.fx;
    float vv = y;
    vv -= c.cy;
    vv /= c.fy;
    const float r2 = powf(uu, 2) + powf(vv, 2);
    const float kr_poly = 1 + c.k1 * r2 + c.k2 * powf(r2, 2) + c.k3 * powf(r2, 3);
    uu *= kr_poly;
    uu *= c.fx;
    uu += c.cx;
    vv *= kr_poly;
    vv *= c.fy;
    vv += c.cy;
    u = uu;
    v = vv;
}
__global__ void __launch_bounds__(256) undistort4_kernel(const float4* __restrict__ x, const float4* __restrict__ y, long long n4,
                                                         const float* __restrict__ distortion_coeffs,
                                                         const float* __restrict__ camera_matrix, float4* __restrict__ u,
                                                         float4* __restrict__ v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const Camera c = load_camera(distortion_coeffs, camera_matrix);
    const float4 a = __ldg(x + i), b = __ldg(y + i);
    float4 uu, vv;
    undistort_one(c, a.x, b.x, uu.x, vv.x);
    undistort_one(c, a.y, b.y, uu.y, vv.y);
    undistort_one(c, a.z, b.z, uu.z, vv.z);
    undistort_one(c, a.w, b.w, uu.w, vv.w);
    u[i] = uu;
    v[i] = vv;
}
__global__ void __launch_bounds__(256) undistort1_kernel(const float* __restrict__ x, const float* __restrict__ y, long long first,
                                                         long long n, const float* __restrict__ distortion_coeffs,
                                                         const float* __restrict__ camera_matrix, float* __restrict__ u,
                                                         float* __restrict__ v)
{
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Camera c = load_camera(distortion_coeffs, camera_matrix);
    undistort_one(c, x[i], y[i], u[i], v[i]);
}

// resample_2D<float> (resample.cu:104-117): the caller's texture decides filtering and addressing
__global__ void __launch_bounds__(256) resample_kernel(float* __restrict__ result, cudaTextureObject_t tex, long long n,
                                                       const float* __restrict__ x, const float* __restrict__ y)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float res = tex2D<float>(tex, x[i] + 0.5f, y[i] + 0.5f);
    result[i] = res * 255.9999f;
}

// extract_channel / put_channel / set_alpha_to_const (bgra_2_gray.cu:33-112): one 32-bit word per pixel
__global__ void __launch_bounds__(256) extract_channel4_kernel(const uint4* __restrict__ bgra, float4* __restrict__ out, long long n4,
                                                               int channel)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const uint4 p = __ldg(bgra + i);
    const int sh = 8 * channel;
    out[i] = make_float4((float)((p.x >> sh) & 0xffu), (float)((p.y >> sh) & 0xffu), (float)((p.z >> sh) & 0xffu),
                         (float)((p.w >> sh) & 0xffu));
}
__global__ void __launch_bounds__(256) extract_channel_kernel(const unsigned* __restrict__ bgra, float* __restrict__ out, long long first,
                                                              long long n, int channel)
{
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)((__ldg(bgra + i) >> (8 * channel)) & 0xffu);
}
__global__ void __launch_bounds__(256) put_channel_kernel(unsigned* __restrict__ bgra, const float* __restrict__ in, long long n,
                                                          int channel)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned v = channel == 3 ? 255u : (unsigned)(unsigned char)in[i];      // alpha is set to 255 (bgra_2_gray.cu:71)
    bgra[i] = (bgra[i] & ~(0xffu << (8 * channel))) | (v << (8 * channel));
}
__global__ void __launch_bounds__(256) set_alpha_kernel(unsigned* __restrict__ bgra, long long n, unsigned val)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bgra[i] = (bgra[i] & 0x00ffffffu) | (val << 24);
}

} // namespace

int nm_grayscale_launch(const void* bgra, float* out, long long n, cudaStream_t st)
{
    const bool vec = !(reinterpret_cast<uintptr_t>(bgra) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15);
    const long long n4 = vec ? n / 4 : 0;
    if (n4) {
        gray4_kernel<<<(unsigned)nm_div_up64(n4, 256), 256, 0, st>>>(static_cast<const uint4*>(bgra), reinterpret_cast<float4*>(out), n4);
        NM_LAUNCH_CHECK();
    }
    if (n4 * 4 < n) {
        gray1_kernel<<<(unsigned)nm_div_up64(n - n4 * 4, 256), 256, 0, st>>>(static_cast<const unsigned*>(bgra), out, n4 * 4, n);
        NM_LAUNCH_CHECK();
    }
    return NM_OK;
}

extern "C" int nm_grayscale_bgra_f32(const void* bgra, float* output, int width, int height, nm_stream_t stream)
{
    if (width < 0 || height < 0) return NM_ERR_INVALID;
    const long long n = (long long)width * height;
    if (n == 0) return NM_OK;
    if (!bgra || !output || (reinterpret_cast<uintptr_t>(bgra) & 3)) return NM_ERR_INVALID;
    return nm_grayscale_launch(bgra, output, n, (cudaStream_t)stream);
}

extern "C" int nm_cast_f32_u8(const float* src, int cols, int rows, unsigned char* dst, unsigned char max_val,
                              nm_stream_t stream)
{
    if (cols < 0 || rows < 0) return NM_ERR_INVALID;
    const long long n = (long long)cols * rows;
    if (n == 0) return NM_OK;
    if (!src || !dst) return NM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = !(reinterpret_cast<uintptr_t>(src) & 15) && !(reinterpret_cast<uintptr_t>(dst) & 3);
    const long long n4 = vec ? n / 4 : 0;
    if (n4) {
        cast4_kernel<<<(unsigned)nm_div_up64(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uchar4*>(dst), n4,
                                                                    max_val);
        NM_LAUNCH_CHECK();
    }
    if (n4 * 4 < n) {
        cast1_kernel<<<(unsigned)nm_div_up64(n - n4 * 4, 256), 256, 0, st>>>(src, dst, n4 * 4, n, max_val);
        NM_LAUNCH_CHECK();
    }
    return NM_OK;
}

extern "C" int nm_undistort_map_f32(const float* x, const float* y, int cols, int rows, const float* camera_matrix,
                                    const float* distortion_coeffs, float* u, float* v, nm_stream_t stream)
{
    if (cols < 0 || rows < 0) return NM_ERR_INVALID;
    const long long n = (long long)cols * rows;
    if (n == 0) return NM_OK;
    if (!x || !y || !camera_matrix || !distortion_coeffs || !u || !v) return NM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(u) |
                        reinterpret_cast<uintptr_t>(v)) & 15);
    const long long n4 = vec ? n / 4 : 0;
    if (n4) {
        undistort4_kernel<<<(unsigned)nm_div_up64(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y),
                                                                         n4, distortion_coeffs, camera_matrix, reinterpret_cast<float4*>(u),
                                                                         reinterpret_cast<float4*>(v));
        NM_LAUNCH_CHECK();
    }
    if (n4 * 4 < n) {
        undistort1_kernel<<<(unsigned)nm_div_up64(n - n4 * 4, 256), 256, 0, st>>>(x, y, n4 * 4, n, distortion_coeffs, camera_matrix, u, v);
        NM_LAUNCH_CHECK();
    }
    return NM_OK;
}

extern "C" int nm_resample_tex_f32(unsigned long long tex, const float* x, const float* y, int cols, int rows,
                                   float* result, nm_stream_t stream)
{
    if (cols < 0 || rows < 0) return NM_ERR_INVALID;
    const long long n = (long long)cols * rows;
    if (n == 0) return NM_OK;
    if (!tex || !x || !y || !result) return NM_ERR_INVALID;
    resample_kernel<<<(unsigned)nm_div_up64(n, 256), 256, 0, (cudaStream_t)stream>>>(result, (cudaTextureObject_t)tex, n, x, y);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

extern "C" int nm_bgra_extract_channel_f32(const void* bgra, float* output, int width, int height, int channel, nm_stream_t stream)
{
    if (width < 0 || height < 0) return NM_ERR_INVALID;
    const long long n = (long long)width * height;
    if (n == 0 || channel < 0 || channel > 3) return NM_OK;          // the reference writes nothing for other channel numbers
    if (!bgra || !output || (reinterpret_cast<uintptr_t>(bgra) & 3)) return NM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = !((reinterpret_cast<uintptr_t>(bgra) | reinterpret_cast<uintptr_t>(output)) & 15);
    const long long n4 = vec ? n / 4 : 0;
    if (n4) {
        extract_channel4_kernel<<<(unsigned)nm_div_up64(n4, 256), 256, 0, st>>>(static_cast<const uint4*>(bgra),
                                                                               reinterpret_cast<float4*>(output), n4, channel);
        NM_LAUNCH_CHECK();
    }
    if (n4 * 4 < n) {
        extract_channel_kernel<<<(unsigned)nm_div_up64(n - n4 * 4, 256), 256, 0, st>>>(static_cast<const unsigned*>(bgra), output, n4 * 4, n,
                                                                                      channel);
        NM_LAUNCH_CHECK();
    }
    return NM_OK;
}

extern "C" int nm_bgra_put_channel_f32(void* bgra, const float* input, int width, int height, int channel, nm_stream_t stream)
{
    if (width < 0 || height < 0) return NM_ERR_INVALID;
    const long long n = (long long)width * height;
    if (n == 0 || channel < 0 || channel > 3) return NM_OK;
    if (!bgra || !input || (reinterpret_cast<uintptr_t>(bgra) & 3)) return NM_ERR_INVALID;
    put_channel_kernel<<<(unsigned)nm_div_up64(n, 256), 256, 0, (cudaStream_t)stream>>>(static_cast<unsigned*>(bgra), input, n, channel);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

extern "C" int nm_bgra_set_alpha(void* bgra, int width, int height, unsigned char val, nm_stream_t stream)
{
    if (width < 0 || height < 0) return NM_ERR_INVALID;
    const long long n = (long long)width * height;
    if (n == 0) return NM_OK;
    if (!bgra || (reinterpret_cast<uintptr_t>(bgra) & 3)) return NM_ERR_INVALID;
    set_alpha_kernel<<<(unsigned)nm_div_up64(n, 256), 256, 0, (cudaStream_t)stream>>>(static_cast<unsigned*>(bgra), n, val);
    NM_LAUNCH_CHECK();
    return NM_OK;
}
