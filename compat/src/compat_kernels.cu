// compat_kernels.cu -- libkernels.a of the drop-in layer: the reference's kernel launchers
// (src/gpu/kernels/*.h) as thin wrappers over the C-ABI of libnm_b200 (include/nm_b200.h).
// No kernel lives here: a wrapper checks nothing the C-ABI does not check, forwards the caller's
// stream, and turns a non-zero status into the reference's RUNTIME_EXCEPTION.
#include "nm_compat.hpp"
#include "nm_b200.h"

namespace {
inline void nm_check(int rc, const char* what)
{
    if (rc != NM_OK) RUNTIME_EXCEPTION(std::string(what) + ": " + nm_strerror(rc));
}
} // namespace

// convolution.h:20
template <>
void convolve<float>(float* result, const float* image, float* buffer, const int width, const int height,
                     const float* kernel, const int kernel_radius, cudaStream_t stream)
{
    nm_check(nm_blur_f32(result, image, buffer, width, height, kernel, kernel_radius, stream), "convolve");
}

// downsample.h (float only: the uchar4 instantiation serves the out-of-scope colour path)
template <>
void downsample_by_2<float>(float* result, const int result_width, const int result_height, const float* source,
                            const int source_width, const int source_height, cudaStream_t stream)
{
    nm_check(nm_downsample2_f32(result, result_width, result_height, source, source_width, source_height, stream),
             "downsample_by_2");
}

// cudamath.h:18-45
extern "C" int DivUp(int a, int b) { return (a + b - 1) / b; }
extern "C" int DivDown(int a, int b) { return a / b; }
extern "C" int AlignUp(int a, int b) { return DivUp(a, b) * b; }
extern "C" int AlignDown(int a, int b) { return DivDown(a, b) * b; }

template <>
void subtract<float>(const float* A, const float* B, float* C, const int width, const int height, cudaStream_t stream)
{
    nm_check(nm_subtract_f32(A, B, C, width, height, stream), "subtract");
}

template <>
void gradient<float>(const float* source, float2* result, const int width, const int height, cudaStream_t stream)
{
    nm_check(nm_gradient_f32(source, reinterpret_cast<float*>(result), width, height, stream), "gradient");
}

// keypoint.h:25, :52
void find_keypoints(cudaTextureObject_t current, cudaTextureObject_t down, cudaTextureObject_t up, const int width,
                    const int height, const float peak_threshold, const float edge_threshold, const float xper,
                    const float sigma_0, const int num_dogs, const int dog, float4* result, cudaStream_t stream)
{
    nm_check(nm_keypoints_dense_tex(current, 0, down, up, width, height, peak_threshold, edge_threshold, xper, sigma_0,
                                    num_dogs, dog, reinterpret_cast<float*>(result), stream),
             "find_keypoints");
}

void find_keypoints(cudaTextureObject_t current, cudaTextureObject_t mask, cudaTextureObject_t down,
                    cudaTextureObject_t up, const int width, const int height, const float peak_threshold,
                    const float edge_threshold, const float xper, const float sigma_0, const int num_dogs,
                    const int dog, float4* result, cudaStream_t stream)
{
    nm_check(nm_keypoints_dense_tex(current, mask, down, up, width, height, peak_threshold, edge_threshold, xper,
                                    sigma_0, num_dogs, dog, reinterpret_cast<float*>(result), stream),
             "find_keypoints (masked)");
}

// orientation.h:19 -- zero keypoints is a no-op here (the reference would launch an empty grid and abort)
void detect_orientations(const float4* key_pts, const float2* grad, const int num_pts, const int octave_width,
                         const int octave_height, float gauss_factor, const float xper, float2* result,
                         cudaStream_t stream)
{
    if (num_pts <= 0) return;
    nm_check(nm_orientations_f32(reinterpret_cast<const float*>(key_pts), reinterpret_cast<const float*>(grad), num_pts,
                                 octave_width, octave_height, gauss_factor, xper, reinterpret_cast<float*>(result),
                                 stream),
             "detect_orientations");
}

// descriptor.h:25
void compute_sift_descriptors(const float4* key_pts, const float2* orients, const float2* grad, const int num_pts,
                              const int octave_width, const int octave_height, const int num_dogs, const float xper,
                              float* desc, float* x, float* y, cudaStream_t stream)
{
    if (num_pts <= 0) return;
    nm_check(nm_descriptors_f32(reinterpret_cast<const float*>(key_pts), reinterpret_cast<const float*>(orients),
                                reinterpret_cast<const float*>(grad), num_pts, octave_width, octave_height, num_dogs,
                                xper, desc, x, y, stream),
             "compute_sift_descriptors");
}

// match.h:19, :41 and transpose.h:17
template <>
void compute_brute_force_distance<float>(const float* A, const int size_A, const float* B, const int size_B,
                                         const int sift_vector_size, float* result, cudaStream_t stream)
{
    nm_check(nm_dist2_f32(A, size_A, B, size_B, sift_vector_size, result, stream), "compute_brute_force_distance");
}

template <>
void get_sift_matches<float>(const float* distance, const int rows, const int cols, const int buffer_width, int* result,
                             float ambiguity, cudaStream_t stream)
{
    nm_check(nm_set_matches_f32(distance, rows, cols, buffer_width, result, ambiguity, stream), "get_sift_matches");
}

template <>
void transpose<float>(float* odata, const float* idata, int width, int height, cudaStream_t stream)
{
    nm_check(nm_transpose_f32(odata, idata, width, height, stream), "transpose");
}

// ransac.h:8-22 (ransac.cu:50-59, :526-694)
#include <random>
void align_points(const float* src_x, const float* src_y, const float* dst_x, const float* dst_y, float* c_src_x,
                  float* c_src_y, float* c_dst_x, float* c_dst_y, const int* matches, const int num_pts, cudaStream_t stream)
{
    nm_check(nm_align_points_f32(src_x, src_y, dst_x, dst_y, c_src_x, c_src_y, c_dst_x, c_dst_y, matches, num_pts, stream),
             "align_points");
}

namespace {
bool ransac_any(int kind, const float* src_x, const float* src_y, const float* dst_x, const float* dst_y, int src_size,
                float inlier_threshold, int iterations, float* homography, cudaStream_t stream, const char* what)
{
    unsigned long long seed;
    if (const char* e = std::getenv("NM_RANSAC_SEED")) seed = std::strtoull(e, nullptr, 0);
    else {
        std::random_device seeder;                       // ransac.cu:546: a fresh seed per call
        seed = ((unsigned long long)seeder() << 32) | seeder();
    }
    int* status = nullptr;
    if (cudaMallocAsync(&status, 3 * sizeof(int), stream) != cudaSuccess) RUNTIME_EXCEPTION(std::string(what) + ": allocation failed");
    const int rc = nm_ransac_f32(kind, src_x, src_y, dst_x, dst_y, src_size, inlier_threshold, iterations, seed, homography,
                                 status, stream);
    int h[3] = {0, 0, 0};
    cudaMemcpyAsync(h, status, sizeof(h), cudaMemcpyDeviceToHost, stream);
    cudaFreeAsync(status, stream);
    cudaStreamSynchronize(stream);
    nm_check(rc, what);
    return h[0] != 0;
}
} // namespace

bool ransac_homography(float* src_x, float* src_y, float* dst_x, float* dst_y, const int src_size, const int,
                       float inlier_threshold, int iterations, float* homography, cudaStream_t stream)
{
    return ransac_any(NM_RANSAC_HOMOGRAPHY, src_x, src_y, dst_x, dst_y, src_size, inlier_threshold, iterations, homography,
                      stream, "ransac_homography");
}
bool ransac_translation(float* src_x, float* src_y, float* dst_x, float* dst_y, const int src_size, const int,
                        float inlier_threshold, int iterations, float* homography, cudaStream_t stream)
{
    return ransac_any(NM_RANSAC_TRANSLATION, src_x, src_y, dst_x, dst_y, src_size, inlier_threshold, iterations, homography,
                      stream, "ransac_translation");
}
bool ransac_similarity(float* src_x, float* src_y, float* dst_x, float* dst_y, const int src_size, const int,
                       float inlier_threshold, int iterations, float* homography, cudaStream_t stream)
{
    return ransac_any(NM_RANSAC_SIMILARITY, src_x, src_y, dst_x, dst_y, src_size, inlier_threshold, iterations, homography,
                      stream, "ransac_similarity");
}

// bgra_2_gray.h / cast.h / undistort.h / resample.h
template <>
void cuda_grayscale<float>(const uchar4* bgra, float* output, const int width, const int height, cudaStream_t stream)
{
    nm_check(nm_grayscale_bgra_f32(bgra, output, width, height, stream), "cuda_grayscale");
}
template <>
void cuda_cast<float, unsigned char>(const float* src, const size_t cols, const size_t rows, unsigned char* dst,
                                     unsigned char max_val, cudaStream_t stream)
{
    nm_check(nm_cast_f32_u8(src, (int)cols, (int)rows, dst, max_val, stream), "cuda_cast");
}
void cuda_undistort(const float* x, const float* y, const size_t cols, const size_t rows, const float* camera_matrix,
                    const float* distortion_coeffs, float* u, float* v, cudaStream_t stream)
{
    nm_check(nm_undistort_map_f32(x, y, (int)cols, (int)rows, camera_matrix, distortion_coeffs, u, v, stream), "cuda_undistort");
}
void resample_undistort(cudaTextureObject_t tex, const float* x, const float* y, const size_t cols, const size_t rows,
                        float* undistorted, cudaStream_t stream)
{
    nm_check(nm_resample_tex_f32(tex, x, y, (int)cols, (int)rows, undistorted, stream), "resample_undistort");
}

// resample.h:7-23
void resample_perspective_transform(uchar4* result, cudaTextureObject_t text, const int cols, const int rows, float* x_pos,
                                    float* y_pos, const float* mat3x3, bool inverse, cudaStream_t stream)
{
    nm_check(nm_resample_perspective_bgra(result, text, cols, rows, x_pos, y_pos, mat3x3, inverse ? 1 : 0, stream),
             "resample_perspective_transform");
}
void resample_mask(unsigned char* result, cudaTextureObject_t text, const int cols, const int rows, const float* x_pos,
                   const float* y_pos, const float threshold, cudaStream_t stream)
{
    nm_check(nm_resample_mask_tex_u8(result, text, cols, rows, x_pos, y_pos, threshold, stream), "resample_mask");
}
void transform_blend(uchar4* canvas, const int cw, const int ch, cudaTextureObject_t frame, const int fw, const int fh,
                     const int nw, const int nh, const float* mat3x3, const int tx, const int ty, cudaTextureObject_t frame_mask,
                     float* canvas_wts, cudaTextureObject_t frame_wts, cudaStream_t stream)
{
    nm_check(nm_transform_blend_bgra(canvas, cw, ch, frame, fw, fh, nw, nh, mat3x3, tx, ty, frame_mask, canvas_wts, frame_wts, stream),
             "transform_blend");
}

// bgra_2_gray.h: colour-channel helpers; downsample.h: the uchar4 instantiation (a 32-bit word copy, like float)
template <>
void cuda_extract_channel<float>(const uchar4* bgra, float* output, const int width, const int height, const int channel,
                                 cudaStream_t stream)
{
    nm_check(nm_bgra_extract_channel_f32(bgra, output, width, height, channel, stream), "cuda_extract_channel");
}
template <>
void cuda_put_channel<float>(uchar4* bgra, const float* input, const int width, const int height, const int channel,
                             cudaStream_t stream)
{
    nm_check(nm_bgra_put_channel_f32(bgra, input, width, height, channel, stream), "cuda_put_channel");
}
void cuda_set_alpha_to_const(uchar4* bgra, const int width, const int height, const unsigned char val, cudaStream_t stream)
{
    nm_check(nm_bgra_set_alpha(bgra, width, height, val, stream), "cuda_set_alpha_to_const");
}
template <>
void downsample_by_2<uchar4>(uchar4* result, const int result_width, const int result_height, const uchar4* source,
                             const int source_width, const int source_height, cudaStream_t stream)
{
    nm_check(nm_downsample2_f32(reinterpret_cast<float*>(result), result_width, result_height,
                                reinterpret_cast<const float*>(source), source_width, source_height, stream),
             "downsample_by_2<uchar4>");
}
