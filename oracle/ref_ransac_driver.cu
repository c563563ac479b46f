// oracle/ref_ransac_driver.cu -- TEST INFRASTRUCTURE ONLY (never part of the product path).
//
// Host-array entry points around the reference's registration code (gpu/kernels/ransac.h:8-22),
// compiled by oracle/build_ref.sh into oracle/_ref/libnmref.so.  The reference's estimators draw their
// index list from std::mt19937 seeded by std::random_device (ransac.cu:546-555), so their result is not
// reproducible; to pin parity the reference's OWN hypothesis kernels (translation_kernel,
// similarity_transformation_kernel, homography_kernel, ransac.cu:437-520) are launched here on an index
// list the test supplies -- the source is included textually (resolved at build time from
// /root/reference/src/gpu/kernels, not copied) because the kernels are not declared in any header.
// nmref_ransac calls the public functions unchanged (random list) for the statistical comparison.
//
// -DNM_COMPAT_BUILD compiles the public-API part against the drop-in headers of this repository
// (entry points nmcompat_*).
#ifndef NM_COMPAT_BUILD
#include "ransac.cu"
#define NMREF(name) nmref_##name
#else
#include "ransac.h"
#define NMREF(name) nmcompat_##name
#endif
#include <cuda_runtime.h>
#include <vector>

namespace {
struct DevPts {
    float *sx = nullptr, *sy = nullptr, *dx = nullptr, *dy = nullptr;
    DevPts(const float* hsx, const float* hsy, const float* hdx, const float* hdy, int n_src, int n_dst)
    {
        cudaMalloc(&sx, sizeof(float) * n_src); cudaMalloc(&sy, sizeof(float) * n_src);
        cudaMalloc(&dx, sizeof(float) * n_dst); cudaMalloc(&dy, sizeof(float) * n_dst);
        cudaMemcpy(sx, hsx, sizeof(float) * n_src, cudaMemcpyHostToDevice);
        cudaMemcpy(sy, hsy, sizeof(float) * n_src, cudaMemcpyHostToDevice);
        cudaMemcpy(dx, hdx, sizeof(float) * n_dst, cudaMemcpyHostToDevice);
        cudaMemcpy(dy, hdy, sizeof(float) * n_dst, cudaMemcpyHostToDevice);
    }
    ~DevPts() { cudaFree(sx); cudaFree(sy); cudaFree(dx); cudaFree(dy); }
};
} // namespace

extern "C" {

// align_points on host arrays: src has n_src points (= length of matches), dst n_dst.
int NMREF(align_points)(const float* src_x, const float* src_y, int n_src, const float* dst_x, const float* dst_y,
                        int n_dst, const int* matches, float* c_src_x, float* c_src_y, float* c_dst_x, float* c_dst_y)
{
    DevPts p(src_x, src_y, dst_x, dst_y, n_src, n_dst);
    int* m = nullptr;
    float* c = nullptr;
    cudaMalloc(&m, sizeof(int) * n_src);
    cudaMalloc(&c, sizeof(float) * 4 * n_src);
    cudaMemcpy(m, matches, sizeof(int) * n_src, cudaMemcpyHostToDevice);
    align_points(p.sx, p.sy, p.dx, p.dy, c, c + n_src, c + 2 * n_src, c + 3 * n_src, m, n_src, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(c_src_x, c, sizeof(float) * n_src, cudaMemcpyDeviceToHost);
    cudaMemcpy(c_src_y, c + n_src, sizeof(float) * n_src, cudaMemcpyDeviceToHost);
    cudaMemcpy(c_dst_x, c + 2 * n_src, sizeof(float) * n_src, cudaMemcpyDeviceToHost);
    cudaMemcpy(c_dst_y, c + 3 * n_src, sizeof(float) * n_src, cudaMemcpyDeviceToHost);
    cudaFree(m); cudaFree(c);
    return e == cudaSuccess ? 0 : 1000 + (int)e;
}

// The public estimators, unchanged (random index list): returns 1 / 0 = the reference's bool, or < 0.
int NMREF(ransac)(int kind, const float* src_x, const float* src_y, const float* dst_x, const float* dst_y, int n,
                  float thr, int iterations, float* H9)
{
    DevPts p(src_x, src_y, dst_x, dst_y, n, n);
    float* H = nullptr;
    cudaMalloc(&H, sizeof(float) * 9);
    cudaMemset(H, 0, sizeof(float) * 9);
    bool ok;
    if (kind == 0) ok = ransac_translation(p.sx, p.sy, p.dx, p.dy, n, n, thr, iterations, H, 0);
    else if (kind == 1) ok = ransac_similarity(p.sx, p.sy, p.dx, p.dy, n, n, thr, iterations, H, 0);
    else ok = ransac_homography(p.sx, p.sy, p.dx, p.dy, n, n, thr, iterations, H, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(H9, H, sizeof(float) * 9, cudaMemcpyDeviceToHost);
    cudaFree(H);
    if (e != cudaSuccess) return -(1000 + (int)e);
    return ok ? 1 : 0;
}

#ifndef NM_COMPAT_BUILD
// The reference's hypothesis kernels on the caller's index list, buffers zero-filled like
// thrust::device_vector<...>(n, 0) (ransac.cu:556-561).
int NMREF(ransac_hypotheses)(int kind, const float* src_x, const float* src_y, const float* dst_x, const float* dst_y,
                             int n, const int* rand_list, int iterations, float thr, float* H_out, int* inliers_out)
{
    const int m = kind == 0 ? 1 : kind == 1 ? 2 : 4;
    DevPts p(src_x, src_y, dst_x, dst_y, n, n);
    int *rl = nullptr, *inl = nullptr;
    float* H = nullptr;
    cudaMalloc(&rl, sizeof(int) * iterations * m);
    cudaMalloc(&inl, sizeof(int) * iterations);
    cudaMalloc(&H, sizeof(float) * 9 * iterations);
    cudaMemcpy(rl, rand_list, sizeof(int) * iterations * m, cudaMemcpyHostToDevice);
    cudaMemset(inl, 0, sizeof(int) * iterations);
    cudaMemset(H, 0, sizeof(float) * 9 * iterations);
    const int threads = 256, blocks = DivUp(iterations, threads);
    if (kind == 0) translation_kernel<<<blocks, threads>>>(p.sx, p.sy, p.dx, p.dy, n, H, inl, rl, iterations, thr);
    else if (kind == 1) similarity_transformation_kernel<<<blocks, threads>>>(p.sx, p.sy, p.dx, p.dy, n, H, inl, rl, iterations, thr);
    else homography_kernel<<<blocks, threads>>>(p.sx, p.sy, p.dx, p.dy, n, H, inl, rl, iterations, thr);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(H_out, H, sizeof(float) * 9 * iterations, cudaMemcpyDeviceToHost);
    cudaMemcpy(inliers_out, inl, sizeof(int) * iterations, cudaMemcpyDeviceToHost);
    cudaFree(rl); cudaFree(inl); cudaFree(H);
    return e == cudaSuccess ? 0 : 1000 + (int)e;
}
#endif

} // extern "C"
