// oracle/ref_driver.cu -- TEST INFRASTRUCTURE ONLY (never part of the product path).
//
// A client loop around the UNMODIFIED reference library (gift-surg/NiftyMatch),
// compiled by oracle/build_ref.sh from the sources where they lie under
// /root/reference into oracle/_ref/libnmref.so.  The reference ships no
// "run SIFT on an image" function (clients drive the per-octave calls themselves,
// reference README.md:20-22), so this file is the smallest such client: it only
// calls the reference's public API (src/gpu/sift/siftfunctions.h:19-101,
// src/gpu/kernels/convolution.h:20, downsample.h, match.h) in the order the
// container shapes dictate and copies intermediate
// buffers out so that tests can compare stage by stage.
//
// The library built from this file is the parity pin for oracle/nm_oracle.c and
// for the CUDA product: tests/ and bench.py --impl reference are its only users.
#include "siftfunctions.h"
#include "convolution.h"
#include "downsample.h"
#include "cudamath.h"
#include "match.h"
#include "transpose.h"
#include "cudatex2D.h"
// Textual include of the reference source (not a copy: resolved at build time from
// /root/reference/src/gpu/kernels): gives this client access to BOTH orientation kernels
// the reference ships.  Needed because the kernel its public API launches
// (kernel_orientations_optim via detect_orientations, orientation.cu:11-129,219) calls
// __syncthreads() inside a divergent branch (:68-86) and DEADLOCKS on every sm_70+ GPU
// (independent thread scheduling), B200 included -- observed on the GPU box, see
// DESIGN.md.  The reference's other kernel, kernel_orientations_naive (:132-216), has the
// same arithmetic without the 10-pixel window clamp and runs fine.
//
// -DNM_COMPAT_BUILD compiles this SAME client against the drop-in headers/libraries of this
// repository (compat/, `make compat-client` -> build/compat/libnmcompat.so, entry points
// nmcompat_*): proof that reference client code builds and runs unchanged on the new library.
// In that build only the public API is used (orientation mode 0 works there: no deadlock).
#ifndef NM_COMPAT_BUILD
#include "orientation.cu"
// The reference's public kernel with its two divergent barriers hoisted out of the branch (a patched temporary
// copy of the same source, made and compiled by oracle/build_ref.sh): orientation mode 3.
void detect_orientations_hoisted(const float4* key_pts, const float2* grad, const int num_pts, const int octave_width,
                                 const int octave_height, float gauss_factor, const float xper, float2* result,
                                 cudaStream_t stream);
#define NMREF(name) nmref_##name
#else
#define NMREF(name) nmcompat_##name
#endif

#include <thrust/fill.h>
#include <thrust/copy.h>
#include <thrust/device_vector.h>
#include <cstring>
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace {

// NMREF_VERBOSE=1: synchronise and report after every reference call (to locate hangs).
void trace(const char* what, int o)
{
    static const bool on = std::getenv("NMREF_VERBOSE") != nullptr;
    if (!on) return;
    cudaError_t e = cudaDeviceSynchronize();
    std::fprintf(stderr, "[nmref] octave %d: %s done (%s)\n", o, what, cudaGetErrorString(e));
    std::fflush(stderr);
}

struct Cfg {
    float peak_threshold;
    float edge_threshold;   // <= 0: keep the reference default (10)
    int   num_octaves;      // <= 0: keep the reference default
    int   capacity;         // SiftData capacity (<= 0: MAX_DESCRIPTORS)
    int   clear_grad;       // 1: client zeroes PyramidData::_grad before each octave (defined borders)
    int   orient_mode;      // 0: detect_orientations (public API; deadlocks on sm_70+, refused unless
                            //    NMREF_ALLOW_DEADLOCK is set), 1: the reference's kernel_orientations_naive,
                            // 2: orientations injected by the caller (orient_in),
                            // 3: the reference's public kernel with the barriers hoisted (build_ref.sh; reference build only)
};

// compute_orientations of the reference (siftfunctions.cu:136-152) with a selectable kernel.
void orientations_step(PyramidData& py, const SiftParams& P, int o, int ow, int oh, int mode,
                       const float* orient_in, int* inject_off)
{
    const float xper = std::pow(2.0, o);
    const int n_pix = ow * oh;
    for (int i = 0; i < P._num_dog_levels; ++i) {
        py.gpu_collate_keypoints_for_level(i, n_pix);             // public method (pyramidata.cu:84)
        const int n = (int)py._orientations[i].size();
        if (n == 0) return;                                       // siftfunctions.cu:145
        float4* key_pts = thrust::raw_pointer_cast(py._collated_kpts[i].data());
        float2* orient = thrust::raw_pointer_cast(py._orientations[i].data());
        float2* grad = thrust::raw_pointer_cast(py._grad.data());
        if (mode == 1) {
#ifndef NM_COMPAT_BUILD
            kernel_orientations_naive<<<(n + 127) / 128, 128>>>(key_pts, grad, n, ow, oh, 1.5f, xper, orient);
#endif
        } else if (mode == 2) {
            cudaMemcpy(orient, orient_in + 2 * (size_t)(*inject_off), n * sizeof(float2), cudaMemcpyHostToDevice);
            *inject_off += n;
        } else if (mode == 3) {
#ifndef NM_COMPAT_BUILD
            detect_orientations_hoisted(key_pts, grad, n, ow, oh, 1.5f, xper, orient, 0);
#endif
        } else {
            detect_orientations(key_pts, grad, n, ow, oh, 1.5f, xper, orient);
        }
    }
}

template <typename T>
T* raw(thrust::device_vector<T>& v) { return thrust::raw_pointer_cast(v.data()); }

SiftParams make_params(int w, int h, const Cfg& c)
{
    SiftParams P(w, h);
    P._peak_threshold = c.peak_threshold;
    if (c.edge_threshold > 0.f) P._edge_threshold = c.edge_threshold;
    if (c.num_octaves > 0) P._num_octaves = c.num_octaves;
    return P;
}

// One frame through the reference, image already on the device.
// dump pointers may be null.
void run_frame(const float* image_dev, int w, int h, const SiftParams& P, PyramidData& py,
               SiftData& data, int clear_grad, int orient_mode, const float* orient_in,
               float* levels_out, float* kpts_out, float* orient_out, int* seg_counts, int kp_cap,
               float* grad_out, cudaTextureObject_t mask = 0)
{
    data._num_items = 0;
    int inject_off = 0;
    convolve<float>(raw(py._octave[0]), image_dev, raw(py._buffer), w, h,
                    raw(py._base_kernel), py._base_radius);
    size_t lev_off = 0, grad_off = 0;
    int kp_off = 0;
    for (int o = 0; o < P._num_octaves; ++o) {
        const int ow = w >> o, oh = h >> o;
        const size_t n = (size_t)ow * oh;
        for (int i = 0; i < py._num_kernels; ++i)
            convolve<float>(raw(py._octave[i + 1]), raw(py._octave[i]), raw(py._buffer), ow, oh,
                            raw(py._kernels[i]), py._kernel_radii[i]);
        trace("convolve x5", o);
        compute_dog(py, ow, oh);
        trace("compute_dog", o);
        if (clear_grad) thrust::fill(py._grad.begin(), py._grad.end(), make_float2(0.f, 0.f));
        compute_gradients(py, P, ow, oh);
        trace("compute_gradients", o);
        if (mask) {
            SiftParams Pm = P;                                   // the masked entry takes a non-const reference
            compute_keypoints_with_mask(py, Pm, mask, o, ow, oh);
        } else {
            compute_keypoints(py, P, o, ow, oh);
        }
        trace("compute_keypoints", o);
        if (orient_mode == 0) compute_orientations(py, P, o, ow, oh);
        else orientations_step(py, P, o, ow, oh, orient_mode, orient_in, &inject_off);
        trace("compute_orientations", o);
        compute_descriptors(py, P, o, ow, oh, data);
        trace("compute_descriptors", o);
        cudaDeviceSynchronize();

        if (levels_out) {
            for (int i = 0; i < py._num_octaves; ++i) {
                cudaMemcpy(levels_out + lev_off, raw(py._octave[i]), n * sizeof(float),
                           cudaMemcpyDeviceToHost);
                lev_off += n;
            }
        }
        if (grad_out) {
            cudaMemcpy(grad_out + grad_off, raw(py._grad), 3 * n * sizeof(float2),
                       cudaMemcpyDeviceToHost);
            grad_off += 3 * n * 2;
        }
        if (seg_counts) {
            // Levels after the first empty one are never collated in this octave
            // (reference siftfunctions.cu:145 returns): report them as 0.
            bool stopped = false;
            for (int l = 0; l < P._num_dog_levels; ++l) {
                int cnt = stopped ? 0 : (int)py._orientations[l].size();
                if (cnt == 0) stopped = true;
                seg_counts[o * P._num_dog_levels + l] = cnt;
                int take = cnt;
                if (kp_off + take > kp_cap) take = kp_cap - kp_off;
                if (take > 0 && kpts_out)
                    cudaMemcpy(kpts_out + 4 * (size_t)kp_off, raw(py._collated_kpts[l]),
                               take * sizeof(float4), cudaMemcpyDeviceToHost);
                if (take > 0 && orient_out)
                    cudaMemcpy(orient_out + 2 * (size_t)kp_off, raw(py._orientations[l]),
                               take * sizeof(float2), cudaMemcpyDeviceToHost);
                if (take > 0) kp_off += take;
            }
        }
        if (o + 1 < P._num_octaves)
            downsample_by_2<float>(raw(py._octave[0]), ow / 2, oh / 2,
                                   raw(py._octave[P._num_dog_levels]), ow, oh);
    }
    cudaDeviceSynchronize();
}

} // namespace

extern "C" {

// Parameter derivation of the reference (siftparams.h:30-51), for cross-checking ports.
int NMREF(params)(int w, int h, int* num_octaves, float* sigma_k, float* sigma_0, float* sigma_d_0,
                 float* base_smooth, float* sigmas5)
{
    SiftParams P(w, h);
    *num_octaves = P._num_octaves; *sigma_k = P._sigma_k; *sigma_0 = P._sigma_0;
    *sigma_d_0 = P._sigma_d_0; *base_smooth = P._base_smooth;
    for (size_t i = 0; i < P._sigmas.size() && i < 5; ++i) sigmas5[i] = P._sigmas[i];
    return (int)P._sigmas.size();
}

// Gaussian taps as PyramidData builds them (pyramidata.cu:105-123).
// which = -1: base kernel, 0..4: level kernels.  Returns the radius.
int NMREF(taps)(int w, int h, int which, float* taps_out)
{
    SiftParams P(w, h);
    PyramidData py(P);
    thrust::device_vector<float>& k = which < 0 ? py._base_kernel : py._kernels[which];
    int r = which < 0 ? py._base_radius : py._kernel_radii[which];
    cudaMemcpy(taps_out, raw(k), k.size() * sizeof(float), cudaMemcpyDeviceToHost);
    return r;
}

int NMREF(convolve)(float* result_host, const float* image_host, int w, int h,
                   const float* taps_host, int radius)
{
    thrust::device_vector<float> img(image_host, image_host + (size_t)w * h);
    // one extra row of slack: the reference column kernel touches one row past the
    // image when width % 16 != 0
    thrust::device_vector<float> buf((size_t)w * (h + 1) + 64, 0.f), res((size_t)w * (h + 1) + 64, 0.f);
    thrust::device_vector<float> taps(taps_host, taps_host + 2 * radius + 1);
    convolve<float>(raw(res), raw(img), raw(buf), w, h, raw(taps), radius);
    cudaDeviceSynchronize();
    cudaMemcpy(result_host, raw(res), (size_t)w * h * sizeof(float), cudaMemcpyDeviceToHost);
    return 0;
}

// Whole frame, host in / host out, with optional stage dumps.
//   levels_out : per octave, 6 levels of ow*oh floats, concatenated
//   kpts_out   : float4 per keypoint, concatenated over (octave, level) segments
//   orient_out : float2 per keypoint, same order
//   seg_counts : [num_octaves*3]
//   grad_out   : per octave, 3*ow*oh float2
//   orient_in  : (orient_mode 2) float2 per keypoint in segment order, injected
//   cfg6 = {peak, edge(<=0 default), num_octaves(<=0 default), capacity, clear_grad, orient_mode}
int NMREF(sift_frame)(const float* image_host, int w, int h, const float* cfg6,
                     float* desc, float* x, float* y, int* num_items,
                     float* levels_out, float* kpts_out, float* orient_out, int* seg_counts,
                     int kp_cap, float* grad_out, const float* orient_in)
{
    Cfg c = { cfg6[0], cfg6[1], (int)cfg6[2], (int)cfg6[3], (int)cfg6[4], (int)cfg6[5] };
#ifndef NM_COMPAT_BUILD
    if (c.orient_mode == 0 && !std::getenv("NMREF_ALLOW_DEADLOCK")) return -1;
#else
    if (c.orient_mode == 1 || c.orient_mode == 3) return -3;     // the reference's non-public / patched kernels do not exist here
#endif
    if (c.orient_mode == 2 && !orient_in) return -2;
    SiftParams P = make_params(w, h, c);
    PyramidData py(P);
    SiftData data(c.capacity > 0 ? c.capacity : MAX_DESCRIPTORS);
    thrust::device_vector<float> img(image_host, image_host + (size_t)w * h);
    run_frame(raw(img), w, h, P, py, data, c.clear_grad, c.orient_mode, orient_in, levels_out, kpts_out,
              orient_out, seg_counts, kp_cap, grad_out);
    const int n = data._num_items;
    *num_items = n;
    if (n > 0) {
        if (desc) cudaMemcpy(desc, raw(data._desc), (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost);
        if (x) cudaMemcpy(x, raw(data._x), n * sizeof(float), cudaMemcpyDeviceToHost);
        if (y) cudaMemcpy(y, raw(data._y), n * sizeof(float), cudaMemcpyDeviceToHost);
    }
    return P._num_octaves;
}

// The same with the reference's masked detector (compute_keypoints_with_mask, siftfunctions.cu:65-98):
// mask_host = w*h floats, bound the way a client binds it -- a float cudaArray behind the reference's own
// CudaTex2D (linear filter, border addressing, unnormalised coordinates, element-type reads, cudatex2D.cu:4-21).
int NMREF(sift_frame_masked)(const float* image_host, int w, int h, const float* cfg6,
                            float* desc, float* x, float* y, int* num_items,
                            float* levels_out, float* kpts_out, float* orient_out, int* seg_counts,
                            int kp_cap, float* grad_out, const float* orient_in, const float* mask_host)
{
    Cfg c = { cfg6[0], cfg6[1], (int)cfg6[2], (int)cfg6[3], (int)cfg6[4], (int)cfg6[5] };
#ifndef NM_COMPAT_BUILD
    if (c.orient_mode == 0 && !std::getenv("NMREF_ALLOW_DEADLOCK")) return -1;
#else
    if (c.orient_mode == 1 || c.orient_mode == 3) return -3;
#endif
    if (c.orient_mode == 2 && !orient_in) return -2;
    if (!mask_host) return -4;
    SiftParams P = make_params(w, h, c);
    PyramidData py(P);
    SiftData data(c.capacity > 0 ? c.capacity : MAX_DESCRIPTORS);
    thrust::device_vector<float> img(image_host, image_host + (size_t)w * h);
    cudaArray* arr = nullptr;
    cudaChannelFormatDesc desc_f = cudaCreateChannelDesc<float>();
    if (cudaMallocArray(&arr, &desc_f, w, h) != cudaSuccess) return -5;
    cudaMemcpy2DToArray(arr, 0, 0, mask_host, (size_t)w * sizeof(float), (size_t)w * sizeof(float), h,
                        cudaMemcpyHostToDevice);
    int n = 0;
    {
        CudaTex2D tex;
        tex.set(arr, cudaReadModeElementType);                   // float texels read as they are
        run_frame(raw(img), w, h, P, py, data, c.clear_grad, c.orient_mode, orient_in, levels_out, kpts_out,
                  orient_out, seg_counts, kp_cap, grad_out, (cudaTextureObject_t)tex);
        n = data._num_items;
        *num_items = n;
        if (n > 0) {
            if (desc) cudaMemcpy(desc, raw(data._desc), (size_t)n * 128 * sizeof(float), cudaMemcpyDeviceToHost);
            if (x) cudaMemcpy(x, raw(data._x), n * sizeof(float), cudaMemcpyDeviceToHost);
            if (y) cudaMemcpy(y, raw(data._y), n * sizeof(float), cudaMemcpyDeviceToHost);
        }
    }
    cudaFreeArray(arr);
    return P._num_octaves;
}

// Timing arm: frames resident on the device (frames_dev: n_frames * w*h floats).
// Returns total milliseconds (cudaEvent) for `iters` passes over the batch, one
// PyramidData / SiftData reused across frames as a client would.
int NMREF(sift_bench)(const float* frames_dev, int n_frames, int w, int h, const float* cfg6,
                     int iters, float* ms_out, long long* total_items)
{
    Cfg c = { cfg6[0], cfg6[1], (int)cfg6[2], (int)cfg6[3], (int)cfg6[4], (int)cfg6[5] };
#ifndef NM_COMPAT_BUILD
    if (c.orient_mode != 1 && c.orient_mode != 3) return -1;     // only the runnable configurations can be timed
#else
    if (c.orient_mode != 0) return -1;
#endif
    SiftParams P = make_params(w, h, c);
    PyramidData py(P);
    SiftData data(c.capacity > 0 ? c.capacity : MAX_DESCRIPTORS);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    long long items = 0;
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for (int it = 0; it < iters; ++it)
        for (int f = 0; f < n_frames; ++f) {
            run_frame(frames_dev + (size_t)f * w * h, w, h, P, py, data, c.clear_grad, c.orient_mode, nullptr,
                      nullptr, nullptr, nullptr, nullptr, 0, nullptr);
            items += data._num_items;
        }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(ms_out, e0, e1);
    *total_items = items;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

// compute_sift_matches on host data.  match_io is in/out (the reference leaves entries
// untouched when min2 <= 0, match.cu:107).  dist_out may be null (nA*nB floats otherwise).
int NMREF(match)(const float* A_host, int nA, const float* B_host, int nB, float ambiguity,
                int* match_io, float* dist_out)
{
    SiftData A(nA), B(nB);
    cudaMemcpy(raw(A._desc), A_host, (size_t)nA * 128 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(raw(B._desc), B_host, (size_t)nB * 128 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(raw(A._match_indexes), match_io, nA * sizeof(int), cudaMemcpyHostToDevice);
    A._num_items = nA; B._num_items = nB;
    thrust::device_vector<float> dist((size_t)nA * nB);
    compute_sift_matches(&A, &B, raw(dist), ambiguity);
    cudaDeviceSynchronize();
    cudaMemcpy(match_io, raw(A._match_indexes), nA * sizeof(int), cudaMemcpyDeviceToHost);
    if (dist_out) cudaMemcpy(dist_out, raw(dist), (size_t)nA * nB * sizeof(float), cudaMemcpyDeviceToHost);
    return 0;
}

// Timing arm for the matcher: descriptors resident on the device.
int NMREF(match_bench)(const float* A_dev, int nA, const float* B_dev, int nB, float ambiguity,
                      int iters, float* ms_out)
{
    SiftData A(nA), B(nB);
    cudaMemcpy(raw(A._desc), A_dev, (size_t)nA * 128 * sizeof(float), cudaMemcpyDeviceToDevice);
    cudaMemcpy(raw(B._desc), B_dev, (size_t)nB * 128 * sizeof(float), cudaMemcpyDeviceToDevice);
    A._num_items = nA; B._num_items = nB;
    thrust::device_vector<float> dist((size_t)nA * nB);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    compute_sift_matches(&A, &B, raw(dist), ambiguity);   // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for (int it = 0; it < iters; ++it) compute_sift_matches(&A, &B, raw(dist), ambiguity);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

} // extern "C"
