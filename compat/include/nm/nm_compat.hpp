// nm_compat.hpp -- the public C++ surface of gift-surg/NiftyMatch's feature pipeline, served by
// the B200-native C-ABI library (nm_b200.h).
//
// Downstream projects include the reference's header NAMES (siftfunctions.h, pyramidata.h,
// convolution.h, match.h, ...): each of those files in this directory is a one-line forwarder
// to this header, which declares the same classes (same public members, same names and types)
// and the same free functions (same signatures and default arguments) as the reference:
//   classes   src/gpu/sift/{siftparams.h:14-99, siftdata.h:20-111, pyramidata.h:15-131},
//             src/gpu/utils/{cudatex2D.h:11-53, cudatimer.h:14-40, exception.h:27-86}
//   functions src/gpu/sift/siftfunctions.h:19-101, src/gpu/kernels/{convolution.h:20,
//             downsample.h, cudamath.h:18-87, keypoint.h:25-63, orientation.h:19,
//             descriptor.h:25, match.h:19-46, transpose.h:17, ransac.h:8-22}
// The bodies (compat/src/*.cu) unwrap thrust vectors to raw pointers, call the C-ABI and turn a
// non-zero status into the reference's exception type.  ransac.h (align_points, ransac_*) is served
// too, and the pipeline-input functions of bgra_2_gray.h, cast.h, undistort.h and
// resample.h (ranks 2 and 4), cudautils.h.
#ifndef NM_COMPAT_HPP
#define NM_COMPAT_HPP

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_runtime_api.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

// ---- utils/macros.h ------------------------------------------------------------------------
#ifndef DISALLOW_COPY_AND_ASSIGNMENT
#define DISALLOW_COPY_AND_ASSIGNMENT(TypeName) \
    TypeName(const TypeName&) = delete;        \
    void operator=(const TypeName&) = delete
#endif

// ---- gpu/utils/exception.h -------------------------------------------------------------------
// Exception<E>::throw_it(file, line, text) throws an E-derived exception whose what() names the
// location; the three macros are what library code and clients use.
template <class Std_Exception>
class Exception : public Std_Exception
{
public:
    static void throw_it(const char* file, const int line, const char* detailed = "-")
    {
        std::ostringstream text;
        text << "Exception in file '" << file << "' in line " << line << "\n"
             << "Detailed description: " << detailed << "\n";
        throw Exception(text.str());
    }
    static void throw_it(const char* file, const int line, const std::string& detailed)
    {
        throw_it(file, line, detailed.c_str());
    }
    virtual ~Exception() throw() {}

private:
    Exception() : Std_Exception("Unknown Exception.\n") {}
    explicit Exception(const std::string& what) : Std_Exception(what) {}
};

template <class Exception_Typ>
inline void handleException(const Exception_Typ& ex)
{
    std::cerr << ex.what() << std::endl;
    std::exit(EXIT_FAILURE);
}

#define RUNTIME_EXCEPTION(msg) Exception<std::runtime_error>::throw_it(__FILE__, __LINE__, msg)
#define LOGIC_EXCEPTION(msg) Exception<std::logic_error>::throw_it(__FILE__, __LINE__, msg)
#define RANGE_EXCEPTION(msg) Exception<std::range_error>::throw_it(__FILE__, __LINE__, msg)

// ---- gpu/utils/cudatex2D.h, cudatimer.h ------------------------------------------------------
// RAII texture object over a cudaArray: linear filtering, border addressing, unnormalised
// coordinates (what the reference's keypoint detector and mask sampling expect).
class CudaTex2D
{
public:
    CudaTex2D() : _tex(0) {}
    CudaTex2D(cudaArray* array);
    ~CudaTex2D();
    void set(cudaArray* array, cudaTextureReadMode read_mode = cudaReadModeNormalizedFloat);
    void release();
    operator cudaTextureObject_t() const { return _tex; }

private:
    cudaTextureObject_t _tex;
    DISALLOW_COPY_AND_ASSIGNMENT(CudaTex2D);
};

// cudaEvent stopwatch on a stream; stop() returns milliseconds.
class CudaTimer
{
public:
    CudaTimer(cudaStream_t stream = 0);
    ~CudaTimer();
    void start();
    float stop();

private:
    cudaEvent_t  _start;
    cudaEvent_t  _stop;
    cudaStream_t _stream;
    DISALLOW_COPY_AND_ASSIGNMENT(CudaTimer);
};

// ---- gpu/sift/siftparams.h -------------------------------------------------------------------
#define MINIMUM_OCTAVE_SIZE 32

// All members are public and mutable, as clients set thresholds / octave counts directly.
// The derivation lives in the C-ABI (nm_sift_params_init) so that both sides agree bit for bit.
class SiftParams
{
public:
    SiftParams() : _width(0), _height(0) {}
    SiftParams(int width, int height);

    int   _width;
    int   _height;
    int   _num_octaves;
    int   _num_dog_levels;
    int   _level_max;
    int   _level_min;
    float _sigma_d_0;
    float _sigma_k;
    float _sigma_0;
    float _sigma_n;
    float _base_smooth;
    std::vector<float> _sigmas;
    float _peak_threshold;
    float _edge_threshold;
};

// ---- gpu/sift/siftdata.h ---------------------------------------------------------------------
#define SIFT_VECTOR_SIZE 128
#define MAX_DESCRIPTORS 2048

// Output container: descriptors (row-major, 128 per keypoint), absolute coordinates, match indices.
struct SiftData
{
    thrust::device_vector<float> _desc;
    thrust::device_vector<int>   _match_indexes;
    thrust::device_vector<float> _x;
    thrust::device_vector<float> _y;
    float* _x_ptr;
    float* _y_ptr;
    int*   _match_indexes_ptr;
    int    _num_items;
    int    _capacity;

    SiftData() {}
    SiftData(int capacity);
    ~SiftData();
    void copy_from(const SiftData& in);
    void initialize_data(int capacity = MAX_DESCRIPTORS);
    void clear_data();
};

// ---- gpu/sift/pyramidata.h -------------------------------------------------------------------
#define MAX_KERNEL_LENGTH 91

// Per-octave working set the client loop passes to the sift functions; every buffer is sized for
// the full-resolution image and reused by all octaves (public, like the reference's).
class PyramidData
{
public:
    PyramidData() : _num_octaves(0), _num_dogs(0), _num_kernels(0) {}
    PyramidData(const SiftParams& params);
    ~PyramidData() {}

    void initialize(const SiftParams& params);
    void clear();
    void gpu_collate_keypoints_for_level(int level, int num_pixels);

public:
    thrust::device_vector<float>  _octave[20];        // the Gaussian levels of the current octave
    thrust::device_vector<float>  _dog[19];
    thrust::device_vector<float4> _key_pts[19];       // dense, one entry per pixel
    thrust::device_vector<float2> _orientations[19];
    thrust::device_vector<float>  _base_kernel;
    int                           _base_radius;
    thrust::device_vector<float>  _kernels[20];
    std::vector<int>              _kernel_radii;
    thrust::device_vector<float>  _buffer;
    thrust::device_vector<float2> _grad;
    thrust::device_vector<float4> _collated_kpts[19];
    int _num_octaves;                                 // number of levels per octave (sic)
    int _num_dogs;
    int _num_kernels;

private:
    void generate_kernels(const SiftParams& params);
    void create_kernel_for_sigma(float sigma, thrust::device_vector<float>& result, int& radius);
};

// ---- gpu/kernels: free functions ---------------------------------------------------------------
template <typename TYPE>
void convolve(TYPE* result, const TYPE* image, TYPE* buffer, const int width, const int height,
              const float* kernel, const int kernel_radius, cudaStream_t stream = 0);

template <typename DataType>
void downsample_by_2(DataType* result, const int result_width, const int result_height,
                     const DataType* source, const int source_width, const int source_height,
                     cudaStream_t stream = 0);

extern "C" int DivUp(int a, int b);
extern "C" int DivDown(int a, int b);
extern "C" int AlignUp(int a, int b);
extern "C" int AlignDown(int a, int b);

template <typename TYPE>
void subtract(const TYPE* A, const TYPE* B, TYPE* C, const int width, const int height,
              cudaStream_t stream = 0);

template <typename TYPE>
void gradient(const TYPE* source, float2* result, const int width, const int height,
              cudaStream_t stream = 0);

// wraps into [0, 2 pi]; note the strict comparison on the upper side
inline __host__ __device__ float mod_2pi_f(float x)
{
    const float two_pi = (float)(2 * 3.14159265358979323846);
    while (x > two_pi) x -= two_pi;
    while (x < 0.0F) x += two_pi;
    return x;
}

void find_keypoints(cudaTextureObject_t current, cudaTextureObject_t down, cudaTextureObject_t up,
                    const int width, const int height, const float peak_threshold,
                    const float edge_threshold, const float xper, const float sigma_0,
                    const int num_dogs, const int dog, float4* result, cudaStream_t stream = 0);

void find_keypoints(cudaTextureObject_t current, cudaTextureObject_t mask, cudaTextureObject_t down,
                    cudaTextureObject_t up, const int width, const int height,
                    const float peak_threshold, const float edge_threshold, const float xper,
                    const float sigma_0, const int num_dogs, const int dog, float4* result,
                    cudaStream_t stream = 0);

void detect_orientations(const float4* key_pts, const float2* grad, const int num_pts,
                         const int octave_width, const int octave_height, float gauss_factor,
                         const float xper, float2* result, cudaStream_t stream = 0);

void compute_sift_descriptors(const float4* key_pts, const float2* orients, const float2* grad,
                              const int num_pts, const int octave_width, const int octave_height,
                              const int num_dogs, const float xper, float* desc, float* x, float* y,
                              cudaStream_t stream = 0);

// A is dimension-major (sift_vector_size x size_A); result is the TRANSPOSED distance matrix
template <typename TYPE>
void compute_brute_force_distance(const TYPE* A, const int size_A, const TYPE* B, const int size_B,
                                  const int sift_vector_size, TYPE* result, cudaStream_t stream = 0);

template <typename TYPE>
void get_sift_matches(const TYPE* distance, const int rows, const int cols, const int buffer_width,
                      int* result, float ambiguity = 0.8f, cudaStream_t stream = 0);

template <typename TYPE>
void transpose(TYPE* odata, const TYPE* idata, int width, int height, cudaStream_t stream = 0);

// ---- gpu/kernels/bgra_2_gray.h (cuda_grayscale), cast.h, undistort.h:29, resample.h:36 -----------------
// Served instantiations: the ones the reference instantiates (<float>; cuda_cast<float, unsigned char>).
template <typename OutputType>
void cuda_grayscale(const uchar4* bgra, OutputType* output, const int width, const int height, cudaStream_t stream = 0);
template <typename OutputType>
void cuda_extract_channel(const uchar4* bgra, OutputType* output, const int width, const int height, const int channel,
                          cudaStream_t stream = 0);
template <typename InputType>
void cuda_put_channel(uchar4* bgra, const InputType* input, const int width, const int height, const int channel,
                      cudaStream_t stream = 0);
void cuda_set_alpha_to_const(uchar4* bgra, const int width, const int height, const unsigned char val = 255,
                             cudaStream_t stream = 0);
template <typename FROM, typename TO>
void cuda_cast(const FROM* src, const size_t cols, const size_t rows, TO* dst, TO max_val = 0, cudaStream_t stream = 0);
void cuda_undistort(const float* x, const float* y, const size_t cols, const size_t rows, const float* camera_matrix,
                    const float* distortion_coeffs, float* u, float* v, cudaStream_t stream = 0);
void resample_undistort(cudaTextureObject_t tex, const float* x, const float* y, const size_t cols, const size_t rows,
                        float* undistorted, cudaStream_t stream = 0);

// resample.h:7-23 (mosaic rendering)
void resample_perspective_transform(uchar4* result, cudaTextureObject_t text, const int cols, const int rows, float* x_pos,
                                    float* y_pos, const float* mat3x3, bool inverse = true, cudaStream_t stream = 0);
void resample_mask(unsigned char* result, cudaTextureObject_t text, const int cols, const int rows, const float* x_pos,
                   const float* y_pos, const float threshold = 0.5f, cudaStream_t stream = 0);
void transform_blend(uchar4* canvas, const int cw, const int ch, cudaTextureObject_t frame, const int fw, const int fh,
                     const int nw, const int nh, const float* mat3x3, const int tx, const int ty,
                     cudaTextureObject_t frame_mask, float* canvas_wts, cudaTextureObject_t frame_wts, cudaStream_t stream = 0);

// ---- gpu/utils/cudautils.h ---------------------------------------------------------------------------
class CudaUtils {
public:
    static int get_max_flops_device_id();      // the device with the most SMs x clock (helper_cuda's gpuGetMaxGflopsDeviceId)
    static void setup_CUDA(int device_id);     // cudaSetDevice; the reference's cudaGLSetGLDevice (deprecated interop) is not called
private:
    static int _max_gflops_device_id;
};

// ---- gpu/kernels/ransac.h:8-22 -------------------------------------------------------------------
void align_points(const float* src_x, const float* src_y, const float* dst_x, const float* dst_y,
                  float* c_src_x, float* c_src_y, float* c_dst_x, float* c_dst_y,
                  const int* matches, const int num_pts, cudaStream_t stream = 0);
// The reference seeds std::mt19937 from std::random_device on every call; so do these wrappers, unless the
// environment variable NM_RANSAC_SEED holds a number (reproducible runs).  They block on `stream` to
// return the reference's bool (false: fewer than 2 / 4 usable correspondences, `homography` untouched).
bool ransac_homography(float* src_x, float* src_y, float* dst_x, float* dst_y,
                       const int src_size, const int dst_size, float inlier_threshold,
                       int iterations, float* homography, cudaStream_t stream = 0);
bool ransac_translation(float* src_x, float* src_y, float* dst_x, float* dst_y,
                        const int src_size, const int dst_size, float inlier_threshold,
                        int iterations, float* homography, cudaStream_t stream = 0);
bool ransac_similarity(float* src_x, float* src_y, float* dst_x, float* dst_y,
                       const int src_size, const int dst_size, float inlier_threshold,
                       int iterations, float* homography, cudaStream_t stream = 0);

// ---- gpu/sift/siftfunctions.h ------------------------------------------------------------------
void compute_sift_matches(SiftData* A, SiftData* B, float* distance, float ambiguity = 0.8f,
                          cudaStream_t stream = 0);
void compute_dog(PyramidData& pydata, const int octave_width, const int octave_height,
                 cudaStream_t stream = 0);
void compute_gradients(PyramidData& pydata, const SiftParams& params, const int octave_width,
                       const int octave_height, cudaStream_t stream = 0);
void compute_keypoints(PyramidData& pydata, const SiftParams& params, const int octave,
                       const int octave_width, const int octave_height, cudaStream_t stream = 0);
void compute_keypoints_with_mask(PyramidData& pydata, SiftParams& params, cudaTextureObject_t mask,
                                 const int octave, const int octave_width, const int octave_height,
                                 cudaStream_t stream = 0);
void compute_orientations(PyramidData& pydata, const SiftParams& params, const int octave,
                          const int octave_width, const int octave_height, cudaStream_t stream = 0);
void compute_descriptors(PyramidData& pydata, const SiftParams& params, const int octave,
                         const int octave_width, const int octave_height, SiftData& data,
                         cudaStream_t stream = 0);

#endif // NM_COMPAT_HPP
