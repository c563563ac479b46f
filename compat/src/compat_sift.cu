// compat_sift.cu -- libsift.a of the drop-in layer: SiftParams, SiftData, PyramidData and the seven
// functions of the reference's siftfunctions.h, with the reference's observable behaviour
// (which buffers are filled, early returns, capacity truncation; src/gpu/sift/siftfunctions.cu:15-181,
// pyramidata.cu:24-123, siftdata.cu:3-57) on top of the C-ABI.  What changes underneath:
//   * compute_keypoints needs no cudaArray copies / texture objects (5 allocations + copies per
//     octave in the reference): the detector reads the linear DoG buffers;
//   * compute_sift_matches is one call: no transposes, no second N_A x N_B matrix;
//   * every call honours the caller's stream.
#include <cstdlib>
#include "nm_compat.hpp"
#include "nm_b200.h"

#include <thrust/fill.h>

namespace {
inline void nm_check(int rc, const char* what)
{
    if (rc != NM_OK) RUNTIME_EXCEPTION(std::string(what) + ": " + nm_strerror(rc));
}
template <typename T>
inline T* dev(thrust::device_vector<T>& v) { return thrust::raw_pointer_cast(v.data()); }
} // namespace

// ---- SiftParams ---------------------------------------------------------------------------------
SiftParams::SiftParams(int width, int height) : _width(width), _height(height)
{
    nm_sift_params p;
    nm_check(nm_sift_params_init(&p, width, height), "SiftParams");
    _num_octaves = p.num_octaves; _num_dog_levels = p.num_dog_levels;
    _level_max = p.level_max; _level_min = p.level_min;
    _sigma_d_0 = p.sigma_d_0; _sigma_k = p.sigma_k; _sigma_0 = p.sigma_0; _sigma_n = p.sigma_n;
    _base_smooth = p.base_smooth;
    _sigmas.assign(p.sigmas, p.sigmas + p.num_sigmas);
    _peak_threshold = p.peak_threshold; _edge_threshold = p.edge_threshold;
}

// ---- SiftData -----------------------------------------------------------------------------------
SiftData::SiftData(int capacity)
{
    if (capacity <= 0) throw std::runtime_error("Invalid initialization of SIFT data");
    initialize_data(capacity);
}

SiftData::~SiftData() { clear_data(); }

void SiftData::initialize_data(int capacity)
{
    clear_data();
    _desc.assign((size_t)SIFT_VECTOR_SIZE * capacity, 0.f);
    _match_indexes.assign(capacity, -1);
    _x.resize(capacity);
    _y.resize(capacity);
    _match_indexes_ptr = dev(_match_indexes);
    _x_ptr = dev(_x);
    _y_ptr = dev(_y);
    _capacity = capacity;
    _num_items = 0;
}

void SiftData::clear_data()
{
    _desc.clear(); _match_indexes.clear(); _x.clear(); _y.clear();
    _match_indexes_ptr = NULL;
    _x_ptr = _y_ptr = NULL;
    _num_items = _capacity = 0;
}

// a deep copy (the reference's is one too, whatever its comment says)
void SiftData::copy_from(const SiftData& in)
{
    _desc = in._desc; _x = in._x; _y = in._y; _match_indexes = in._match_indexes;
    _x_ptr = dev(_x); _y_ptr = dev(_y); _match_indexes_ptr = dev(_match_indexes);
    _capacity = in._capacity;
    _num_items = in._num_items;
}

// ---- PyramidData --------------------------------------------------------------------------------
PyramidData::PyramidData(const SiftParams& params) : _num_octaves(0), _num_dogs(0), _num_kernels(0)
{
    initialize(params);
}

void PyramidData::initialize(const SiftParams& params)
{
    clear();
    _num_octaves = params._level_max - params._level_min + 1;       // levels per octave
    if (_num_octaves > 20) RUNTIME_EXCEPTION("Maximum bumber of levels is 20.");
    const size_t n = (size_t)params._width * params._height;
    _num_dogs = params._level_max - params._level_min;
    for (int i = 0; i < _num_octaves; ++i) _octave[i].resize(n);
    for (int i = 0; i < _num_dogs; ++i) _dog[i].resize(n);
    for (int i = 0; i < params._num_dog_levels; ++i) {
        _key_pts[i].assign(n, make_float4(-1.f, -1.f, -1.f, -1.f));
        _collated_kpts[i].assign(n, make_float4(-1.f, -1.f, -1.f, -1.f));
    }
    _grad.assign(n * _num_dogs, make_float2(0.f, 0.f));
    _buffer.resize(n);
    generate_kernels(params);
}

void PyramidData::clear()
{
    for (int i = 0; i < 20; ++i) { _octave[i].clear(); _kernels[i].clear(); }
    for (int i = 0; i < 19; ++i) { _dog[i].clear(); _key_pts[i].clear(); _collated_kpts[i].clear(); _orientations[i].clear(); }
    _grad.clear(); _buffer.clear(); _base_kernel.clear(); _kernel_radii.clear();
    _num_octaves = _num_dogs = _num_kernels = 0;
}

void PyramidData::create_kernel_for_sigma(float sigma, thrust::device_vector<float>& result, int& radius)
{
    float taps[96];
    nm_check(nm_gaussian_taps(sigma, taps, &radius), "create_kernel_for_sigma");
    result.assign(taps, taps + 2 * radius + 1);
}

void PyramidData::generate_kernels(const SiftParams& params)
{
    create_kernel_for_sigma(params._base_smooth, _base_kernel, _base_radius);
    _num_kernels = (int)params._sigmas.size();
    _kernel_radii.assign(_num_kernels, 0);
    for (int i = 0; i < _num_kernels; ++i) create_kernel_for_sigma(params._sigmas[i], _kernels[i], _kernel_radii[i]);
}

// stable compaction of the valid entries (w >= 0) in raster order; sizes _orientations[level]
void PyramidData::gpu_collate_keypoints_for_level(int level, int num_pixels)
{
    thrust::device_vector<int> count(1, 0);
    nm_check(nm_collate_f32(reinterpret_cast<const float*>(dev(_key_pts[level])), num_pixels,
                            reinterpret_cast<float*>(dev(_collated_kpts[level])), dev(count), 0),
             "gpu_collate_keypoints_for_level");
    const int n = count[0];                                        // host-visible count, as in the reference
    _orientations[level].assign(n, make_float2(-1.f, -1.f));
}

// ---- siftfunctions.h ----------------------------------------------------------------------------
// The reference fills the caller's N_A x N_B `distance` buffer (siftfunctions.cu:25-33), so a non-null buffer is
// filled here too -- bitwise, which means 128 sequential fp32 multiply-adds per pair on the CUDA cores (the exact
// engine; the matrix cannot come out of a tensor-core contraction bit for bit).  Clients that only read
// _match_indexes can skip it and get the tcgen05 engine (same indices): pass distance = nullptr (an extension: the
// reference would dereference it), or set NM_COMPAT_SKIP_DISTANCE=1 to leave an unmodified client's buffer
// untouched.
void compute_sift_matches(SiftData* A, SiftData* B, float* distance, float ambiguity, cudaStream_t stream)
{
    static const bool skip_distance = std::getenv("NM_COMPAT_SKIP_DISTANCE") != nullptr;
    nm_check(nm_match_f32(dev(A->_desc), A->_num_items, dev(B->_desc), B->_num_items, ambiguity, dev(A->_match_indexes),
                          skip_distance ? nullptr : distance, stream),
             "compute_sift_matches");
}

void compute_dog(PyramidData& pydata, const int octave_width, const int octave_height, cudaStream_t stream)
{
    for (int i = 0; i < pydata._num_dogs; ++i)
        subtract<float>(dev(pydata._octave[i + 1]), dev(pydata._octave[i]), dev(pydata._dog[i]), octave_width,
                        octave_height, stream);
}

void compute_gradients(PyramidData& pydata, const SiftParams& params, const int octave_width,
                       const int octave_height, cudaStream_t stream)
{
    const size_t plane = (size_t)octave_width * octave_height;
    for (int i = params._level_min + 1; i <= params._level_max - 2; ++i)
        gradient<float>(dev(pydata._octave[i + 1]), dev(pydata._grad) + i * plane, octave_width, octave_height, stream);
}

namespace {
void keypoints_impl(PyramidData& pydata, const SiftParams& params, cudaTextureObject_t mask, int octave, int ow, int oh,
                    cudaStream_t stream)
{
    const float xper = std::pow(2.0, octave);
    for (int i = 1; i < pydata._num_dogs - 1; ++i) {
        // the dense map is reset over its whole (full-resolution) extent, like the reference
        thrust::fill(thrust::cuda::par.on(stream), pydata._key_pts[i - 1].begin(), pydata._key_pts[i - 1].end(),
                     make_float4(-1.f, -1.f, -1.f, -1.f));
        nm_check(nm_keypoints_dense_masked_f32(dev(pydata._dog[i]), dev(pydata._dog[i - 1]), dev(pydata._dog[i + 1]), mask,
                                               ow, oh, params._peak_threshold, params._edge_threshold, xper,
                                               params._sigma_0, params._num_dog_levels, i - 1,
                                               reinterpret_cast<float*>(dev(pydata._key_pts[i - 1])), stream),
                 "compute_keypoints");
    }
}
} // namespace

void compute_keypoints(PyramidData& pydata, const SiftParams& params, const int octave, const int octave_width,
                       const int octave_height, cudaStream_t stream)
{
    keypoints_impl(pydata, params, 0, octave, octave_width, octave_height, stream);
}

void compute_keypoints_with_mask(PyramidData& pydata, SiftParams& params, cudaTextureObject_t mask, const int octave,
                                 const int octave_width, const int octave_height, cudaStream_t stream)
{
    keypoints_impl(pydata, params, mask, octave, octave_width, octave_height, stream);
}

void compute_orientations(PyramidData& pydata, const SiftParams& params, const int octave, const int octave_width,
                          const int octave_height, cudaStream_t stream)
{
    const float xper = std::pow(2.0, octave);
    const int n_pix = octave_width * octave_height;
    cudaStreamSynchronize(stream);                                  // the collation below counts on the host
    for (int i = 0; i < params._num_dog_levels; ++i) {
        pydata.gpu_collate_keypoints_for_level(i, n_pix);
        const int n = (int)pydata._orientations[i].size();
        if (n == 0) return;                                         // the first empty level ends the octave
        detect_orientations(dev(pydata._collated_kpts[i]), dev(pydata._grad), n, octave_width, octave_height, 1.5f, xper,
                            dev(pydata._orientations[i]), stream);
    }
}

void compute_descriptors(PyramidData& pydata, const SiftParams& params, const int octave, const int octave_width,
                         const int octave_height, SiftData& data, cudaStream_t stream)
{
    const float xper = std::pow(2.0, octave);
    for (int i = 0; i < params._num_dog_levels; ++i) {
        int n = (int)pydata._orientations[i].size();
        if (n == 0) return;
        const int capacity = (int)(data._desc.size() / SIFT_VECTOR_SIZE);
        if (n + data._num_items > capacity) n = capacity - data._num_items;   // silent truncation
        if (n <= 0) continue;
        compute_sift_descriptors(dev(pydata._collated_kpts[i]), dev(pydata._orientations[i]), dev(pydata._grad), n,
                                 octave_width, octave_height, params._num_dog_levels, xper,
                                 dev(data._desc) + (size_t)data._num_items * SIFT_VECTOR_SIZE, dev(data._x) + data._num_items,
                                 dev(data._y) + data._num_items, stream);
        data._num_items += n;
    }
}
