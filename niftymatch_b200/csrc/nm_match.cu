// nm_match.cu -- brute-force k=2 ratio-test matching: exact fp32 engine, record merge,
// ratio rule, and the compat operators (transpose, distance matrix, set_matches).
//
// Replaces compute_sift_matches (gpu/sift/siftfunctions.cu:15-40), brute_force_distance /
// set_matches (gpu/kernels/match.cu:14-117) and transpose (gpu/kernels/transpose.cu:9-40).
//
// The reference materialises two N_A x N_B fp32 matrices and two transposes per call and
// scans rows with one thread each.  Here the scan keeps a running top-2 per query row in
// registers; nothing of size N_A x N_B touches HBM unless the caller asks for `distance`.
//
// Exactness contract (shared with the tensor-core engine in nm_match_tc.cu): a record
// holds the TRUE two smallest squared distances of the row, each computed exactly as the
// reference does (i = 0..127 sequential, t = a-b, acc = fma(t,t,acc); match.cu:36-42 is
// sub + FFMA in its SASS), ties resolved to the lowest index.  The reference's sequential
// scan (match.cu:88-105) is then reproduced from the record:
//     min1 = d1, idx = i1,  min2 = (i1 == 0) ? min(2139095040.0f, d2) : d2
// (when the first column is the minimum, min2 never loses its odd start value
// 0x7f800000-as-int; otherwise the first displacement overwrites it -- see DESIGN.md).
#include "nm_match.cuh"
#include <math.h>

namespace {

constexpr int MT = 64;            // tile: 64 query rows x 64 database rows
constexpr int FB_SMALL_MAX = 256; // row-list fallback: lists up to this length take the warp-per-row kernel
constexpr int FB_CHUNK = 64;      // ... with one block per chunk of this many database rows
constexpr int MK = 32;            // dims per staged chunk
constexpr int MP = MT + 4;        // smem pitch

__device__ __forceinline__ void rec_update(float& t1, int& i1, float& t2, float v, int idx)
{
    // idx is increasing within a thread, so strict < keeps the lowest index on ties
    if (v < t1) { t2 = t1; t1 = v; i1 = idx; }
    else if (v < t2) t2 = v;
}

__device__ __forceinline__ void rec_merge(float& t1, int& i1, float& t2, float u1, int j1, float u2)
{
    if (u1 < t1 || (u1 == t1 && j1 < i1)) {
        const float a = t1; const int ai = i1; const float b = t2;
        t1 = u1; i1 = j1; t2 = u2; u1 = a; j1 = ai; u2 = b;
    }
    t2 = fminf(t2, u1);
    (void)j1; (void)u2;
}

// ROWS = false: query rows are 0..nA-1, one row tile per blockIdx.x.
// ROWS = true : query rows are row_list[0 .. *row_count) (device-side count, e.g. the rows the
//               tensor-core engine could not certify); blocks stride over the listed row tiles
//               and records are written by LIST POSITION (part[split * nA + k]).
template <bool WRITE_D, bool ROWS>
__global__ void __launch_bounds__(256) scan_exact_kernel(const float* __restrict__ A, long long a_sa,
                                                         long long a_sk, int nA, const float* __restrict__ B,
                                                         int nB, int dim, int index_offset, int b_tiles_per_split,
                                                         float4* __restrict__ part, float* __restrict__ D,
                                                         long long d_sa, long long d_sb,
                                                         const int* __restrict__ row_list,
                                                         const int* __restrict__ row_count,
                                                         unsigned long long* __restrict__ key1,
                                                         unsigned* __restrict__ key2)
{
    __shared__ __align__(16) float As[MK][MP];
    __shared__ __align__(16) float Bs[MK][MP];
    __shared__ float4 s_rec[MT][17];
    __shared__ int s_rows[MT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int split = blockIdx.y;
    const int n_btiles = (nB + MT - 1) / MT;
    const int bt_beg = split * b_tiles_per_split;
    const int bt_end = min(bt_beg + b_tiles_per_split, n_btiles);
    const int n_rows = ROWS ? min(*row_count, nA) : nA;
    if (ROWS && n_rows <= FB_SMALL_MAX) return;        // short lists: scan_rows_small_kernel

  for (int a0 = blockIdx.x * MT; a0 < n_rows; a0 += gridDim.x * MT) {
    if (ROWS) {
        __syncthreads();
        if (tid < MT) s_rows[tid] = a0 + tid < n_rows ? row_list[a0 + tid] : -1;
        __syncthreads();
    }
    float t1[4], t2[4]; int i1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { t1[i] = INFINITY; t2[i] = INFINITY; i1[i] = 0x7fffffff; }

    for (int bt = bt_beg; bt < bt_end; ++bt) {
        const int b0 = bt * MT;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < dim; k0 += MK) {
            // stage 64 rows x 32 dims of A and B (zero beyond the ends)
#pragma unroll
            for (int e = 0; e < (MT * MK) / 256; ++e) {
                const int idx = e * 256 + tid;
                const int k = idx & (MK - 1), r = idx >> 5;
                float va = 0.f, vb = 0.f;
                if (k0 + k < dim) {
                    const int ar = ROWS ? s_rows[r] : (a0 + r < nA ? a0 + r : -1);
                    if (ar >= 0) va = __ldg(A + (long long)ar * a_sa + (long long)(k0 + k) * a_sk);
                    if (b0 + r < nB) vb = __ldg(B + (long long)(b0 + r) * dim + (k0 + k));
                }
                As[k][r] = va; Bs[k][r] = vb;
            }
            __syncthreads();
            const int kmax = min(MK, dim - k0);
            // row-list mode: the last listed tile is usually almost empty (a handful of uncertified
            // rows); threads whose 4 rows are all padding skip the arithmetic
            if (!ROWS || a0 + ty * 4 < n_rows)
            for (int k = 0; k < kmax; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float t = __fsub_rn(av[i], bv[j]);          // match.cu:39
                        acc[i][j] = __fmaf_rn(t, t, acc[i][j]);           // match.cu:40
                    }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int a = a0 + ty * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = b0 + tx * 4 + j;
                if (b < nB) {
                    rec_update(t1[i], i1[i], t2[i], acc[i][j], b);
                    if (WRITE_D && a < nA) D[(long long)a * d_sa + (long long)b * d_sb] = acc[i][j];
                }
            }
        }
    }
    // merge the 16 column-threads of each row
#pragma unroll
    for (int i = 0; i < 4; ++i) s_rec[ty * 4 + i][tx] = make_float4(t1[i], __int_as_float(i1[i]), t2[i], 0.f);
    __syncthreads();
    if (tid < MT) {
        const int a = a0 + tid;
        float4 r = s_rec[tid][0];
        float m1 = r.x, m2 = r.z; int mi = __float_as_int(r.y);
        for (int k = 1; k < 16; ++k) {
            const float4 q = s_rec[tid][k];
            rec_merge(m1, mi, m2, q.x, __float_as_int(q.y), q.z);
        }
        if (a < n_rows) {
            if (mi != 0x7fffffff) mi += index_offset; else mi = -1;
            if (!ROWS) {
                part[(long long)split * nA + a] = make_float4(m1, __int_as_float(mi), m2, 0.f);
            } else if (mi >= 0) {
                // lock-free merge over the splits (distances are >= 0, so their bit patterns order like
                // the values): key1 = (d1, index) minimum = best column, lowest index on ties; key2 = the
                // minimum over every other value seen (each split's d2 and every displaced d1)
                const unsigned long long mine = ((unsigned long long)__float_as_uint(m1) << 32) | (unsigned)mi;
                const unsigned long long old = atomicMin(key1 + a, mine);
                const unsigned long long loser = old > mine ? old : mine;
                atomicMin(key2 + a, min((unsigned)(loser >> 32), __float_as_uint(m2)));
            }
        }
    }
    __syncthreads();                                              // s_rec reused by the next row tile
  }
}

// Few listed rows (the usual case: a handful of uncertified rows out of 100 000): one block per
// 64-column chunk of the database stages the chunk in shared memory ONCE and runs every listed row
// against it (a warp per row, a lane per column, the reference's sequential fp32 arithmetic), so the
// database is read once however many rows are listed; merged with the same keys as the tiled kernel.
__global__ void __launch_bounds__(128) scan_rows_small_kernel(const float* __restrict__ A, const float* __restrict__ B, int nB,
                                                              int index_offset, const int* __restrict__ row_list,
                                                              const int* __restrict__ row_count, int nA,
                                                              unsigned long long* __restrict__ key1,
                                                              unsigned* __restrict__ key2)
{
    __shared__ float s_b[FB_CHUNK][129];               // pitch 129: lanes walk rows conflict free
    __shared__ __align__(16) float s_a[4][128];
    const int n_rows = min(*row_count, nA);
    if (n_rows == 0 || n_rows > FB_SMALL_MAX) return;  // long lists: the tiled kernel
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * FB_CHUNK;
    for (int e = threadIdx.x; e < FB_CHUNK * 32; e += 128) {
        const int r = e >> 5, q = e & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + r < nB) v = __ldg(reinterpret_cast<const float4*>(B + (size_t)(c0 + r) * 128) + q);
        s_b[r][4 * q] = v.x; s_b[r][4 * q + 1] = v.y; s_b[r][4 * q + 2] = v.z; s_b[r][4 * q + 3] = v.w;
    }
    __syncthreads();
    for (int k = wid; k < n_rows; k += 4) {
        const int row = row_list[k];
        __syncwarp();
        *reinterpret_cast<float4*>(&s_a[wid][lane * 4]) = __ldg(reinterpret_cast<const float4*>(A + (size_t)row * 128) + lane);
        __syncwarp();
        float acc0 = 0.f, acc1 = 0.f;                  // columns c0 + lane and c0 + lane + 32
#pragma unroll 8
        for (int i = 0; i < 128; ++i) {
            const float a = s_a[wid][i];
            float t;
            t = __fsub_rn(a, s_b[lane][i]); acc0 = __fmaf_rn(t, t, acc0);          // match.cu:39-40, i ascending
            t = __fsub_rn(a, s_b[lane + 32][i]); acc1 = __fmaf_rn(t, t, acc1);
        }
        float t1 = INFINITY, t2 = INFINITY; int i1 = 0x7fffffff;
        if (c0 + lane < nB) rec_update(t1, i1, t2, acc0, c0 + lane);
        if (c0 + lane + 32 < nB) rec_update(t1, i1, t2, acc1, c0 + lane + 32);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float u1 = __shfl_xor_sync(0xffffffffu, t1, d);
            const int j1 = __shfl_xor_sync(0xffffffffu, i1, d);
            const float u2 = __shfl_xor_sync(0xffffffffu, t2, d);
            rec_merge(t1, i1, t2, u1, j1, u2);
        }
        if (lane == 0 && i1 != 0x7fffffff) {
            const unsigned long long mine = ((unsigned long long)__float_as_uint(t1) << 32) | (unsigned)(i1 + index_offset);
            const unsigned long long old = atomicMin(key1 + k, mine);
            const unsigned long long loser = old > mine ? old : mine;
            atomicMin(key2 + k, min((unsigned)(loser >> 32), __float_as_uint(t2)));
        }
    }
}

// Records of a row list from the merged keys: out_rec[row_list[k]] for k < *row_count.
__global__ void finish_rows_kernel(const unsigned long long* __restrict__ key1, const unsigned* __restrict__ key2, int nA,
                                   const int* __restrict__ row_list, const int* __restrict__ row_count,
                                   float4* __restrict__ out_rec)
{
    const int n_rows = min(*row_count, nA);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_rows; k += gridDim.x * blockDim.x) {
        const unsigned long long k1 = key1[k];
        const unsigned k2 = key2[k];
        const bool any = k1 != ~0ull;
        out_rec[row_list[k]] = make_float4(any ? __uint_as_float((unsigned)(k1 >> 32)) : INFINITY,
                                           __int_as_float(any ? (int)(unsigned)k1 : -1),
                                           k2 != ~0u ? __uint_as_float(k2) : INFINITY, 0.f);
    }
}

// Merge shard/split-major record arrays and (optionally) apply the ratio rule.
__global__ void merge_kernel(const float4* __restrict__ recs, int n_shards, int nA, float4* __restrict__ out_rec,
                             int apply_rule, float ambiguity, int* __restrict__ match_io, long long shard_stride = 0)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nA) return;
    if (shard_stride <= 0) shard_stride = nA;
    float m1 = INFINITY, m2 = INFINITY; int mi = 0x7fffffff;
    for (int s = 0; s < n_shards; ++s) {
        const float4 q = recs[s * shard_stride + a];
        int qi = __float_as_int(q.y);
        if (qi < 0) qi = 0x7fffffff;
        rec_merge(m1, mi, m2, q.x, qi, q.z);
    }
    if (out_rec) out_rec[a] = make_float4(m1, __int_as_float(mi == 0x7fffffff ? -1 : mi), m2, 0.f);
    if (apply_rule && mi != 0x7fffffff) {
        const float min2 = (mi == 0) ? fminf(NM_MIN2_INIT, m2) : m2;     // match.cu:91 + scan order
        if (min2 > 0.f) {                                                // match.cu:107
            const float r = __fdiv_rn(m1, min2);
            match_io[a] = (r < ambiguity) ? mi : -1;                     // match.cu:109-114
        }
    }
}

// Warp per row over a materialised matrix (compat get_sift_matches).
__global__ void __launch_bounds__(256) set_matches_kernel(const float* __restrict__ distance, int rows, int cols,
                                                          int buffer_width, int* __restrict__ result, float ambiguity)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = distance + (long long)row * buffer_width;
    float t1 = INFINITY, t2 = INFINITY; int i1 = 0x7fffffff;
    for (int j = lane; j < cols; j += 32) rec_update(t1, i1, t2, __ldg(p + j), j);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const float u1 = __shfl_xor_sync(0xffffffffu, t1, d);
        const int j1 = __shfl_xor_sync(0xffffffffu, i1, d);
        const float u2 = __shfl_xor_sync(0xffffffffu, t2, d);
        rec_merge(t1, i1, t2, u1, j1, u2);
    }
    if (lane == 0 && i1 != 0x7fffffff) {
        const float min2 = (i1 == 0) ? fminf(NM_MIN2_INIT, t2) : t2;
        if (min2 > 0.f) result[row] = (__fdiv_rn(t1, min2) < ambiguity) ? i1 : -1;
    }
}

__global__ void transpose_kernel(float* __restrict__ odata, const float* __restrict__ idata, int width, int height)
{
    __shared__ float tile[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    for (int k = 0; k < 32; k += 8)
        if (x < width && y + k < height) tile[threadIdx.y + k][threadIdx.x] = idata[(long long)(y + k) * width + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x; y = blockIdx.x * 32 + threadIdx.y;
    for (int k = 0; k < 32; k += 8)
        if (x < height && y + k < width) odata[(long long)(y + k) * height + x] = tile[threadIdx.x][threadIdx.y + k];
}

int g_engine = -1;     // -1: auto

} // namespace

int nm_match_scan_exact(const float* A, long long a_sa, long long a_sk, int nA, const float* B, int nB,
                        int dim, int index_offset, float4* rec4, float* D, long long d_sa, long long d_sb,
                        cudaStream_t stream)
{
    if (nA <= 0 || nB < 0 || dim <= 0) return NM_ERR_INVALID;
    const int a_tiles = nm_div_up(nA, MT), b_tiles = nm_div_up(nB > 0 ? nB : 1, MT);
    // split the database so that small query sets still fill the machine (>= ~2 waves)
    int splits = 1;
    if (a_tiles < 592) splits = min(b_tiles, nm_div_up(592, a_tiles));
    const int per = nm_div_up(b_tiles, splits);
    splits = nm_div_up(b_tiles, per);
    float4* part = rec4;
    if (splits > 1) NM_CUDA_TRY(nm_ws_alloc(&part, sizeof(float4) * (size_t)splits * nA, stream));
    dim3 grid(a_tiles, splits);
    if (D)
        scan_exact_kernel<true, false><<<grid, 256, 0, stream>>>(A, a_sa, a_sk, nA, B, nB, dim, index_offset, per, part, D, d_sa, d_sb, nullptr, nullptr, nullptr, nullptr);
    else
        scan_exact_kernel<false, false><<<grid, 256, 0, stream>>>(A, a_sa, a_sk, nA, B, nB, dim, index_offset, per, part, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr);
    cudaError_t e = cudaGetLastError();
    if (splits > 1) {
        if (e == cudaSuccess) {
            merge_kernel<<<nm_div_up(nA, 256), 256, 0, stream>>>(part, splits, nA, rec4, 0, 0.f, nullptr);
            e = cudaGetLastError();
        }
        cudaFreeAsync(part, stream);
    }
    return nm_cuda_err(e);
}

// Exact records for the rows row_list[0 .. *row_count) only (count lives on the device, no host
// synchronisation).  The database is cut into 256-column splits so that even a handful of rows
// spreads over the whole machine; a fixed-size grid strides over the listed row tiles and the
// splits merge lock-free into two key arrays (no per-split scratch).
int nm_match_scan_exact_rows(const float* A, int nA, const float* B, int nB, int dim, int index_offset,
                             const int* row_list, const int* row_count, float4* rec4, cudaStream_t stream)
{
    if (nA <= 0 || nB <= 0 || dim <= 0 || !row_list || !row_count) return NM_ERR_INVALID;
    // scan_rows_small_kernel is written for 128-D rows read with 16-byte loads
    if (dim != 128 || ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15)) return NM_ERR_INVALID;
    const int b_tiles = nm_div_up(nB, MT);
    const int per = 4;                                            // 4 x 64 database rows per block
    const int splits = nm_div_up(b_tiles, per);
    if (splits > 65535) return NM_ERR_OVERFLOW;
    const int row_blocks = min(nm_div_up(nA, MT), 32);
    unsigned char* keys = nullptr;
    NM_CUDA_TRY(nm_ws_alloc(&keys, 12 * (size_t)nA, stream));
    unsigned long long* key1 = reinterpret_cast<unsigned long long*>(keys);
    unsigned* key2 = reinterpret_cast<unsigned*>(keys + 8 * (size_t)nA);
    cudaError_t e = cudaMemsetAsync(keys, 0xff, 12 * (size_t)nA, stream);
    if (e == cudaSuccess) {
        // both regimes are launched; each kernel looks at the device-side count and leaves if the
        // list is not its size
        scan_rows_small_kernel<<<nm_div_up(nB, FB_CHUNK), 128, 0, stream>>>(
            A, B, nB, index_offset, row_list, row_count, nA, key1, key2);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        scan_exact_kernel<false, true><<<dim3(row_blocks, splits), 256, 0, stream>>>(
            A, dim, 1, nA, B, nB, dim, index_offset, per, nullptr, nullptr, 0, 0, row_list, row_count, key1, key2);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {
        finish_rows_kernel<<<min(nm_div_up(nA, 256), 64), 256, 0, stream>>>(key1, key2, nA, row_list, row_count, rec4);
        e = cudaGetLastError();
    }
    cudaFreeAsync(keys, stream);
    return nm_cuda_err(e);
}

int nm_match_finalize(const float4* recs, int n_shards, int nA, float ambiguity, int* match_io, cudaStream_t stream,
                      long long shard_stride)
{
    merge_kernel<<<nm_div_up(nA, 256), 256, 0, stream>>>(recs, n_shards, nA, nullptr, 1, ambiguity, match_io, shard_stride);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

// Auto policy: both engines return the same exact records, so the choice is purely a matter of
// speed -- the tensor-core engine pays ~6 extra small launches (pack, rerank, fallback) and wins
// once there are a few million pairs to scan.
static int pick_engine(long long pairs = -1)
{
    if (g_engine >= 0) return g_engine;
    if (!nm_match_tc_available()) return 0;
    return (pairs < 0 || pairs >= (1ll << 22)) ? 1 : 0;
}

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------
extern "C" int nm_match_set_engine(int engine)
{
    if (engine != 0 && engine != 1 && engine != -1) return NM_ERR_INVALID;
    if (engine == 1 && !nm_match_tc_available()) return NM_ERR_UNSUPPORTED;
    g_engine = engine;
    return NM_OK;
}
extern "C" int nm_match_get_engine(void) { return pick_engine(); }

extern "C" int nm_match_top2_f32(const float* A, int nA, const float* B, int nB, int index_offset, float* rec4,
                                 nm_stream_t stream)
{
    if (!A || !rec4 || nA <= 0 || nB < 0 || (nB > 0 && !B)) return NM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    // the tensor-core engine reads rows with 16-byte loads; unaligned inputs take the exact engine
    // (both return the same records)
    const bool aligned = !((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15);
    if (nB > 0 && aligned && pick_engine((long long)nA * nB) == 1)
        return nm_match_scan_tc(A, nA, B, nB, index_offset, reinterpret_cast<float4*>(rec4), st);
    return nm_match_scan_exact(A, 128, 1, nA, B, nB, 128, index_offset, reinterpret_cast<float4*>(rec4), nullptr, 0, 0, st);
}

extern "C" int nm_match_merge_top2(const float* recs4, int n_shards, int nA, float ambiguity, int* match_io,
                                   nm_stream_t stream)
{
    if (!recs4 || !match_io || n_shards <= 0 || nA <= 0) return NM_ERR_INVALID;
    return nm_match_finalize(reinterpret_cast<const float4*>(recs4), n_shards, nA, ambiguity, match_io, (cudaStream_t)stream);
}

extern "C" int nm_match_merge_top2_strided(const float* recs4, int n_shards, long long shard_stride_rows, int nA,
                                           float ambiguity, int* match_io, nm_stream_t stream)
{
    if (!recs4 || !match_io || n_shards <= 0 || nA <= 0 || shard_stride_rows < nA) return NM_ERR_INVALID;
    return nm_match_finalize(reinterpret_cast<const float4*>(recs4), n_shards, nA, ambiguity, match_io, (cudaStream_t)stream,
                             shard_stride_rows);
}

extern "C" int nm_match_f32(const float* A, int nA, const float* B, int nB, float ambiguity, int* match_io,
                            float* distance, nm_stream_t stream)
{
    if (!A || !B || !match_io || nA <= 0 || nB <= 0) return NM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    float4* rec = nullptr;
    NM_CUDA_TRY(nm_ws_alloc(&rec, sizeof(float4) * (size_t)nA, st));
    int rc;
    const bool aligned = !((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15);
    if (distance || !aligned || pick_engine((long long)nA * nB) == 0)
        rc = nm_match_scan_exact(A, 128, 1, nA, B, nB, 128, 0, rec, distance, nB, 1, st);
    else
        rc = nm_match_scan_tc(A, nA, B, nB, 0, rec, st);
    if (rc == NM_OK) rc = nm_match_finalize(rec, 1, nA, ambiguity, match_io, st);
    cudaFreeAsync(rec, st);
    return rc;
}

extern "C" int nm_transpose_f32(float* odata, const float* idata, int width, int height, nm_stream_t stream)
{
    if (!odata || !idata || width <= 0 || height <= 0) return NM_ERR_INVALID;
    dim3 block(32, 8), grid(nm_div_up(width, 32), nm_div_up(height, 32));
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(odata, idata, width, height);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

extern "C" int nm_dist2_f32(const float* A_t, int size_A, const float* B, int size_B, int vector_dim,
                            float* result_t, nm_stream_t stream)
{
    if (!A_t || !B || !result_t || size_A <= 0 || size_B <= 0 || vector_dim <= 0) return NM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    float4* rec = nullptr;
    NM_CUDA_TRY(nm_ws_alloc(&rec, sizeof(float4) * (size_t)size_A, st));
    // A(a,k) = A_t[k*size_A + a];  D^T[b*size_A + a]   (match.h:7-23)
    int rc = nm_match_scan_exact(A_t, 1, size_A, size_A, B, size_B, vector_dim, 0, rec, result_t, 1, size_A, st);
    cudaFreeAsync(rec, st);
    return rc;
}

extern "C" int nm_set_matches_f32(const float* distance, int rows, int cols, int buffer_width, int* result,
                                  float ambiguity, nm_stream_t stream)
{
    if (!distance || !result || rows <= 0 || cols <= 0 || buffer_width < cols) return NM_ERR_INVALID;
    set_matches_kernel<<<nm_div_up(rows, 8), 256, 0, (cudaStream_t)stream>>>(distance, rows, cols, buffer_width, result, ambiguity);
    NM_LAUNCH_CHECK();
    return NM_OK;
}
