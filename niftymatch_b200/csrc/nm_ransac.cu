// nm_ransac.cu -- frame-to-frame registration after matching.
//
// Replaces the reference's align_points / ransac_translation / ransac_similarity / ransac_homography
// (gpu/kernels/ransac.cu:50-59, :526-694) and the one-sided Jacobi SVD they call (gpu/kernels/svd.cu:
// 200-360, the reference's port of GSL's gsl_linalg_SV_decomp_jacobi).
//
// Provenance note: the reference's svd.cu carries a GNU GPL header of its own (svd.cu:1-20, derived from GSL),
// unlike the rest of the reference.  jacobi_right_vectors below is a restructured restatement of that published
// algorithm (fused dot / norm loop, lane-interleaved shared-memory matrices, templated sizes, no singular-value
// tail), but it follows the same order of operations -- bitwise hypotheses require it -- so whoever ships this
// file should treat it under the terms that apply to svd.cu.
//
// What is kept: the arithmetic of a hypothesis (normalised 4-point DLT / 2-point similarity / 1-point
// translation, Jacobi sweeps with GSL's error-estimate skip rule, the expanded de-normalisation) in the
// reference's order of operations, the inlier rule (squared reprojection error < threshold over the
// correspondences whose src_x >= 0), "first maximum wins", duplicate index draws score 0 with H = 0.
//
// What is different (B200 design):
//  * the reference copies src_x to the HOST, filters the valid indices there, draws the random list with
//    std::mt19937 seeded from std::random_device (not reproducible) and copies it back: three blocking
//    transfers per call.  Here the valid-index compaction, a counter-based generator (splitmix64 of
//    (seed, draw)), the hypotheses, the scores and the arg-max all run on the caller's stream with no host
//    synchronisation; the seed is a parameter, so a run is reproducible;
//  * the reference scores with one thread per hypothesis looping over every correspondence in global
//    memory; here a CTA stages a block of correspondences in shared memory once and each of its threads
//    scores one hypothesis against it (grid = hypothesis blocks x correspondence blocks, integer atomics,
//    so the counts are exact and order independent);
//  * nm_ransac_batch_f32 estimates many frame pairs in one launch sequence (every kernel indexes the pair by a
//    grid dimension): a hypothesis is a long dependent chain (5-9 Jacobi sweeps of 36 rotations, ~1 ms), so
//    one pair at a time leaves the GPU idle; 64 pairs cost 43 us each;
//  * nm_ransac_hypotheses_f32 takes the caller's index list: that is the entry the parity tests drive with
//    the list they also hand to the reference's own kernels (homographies and inlier counts are bitwise the
//    reference's on the same lists, tests/test_gpu_ransac.py).
#include "nm_common.cuh"

namespace {

constexpr float kEps = 1.1920928955078125e-07f;      // svd.cu:33 (FLT_EPSILON)

// ---- one-sided Jacobi SVD ---------------------------------------------------------------------------
// svd.cu:133-157
__device__ float hyp(float x, float y)
{
    const float xa = fabsf(x), ya = fabsf(y);
    const float mn = fminf(xa, ya), mx = fmaxf(xa, ya);
    if (mn == 0.f) return mx;
    const float u = mn / mx;
    return mx * sqrtf(1.f + u * u);
}

// A thread's matrix lives in shared memory, element e of lane l at [e * kHypThreads + l]: dynamic column
// indices (j, k) would put a thread-local array in local memory; this layout is bank-conflict free and the
// rotations read and write it at shared-memory latency.
constexpr int kHypThreads = 32;
struct SMat {
    float* p;
    __device__ __forceinline__ float& operator[](int e) const { return p[e * kHypThreads]; }
};

// Right singular vectors Q (COLS x COLS) of A (ROWS x COLS, destroyed); svd.cu:200-316.  The singular
// values and the column normalisation of A that follow in the reference (:318-352) are not read by any
// caller and are not computed.  err: COLS column error estimates (the reference's S).
template <int ROWS, int COLS>
__device__ void jacobi_right_vectors(SMat A, SMat Q, SMat err)
{
    const float tol = (float)(10 * ROWS) * kEps;
    const int sweepmax = 5 * COLS > 12 ? 5 * COLS : 12;
    for (int i = 0; i < COLS * COLS; ++i) Q[i] = 0.f;
    for (int i = 0; i < COLS; ++i) Q[i * COLS + i] = 1.f;
    for (int j = 0; j < COLS; ++j) {
        // dnrm2 (svd.cu:159-198)
        float scale = 0.f, ssq = 1.f;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const float x = A[i * COLS + j];
            if (x != 0.f) {
                const float ax = fabsf(x);
                if (scale < ax) { ssq = 1.f + ssq * (scale / ax) * (scale / ax); scale = ax; }
                else ssq += (ax / scale) * (ax / scale);
            }
        }
        err[j] = kEps * (scale * sqrtf(ssq));
    }
    int count = 1, sweep = 0;
    while (count > 0 && sweep <= sweepmax) {
        count = COLS * (COLS - 1) / 2;
#pragma unroll 1
        for (int j = 0; j < COLS - 1; ++j)
#pragma unroll 1
            for (int k = j + 1; k < COLS; ++k) {
                // dot product and the two scaled norms in ONE pass over the rows: three independent dependency
                // chains in flight instead of three loops back to back (each quantity sees the operations of
                // ddot / dnrm2, svd.cu:123-131 and :159-198, in the same order)
                float cj[ROWS], ck[ROWS];
#pragma unroll
                for (int i = 0; i < ROWS; ++i) { cj[i] = A[i * COLS + j]; ck[i] = A[i * COLS + k]; }
                float p = 0.f, sa = 0.f, qa = 1.f, sb = 0.f, qb = 1.f;
#pragma unroll
                for (int i = 0; i < ROWS; ++i) {
                    const float xj = cj[i], xk = ck[i];
                    p += xj * xk;
                    if (xj != 0.f) {
                        const float ax = fabsf(xj);
                        if (sa < ax) { qa = 1.f + qa * (sa / ax) * (sa / ax); sa = ax; }
                        else qa += (ax / sa) * (ax / sa);
                    }
                    if (xk != 0.f) {
                        const float ax = fabsf(xk);
                        if (sb < ax) { qb = 1.f + qb * (sb / ax) * (sb / ax); sb = ax; }
                        else qb += (ax / sb) * (ax / sb);
                    }
                }
                p *= 2.0f;
                const float a = sa * sqrtf(qa), b = sb * sqrtf(qb);
                const float q = a * a - b * b;
                const float v = hyp(p, q);
                const float ea = err[j], eb = err[k];
                const bool sorted = a >= b;
                const bool orthog = fabsf(p) <= tol * (a * b);
                if (sorted && (orthog || a < ea || b < eb)) { --count; continue; }
                float c, s;
                if (v == 0.f || !sorted) { c = 0.f; s = 1.f; }
                else {
                    c = (float)sqrt((double)(v + q) / (2.0 * (double)v));
                    s = (float)((double)p / (2.0 * (double)v * (double)c));
                }
#pragma unroll
                for (int i = 0; i < ROWS; ++i) {
                    const float Aik = ck[i], Aij = cj[i];
                    A[i * COLS + j] = Aij * c + Aik * s;
                    A[i * COLS + k] = -Aij * s + Aik * c;
                }
                err[j] = fabsf(c) * ea + fabsf(s) * eb;
                err[k] = fabsf(s) * ea + fabsf(c) * eb;
#pragma unroll
                for (int i = 0; i < COLS; ++i) {
                    const float Qij = Q[i * COLS + j], Qik = Q[i * COLS + k];
                    Q[i * COLS + j] = Qij * c + Qik * s;
                    Q[i * COLS + k] = -Qij * s + Qik * c;
                }
            }
        ++sweep;
    }
}

// inv(dst_transform) * H * src_transform, expanded (ransac.cu:201-212, :424-434)
__device__ void denormalise(const float* H, float s1, float s2, float tx1, float ty1, float tx2, float ty2, float* R)
{
    R[0] = s1 * tx2 * H[6] + s1 * H[0] / s2;
    R[1] = s1 * tx2 * H[7] + s1 * H[1] / s2;
    R[2] = tx2 * (H[8] - s1 * ty1 * H[7] - s1 * tx1 * H[6]) + (H[2] - s1 * ty1 * H[1] - s1 * tx1 * H[0]) / s2;
    R[3] = s1 * ty2 * H[6] + s1 * H[3] / s2;
    R[4] = s1 * ty2 * H[7] + s1 * H[4] / s2;
    R[5] = ty2 * (H[8] - s1 * ty1 * H[7] - s1 * tx1 * H[6]) + (H[5] - s1 * ty1 * H[4] - s1 * tx1 * H[3]) / s2;
    R[6] = s1 * H[6];
    R[7] = s1 * H[7];
    R[8] = H[8] - s1 * ty1 * H[7] - s1 * tx1 * H[6];
}

// centroid, mean squared distance, sqrt(2)/rms scale of M points (ransac.cu:100-118, :326-341)
template <int M>
__device__ void normaliser(const float2* p, float& mx, float& my, float& scale)
{
    float sx = p[0].x, sy = p[0].y;
#pragma unroll
    for (int i = 1; i < M; ++i) { sx += p[i].x; sy += p[i].y; }
    mx = sx * (1.0f / M); my = sy * (1.0f / M);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < M; ++i) var += (p[i].x - mx) * (p[i].x - mx) + (p[i].y - my) * (p[i].y - my);
    if (M == 2) var = (float)((double)var * 0.5);       // ransac.cu:336-337 multiply by a double literal
    else var *= 0.25f;
    scale = sqrtf(2.0f) / sqrtf(var);
}

// 4-point homography, ransac.cu:84-214: nine DLT rows (the two of every point + the third of point 3)
__device__ void homography4(const float2* src, const float2* dst, float* R, SMat X, SMat V, SMat err)
{
    float smx, smy, s1, dmx, dmy, s2;
    normaliser<4>(src, smx, smy, s1);
    normaliser<4>(dst, dmx, dmy, s2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = (src[i].x - smx) * s1, b = (src[i].y - smy) * s1;
        const float u = (dst[i].x - dmx) * s2, w = (dst[i].y - dmy) * s2;
        const float r1[9] = {0.f, 0.f, 0.f, -a, -b, -1.f, w * a, w * b, w};
        const float r2[9] = {a, b, 1.f, 0.f, 0.f, 0.f, -u * a, -u * b, -u};
#pragma unroll
        for (int e = 0; e < 9; ++e) { X[18 * i + e] = r1[e]; X[18 * i + 9 + e] = r2[e]; }
        if (i == 3) {
            const float r3[9] = {-w * a, -w * b, -w, u * a, u * b, u, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 9; ++e) X[72 + e] = r3[e];
        }
    }
    jacobi_right_vectors<9, 9>(X, V, err);
    float H[9];
    const float div = V[80];
#pragma unroll
    for (int i = 0; i < 8; ++i) H[i] = V[i * 9 + 8] / div;
    H[8] = 1.f;
    denormalise(H, s1, s2, smx, smy, dmx, dmy, R);
}

// 2-point similarity, ransac.cu:320-435: 4 x 5 system in (a, tx, b, ty, 1)
__device__ void similarity2(const float2* src, const float2* dst, float* R, SMat X, SMat V, SMat err)
{
    float smx, smy, s1, dmx, dmy, s2;
    normaliser<2>(src, smx, smy, s1);
    normaliser<2>(dst, dmx, dmy, s2);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float a = (src[i].x - smx) * s1, b = (src[i].y - smy) * s1;
        const float u = (dst[i].x - dmx) * s2, w = (dst[i].y - dmy) * s2;
        const float r1[5] = {a, 1.f, -b, 0.f, u};
        const float r2[5] = {b, 0.f, a, 1.f, w};
#pragma unroll
        for (int e = 0; e < 5; ++e) { X[10 * i + e] = r1[e]; X[10 * i + 5 + e] = r2[e]; }
    }
    jacobi_right_vectors<4, 5>(X, V, err);
    const float div = V[24];
    const float a0 = -V[4] / div, a1 = -V[9] / div, b0 = -V[14] / div, b1 = -V[19] / div;
    const float H[9] = {a0, -b0, a1, b0, a0, b1, 0.f, 0.f, 1.f};
    denormalise(H, s1, s2, smx, smy, dmx, dmy, R);
}

// ---- kernels -----------------------------------------------------------------------------------
// establish_correspondences (ransac.cu:29-48)
__global__ void align_kernel(const float* __restrict__ src_x, const float* __restrict__ src_y,
                             const float* __restrict__ dst_x, const float* __restrict__ dst_y,
                             float* __restrict__ c_src_x, float* __restrict__ c_src_y,
                             float* __restrict__ c_dst_x, float* __restrict__ c_dst_y,
                             const int* __restrict__ matches, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = matches[i];
    const bool ok = m != -1;
    c_src_x[i] = ok ? src_x[i] : -1.f;
    c_src_y[i] = ok ? src_y[i] : -1.f;
    c_dst_x[i] = ok ? dst_x[m] : -1.f;
    c_dst_y[i] = ok ? dst_y[m] : -1.f;
}

// align_points for the consecutive pairs of a SIFT batch: pair p = (frame p, frame p + 1), sizes on the device
__global__ void align_pairs_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ matches,
                                   const int* __restrict__ counts, int capacity, float* __restrict__ c_src_x,
                                   float* __restrict__ c_src_y, float* __restrict__ c_dst_x, float* __restrict__ c_dst_y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (i >= capacity) return;
    const long long o = (long long)p * capacity + i;
    const int nA = min(counts[p], capacity), nB = min(counts[p + 1], capacity);
    const int m = i < nA ? matches[o] : -1;
    const bool ok = m >= 0 && m < nB;
    c_src_x[o] = ok ? x[o] : -1.f;
    c_src_y[o] = ok ? y[o] : -1.f;
    c_dst_x[o] = ok ? x[o + capacity - i + m] : -1.f;
    c_dst_y[o] = ok ? y[o + capacity - i + m] : -1.f;
}

// ---- batched estimator: every kernel indexes a PAIR of frames by a grid dimension ------------------------
// Pair p reads its correspondences at (sx, sy, dx, dy) + p * stride, n(p) = counts ? min(counts[p], max_pts) :
// max_pts of them; per-pair workspace slices: valid[p * max_pts], state[p * 4], rand[p * draws], H[p * it * 9],
// inliers / skip [p * it].  A single estimate is the batch of one.
struct Pairs {
    const float *sx, *sy, *dx, *dy;
    long long stride;
    const int* counts;
    int max_pts;
};
__device__ __forceinline__ int pair_n(const Pairs& P, int p)
{
    if (P.counts == nullptr) return P.max_pts;
    const int c = P.counts[p];
    return c < 0 ? 0 : (c < P.max_pts ? c : P.max_pts);
}

// Ordered list of the indices with src_x >= 0 (the host loop of ransac.cu:533-538); one CTA per pair.
// state[p * 4] = number of valid correspondences.
__global__ void __launch_bounds__(1024) valid_list_kernel(const Pairs P, int* __restrict__ list_all, int* __restrict__ state_all)
{
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int p = blockIdx.x, n = pair_n(P, p);
    const float* __restrict__ src_x = P.sx + (long long)p * P.stride;
    int* __restrict__ list = list_all + (long long)p * P.max_pts;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += 1024) {
        const int i = start + tid;
        const bool ok = i < n && src_x[i] >= 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_tot[wid] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < wid; ++w) off += warp_tot[w];
        if (ok) list[off + __popc(m & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 32; ++w) t += warp_tot[w];
            base += t;
        }
        __syncthreads();
    }
    if (tid == 0) state_all[p * 4] = base;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// rand_list[d] = valid[u(seed + p, d) scaled to [0, n_valid)], d < draws (the host loop of ransac.cu:551-555)
__global__ void draw_kernel(const int* __restrict__ valid_all, const int* __restrict__ state_all, int max_pts, int min_pts,
                            unsigned long long seed, int draws, int* __restrict__ rand_all)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (d >= draws) return;
    int* rand_list = rand_all + (long long)p * draws;
    const int nv = state_all[p * 4];
    if (nv < min_pts) { rand_list[d] = 0; return; }
    const unsigned r = (unsigned)(splitmix64((seed + (unsigned long long)p) ^ (0xD1B54A32D192ED03ull * (unsigned long long)(d + 1))) >> 32);
    rand_list[d] = valid_all[(long long)p * max_pts + (int)(((unsigned long long)r * (unsigned)nv) >> 32)];
}

// One thread per (pair, iteration): the hypothesis (translation_kernel / similarity_transformation_kernel /
// homography_kernel, ransac.cu:437-520, without their scoring loop).  skip = 1 for an iteration with a
// repeated index: H stays 0 and it scores 0, as in the reference (zero-filled buffers, early return).
// CTAs of kHypThreads = 32: a hypothesis is one long dependent chain (5-9 Jacobi sweeps of 36 rotations), so
// the hypotheses are spread over as many SMs as possible instead of packed 128 or 256 to a CTA.
template <int KIND>
__global__ void __launch_bounds__(kHypThreads) hypothesis_kernel(const Pairs P, const int* __restrict__ rand_all, int iterations,
                                                                 const int* __restrict__ state_all, int min_pts,
                                                                 float* __restrict__ H_all, int* __restrict__ inliers,
                                                                 unsigned char* __restrict__ skip)
{
    constexpr int M = KIND == 0 ? 1 : KIND == 1 ? 2 : 4;
    __shared__ float ws[KIND == 0 ? 1 : (81 + 81 + 9) * kHypThreads];
    const SMat X{ws + threadIdx.x}, V{ws + 81 * kHypThreads + threadIdx.x}, err{ws + 162 * kHypThreads + threadIdx.x};
    const int it = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (it >= iterations) return;
    const long long off = (long long)p * P.stride;
    const float* __restrict__ sx = P.sx + off;
    const float* __restrict__ sy = P.sy + off;
    const float* __restrict__ dx = P.dx + off;
    const float* __restrict__ dy = P.dy + off;
    const int* __restrict__ rand_list = rand_all + (long long)p * iterations * M;
    float H[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = 0.f;
    int r[M];
#pragma unroll
    for (int i = 0; i < M; ++i) r[i] = rand_list[it * M + i];
    bool dup = state_all != nullptr && state_all[p * 4] < min_pts;      // not enough correspondences: nothing is estimated
    const int n = pair_n(P, p);
#pragma unroll
    for (int a = 0; a < M; ++a) {
        dup |= (unsigned)r[a] >= (unsigned)n;                           // a caller's list with an index outside the arrays
#pragma unroll
        for (int b = a + 1; b < M; ++b) dup |= r[a] == r[b];
    }
    if (!dup) {
        float2 src[M], dst[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            src[i] = make_float2(sx[r[i]], sy[r[i]]);
            dst[i] = make_float2(dx[r[i]], dy[r[i]]);
        }
        if (KIND == 0) {                                     // compute_translation, ransac.cu:304-310
            H[0] = H[4] = H[8] = 1.f;
            H[2] = dst[0].x - src[0].x;
            H[5] = dst[0].y - src[0].y;
        } else if (KIND == 1) similarity2(src, dst, H, X, V, err);
        else homography4(src, dst, H, X, V, err);
    }
    const long long o = (long long)p * iterations + it;
#pragma unroll
    for (int i = 0; i < 9; ++i) H_all[o * 9 + i] = H[i];
    inliers[o] = 0;
    skip[o] = dup ? 1 : 0;
}

// eval_transformation (ransac.cu:61-82) for 128 hypotheses x one block of up to 256 correspondences of one pair
// (two IEEE divisions per evaluation make the loop long: short blocks give the grid enough CTAs to fill the SMs).
constexpr int kScoreThreads = 128, kScorePts = 256;
__global__ void __launch_bounds__(kScoreThreads) score_kernel(const Pairs P, const float* __restrict__ H_all,
                                                              const unsigned char* __restrict__ skip, int iterations, float thr,
                                                              int* __restrict__ inliers)
{
    __shared__ float4 pts[kScorePts];
    const int p = blockIdx.z, n = pair_n(P, p);
    const int p0 = blockIdx.y * kScorePts, np = min(kScorePts, n - p0);
    if (np <= 0) return;
    const long long off = (long long)p * P.stride + p0;
    for (int i = threadIdx.x; i < np; i += kScoreThreads) pts[i] = make_float4(P.sx[off + i], P.sy[off + i], P.dx[off + i], P.dy[off + i]);
    __syncthreads();
    const int it = blockIdx.x * kScoreThreads + threadIdx.x;
    const long long o = (long long)p * iterations + it;
    if (it >= iterations || skip[o]) return;
    float H[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = H_all[o * 9 + i];
    int cnt = 0;
    for (int i = 0; i < np; ++i) {
        const float4 q = pts[i];
        if (q.x >= 0.f) {
            float x = H[0] * q.x + H[1] * q.y + H[2];
            float y = H[3] * q.x + H[4] * q.y + H[5];
            const float z = H[6] * q.x + H[7] * q.y + H[8];
            x /= z;
            y /= z;
            const float d2 = (q.z - x) * (q.z - x) + (q.w - y) * (q.w - y);
            if (d2 < thr) ++cnt;
        }
    }
    if (cnt) atomicAdd(inliers + o, cnt);
}

// thrust::max_element (first maximum, ransac.cu:566-570) + the 9-float copy; one CTA per pair.
// status[p*3 + 0] = 1 when a model was written, 0 when there were too few correspondences (the reference returns
// false and leaves `homography` untouched); [1] = inlier count of the chosen hypothesis, [2] = its index.
__global__ void __launch_bounds__(256) select_kernel(const int* __restrict__ inliers_all, const float* __restrict__ H_all, int iterations,
                                                     const int* __restrict__ state_all, int min_pts, float* __restrict__ H_out_all,
                                                     int* __restrict__ status_all)
{
    __shared__ long long best[256];
    const int tid = threadIdx.x, p = blockIdx.x;
    int* status = status_all + p * 3;
    if (state_all[p * 4] < min_pts) {
        if (tid == 0) { status[0] = 0; status[1] = 0; status[2] = -1; }
        return;
    }
    const int* inliers = inliers_all + (long long)p * iterations;
    // key = count in the high word, ~index in the low word: a larger key is a larger count, then a smaller index
    long long key = -1;
    for (int i = tid; i < iterations; i += 256) {
        const long long k = ((long long)inliers[i] << 32) | (unsigned)(0x7fffffff - i);
        key = k > key ? k : key;
    }
    best[tid] = key;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (tid < s && best[tid + s] > best[tid]) best[tid] = best[tid + s];
        __syncthreads();
    }
    const int idx = 0x7fffffff - (int)(best[0] & 0xffffffffll);
    if (tid < 9) H_out_all[p * 9 + tid] = H_all[((long long)p * iterations + idx) * 9 + tid];
    if (tid == 0) { status[0] = 1; status[1] = (int)(best[0] >> 32); status[2] = idx; }
}

int min_points(int kind) { return kind == 2 ? 4 : 2; }        // ransac.cu:541, :606, :656 (translation also asks for 2)
int sample_size(int kind) { return kind == 0 ? 1 : kind == 1 ? 2 : 4; }

int launch_hypotheses(int kind, const Pairs& P, int n_pairs, const int* rand_list, int iterations, float thr, const int* state,
                      float* H_all, int* inliers, unsigned char* skip, cudaStream_t st)
{
    const dim3 grid(nm_div_up(iterations, kHypThreads), n_pairs);
    const int mp = min_points(kind);
    if (kind == 0) hypothesis_kernel<0><<<grid, kHypThreads, 0, st>>>(P, rand_list, iterations, state, mp, H_all, inliers, skip);
    else if (kind == 1) hypothesis_kernel<1><<<grid, kHypThreads, 0, st>>>(P, rand_list, iterations, state, mp, H_all, inliers, skip);
    else hypothesis_kernel<2><<<grid, kHypThreads, 0, st>>>(P, rand_list, iterations, state, mp, H_all, inliers, skip);
    NM_LAUNCH_CHECK();
    const dim3 sg(nm_div_up(iterations, kScoreThreads), nm_div_up(P.max_pts, kScorePts), n_pairs);
    score_kernel<<<sg, kScoreThreads, 0, st>>>(P, H_all, skip, iterations, thr, inliers);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

int ransac_batch(int kind, const Pairs& P, int n_pairs, float thr, int iterations, unsigned long long seed, float* homographies,
                 int* status, cudaStream_t st)
{
    const int m = sample_size(kind), mp = min_points(kind);
    const long long draws = (long long)iterations * m;
    if (draws >= (1LL << 31) || (long long)n_pairs * iterations >= (1LL << 31) || n_pairs > 65535) return NM_ERR_INVALID;
    // one stream-ordered block: valid lists | states | rand lists | homographies | inliers | skip flags
    const size_t np = (size_t)n_pairs;
    const size_t o_valid = 0, o_state = o_valid + sizeof(int) * np * (size_t)P.max_pts, o_rand = o_state + 16 * np,
                 o_H = o_rand + sizeof(int) * np * (size_t)draws, o_inl = o_H + sizeof(float) * 9 * np * (size_t)iterations,
                 o_skip = o_inl + sizeof(int) * np * (size_t)iterations, total = o_skip + np * (size_t)iterations;
    char* ws = nullptr;
    NM_CUDA_TRY(nm_ws_alloc(&ws, total, st));
    int* valid = reinterpret_cast<int*>(ws + o_valid);
    int* state = reinterpret_cast<int*>(ws + o_state);
    int* rand_list = reinterpret_cast<int*>(ws + o_rand);
    float* H_all = reinterpret_cast<float*>(ws + o_H);
    int* inl = reinterpret_cast<int*>(ws + o_inl);
    unsigned char* skip = reinterpret_cast<unsigned char*>(ws + o_skip);
    valid_list_kernel<<<n_pairs, 1024, 0, st>>>(P, valid, state);
    draw_kernel<<<dim3(nm_div_up((int)draws, 256), n_pairs), 256, 0, st>>>(valid, state, P.max_pts, mp, seed, (int)draws, rand_list);
    int rc = cudaGetLastError() == cudaSuccess ? NM_OK : NM_ERR_CUDA_BASE;
    if (rc == NM_OK) rc = launch_hypotheses(kind, P, n_pairs, rand_list, iterations, thr, state, H_all, inl, skip, st);
    if (rc == NM_OK) {
        select_kernel<<<n_pairs, 256, 0, st>>>(inl, H_all, iterations, state, mp, homographies, status);
        rc = cudaGetLastError() == cudaSuccess ? NM_OK : NM_ERR_CUDA_BASE;
    }
    NM_CUDA_TRY(cudaFreeAsync(ws, st));
    return rc;
}

} // namespace

extern "C" int nm_align_points_f32(const float* src_x, const float* src_y, const float* dst_x, const float* dst_y,
                                   float* c_src_x, float* c_src_y, float* c_dst_x, float* c_dst_y,
                                   const int* matches, int num_pts, nm_stream_t stream)
{
    if (num_pts < 0) return NM_ERR_INVALID;
    if (num_pts == 0) return NM_OK;
    if (!src_x || !src_y || !dst_x || !dst_y || !c_src_x || !c_src_y || !c_dst_x || !c_dst_y || !matches) return NM_ERR_INVALID;
    align_kernel<<<nm_div_up(num_pts, 256), 256, 0, (cudaStream_t)stream>>>(src_x, src_y, dst_x, dst_y, c_src_x, c_src_y,
                                                                          c_dst_x, c_dst_y, matches, num_pts);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

extern "C" int nm_align_pairs_f32(const float* x, const float* y, const int* matches, const int* counts_dev, int n_frames,
                                  int capacity, float* c_src_x, float* c_src_y, float* c_dst_x, float* c_dst_y,
                                  nm_stream_t stream)
{
    if (!x || !y || !matches || !counts_dev || !c_src_x || !c_src_y || !c_dst_x || !c_dst_y || n_frames < 2 || capacity <= 0)
        return NM_ERR_INVALID;
    align_pairs_kernel<<<dim3(nm_div_up(capacity, 256), n_frames - 1), 256, 0, (cudaStream_t)stream>>>(
        x, y, matches, counts_dev, capacity, c_src_x, c_src_y, c_dst_x, c_dst_y);
    NM_LAUNCH_CHECK();
    return NM_OK;
}

extern "C" int nm_ransac_hypotheses_f32(int kind, const float* src_x, const float* src_y, const float* dst_x,
                                        const float* dst_y, int num_pts, const int* rand_list, int iterations,
                                        float inlier_threshold, float* homographies, int* inliers, nm_stream_t stream)
{
    if (kind < 0 || kind > 2 || num_pts <= 0 || iterations <= 0) return NM_ERR_INVALID;
    if (!src_x || !src_y || !dst_x || !dst_y || !rand_list || !homographies || !inliers) return NM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* skip = nullptr;
    NM_CUDA_TRY(nm_ws_alloc(&skip, (size_t)iterations, st));
    const Pairs P{src_x, src_y, dst_x, dst_y, 0, nullptr, num_pts};
    const int rc = launch_hypotheses(kind, P, 1, rand_list, iterations, inlier_threshold, nullptr, homographies, inliers, skip, st);
    NM_CUDA_TRY(cudaFreeAsync(skip, st));
    return rc;
}

extern "C" int nm_ransac_f32(int kind, const float* src_x, const float* src_y, const float* dst_x, const float* dst_y,
                             int num_pts, float inlier_threshold, int iterations, unsigned long long seed,
                             float* homography, int* status, nm_stream_t stream)
{
    if (kind < 0 || kind > 2 || num_pts <= 0 || iterations <= 0) return NM_ERR_INVALID;
    if (!src_x || !src_y || !dst_x || !dst_y || !homography || !status) return NM_ERR_INVALID;
    const Pairs P{src_x, src_y, dst_x, dst_y, 0, nullptr, num_pts};
    return ransac_batch(kind, P, 1, inlier_threshold, iterations, seed, homography, status, (cudaStream_t)stream);
}

extern "C" int nm_ransac_batch_f32(int kind, const float* src_x, const float* src_y, const float* dst_x, const float* dst_y,
                                   long long pair_stride, const int* counts, int max_pts, int n_pairs, float inlier_threshold,
                                   int iterations, unsigned long long seed, float* homographies, int* status, nm_stream_t stream)
{
    if (kind < 0 || kind > 2 || max_pts <= 0 || n_pairs <= 0 || iterations <= 0 || pair_stride < 0) return NM_ERR_INVALID;
    if (!src_x || !src_y || !dst_x || !dst_y || !homographies || !status) return NM_ERR_INVALID;
    const Pairs P{src_x, src_y, dst_x, dst_y, pair_stride, counts, max_pts};
    return ransac_batch(kind, P, n_pairs, inlier_threshold, iterations, seed, homographies, status, (cudaStream_t)stream);
}
