/* oracle/nm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A scalar CPU restatement of the one hot path of gift-surg/NiftyMatch that this
 * repository accelerates (SIFT detect+describe and brute-force k=2 ratio-test
 * matching).  The reference has no CPU path; every function below follows the
 * reference's CUDA code operation by operation and cites the file:line it follows
 * (paths relative to /root/reference/src).  It is the checker for tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg -- nothing in the product
 * path (niftymatch_b200/) may import, link or execute it.
 *
 * Parity pin: the reference ships no golden vectors, so this port is
 * pinned against outputs of the reference's own CUDA code compiled for sm_100a
 * (oracle/build_ref.sh -> oracle/_ref/libnmref.so) on identical inputs; the captured
 * outputs live in tests/golden/ (see tests/golden/README.md for the capture status).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see Makefile).  FMA is
 * written explicitly (fmaf) exactly where nvcc/ptxas contract in the reference build
 * (checked in the SASS of oracle/_ref), nowhere else.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { float x, y, z, w; } f4;
typedef struct { float x, y; } f2;

/* ------------------------------------------------------------------------- */
/* SiftParams (gpu/sift/siftparams.h:30-51)                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    int   width, height;
    int   num_octaves, num_dog_levels, level_max, level_min;
    float sigma_d_0, sigma_k, sigma_0, sigma_n, base_smooth;
    float sigmas[8];
    int   num_sigmas;
    float peak_threshold, edge_threshold;
} orc_params;

void orc_params_init(orc_params* P, int width, int height)
{
    memset(P, 0, sizeof(*P));
    P->width = width; P->height = height;
    P->num_dog_levels = 3;                       /* siftparams.h:31 */
    P->sigma_n = 0.5f; P->peak_threshold = 0.f; P->edge_threshold = 10.f;
    P->level_max = P->num_dog_levels + 1;        /* :34 */
    P->level_min = -1;                           /* :35 */
    int m = width < height ? width : height;
    P->num_octaves = (int)floor(log(m * 2.0 / 32) / log(2.0));   /* :36 */
    if (P->num_octaves <= 0) P->num_octaves = 1;
    /* :39 std::pow(float,float) is the float overload */
    P->sigma_k = powf(2.0f, 1.0f / P->num_dog_levels);
    P->sigma_0 = 1.6f * P->sigma_k;              /* :40 */
    /* :41 float product widened, rest in double */
    P->sigma_d_0 = (float)((double)P->sigma_0 *
                           sqrt(1.0 - 1.0 / (double)(P->sigma_k * P->sigma_k)));
    /* :43 std::pow(float,int) promotes to double */
    float sa = (float)((double)P->sigma_0 * pow((double)P->sigma_k, (double)P->level_min));
    float sb = P->sigma_n;
    if (sa > sb) P->base_smooth = sqrtf(sa * sa - sb * sb);      /* :47 */
    P->num_sigmas = 0;
    for (int i = P->level_min + 1; i <= P->level_max; ++i)       /* :50 */
        P->sigmas[P->num_sigmas++] =
            (float)((double)P->sigma_d_0 * pow((double)P->sigma_k, (double)i));
}

/* Same outputs as nmref_params in oracle/ref_driver.cu. */
int orc_params_query(int w, int h, int* num_octaves, float* sigma_k, float* sigma_0,
                     float* sigma_d_0, float* base_smooth, float* sigmas5)
{
    orc_params P; orc_params_init(&P, w, h);
    *num_octaves = P.num_octaves; *sigma_k = P.sigma_k; *sigma_0 = P.sigma_0;
    *sigma_d_0 = P.sigma_d_0; *base_smooth = P.base_smooth;
    for (int i = 0; i < P.num_sigmas && i < 5; ++i) sigmas5[i] = P.sigmas[i];
    return P.num_sigmas;
}

/* Gaussian taps (gpu/sift/pyramidata.cu:105-123).  taps must hold 2*radius+1 floats
 * (<= 91, pyramidata.h:9).  Returns the radius. */
int orc_make_taps(float sigma, float* taps)
{
    const int radius = (int)ceilf(sigma * 4);                    /* :108 */
    const int len = 2 * radius + 1;
    float sum = 0.f;
    for (int j = 0; j < len; ++j) {
        float val = ((float)j - radius) / sigma;                 /* :114 */
        val = (float)exp(-0.5 * (double)(val * val));            /* :115 */
        taps[j] = val;
        sum += val;                                              /* :117 */
    }
    for (int j = 0; j < len; ++j) taps[j] = taps[j] / sum;       /* :119-120 */
    return radius;
}

/* which = -1: base kernel, 0..4: level kernels (pyramidata.cu:94-103). */
int orc_taps(int w, int h, int which, float* taps_out)
{
    orc_params P; orc_params_init(&P, w, h);
    return orc_make_taps(which < 0 ? P.base_smooth : P.sigmas[which], taps_out);
}

/* ------------------------------------------------------------------------- */
/* convolve<float> (gpu/kernels/convolution.cu:16-159)                        */
/* rows then columns, zero padding, k = -R..R, sum = fma(data, tap, sum),     */
/* fp32 row intermediate in `buffer`.                                         */
/* ------------------------------------------------------------------------- */
void orc_convolve(float* result, const float* image, float* buffer, int w, int h,
                  const float* taps, int R)
{
    for (int y = 0; y < h; ++y) {                                /* convolve_rows :16-74 */
        const float* row = image + (size_t)y * w;
        float* out = buffer + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float sum = 0.f;
            for (int k = -R; k <= R; ++k) {                      /* :69-70 */
                int xx = x + k;
                float d = (xx >= 0 && xx < w) ? row[xx] : 0.f;   /* :51-52 */
                sum = fmaf(d, taps[R - k], sum);
            }
            out[x] = sum;
        }
    }
    for (int y = 0; y < h; ++y) {                                /* convolve_cols :78-137 */
        float* out = result + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float sum = 0.f;
            for (int k = -R; k <= R; ++k) {                      /* :130-131 */
                int yy = y + k;
                float d = (yy >= 0 && yy < h) ? buffer[(size_t)yy * w + x] : 0.f;  /* :110-111 */
                sum = fmaf(d, taps[R - k], sum);
            }
            out[x] = sum;
        }
    }
}

/* downsample_by_2<float> (gpu/kernels/downsample.cu:6-17) */
void orc_downsample2(float* result, int rw, int rh, const float* source, int sw, int sh)
{
    (void)sh;
    for (int y = 0; y < rh; ++y)
        for (int x = 0; x < rw; ++x)
            result[(size_t)y * rw + x] = source[(size_t)(y * 2) * sw + x * 2];
}

/* subtract<float> (gpu/kernels/cudamath.cu:26-35): result = A - B */
void orc_subtract(const float* A, const float* B, float* result, int w, int h)
{
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; ++i) result[i] = A[i] - B[i];
}

/* mod_2pi_f (gpu/kernels/cudamath.h:82-87) */
static float mod_2pi_f(float x)
{
    while (x > (float)(2 * M_PI)) x -= (float)(2 * M_PI);
    while (x < 0.0f) x += (float)(2 * M_PI);
    return x;
}

/* gradient<float> (gpu/kernels/cudamath.cu:38-54): interior pixels only, the border
 * of `grad` is NOT written.  dx*dx+dy*dy is FMUL(dy,dy) then FFMA(dx,dx,.) in SASS. */
void orc_gradient(const float* source, f2* grad, int w, int h)
{
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            float nx = source[(size_t)y * w + x + 1], px = source[(size_t)y * w + x - 1];
            float ny = source[(size_t)(y + 1) * w + x], py = source[(size_t)(y - 1) * w + x];
            float dx = nx - px, dy = ny - py;
            float g = (float)(0.5 * (double)sqrtf(fmaf(dx, dx, dy * dy)));       /* :51 */
            float r = (g == 0.0f) ? 0.0f
                                  : mod_2pi_f((float)((double)atan2f(dy, dx) + 2 * M_PI)); /* :52 */
            grad[(size_t)y * w + x].x = g;
            grad[(size_t)y * w + x].y = r;
        }
}

/* ------------------------------------------------------------------------- */
/* detect_keypoints / is_maxima / subpixel_refinement                         */
/* (gpu/kernels/keypoint.cu:19-200).  Texture fetches are exact texels        */
/* (linear filter sampled at texel centres, gpu/utils/cudatex2D.cu:15-19).    */
/* ------------------------------------------------------------------------- */
static void refine(int x, int y, const float* cur, const float* down, const float* up, int w,
                   float peak, float edge, float xper, float sigma_0, int num_dogs, int level,
                   f4* result)
{
#define T0(dx_, dy_) cur[(size_t)(y + (dy_)) * w + (x + (dx_))]
#define TU(dx_, dy_) up[(size_t)(y + (dy_)) * w + (x + (dx_))]
#define TD(dx_, dy_) down[(size_t)(y + (dy_)) * w + (x + (dx_))]
    const float c = T0(0, 0);
    /* :119-121 (0.5 * float is exact in either precision) */
    const float fx = 0.5f * (T0(1, 0) - T0(-1, 0));
    const float fy = 0.5f * (T0(0, 1) - T0(0, -1));
    const float fs = 0.5f * (TU(0, 0) - TD(0, 0));
    /* :124-126 float sum widened, minus 2.0*c in double */
    const float fxx = (float)((double)(T0(1, 0) + T0(-1, 0)) - 2.0 * (double)c);
    const float fyy = (float)((double)(T0(0, 1) + T0(0, -1)) - 2.0 * (double)c);
    const float fss = (float)((double)(TU(0, 0) + TD(0, 0)) - 2.0 * (double)c);
    /* :128-135 */
    const float fxy = 0.25f * (((T0(1, 1) + T0(-1, -1)) - T0(-1, 1)) - T0(1, -1));
    const float fxs = 0.25f * (((TU(1, 0) + TD(-1, 0)) - TU(-1, 0)) - TD(1, 0));
    const float fys = 0.25f * (((TU(0, 1) + TD(0, -1)) - TU(0, -1)) - TD(0, 1));
#undef T0
#undef TU
#undef TD
    float A0[4], A1[4], A2[4], tmp[4];
    /* :137-139 */
    if (fxx > 0) { A0[0] = fxx; A0[1] = fxy; A0[2] = fxs; A0[3] = -fx; }
    else         { A0[0] = -fxx; A0[1] = -fxy; A0[2] = -fxs; A0[3] = fx; }
    if (fxy > 0) { A1[0] = fxy; A1[1] = fyy; A1[2] = fys; A1[3] = -fy; }
    else         { A1[0] = -fxy; A1[1] = -fyy; A1[2] = -fys; A1[3] = fy; }
    if (fxs > 0) { A2[0] = fxs; A2[1] = fys; A2[2] = fss; A2[3] = -fs; }
    else         { A2[0] = -fxs; A2[1] = -fys; A2[2] = -fss; A2[3] = fs; }

    const float max_a = fmaxf(fmaxf(A0[0], A1[0]), A2[0]);       /* :142 */
    if (!((double)max_a >= 1e-10)) return;                       /* :143 */
    if (max_a == A1[0]) { memcpy(tmp, A1, 16); memcpy(A1, A0, 16); memcpy(A0, tmp, 16); }
    else if (max_a == A2[0]) { memcpy(tmp, A2, 16); memcpy(A2, A0, 16); memcpy(A0, tmp, 16); }
    /* :150-152, "a -= b*c" is a single FFMA in the reference SASS */
    A0[1] /= A0[0]; A0[2] /= A0[0]; A0[3] /= A0[0];
    A1[1] = fmaf(-A1[0], A0[1], A1[1]); A1[2] = fmaf(-A1[0], A0[2], A1[2]); A1[3] = fmaf(-A1[0], A0[3], A1[3]);
    A2[1] = fmaf(-A2[0], A0[1], A2[1]); A2[2] = fmaf(-A2[0], A0[2], A2[2]); A2[3] = fmaf(-A2[0], A0[3], A2[3]);
    if (fabsf(A2[1]) > fabsf(A1[1])) { memcpy(tmp, A2, 16); memcpy(A2, A1, 16); memcpy(A1, tmp, 16); } /* :154 */
    if (!((double)fabsf(A1[1]) >= 1e-10)) return;                /* :158 */
    A1[2] /= A1[1]; A1[3] /= A1[1];                              /* :159 */
    A2[2] = fmaf(-A2[1], A1[2], A2[2]); A2[3] = fmaf(-A2[1], A1[3], A2[3]);     /* :160 */
    if (!((double)fabsf(A2[2]) >= 1e-10)) return;                /* :161 */
    const float ds = A2[3] / A2[2];                              /* :162 */
    const float dy = fmaf(-ds, A1[2], A1[3]);                    /* :163 */
    const float dx = fmaf(-dy, A0[1], fmaf(-ds, A0[2], A0[3]));  /* :164 */
    /* :165 inner sum float: mul, fma, fma; then one double fma */
    const float inner = fmaf(fs, ds, fmaf(fy, dy, fx * dx));
    const float v = (float)fma((double)inner, 0.5, (double)c);
    /* :166 det = FFMA(fxx, fyy, -(fxy*fxy)) */
    const float tr = fxx + fyy;
    const float s = (tr * tr) / fmaf(fxx, fyy, -(fxy * fxy));
    const float thr = ((edge + 1) * (edge + 1)) / edge;          /* :169 */
    if (fabsf(v) > peak && s < thr && fabsf(dx) < 1 && fabsf(dy) < 1 && fabsf(ds) < 1) {
        f4* r = &result[(size_t)y * w + x];
        r->x = ((float)x + dx) * xper;                           /* :172 */
        r->y = ((float)y + dy) * xper;
        r->z = (float)((double)sigma_0 * pow(2.0, (double)((float)level + ds) / num_dogs) *
                       (double)xper);                            /* :174 */
        r->w = (float)level;
    }
}

/* tex2D<float> of a linear-filter, border-addressed, unnormalised-coordinate texture
 * (cudatex2D.cu:12-19) over a mw x mh float image, CUDA programming guide "Linear Filtering":
 * xB = x - 0.5, i = floor(xB), alpha = frac(xB) held in 1.8 fixed point, texels outside the
 * image are 0.  The detector samples the mask at ((x+.5)*xper, (y+.5)*xper) with xper = 2^o
 * (keypoint.cu:214): alpha = beta = 0 for octave 0 (the texel itself), exactly 1/2 for the others
 * (the mean of the four texels around the block centre), so the weights are exact here. */
static float tex2d_linear_border(const float* img, int mw, int mh, float x, float y)
{
    const float xb = x - 0.5f, yb = y - 0.5f;
    const float fi = floorf(xb), fj = floorf(yb);
    const float a = floorf((xb - fi) * 256.0f + 0.5f) / 256.0f;   /* 8 fractional bits */
    const float b = floorf((yb - fj) * 256.0f + 0.5f) / 256.0f;
    const int i = (int)fi, j = (int)fj;
#define NMO_T(ii, jj) (((ii) >= 0 && (ii) < mw && (jj) >= 0 && (jj) < mh) ? img[(size_t)(jj) * mw + (ii)] : 0.0f)
    const float t00 = NMO_T(i, j), t10 = NMO_T(i + 1, j), t01 = NMO_T(i, j + 1), t11 = NMO_T(i + 1, j + 1);
#undef NMO_T
    return (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
}

/* find_keypoints (keypoint.cu:183-200, 240-251; masked variant :204-222, 226-237 when mask != NULL:
 * pixels whose mask sample is < 1 are skipped, :214).  `result` is the dense per-pixel map; the
 * caller pre-fills it with (-1,-1,-1,-1) (siftfunctions.cu:120 / :86). */
void orc_find_keypoints_masked(const float* cur, const float* down, const float* up, int w, int h,
                               float peak, float edge, float xper, float sigma_0, int num_dogs,
                               int level, f4* result, const float* mask, int mw, int mh)
{
    const float t = 0.8f * peak;                                 /* :195 */
    for (int y = 1; y <= h - 2; ++y)
        for (int x = 1; x <= w - 2; ++x) {
            if (mask && tex2d_linear_border(mask, mw, mh, ((float)x + 0.5f) * xper, ((float)y + 0.5f) * xper) < 1.0f)
                continue;                                        /* :214 */
            const float c = cur[(size_t)y * w + x];
            int is_min = c <= t, is_max = c >= t, ok = 0;
            if (is_min) {
                ok = 1;
                for (int dy = -1; dy <= 1 && ok; ++dy)
                    for (int dx = -1; dx <= 1 && ok; ++dx) {
                        size_t q = (size_t)(y + dy) * w + (x + dx);
                        if ((dx || dy) && !(c < cur[q])) ok = 0;
                        if (!(c < down[q])) ok = 0;
                        if (!(c < up[q])) ok = 0;
                    }
            }
            if (!ok && is_max) {
                ok = 1;
                for (int dy = -1; dy <= 1 && ok; ++dy)
                    for (int dx = -1; dx <= 1 && ok; ++dx) {
                        size_t q = (size_t)(y + dy) * w + (x + dx);
                        if ((dx || dy) && !(c > cur[q])) ok = 0;
                        if (!(c > down[q])) ok = 0;
                        if (!(c > up[q])) ok = 0;
                    }
            }
            if (ok) refine(x, y, cur, down, up, w, peak, edge, xper, sigma_0, num_dogs, level, result);
        }
}

void orc_find_keypoints(const float* cur, const float* down, const float* up, int w, int h,
                        float peak, float edge, float xper, float sigma_0, int num_dogs,
                        int level, f4* result)
{
    orc_find_keypoints_masked(cur, down, up, w, h, peak, edge, xper, sigma_0, num_dogs, level, result, NULL, 0, 0);
}

/* gpu_collate_keypoints_for_level (gpu/sift/pyramidata.cu:9-15,84-91): stable
 * compaction of entries with w >= 0, raster order. */
int orc_collate(const f4* dense, int num_pixels, f4* out)
{
    int n = 0;
    for (int i = 0; i < num_pixels; ++i)
        if (dense[i].w >= 0) out[n++] = dense[i];
    return n;
}

/* ------------------------------------------------------------------------- */
/* detect_orientations (gpu/kernels/orientation.cu:11-129, launch :219-230).  */
/* Histogram smoothing follows the intended Jacobi semantics (= the _naive    */
/* kernel :181-192); accumulation order is raster (the reference's is         */
/* undefined: shared-memory float atomics, :58).  `grad` is the whole 3-level  */
/* gradient buffer of the octave; result is pre-filled with (-1,-1).          */
/* ------------------------------------------------------------------------- */
void orc_orientations(const f4* kp, const f2* grad, int n, int ow, int oh, float gauss_factor,
                      float xper, f2* result, int wmax)
{
    for (int p = 0; p < n; ++p) {
        if (kp[p].w < 0) continue;                               /* :17 */
        const float x = kp[p].x / xper, y = kp[p].y / xper, s = kp[p].z / xper;   /* :19-21 */
        const int xi = (int)((double)x + 0.5), yi = (int)((double)y + 0.5);       /* :23-24 */
        const float sigma_w = gauss_factor * s;                  /* :26 */
        int W = (int)floorf(3 * sigma_w); if (W < 1) W = 1;      /* :27 */
        if (W > wmax) W = wmax;                                  /* :29-30: wmax = 10 (block 22x22) */
        /* :32 index evaluated in float */
        const int gi = (int)(((kp[p].w * (float)oh + (float)yi) * (float)ow) + (float)xi);
        const f2* g = grad + gi;
        float hist[36];
        for (int i = 0; i < 36; ++i) hist[i] = 0.f;
        const int xmin = -W > -xi ? -W : -xi, xmax = W < ow - 1 - xi ? W : ow - 1 - xi;  /* :43-46 */
        const int ymin = -W > -yi ? -W : -yi, ymax = W < oh - 1 - yi ? W : oh - 1 - yi;
        for (int cy = ymin; cy <= ymax; ++cy)
            for (int cx = xmin; cx <= xmax; ++cx) {
                const float dx = (float)(cx + xi) - x, dy = (float)(cy + yi) - y;  /* :52-53 */
                const float r2 = fmaf(dx, dx, dy * dy);          /* :54 (FMUL+FFMA in SASS) */
                if ((double)r2 < (double)(W * W) + 0.6) {        /* :55 */
                    const float wgt = expf(r2 / (2 * sigma_w * sigma_w));          /* :56 (+ exponent) */
                    const f2 gv = g[cy * ow + cx];
                    int bin = (int)floorf((float)((double)(36 * gv.y) / (2 * M_PI)));   /* :57 */
                    bin %= 36; if (bin < 0) bin += 36;
                    hist[bin] += gv.x * wgt;                     /* :58 */
                }
            }
        for (int iter = 0; iter < 6; ++iter) {                   /* :71-85 / :181-192 */
            float t[36];
            for (int i = 0; i < 36; ++i) {
                const float prev = hist[(i + 35) % 36], next = hist[(i + 1) % 36];
                t[i] = (float)((double)((prev + hist[i]) + next) / 3.0);
            }
            memcpy(hist, t, sizeof(t));
        }
        float maxh = 0.f;
        for (int i = 0; i < 36; ++i) maxh = fmaxf(maxh, hist[i]);                 /* :93-95 */
        const float thr = (float)((double)maxh * 0.8);           /* :96 */
        int nangles = 0;
        for (int i = 0; i < 36 && nangles < 2; ++i) {            /* :103-127 */
            const float h0 = hist[i], hm = hist[(i + 35) % 36], hp = hist[(i + 1) % 36];
            if (h0 > thr && h0 > hm && h0 > hp) {
                const float di = (float)(-0.5 * (double)(hp - hm) / (double)((hp + hm) - 2 * h0));  /* :108 */
                const float th = (float)(2 * M_PI * ((double)((float)i + di) + 0.5) / 36);          /* :109 */
                if (nangles == 0) result[p].x = th; else result[p].y = th;
                ++nangles;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* compute_sift_descriptors (gpu/kernels/descriptor.cu:32-145, launch :243).  */
/* First orientation only; positive-exponent window; only the diagonal 16x16  */
/* chunks of the window are visited (:94-97,142-143); no normalisation.       */
/* ------------------------------------------------------------------------- */
void orc_descriptors(const f4* kp, const f2* orients, const f2* grad, int n, int ow, int oh,
                     int num_dogs, float xper, float* desc, float* xp, float* yp)
{
    for (int p = 0; p < n; ++p) {
        const float x = kp[p].x / xper, y = kp[p].y / xper, s = kp[p].z / xper;   /* :41-43 */
        const int xi = (int)((double)x + 0.5), yi = (int)((double)y + 0.5);
        const int si = (int)kp[p].w;                             /* :47 */
        if (xi < 0 || xi >= ow || yi < 0 || yi >= oh || si < 0 || si >= num_dogs) continue;  /* :49 */
        const float SBP = (float)((double)(3 * s) + 1.e-07);     /* :54 */
        const int W = (int)floor(sqrt(2.0) * (double)SBP * 5 / 2.0 + 0.5);        /* :55 */
        const int xmin = -W > -xi ? -W : -xi, xmax = W < ow - 1 - xi ? W : ow - 1 - xi;
        const int ymin = -W > -yi ? -W : -yi, ymax = W < oh - 1 - yi ? W : oh - 1 - yi;
        const int max_dims = (xmax - xmin) > (ymax - ymin) ? (xmax - xmin) : (ymax - ymin);
        const int chunks = (int)ceilf((max_dims + 1.f) / 16);    /* :65 */
        float* d = desc + (size_t)p * 128;
        for (int i = 0; i < 128; ++i) d[i] = 0.f;                /* :73 */
        xp[p] = kp[p].x; yp[p] = kp[p].y;                        /* :76 */
        float* pix_d = d + 2 * 32 + 2 * 8;                       /* :81 */
        const f2* g = grad + ((size_t)(si * oh + yi) * ow + xi); /* :83-84 */
        const float th0 = orients[p].x;
        const double st0 = (double)sinf(th0), ct0 = (double)cosf(th0);            /* :90-91 */
        for (int c = 0; c < chunks; ++c)
            for (int ty = 0; ty < 16; ++ty)
                for (int tx = 0; tx < 16; ++tx) {
                    const int cx = xmin + tx + 16 * c, cy = ymin + ty + 16 * c;
                    if (!(cx <= xmax && cy <= ymax)) continue;   /* :96 */
                    const float mod = g[cy * ow + cx].x, ang = g[cy * ow + cx].y;
                    const float theta = mod_2pi_f(ang - th0);    /* :100 */
                    const float dx = (float)(xi + cx) - x, dy = (float)(yi + cy) - y;
                    /* :104-105 double; the products contract to DFMA in the reference SASS */
                    const float nx = (float)(fma(ct0, (double)dx, st0 * (double)dy) / (double)SBP);
                    const float ny = (float)(fma(ct0, (double)dy, -(st0 * (double)dx)) / (double)SBP);
                    const float nt = (float)((double)(8 * theta) / (2 * M_PI));   /* :107 */
                    const float win = (float)exp((double)fmaf(nx, nx, ny * ny) / 8.0);  /* :108 */
                    const int binx = (int)floor((double)nx - 0.5);
                    const int biny = (int)floor((double)ny - 0.5);
                    const int bint = (int)floorf(nt);
                    const float rbinx = (float)((double)nx - ((double)binx + 0.5));
                    const float rbiny = (float)((double)ny - ((double)biny + 0.5));
                    const float rbint = nt - (float)bint;
                    for (int dbx = 0; dbx < 2; ++dbx)
                        for (int dby = 0; dby < 2; ++dby)
                            for (int dbt = 0; dbt < 2; ++dbt) {
                                if (binx + dbx >= -2 && binx + dbx < 2 && biny + dby >= -2 && biny + dby < 2) {
                                    const float wt = win * mod * fabsf(1.f - dbx - rbinx) *
                                                     fabsf(1.f - dby - rbiny) * fabsf(1.f - dbt - rbint);
                                    const int loc = (binx + dbx) * 8 + (biny + dby) * 32 + ((bint + dbt) % 8);
                                    pix_d[loc] += wt;            /* :135 */
                                }
                            }
                }
    }
}

/* ------------------------------------------------------------------------- */
/* Matching (gpu/kernels/match.cu:14-117, gpu/sift/siftfunctions.cu:15-40)    */
/* ------------------------------------------------------------------------- */
static float dist2(const float* a, const float* b, int dim)
{
    float acc = 0.f;
    for (int i = 0; i < dim; ++i) {                              /* match.cu:36-42 */
        float t = a[i] - b[i];
        acc = fmaf(t, t, acc);
    }
    return acc;
}

/* D[a][b], row-major nA x nB (what compute_sift_matches leaves in `distance`). */
void orc_dist2(const float* A, int nA, const float* B, int nB, int dim, float* D)
{
#pragma omp parallel for schedule(static)
    for (int a = 0; a < nA; ++a)
        for (int b = 0; b < nB; ++b)
            D[(size_t)a * nB + b] = dist2(A + (size_t)a * dim, B + (size_t)b * dim, dim);
}

/* set_matches rule on one row given a distance getter (match.cu:88-116). */
static void row_rule(float m1, int i1, float m2, float ambiguity, int* out)
{
    if (m2 > 0) {                                                /* :107 */
        float a = m1 / m2;
        *out = (a < ambiguity) ? i1 : -1;
    }
}

/* get_sift_matches on a materialised matrix. */
void orc_set_matches(int* result, const float* distance, int rows, int cols, int buffer_width,
                     float ambiguity)
{
    for (int i = 0; i < rows; ++i) {
        const float* row = distance + (size_t)i * buffer_width;
        float m1 = row[0], m2 = 2139095040.0f;                   /* :90-91 (0x7f800000 as an int) */
        int i1 = 0;
        for (int j = 1; j < cols; ++j) {
            float cur = row[j];
            if (cur < m1) { m2 = m1; i1 = j; m1 = cur; }
            else if (cur < m2) m2 = cur;
        }
        row_rule(m1, i1, m2, ambiguity, &result[i]);
    }
}

/* Fused: per-row top-2 records (d1, i1, d2) without materialising D. */
void orc_match_top2(const float* A, int nA, const float* B, int nB, float* d1, int* i1, float* d2)
{
#pragma omp parallel for schedule(static)
    for (int a = 0; a < nA; ++a) {
        const float* av = A + (size_t)a * 128;
        float m1 = dist2(av, B, 128), m2 = 2139095040.0f;
        int idx = 0;
        for (int b = 1; b < nB; ++b) {
            float cur = dist2(av, B + (size_t)b * 128, 128);
            if (cur < m1) { m2 = m1; idx = b; m1 = cur; }
            else if (cur < m2) m2 = cur;
        }
        d1[a] = m1; i1[a] = idx; d2[a] = m2;
    }
}

/* --- sharded-database helpers (checker for the multi-GPU path; no reference equivalent) ---
 * True per-row top-2 (d1, i1 + index_offset, d2) with +inf for missing entries, ties to
 * the lowest index; rec4[a] = {d1, bits(i1), d2, 0}. */
void orc_match_top2_true(const float* A, int nA, const float* B, int nB, int index_offset, float* rec4)
{
#pragma omp parallel for schedule(static)
    for (int a = 0; a < nA; ++a) {
        float m1 = INFINITY, m2 = INFINITY; int idx = -1;
        for (int b = 0; b < nB; ++b) {
            float cur = dist2(A + (size_t)a * 128, B + (size_t)b * 128, 128);
            if (cur < m1) { m2 = m1; idx = b; m1 = cur; }
            else if (cur < m2) m2 = cur;
        }
        int gi = idx < 0 ? -1 : idx + index_offset;
        rec4[4 * a] = m1; memcpy(&rec4[4 * a + 1], &gi, 4); rec4[4 * a + 2] = m2; rec4[4 * a + 3] = 0.f;
    }
}

/* Merge shard-major records and apply the reference's rule.  The sequential scan of
 * match.cu:88-105 ends with min2 = (idx == 0) ? min(2139095040.0f, d2) : d2 where
 * (d1, idx, d2) is the true top-2: the odd start value of min2 only survives while column
 * 0 stays the minimum (the first displacement overwrites it with D[0]). */
void orc_merge_top2(const float* recs4, int n_shards, int nA, float ambiguity, int* match_io)
{
    for (int a = 0; a < nA; ++a) {
        float m1 = INFINITY, m2 = INFINITY; int mi = -1;
        for (int s = 0; s < n_shards; ++s) {
            const float* r = recs4 + 4 * ((size_t)s * nA + a);
            float u1 = r[0], u2 = r[2]; int j1; memcpy(&j1, &r[1], 4);
            if (j1 < 0) continue;
            if (mi < 0 || u1 < m1 || (u1 == m1 && j1 < mi)) {
                float t = m1; m1 = u1; u1 = t;
                t = m2; m2 = u2; u2 = t;
                mi = j1;
            }
            if (u1 < m2) m2 = u1;
        }
        if (mi < 0) continue;
        float min2 = (mi == 0) ? fminf(2139095040.0f, m2) : m2;
        row_rule(m1, mi, min2, ambiguity, &match_io[a]);
    }
}

/* compute_sift_matches semantics (match_io in/out). */
void orc_match(const float* A, int nA, const float* B, int nB, float ambiguity, int* match_io)
{
    float* d1 = (float*)malloc(sizeof(float) * nA);
    float* d2 = (float*)malloc(sizeof(float) * nA);
    int* i1 = (int*)malloc(sizeof(int) * nA);
    orc_match_top2(A, nA, B, nB, d1, i1, d2);
    for (int a = 0; a < nA; ++a) row_rule(d1[a], i1[a], d2[a], ambiguity, &match_io[a]);
    free(d1); free(d2); free(i1);
}

/* ------------------------------------------------------------------------- */
/* Whole frame: the client loop of SURVEY.md 3.1 around                        */
/* compute_dog/_gradients/_keypoints/_orientations/_descriptors                */
/* (gpu/sift/siftfunctions.cu:42-181).  Same signature and dump layout as      */
/* nmref_sift_frame(_masked) in oracle/ref_driver.cu; mask = NULL or a w x h   */
/* float image for compute_keypoints_with_mask (siftfunctions.cu:65-98).       */
/* cfg6 = {peak_threshold, edge_threshold(<=0 default), num_octaves(<=0        */
/*         default), capacity(<=0: 2048), clear_grad, orient_mode}             */
/* orient_mode 0: public-API orientation (window clamp 10); 1: arithmetic of   */
/* kernel_orientations_naive (no clamp); 2: orientations injected (orient_in,  */
/* float2 per keypoint in segment order).                                      */
/* ------------------------------------------------------------------------- */
int orc_sift_frame_masked(const float* image, int w, int h, const float* cfg5,
                          float* desc, float* xo, float* yo, int* num_items,
                          float* levels_out, float* kpts_out, float* orient_out, int* seg_counts,
                          int kp_cap, float* grad_out, const float* orient_in, const float* mask)
{
    const int orient_mode = (int)cfg5[5];
    int inject_off = 0;
    orc_params P; orc_params_init(&P, w, h);
    P.peak_threshold = cfg5[0];
    if (cfg5[1] > 0.f) P.edge_threshold = cfg5[1];
    if ((int)cfg5[2] > 0) P.num_octaves = (int)cfg5[2];
    int capacity = (int)cfg5[3] > 0 ? (int)cfg5[3] : 2048;
    const int clear_grad = (int)cfg5[4];

    const size_t N = (size_t)w * h;
    float* oct[6]; float* dog[5];
    for (int i = 0; i < 6; ++i) oct[i] = (float*)calloc(N, sizeof(float));
    for (int i = 0; i < 5; ++i) dog[i] = (float*)calloc(N, sizeof(float));
    float* buffer = (float*)calloc(N, sizeof(float));
    f2* grad = (f2*)calloc(N * 5, sizeof(f2));                   /* pyramidata.cu:46 */
    f4* dense = (f4*)malloc(N * sizeof(f4));
    f4* coll[3]; f2* orient[3]; int cnt[3];
    for (int i = 0; i < 3; ++i) { coll[i] = (f4*)malloc(N * sizeof(f4)); orient[i] = (f2*)malloc(N * sizeof(f2)); }
    float taps[96]; int R;
    float* dloc = (float*)malloc((size_t)capacity * 128 * sizeof(float));
    float* xl = (float*)malloc((size_t)capacity * sizeof(float));
    float* yl = (float*)malloc((size_t)capacity * sizeof(float));

    int items = 0, kp_off = 0;
    size_t lev_off = 0, grad_off = 0;
    R = orc_make_taps(P.base_smooth, taps);
    orc_convolve(oct[0], image, buffer, w, h, taps, R);
    for (int o = 0; o < P.num_octaves; ++o) {
        const int ow = w >> o, oh = h >> o;
        const size_t n = (size_t)ow * oh;
        const float xper = (float)pow(2.0, o);                   /* siftfunctions.cu:118 */
        for (int i = 0; i < P.num_sigmas; ++i) {
            R = orc_make_taps(P.sigmas[i], taps);
            orc_convolve(oct[i + 1], oct[i], buffer, ow, oh, taps, R);
        }
        for (int i = 0; i < 5; ++i) orc_subtract(oct[i + 1], oct[i], dog[i], ow, oh);   /* :42-51 */
        if (clear_grad) memset(grad, 0, N * 5 * sizeof(f2));
        for (int i = 0; i <= 2; ++i) orc_gradient(oct[i + 1], grad + (size_t)i * n, ow, oh);  /* :53-63 */
        cnt[0] = cnt[1] = cnt[2] = 0;
        int stopped = 0;
        for (int l = 0; l < 3 && !stopped; ++l) {
            for (size_t i = 0; i < n; ++i) { dense[i].x = dense[i].y = dense[i].z = dense[i].w = -1.f; }
            orc_find_keypoints_masked(dog[l + 1], dog[l], dog[l + 2], ow, oh, P.peak_threshold,
                                      P.edge_threshold, xper, P.sigma_0, P.num_dog_levels, l, dense,
                                      mask, w, h);               /* :119-126 / :83-92 with a mask */
            cnt[l] = orc_collate(dense, (int)n, coll[l]);        /* :144 */
            if (cnt[l] == 0) { stopped = 1; break; }             /* :145 `return` */
            for (int i = 0; i < cnt[l]; ++i) { orient[l][i].x = -1.f; orient[l][i].y = -1.f; }
            if (orient_mode == 2) {
                memcpy(orient[l], orient_in + 2 * (size_t)inject_off, cnt[l] * sizeof(f2));
                inject_off += cnt[l];
            } else
                orc_orientations(coll[l], grad, cnt[l], ow, oh, 1.5f, xper, orient[l],
                                 orient_mode == 1 ? 1 << 20 : 10);                      /* :149 */
        }
        for (int l = 0; l < 3; ++l) {                            /* compute_descriptors :154-181 */
            if (cnt[l] == 0) break;                              /* :160 */
            int num = cnt[l];
            if (num + items > capacity) num = capacity - items;  /* :167-169 */
            if (num > 0) {
                orc_descriptors(coll[l], orient[l], grad, num, ow, oh, P.num_dog_levels, xper,
                                dloc + (size_t)items * 128, xl + items, yl + items);
                items += num;
            }
        }
        if (levels_out)
            for (int i = 0; i < 6; ++i) { memcpy(levels_out + lev_off, oct[i], n * sizeof(float)); lev_off += n; }
        if (grad_out) { memcpy(grad_out + grad_off, grad, 3 * n * sizeof(f2)); grad_off += 3 * n * 2; }
        if (seg_counts)
            for (int l = 0; l < 3; ++l) {
                seg_counts[o * 3 + l] = cnt[l];
                int take = cnt[l];
                if (kp_off + take > kp_cap) take = kp_cap - kp_off;
                if (take > 0 && kpts_out) memcpy(kpts_out + 4 * (size_t)kp_off, coll[l], take * sizeof(f4));
                if (take > 0 && orient_out) memcpy(orient_out + 2 * (size_t)kp_off, orient[l], take * sizeof(f2));
                if (take > 0) kp_off += take;
            }
        if (o + 1 < P.num_octaves) orc_downsample2(oct[0], ow / 2, oh / 2, oct[3], ow, oh);
    }
    *num_items = items;
    if (desc) memcpy(desc, dloc, (size_t)items * 128 * sizeof(float));
    if (xo) memcpy(xo, xl, items * sizeof(float));
    if (yo) memcpy(yo, yl, items * sizeof(float));
    for (int i = 0; i < 6; ++i) free(oct[i]);
    for (int i = 0; i < 5; ++i) free(dog[i]);
    for (int i = 0; i < 3; ++i) { free(coll[i]); free(orient[i]); }
    free(buffer); free(grad); free(dense); free(dloc); free(xl); free(yl);
    return P.num_octaves;
}

int orc_sift_frame(const float* image, int w, int h, const float* cfg5,
                   float* desc, float* xo, float* yo, int* num_items,
                   float* levels_out, float* kpts_out, float* orient_out, int* seg_counts,
                   int kp_cap, float* grad_out, const float* orient_in)
{
    return orc_sift_frame_masked(image, w, h, cfg5, desc, xo, yo, num_items, levels_out, kpts_out, orient_out,
                                 seg_counts, kp_cap, grad_out, orient_in, NULL);
}

/* CPU-baseline leg of bench.py: n_frames independent frames on `threads` OpenMP
 * threads (one frame per task).  counts_out[f] = descriptors of frame f. */
int orc_sift_batch(const float* frames, int n_frames, int w, int h, const float* cfg5,
                   int threads, int* counts_out)
{
    int cap = (int)cfg5[3] > 0 ? (int)cfg5[3] : 2048;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int f = 0; f < n_frames; ++f) {
        float* d = (float*)malloc((size_t)cap * 128 * sizeof(float));
        int n = 0;
        orc_sift_frame(frames + (size_t)f * w * h, w, h, cfg5, d, NULL, NULL, &n,
                       NULL, NULL, NULL, NULL, 0, NULL, NULL);
        counts_out[f] = n;
        free(d);
    }
    return 0;
}

/* ========================================================================= */
/* align_points + RANSAC (gpu/kernels/ransac.cu) with  */
/* the one-sided Jacobi SVD of gpu/kernels/svd.cu (the reference's port of    */
/* GSL's gsl_linalg_SV_decomp_jacobi).  Restated from the published           */
/* algorithm (Hestenes / Nash plane rotations with GSL's error-estimate       */
/* skip rule) in the order of operations of the reference; fp32 throughout,   */
/* with explicit fmaf where the reference's GPU build contracts a*b+c*d (the   */
/* forms were established against tests/golden/ransac_400.npz): translation    */
/* and homography hypotheses are bitwise the reference's, similarity within    */
/* one ulp of one element, see tests/test_oracle_golden.py.                    */
/* ========================================================================= */

/* establish_correspondences (ransac.cu:29-48) */
void orc_align_points(const float* src_x, const float* src_y, const float* dst_x, const float* dst_y,
                      float* c_src_x, float* c_src_y, float* c_dst_x, float* c_dst_y,
                      const int* matches, int num_pts)
{
    for (int i = 0; i < num_pts; ++i) {
        const int m = matches[i];
        if (m != -1) { c_src_x[i] = src_x[i]; c_src_y[i] = src_y[i]; c_dst_x[i] = dst_x[m]; c_dst_y[i] = dst_y[m]; }
        else c_src_x[i] = c_src_y[i] = c_dst_x[i] = c_dst_y[i] = -1.f;
    }
}

#define ORC_EPS 1.1920928955078125e-07f          /* svd.cu:33 */

/* scaled 2-norm of column `col` of the rows x cols row-major matrix (svd.cu:159-195, :86-93) */
static float col_norm(const float* A, int rows, int cols, int col)
{
    float scale = 0.f, ssq = 1.f;
    if (rows == 1) return fabsf(A[col]);
    for (int i = 0; i < rows; ++i) {
        const float x = A[i * cols + col];
        if (x != 0.f) {
            const float ax = fabsf(x);
            if (scale < ax) { ssq = fmaf(ssq * (scale / ax), scale / ax, 1.f); scale = ax; }
            else ssq = fmaf(ax / scale, ax / scale, ssq);
        }
    }
    return scale * sqrtf(ssq);
}

/* svd.cu:133-157 */
static float hyp(float x, float y)
{
    const float xa = fabsf(x), ya = fabsf(y);
    const float mn = xa < ya ? xa : ya, mx = xa < ya ? ya : xa;
    if (mn == 0.f) return mx;
    const float u = mn / mx;
    return mx * sqrtf(fmaf(u, u, 1.f));
}

/* linalg_SV_decomp_jacobi (svd.cu:197-360): A (M x N, overwritten), Q (N x N) = right vectors.
 * Returns 1 when the sweeps converged.  The singular values / column normalisation at the end
 * (:318-352) do not touch Q, which is all the callers read, and are omitted. */
static int jacobi_sv(float* A, int M, int N, float* Q)
{
    float S[16];
    int count = 1, sweep = 0, sweepmax = 5 * N;
    const float tol = (float)(10 * M) * ORC_EPS;               /* :211 */
    if (sweepmax < 12) sweepmax = 12;                          /* :214 */
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) Q[i * N + j] = i == j ? 1.f : 0.f;   /* :217 */
    for (int j = 0; j < N; ++j) S[j] = ORC_EPS * col_norm(A, M, N, j);                          /* :222-227 */
    while (count > 0 && sweep <= sweepmax) {                   /* :231 */
        count = N * (N - 1) / 2;
        for (int j = 0; j < N - 1; ++j)
            for (int k = j + 1; k < N; ++k) {
                float p = 0.f;
                for (int i = 0; i < M; ++i) p = fmaf(A[i * N + j], A[i * N + k], p);   /* ddot :123-131 */
                p *= 2.0f;                                     /* :259 */
                const float a = col_norm(A, M, N, j), b = col_norm(A, M, N, k);
                const float q = fmaf(a, a, -(b * b));
                const float v = hyp(p, q);
                const float ea = S[j], eb = S[k];
                const int sorted = a >= b;
                const int orthog = fabsf(p) <= tol * (a * b);
                const int noisya = a < ea, noisyb = b < eb;
                if (sorted && (orthog || noisya || noisyb)) { --count; continue; }   /* :277-281 */
                float c, s;
                if (v == 0.f || !sorted) { c = 0.f; s = 1.f; }  /* :284-288 */
                else {
                    c = (float)sqrt((double)(v + q) / (2.0 * (double)v));     /* :291 (2.0 * v in double) */
                    s = (float)((double)p / (2.0 * (double)v * (double)c));   /* :292 */
                }
                for (int i = 0; i < M; ++i) {                  /* :296-302 */
                    const float Aik = A[i * N + k], Aij = A[i * N + j];
                    A[i * N + j] = fmaf(Aik, s, Aij * c);
                    A[i * N + k] = fmaf(Aik, c, -Aij * s);
                }
                S[j] = fmaf(fabsf(c), ea, fabsf(s) * eb);      /* :304-305 */
                S[k] = fmaf(fabsf(s), ea, fabsf(c) * eb);
                for (int i = 0; i < N; ++i) {                  /* :308-314 */
                    const float Qij = Q[i * N + j], Qik = Q[i * N + k];
                    Q[i * N + j] = fmaf(Qij, c, Qik * s);
                    Q[i * N + k] = fmaf(Qik, c, -Qij * s);
                }
            }
        ++sweep;
    }
    return count > 0 ? 0 : 1;
}

/* inv(dst_transform) * H * src_transform, expanded (ransac.cu:201-212 = :424-434) */
static void denormalise(const float H[9], float s1, float s2, float tx1, float ty1, float tx2, float ty2, float R[9])
{
    /* the GPU build contracts a*b + c*d to fma(a, b, c*d) and x - a*b to fma(-a, b, x) (established on the
     * vectors the reference produced, tests/golden/ransac_400.npz) */
    const float sty = s1 * ty1, stx = s1 * tx1;
    const float w8 = fmaf(-stx, H[6], fmaf(-sty, H[7], H[8]));
    R[0] = fmaf(s1 * tx2, H[6], s1 * H[0] / s2);
    R[1] = fmaf(s1 * tx2, H[7], s1 * H[1] / s2);
    R[2] = fmaf(tx2, w8, fmaf(-stx, H[0], fmaf(-sty, H[1], H[2])) / s2);
    R[3] = fmaf(s1 * ty2, H[6], s1 * H[3] / s2);
    R[4] = fmaf(s1 * ty2, H[7], s1 * H[4] / s2);
    R[5] = fmaf(ty2, w8, fmaf(-stx, H[3], fmaf(-sty, H[4], H[5])) / s2);
    R[6] = s1 * H[6];
    R[7] = s1 * H[7];
    R[8] = w8;
}

/* compute_homography_2 (ransac.cu:84-214): normalised 4-point DLT, null vector by Jacobi SVD */
static void homography4(const float sx[4], const float sy[4], const float dx[4], const float dy[4], float R[9])
{
    const float smx = (sx[0] + sx[1] + sx[2] + sx[3]) * 0.25f, smy = (sy[0] + sy[1] + sy[2] + sy[3]) * 0.25f;
    const float dmx = (dx[0] + dx[1] + dx[2] + dx[3]) * 0.25f, dmy = (dy[0] + dy[1] + dy[2] + dy[3]) * 0.25f;
    float sv = 0.f, dv = 0.f;
    for (int i = 0; i < 4; ++i) {
        sv += fmaf(sx[i] - smx, sx[i] - smx, (sy[i] - smy) * (sy[i] - smy));
        dv += fmaf(dx[i] - dmx, dx[i] - dmx, (dy[i] - dmy) * (dy[i] - dmy));
    }
    sv *= 0.25f; dv *= 0.25f;
    const float s1 = sqrtf(2.0f) / sqrtf(sv), s2 = sqrtf(2.0f) / sqrtf(dv);      /* :117-118 */
    float X[81], V[81];
    for (int i = 0; i < 4; ++i) {
        const float a = (sx[i] - smx) * s1, b = (sy[i] - smy) * s1, u = (dx[i] - dmx) * s2, w = (dy[i] - dmy) * s2;
        float* r1 = X + (2 * i) * 9; float* r2 = X + (2 * i + 1) * 9;
        r1[0] = r1[1] = r1[2] = 0.f; r1[3] = -a; r1[4] = -b; r1[5] = -1.f; r1[6] = w * a; r1[7] = w * b; r1[8] = w;
        r2[0] = a; r2[1] = b; r2[2] = 1.f; r2[3] = r2[4] = r2[5] = 0.f; r2[6] = -u * a; r2[7] = -u * b; r2[8] = -u;
        if (i == 3) {                                                             /* :159-176 */
            float* r3 = X + 72;
            r3[0] = -w * a; r3[1] = -w * b; r3[2] = -w; r3[3] = u * a; r3[4] = u * b; r3[5] = u; r3[6] = r3[7] = r3[8] = 0.f;
        }
    }
    jacobi_sv(X, 9, 9, V);
    float H[9];
    const float div = V[80];                                                      /* :181 */
    for (int i = 0; i < 8; ++i) H[i] = V[i * 9 + 8] / div;
    H[8] = 1.f;
    denormalise(H, s1, s2, smx, smy, dmx, dmy, R);
}

/* compute_similarity_transform (ransac.cu:320-435): X is 4 x 5, V 5 x 5 */
static void similarity2(const float sx[2], const float sy[2], const float dx[2], const float dy[2], float R[9])
{
    const float smx = (sx[0] + sx[1]) * 0.5f, smy = (sy[0] + sy[1]) * 0.5f;
    const float dmx = (dx[0] + dx[1]) * 0.5f, dmy = (dy[0] + dy[1]) * 0.5f;
    float sv = 0.f, dv = 0.f;
    for (int i = 0; i < 2; ++i) {
        sv += fmaf(sx[i] - smx, sx[i] - smx, (sy[i] - smy) * (sy[i] - smy));
        dv += fmaf(dx[i] - dmx, dx[i] - dmx, (dy[i] - dmy) * (dy[i] - dmy));
    }
    sv = (float)((double)sv * 0.5); dv = (float)((double)dv * 0.5);               /* :336-337 (double literal) */
    const float r2 = sqrtf(2.0f);
    const float s1 = r2 / sqrtf(sv), s2 = r2 / sqrtf(dv);
    float X[20], V[25];
    for (int i = 0; i < 2; ++i) {
        const float a = (sx[i] - smx) * s1, b = (sy[i] - smy) * s1, u = (dx[i] - dmx) * s2, w = (dy[i] - dmy) * s2;
        float* r1 = X + (2 * i) * 5; float* r2_ = X + (2 * i + 1) * 5;
        r1[0] = a; r1[1] = 1.f; r1[2] = -b; r1[3] = 0.f; r1[4] = u;               /* :367-372 */
        r2_[0] = b; r2_[1] = 0.f; r2_[2] = a; r2_[3] = 1.f; r2_[4] = w;           /* :374-379 */
    }
    jacobi_sv(X, 4, 5, V);
    const float div = V[24];
    const float a0 = -V[4] / div, a1 = -V[9] / div, b0 = -V[14] / div, b1 = -V[19] / div;   /* :384-388 */
    const float H[9] = {a0, -b0, a1, b0, a0, b1, 0.f, 0.f, 1.f};
    denormalise(H, s1, s2, smx, smy, dmx, dmy, R);
}

/* eval_transformation (ransac.cu:61-82): squared reprojection error < threshold */
static int count_inliers(const float* sx, const float* sy, const float* dx, const float* dy, int n, const float H[9], float thr)
{
    int inl = 0;
    for (int i = 0; i < n; ++i)
        if (sx[i] >= 0) {
            float x = fmaf(H[0], sx[i], H[1] * sy[i]) + H[2];
            float y = fmaf(H[3], sx[i], H[4] * sy[i]) + H[5];
            const float z = fmaf(H[6], sx[i], H[7] * sy[i]) + H[8];
            x /= z; y /= z;
            const float d2 = fmaf(dx[i] - x, dx[i] - x, (dy[i] - y) * (dy[i] - y));
            if (d2 < thr) ++inl;
        }
    return inl;
}

/* The three hypothesis kernels (ransac.cu:437-520) on a caller-supplied random index list
 * (kind 0: translation, 1 index / iteration; 1: similarity, 2; 2: homography, 4).  Iterations with a
 * repeated index keep H = 0 and 0 inliers (:446, :497-502; the buffers start zeroed :556-561). */
void orc_ransac_hypotheses(int kind, const float* sx, const float* sy, const float* dx, const float* dy, int n,
                           const int* rand_list, int iterations, float thr, float* H_out, int* inliers_out)
{
    const int m = kind == 0 ? 1 : kind == 1 ? 2 : 4;
    for (int it = 0; it < iterations; ++it) {
        float* H = H_out + (size_t)it * 9;
        for (int i = 0; i < 9; ++i) H[i] = 0.f;
        inliers_out[it] = 0;
        const int* r = rand_list + (size_t)it * m;
        int dup = 0;
        for (int a = 0; a < m; ++a) for (int b = a + 1; b < m; ++b) if (r[a] == r[b]) dup = 1;
        if (dup) continue;
        float px[4], py[4], qx[4], qy[4];
        for (int a = 0; a < m; ++a) { px[a] = sx[r[a]]; py[a] = sy[r[a]]; qx[a] = dx[r[a]]; qy[a] = dy[r[a]]; }
        if (kind == 0) {                                       /* compute_translation :304-310 */
            H[0] = H[4] = H[8] = 1.f; H[2] = qx[0] - px[0]; H[5] = qy[0] - py[0];
        } else if (kind == 1) similarity2(px, py, qx, qy, H);
        else homography4(px, py, qx, qy, H);
        inliers_out[it] = count_inliers(sx, sy, dx, dy, n, H, thr);
    }
}

/* thrust::max_element over the inlier counts (first maximum, ransac.cu:566-570): best iteration */
int orc_ransac_best(const int* inliers, int iterations)
{
    int best = 0;
    for (int i = 1; i < iterations; ++i) if (inliers[i] > inliers[best]) best = i;
    return best;
}


/* ========================================================================= */
/* input preprocessing.                                 */
/* ========================================================================= */

/* grayscale (gpu/kernels/bgra_2_gray.cu:8-19): double literals, double sum, rounded to float on the store */
void orc_grayscale_bgra(const unsigned char* bgra, float* out, long long n)
{
    for (long long i = 0; i < n; ++i)
        out[i] = (float)(0.07 * bgra[4 * i] + 0.72 * bgra[4 * i + 1] + 0.21 * bgra[4 * i + 2]);
}

/* cast<float, unsigned char> (cast.cu:7-21).  (unsigned char)src of the device code: cvt.rzi.u32.f32 (saturating:
 * negative and NaN -> 0, above 2^32-1 -> 0xffffffff) followed by the truncation to 8 bits. */
void orc_cast_f32_u8(const float* src, long long n, unsigned char* dst, unsigned char max_val)
{
    for (long long i = 0; i < n; ++i) {
        const float v = src[i];
        if (max_val != 0 && v >= (float)max_val) { dst[i] = max_val; continue; }
        unsigned u;
        if (!(v > 0.f)) u = 0u;
        else if (v >= 4294967296.f) u = 0xffffffffu;
        else u = (unsigned)v;
        dst[i] = (unsigned char)(u & 0xffu);
    }
}

/* undistort (undistort.cu:6-47).  powf(a, 2) and powf(a, 3) are evaluated as repeated products here; the
 * device powf is accurate to 2 ulp in general, parity is tolerance based (tests: 2e-6 relative). */
void orc_undistort_map(const float* x, const float* y, long long n, const float* camera_matrix,
                       const float* distortion_coeffs, float* u, float* v)
{
    const float k1 = distortion_coeffs[0], k2 = distortion_coeffs[1], k3 = distortion_coeffs[2];
    const float fx = camera_matrix[0], fy = camera_matrix[1], cx = camera_matrix[2], cy = camera_matrix[3];
    for (long long i = 0; i < n; ++i) {
        float uu = x[i]; uu -= cx; uu /= fx;
        float vv = y[i]; vv -= cy; vv /= fy;
        const float r2 = uu * uu + vv * vv;
        const float kr = 1 + k1 * r2 + k2 * (r2 * r2) + k3 * (r2 * r2 * r2);
        uu *= kr; uu *= fx; uu += cx;
        vv *= kr; vv *= fy; vv += cy;
        u[i] = uu; v[i] = vv;
    }
}

/* ========================================================================= */
/* mosaic rendering (gpu/kernels/resample.cu).          */
/* The texture unit is restated from the CUDA programming guide (linear        */
/* filtering, un-normalised coordinates, border addressing): xB = x - 0.5,      */
/* i = floor(xB), alpha = frac(xB) held in 9-bit fixed point with 8 fractional  */
/* bits; out-of-range texels read 0.  The hardware's internal rounding is not   */
/* documented, so byte results are compared with a tolerance of one LSB        */
/* (tests/test_oracle_golden.py).                                               */
/* ========================================================================= */
static float tex_fetch(const float* img, int w, int h, int nch, int ch, int ix, int iy)
{
    if (ix < 0 || ix >= w || iy < 0 || iy >= h) return 0.f;
    return img[((size_t)iy * w + ix) * nch + ch];
}

/* bilinear sample of channel ch of an interleaved float image at texture coordinates (x, y) */
static float tex2d_linear(const float* img, int w, int h, int nch, int ch, float x, float y)
{
    const float xb = x - 0.5f, yb = y - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    const float a = floorf((xb - fx) * 256.f + 0.5f) / 256.f, b = floorf((yb - fy) * 256.f + 0.5f) / 256.f;
    const int i = (int)fx, j = (int)fy;
    return (1.f - a) * (1.f - b) * tex_fetch(img, w, h, nch, ch, i, j) + a * (1.f - b) * tex_fetch(img, w, h, nch, ch, i + 1, j) +
           (1.f - a) * b * tex_fetch(img, w, h, nch, ch, i, j + 1) + a * b * tex_fetch(img, w, h, nch, ch, i + 1, j + 1);
}

static unsigned char to_byte(float v)               /* (unsigned char)v of the device code, see orc_cast_f32_u8 */
{
    unsigned u;
    if (!(v > 0.f)) u = 0u;
    else if (v >= 4294967296.f) u = 0xffffffffu;
    else u = (unsigned)v;
    return (unsigned char)(u & 0xffu);
}

/* apply_perspective / apply_perspective_inverse (resample.cu:119-191) */
void orc_perspective_coords(const float* mat9, int inverse, int cols, int rows, float* x_pos, float* y_pos)
{
    float t[9];
    if (inverse) {
        const float* m = mat9;
        const float det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        const float id = 1 / det;
        t[0] = (m[4] * m[8] - m[7] * m[5]) * id; t[1] = (m[2] * m[7] - m[1] * m[8]) * id; t[2] = (m[1] * m[5] - m[2] * m[4]) * id;
        t[3] = (m[5] * m[6] - m[3] * m[8]) * id; t[4] = (m[0] * m[8] - m[2] * m[6]) * id; t[5] = (m[3] * m[2] - m[0] * m[5]) * id;
        t[6] = (m[3] * m[7] - m[6] * m[4]) * id; t[7] = (m[6] * m[1] - m[0] * m[7]) * id; t[8] = (m[0] * m[4] - m[3] * m[1]) * id;
    } else for (int i = 0; i < 9; ++i) t[i] = mat9[i];
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const float xp = t[0] * x + t[1] * y + t[2], yp = t[3] * x + t[4] * y + t[5], sp = t[6] * x + t[7] * y + t[8];
            x_pos[(size_t)y * cols + x] = xp / sp;
            y_pos[(size_t)y * cols + x] = yp / sp;
        }
}

/* resample_2D<uchar4> (resample.cu:83-102) on a BGRA byte frame read as normalised floats */
void orc_resample_bgra(const unsigned char* frame, int fw, int fh, const float* x_pos, const float* y_pos, long long n,
                       unsigned char* result)
{
    float* img = (float*)malloc(sizeof(float) * 4 * (size_t)fw * fh);
    for (size_t i = 0; i < (size_t)4 * fw * fh; ++i) img[i] = frame[i] / 255.f;
    for (long long i = 0; i < n; ++i)
        for (int c = 0; c < 4; ++c)
            result[4 * i + c] = to_byte(tex2d_linear(img, fw, fh, 4, c, x_pos[i] + 0.5f, y_pos[i] + 0.5f) * 255.9999f);
    free(img);
}

/* resample_mask_2D (resample.cu:67-81) */
void orc_resample_mask(const unsigned char* mask, int mw, int mh, const float* x_pos, const float* y_pos, long long n,
                       float threshold, unsigned char* result)
{
    float* img = (float*)malloc(sizeof(float) * (size_t)mw * mh);
    for (size_t i = 0; i < (size_t)mw * mh; ++i) img[i] = mask[i] / 255.f;
    for (long long i = 0; i < n; ++i) {
        const float res = tex2d_linear(img, mw, mh, 1, 0, x_pos[i] + 0.5f, y_pos[i] + 0.5f);
        result[i] = res <= threshold ? 0 : to_byte(res * 255.999f);
    }
    free(img);
}

/* transform_and_blend (resample.cu:7-65), one warp of the frame into the canvas */
void orc_transform_blend(unsigned char* canvas, int cw, int ch, const unsigned char* frame, const unsigned char* mask,
                         const float* wts, int fw, int fh, int nw, int nh, const float* t, int tx, int ty, float* canvas_wts)
{
    float* img = (float*)malloc(sizeof(float) * 4 * (size_t)fw * fh);
    float* msk = (float*)malloc(sizeof(float) * (size_t)fw * fh);
    for (size_t i = 0; i < (size_t)4 * fw * fh; ++i) img[i] = frame[i] / 255.f;
    for (size_t i = 0; i < (size_t)fw * fh; ++i) msk[i] = mask[i] / 255.f;
    for (int y = 0; y < nh; ++y)
        for (int x = 0; x < nw; ++x) {
            const int px = x + tx, py = y + ty;
            if (px < 0 || px >= cw || py < 0 || py >= ch) continue;
            float xp = t[0] * x + t[1] * y + t[2], yp = t[3] * x + t[4] * y + t[5];
            const float sp = t[6] * x + t[7] * y + t[8];
            xp /= sp; yp /= sp;
            if (xp >= fw || yp >= fh) continue;
            if (tex2d_linear(msk, fw, fh, 1, 0, xp + 0.5f, yp + 0.5f) <= 0.5f) continue;
            const float nwt = tex2d_linear(wts, fw, fh, 1, 0, xp + 0.5f, yp + 0.5f);
            const size_t idx = (size_t)py * cw + px;
            float res[3];
            for (int c = 0; c < 3; ++c) res[c] = tex2d_linear(img, fw, fh, 4, c, xp + 0.5f, yp + 0.5f);
            if (canvas_wts[idx] == 0) {
                for (int c = 0; c < 3; ++c) canvas[4 * idx + c] = to_byte(res[c] * 255.9999f);
                canvas[4 * idx + 3] = 255;
                canvas_wts[idx] = nwt;
            } else {
                const float cur = canvas_wts[idx], sum = cur + nwt;
                for (int c = 0; c < 3; ++c)
                    canvas[4 * idx + c] = to_byte((res[c] * nwt * 255.9999f + canvas[4 * idx + c] * cur) / sum);
                canvas[4 * idx + 3] = 255;
                canvas_wts[idx] = cur + nwt;
            }
        }
    free(img); free(msk);
}

/* extract_channel / put_channel / set_alpha_to_const (bgra_2_gray.cu:33-112) */
void orc_extract_channel(const unsigned char* bgra, float* out, long long n, int channel)
{
    if (channel < 0 || channel > 3) return;
    for (long long i = 0; i < n; ++i) out[i] = (float)bgra[4 * i + channel];
}
void orc_put_channel(unsigned char* bgra, const float* in, long long n, int channel)
{
    if (channel < 0 || channel > 3) return;
    for (long long i = 0; i < n; ++i) bgra[4 * i + channel] = channel == 3 ? 255 : to_byte(in[i]);
}
void orc_set_alpha(unsigned char* bgra, long long n, unsigned char val)
{
    for (long long i = 0; i < n; ++i) bgra[4 * i + 3] = val;
}
