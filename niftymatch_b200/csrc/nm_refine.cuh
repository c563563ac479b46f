// nm_refine.cuh -- sub-pixel refinement + contrast/edge rejection of one DoG extremum.
//
// Follows subpixel_refinement of the reference (gpu/kernels/keypoint.cu:108-180)
// operation for operation.  Every arithmetic step is written with an explicit rounding
// intrinsic so that the compiler cannot re-associate or contract it; the FMA/non-FMA
// choices reproduce the SASS nvcc 12.9 generates for the reference on sm_100a
// (elimination steps "a -= b*c" are single FFMAs, the determinant is FFMA(fxx,fyy,-fxy^2),
// the value update is a double FMA), so accepted keypoints are bitwise the reference's.
#pragma once
#include "nm_common.cuh"

// F provides  float cur(int dx,int dy), up(int dx,int dy), down(int dx,int dy):
// exact DoG texels around the candidate (the reference samples its linear-filter
// textures at texel centres, gpu/utils/cudatex2D.cu:15-19, which returns exact texels).
template <class F>
__device__ __forceinline__ bool nm_refine(const F& f, int x, int y, float peak, float edge, float xper,
                                          float sigma_0, int num_dogs, int level, float4& out)
{
    const float c = f.cur(0, 0);
    const float cxp = f.cur(1, 0), cxm = f.cur(-1, 0), cyp = f.cur(0, 1), cym = f.cur(0, -1);
    const float u0 = f.up(0, 0), d0 = f.down(0, 0);
    // keypoint.cu:119-121
    const float fx = __fmul_rn(0.5f, __fsub_rn(cxp, cxm));
    const float fy = __fmul_rn(0.5f, __fsub_rn(cyp, cym));
    const float fs = __fmul_rn(0.5f, __fsub_rn(u0, d0));
    // :124-126  float sum widened, minus 2.0*c in double
    const double c2 = __dadd_rn((double)c, (double)c);
    const float fxx = (float)__dsub_rn((double)__fadd_rn(cxp, cxm), c2);
    const float fyy = (float)__dsub_rn((double)__fadd_rn(cyp, cym), c2);
    const float fss = (float)__dsub_rn((double)__fadd_rn(u0, d0), c2);
    // :128-135
    const float fxy = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(f.cur(1, 1), f.cur(-1, -1)), f.cur(-1, 1)), f.cur(1, -1)));
    const float fxs = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(f.up(1, 0), f.down(-1, 0)), f.up(-1, 0)), f.down(1, 0)));
    const float fys = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(f.up(0, 1), f.down(0, -1)), f.up(0, -1)), f.down(0, 1)));

    // :137-139
    float4 A0 = fxx > 0 ? make_float4(fxx, fxy, fxs, -fx) : make_float4(-fxx, -fxy, -fxs, fx);
    float4 A1 = fxy > 0 ? make_float4(fxy, fyy, fys, -fy) : make_float4(-fxy, -fyy, -fys, fy);
    float4 A2 = fxs > 0 ? make_float4(fxs, fys, fss, -fs) : make_float4(-fxs, -fys, -fss, fs);
    float4 tmp;
    const float max_a = fmaxf(fmaxf(A0.x, A1.x), A2.x);          // :142
    if (!((double)max_a >= 1e-10)) return false;                 // :143
    if (max_a == A1.x) { tmp = A1; A1 = A0; A0 = tmp; }
    else if (max_a == A2.x) { tmp = A2; A2 = A0; A0 = tmp; }
    // :150-152
    A0.y = __fdiv_rn(A0.y, A0.x); A0.z = __fdiv_rn(A0.z, A0.x); A0.w = __fdiv_rn(A0.w, A0.x);
    A1.y = __fmaf_rn(-A1.x, A0.y, A1.y); A1.z = __fmaf_rn(-A1.x, A0.z, A1.z); A1.w = __fmaf_rn(-A1.x, A0.w, A1.w);
    A2.y = __fmaf_rn(-A2.x, A0.y, A2.y); A2.z = __fmaf_rn(-A2.x, A0.z, A2.z); A2.w = __fmaf_rn(-A2.x, A0.w, A2.w);
    if (fabsf(A2.y) > fabsf(A1.y)) { tmp = A2; A2 = A1; A1 = tmp; }   // :154
    if (!((double)fabsf(A1.y) >= 1e-10)) return false;           // :158
    A1.z = __fdiv_rn(A1.z, A1.y); A1.w = __fdiv_rn(A1.w, A1.y);  // :159
    A2.z = __fmaf_rn(-A2.y, A1.z, A2.z); A2.w = __fmaf_rn(-A2.y, A1.w, A2.w);   // :160
    if (!((double)fabsf(A2.z) >= 1e-10)) return false;           // :161
    const float ds = __fdiv_rn(A2.w, A2.z);                      // :162
    const float dy = __fmaf_rn(-ds, A1.z, A1.w);                 // :163
    const float dx = __fmaf_rn(-dy, A0.y, __fmaf_rn(-ds, A0.z, A0.w));   // :164
    // :165
    const float inner = __fmaf_rn(fs, ds, __fmaf_rn(fy, dy, __fmul_rn(fx, dx)));
    const float v = (float)__fma_rn((double)inner, 0.5, (double)c);
    // :166
    const float tr = __fadd_rn(fxx, fyy);
    const float s = __fdiv_rn(__fmul_rn(tr, tr), __fmaf_rn(fxx, fyy, -__fmul_rn(fxy, fxy)));
    const float e1 = __fadd_rn(edge, 1.0f);
    const float thr = __fdiv_rn(__fmul_rn(e1, e1), edge);        // :169
    if (!(fabsf(v) > peak && s < thr && fabsf(dx) < 1.0f && fabsf(dy) < 1.0f && fabsf(ds) < 1.0f))
        return false;
    out.x = __fmul_rn(__fadd_rn((float)x, dx), xper);            // :172
    out.y = __fmul_rn(__fadd_rn((float)y, dy), xper);            // :173
    // :174  sigma_0 * pow(2.0, (double)(level + ds)/num_dogs) * xper, all double
    const double e = __ddiv_rn((double)__fadd_rn((float)level, ds), (double)num_dogs);
    out.z = (float)__dmul_rn(__dmul_rn((double)sigma_0, pow(2.0, e)), (double)xper);
    out.w = (float)level;                                        // :175
    return true;
}

// 26-neighbour strict extremum test with the reference's prefilter
// (keypoint.cu:191-196): (c <= 0.8p && c < all) || (c >= 0.8p && c > all).
template <class F>
__device__ __forceinline__ bool nm_is_extremum(const F& f, float peak)
{
    const float c = f.cur(0, 0);
    const float t = __fmul_rn(0.8f, peak);
    float mx = -INFINITY, mn = INFINITY;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const float a = f.down(dx, dy), b = f.up(dx, dy);
            mx = fmaxf(mx, fmaxf(a, b)); mn = fminf(mn, fminf(a, b));
            if (dx != 0 || dy != 0) {
                const float q = f.cur(dx, dy);
                mx = fmaxf(mx, q); mn = fminf(mn, q);
            }
        }
    return (c <= t && c < mn) || (c >= t && c > mx);
}
