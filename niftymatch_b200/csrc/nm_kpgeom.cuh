// nm_kpgeom.cuh -- keypoint geometry shared by the orientation / descriptor kernels and by the
// gradient-need marking of emit_kernel: the SAME arithmetic decides which gradient samples a keypoint
// reads (nm_orient_desc.cu) and which 8 x 32 blocks of the gradient maps get computed (nm_extrema.cu).
#pragma once
#include "nm_common.cuh"

struct KpGeom {
    float x, y, s;
    int xi, yi, level;
};

// orientation.cu:19-24 / descriptor.cu:41-47: octave coordinates of a keypoint and its nearest pixel
__device__ __forceinline__ KpGeom kp_geom(const float4 kp, float xper)
{
    KpGeom g;
    g.x = __fdiv_rn(kp.x, xper);
    g.y = __fdiv_rn(kp.y, xper);
    g.s = __fdiv_rn(kp.z, xper);
    g.xi = __double2int_rz(__dadd_rn((double)g.x, 0.5));
    g.yi = __double2int_rz(__dadd_rn((double)g.y, 0.5));
    g.level = (int)kp.w;
    return g;
}

// Orientation window radius (orientation.cu:26-30; the 22 x 22 block of the reference clamps it to 10).
__device__ __forceinline__ int kp_orient_radius(const KpGeom& g, float& sigma_w)
{
    sigma_w = __fmul_rn(1.5f, g.s);                               // gauss_factor = 1.5 (siftfunctions.cu:150)
    const int W = max((int)floorf(__fmul_rn(3.0f, sigma_w)), 1);
    return min(W, 10);
}

// Descriptor window (descriptor.cu:54-65): SBP, half width W, the clipped offsets and the number of
// 16 x 16 chunks walked along the diagonal (:94-97, :142-143).
struct KpDescWindow {
    float SBP;
    int xmin, xmax, ymin, ymax, chunks;
};
__device__ __forceinline__ KpDescWindow kp_desc_window(const KpGeom& g, int ow, int oh)
{
    KpDescWindow d;
    d.SBP = (float)__dadd_rn((double)__fmul_rn(3.0f, g.s), 1.e-07);                                       // :54
    const int W = (int)floor(__fma_rn(__dmul_rn(__dmul_rn((double)d.SBP, 1.4142135623730951), 5.0), 0.5, 0.5));   // :55
    d.xmin = max(-W, -g.xi); d.xmax = min(W, ow - 1 - g.xi);                                             // :57-60
    d.ymin = max(-W, -g.yi); d.ymax = min(W, oh - 1 - g.yi);
    const int max_dims = max(d.xmax - d.xmin, d.ymax - d.ymin);
    d.chunks = (int)ceilf(__fdiv_rn(__fadd_rn((float)max_dims, 1.f), 16.f));                              // :65
    return d;
}

// Gradient-need map: one byte per 8-row x 32-column block of a gradient map.
#define NM_NEED_ROWS 8
