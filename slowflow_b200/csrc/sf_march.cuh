// sf_march.cuh -- pieces shared by the marching-form kernels (sf_prep.cu: two-frame warp + derivatives + data term;
// sf_wderivs.cu: multi-frame warp + per-frame derivative planes).  A warp owns a strip of 64 columns (lane l = columns
// 2l, 2l+1, every per-row quantity a packed (column 2l, column 2l+1) pair) and marches down a segment of rows; the
// 5-tap-of-5-tap stencil needs +-4 columns = 2 lanes, so lanes 2..29 (56 columns) produce output and consecutive
// strips overlap by 8.  Border semantics follow image.c:400-526 (see sf_prep.cu).
#pragma once
#include "sf_internal.cuh"
#include "sf_pack.cuh"
#include "sf_stencil.cuh"

namespace sf {

constexpr int PR_OUT_LO = 2, PR_OUT_HI = 29;            // output lanes
constexpr int PR_OUT_W = 2 * (PR_OUT_HI - PR_OUT_LO + 1); // 56 output columns per strip

// per-lane geometry of the strip
struct Lane {
    int x0;       // first column of the pair (may be outside the image in edge strips)
    int xc0, xc1; // clamped columns
    int lane_l, lane_r, comp_r; // lanes holding column 0 / column W-1 (edge strips)
};

// (column 2l, column 2l+1) of one row.  Interior strips: one 8-byte load; edge strips: two clamped 4-byte loads.
template <bool EDGE>
__device__ __forceinline__ p64 load_pair(const float *__restrict__ plane, int rowoff, const Lane &L) {
    if (!EDGE) return __ldg(reinterpret_cast<const p64 *>(plane + rowoff + L.x0));
    return pk(__ldg(plane + rowoff + L.xc0), __ldg(plane + rowoff + L.xc1));
}

// horizontal 5-tap [1,-8,0,8,-1]/12 on a pair: columns (c0-2,c1-2) are the left lane's pair, (c0+2,c1+2) the right lane's
__device__ __forceinline__ p64 hconv_pair(p64 v) {
    const p64 l = shfl_up2(v), r = shfl_down2(v);
    p64 t = mul2(splat2(SF_C0), l);
    t = fma2(splat2(SF_C1), pk(hi_of(l), lo_of(v)), t);
    t = fma2(splat2(SF_C3), pk(hi_of(v), lo_of(r)), t);
    return fma2(splat2(SF_C4), r, t);
}
// vertical 5-tap on pairs (the centre tap has weight -0)
__device__ __forceinline__ p64 vconv_pair(p64 m2, p64 m1, p64 p1, p64 p2) {
    return fma2(splat2(SF_C4), p2, fma2(splat2(SF_C3), p1, fma2(splat2(SF_C1), m1, mul2(splat2(SF_C0), m2))));
}
// second-stage replicate border in x: columns outside the image take the VALUE at column 0 / W-1
__device__ __forceinline__ p64 xedge_fix(p64 v, const Lane &L, int W) {
    const float vl = __shfl_sync(0xffffffffu, lo_of(v), L.lane_l);
    const float r0 = __shfl_sync(0xffffffffu, lo_of(v), L.lane_r), r1 = __shfl_sync(0xffffffffu, hi_of(v), L.lane_r);
    const float vr = L.comp_r ? r1 : r0;
    float a = lo_of(v), b = hi_of(v);
    if (L.x0 < 0) a = vl;
    if (L.x0 + 1 < 0) b = vl;
    if (L.x0 > W - 1) a = vr;
    if (L.x0 + 1 > W - 1) b = vr;
    return pk(a, b);
}
// the four bilinear taps of one warped pixel (variational_aux.c:18-52)
struct Taps {
    int o11, o12, o21, o22;
    float w11, w12, w21, w22;
};
__device__ __forceinline__ Taps warp_taps(const Geom &g, float xx, float yy) {
    const float Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
    const float xf = floorf(xx), yf = floorf(yy);
    const float dx = xx - xf, dy = yy - yf;
    // clamp in float first so that huge flows cannot overflow the int conversion
    const int x = (int)fminf(fmaxf(xf, -2.0f), Wm1 + 2.0f), y = (int)fminf(fmaxf(yf, -2.0f), Hm1 + 2.0f);
    const int x1 = clampi(x, 0, g.W - 1), x2 = clampi(x + 1, 0, g.W - 1);
    const int y1 = clampi(y, 0, g.H - 1) * g.S, y2 = clampi(y + 1, 0, g.H - 1) * g.S;
    Taps t;
    t.o11 = y1 + x1; t.o12 = y1 + x2; t.o21 = y2 + x1; t.o22 = y2 + x2;
    // reference order: s11*(1-dx)*(1-dy) + s12*dx*(1-dy) + s21*(1-dx)*dy + s22*dx*dy
    t.w11 = (1.0f - dx) * (1.0f - dy); t.w12 = dx * (1.0f - dy); t.w21 = (1.0f - dx) * dy; t.w22 = dx * dy;
    return t;
}
__device__ __forceinline__ float warp_fetch(const float *__restrict__ src, const Taps &t) {
    return __ldg(src + t.o11) * t.w11 + __ldg(src + t.o12) * t.w12 + __ldg(src + t.o21) * t.w21 + __ldg(src + t.o22) * t.w22;
}
} // namespace sf
