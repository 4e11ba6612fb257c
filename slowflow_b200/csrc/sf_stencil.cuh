// sf_stencil.cuh -- the reference's derivative filters as device functions (image.c:351-526).
#pragma once
#include "sf_internal.cuh"

namespace sf {

// 5-tap derivative filter [1,-8,0,8,-1]/12 (variational.c:118-119, image.c:351-373)
#define SF_C0 (1.0f / 12.0f)
#define SF_C1 (-8.0f / 12.0f)
#define SF_C2 (-0.0f)
#define SF_C3 (8.0f / 12.0f)
#define SF_C4 (-(1.0f / 12.0f))

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// horizontal 5-tap on already clamped samples (image.c:516)
__device__ __forceinline__ float hconv5(float m2, float m1, float s0, float p1, float p2) {
    return SF_C0 * m2 + SF_C1 * m1 + SF_C2 * s0 + SF_C3 * p1 + SF_C4 * p2;
}
// vertical 5-tap with the border folding of image.c:425-458 (row j of H)
__device__ __forceinline__ float vconv5(float m2, float m1, float s0, float p1, float p2, int j, int H) {
    if (j >= 2 && j < H - 2) return SF_C0 * m2 + SF_C1 * m1 + SF_C2 * s0 + SF_C3 * p1 + SF_C4 * p2;
    if (j == 0) return (SF_C0 + SF_C1 + SF_C2) * s0 + SF_C3 * p1 + SF_C4 * p2;
    if (j == 1) return (SF_C0 + SF_C1) * m1 + SF_C2 * s0 + SF_C3 * p1 + SF_C4 * p2;
    if (j == H - 2) return SF_C0 * m2 + SF_C1 * m1 + SF_C2 * s0 + (SF_C3 + SF_C4) * p1;
    if (j == H - 1) return SF_C0 * m2 + SF_C1 * m1 + (SF_C2 + SF_C3 + SF_C4) * s0;
    return SF_C0 * m2 + SF_C1 * m1 + SF_C2 * s0 + SF_C3 * p1 + SF_C4 * p2; // rows outside the image: unused
}


} // namespace sf
