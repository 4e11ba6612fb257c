// sf_pack.cuh -- packed-fp32 (f32x2) helpers for sm_100a.
//
// Blackwell issues fma/add/mul on PAIRS of fp32 values held in one aligned 64-bit register (FFMA2 / FADD2 / FMUL2).
// With three distinct register operands the packed form sustains ~1.65x the flops of scalar FFMA on B200
// (tools/ffma_bench.cu: 69.6 vs 42.2 TFLOP/s), so stencil kernels whose work comes in natural pairs (two rows,
// two columns) are written on `p64` values.  A pair should be DEFINED by a 64-bit instruction (LDG.64 / LDS.64 /
// a packed arithmetic op): a pair that is only assembled from two scalars with mov.b64 is re-assembled by ptxas
// in front of every use.
#pragma once
#include <cuda_runtime.h>

namespace sf {

typedef unsigned long long p64; // (lo, hi) = two fp32

__device__ __forceinline__ p64 pk(float lo, float hi) {
    p64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(p64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)b;
    return a;
}
__device__ __forceinline__ float hi_of(p64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    (void)a;
    return b;
}
__device__ __forceinline__ p64 fma2(p64 a, p64 b, p64 c) {
    p64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ p64 mul2(p64 a, p64 b) {
    p64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ p64 add2(p64 a, p64 b) {
    p64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ p64 splat2(float v) { return pk(v, v); }
__device__ __forceinline__ p64 shfl_up2(p64 v) {
    return pk(__shfl_up_sync(0xffffffffu, lo_of(v), 1), __shfl_up_sync(0xffffffffu, hi_of(v), 1));
}
__device__ __forceinline__ p64 shfl_down2(p64 v) {
    return pk(__shfl_down_sync(0xffffffffu, lo_of(v), 1), __shfl_down_sync(0xffffffffu, hi_of(v), 1));
}

// 1/x: MUFU.RCP (1 ulp) refined by one Newton step on the packed pipe (~0.5 ulp; the reference divides, divps)
__device__ __forceinline__ p64 rcp2(p64 x) {
    float a, b;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(lo_of(x)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(hi_of(x)));
    const p64 r = pk(a, b);
    const p64 e = fma2(mul2(x, splat2(-1.0f)), r, splat2(1.0f)); // 1 - x*r
    return fma2(r, e, r);
}
// a - b on the packed pipe
__device__ __forceinline__ p64 sub2(p64 a, p64 b) { return fma2(b, splat2(-1.0f), a); }
// two consecutive floats (8-byte aligned) through the read-only path
__device__ __forceinline__ p64 ldg2(const float *p) { return __ldg(reinterpret_cast<const p64 *>(p)); }

} // namespace sf
