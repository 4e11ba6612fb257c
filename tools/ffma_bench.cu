// Micro-benchmark: FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a (debug aid, not part of the library).
// mode A: d = d*s + t with s,t shared by all chains (operand-reuse friendly);  mode B: d[i] = a[i]*b[i] + d[i] with all three
// operands distinct registers per chain (what a stencil with per-pixel coefficients does).
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long p64;
__device__ __forceinline__ p64 mk(float a, float b) { return ((p64)__float_as_uint(a) << 32) | __float_as_uint(b); }
template <int DISTINCT> __global__ void __launch_bounds__(256) k_ffma(float *out, int iters, float s, float t) {
    float d[8], a[8], b[8];
    for (int i = 0; i < 8; i++) { d[i] = threadIdx.x * 0.001f + i; a[i] = s + i * 1e-6f; b[i] = t + i * 1e-6f; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (DISTINCT) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d[i]) : "f"(a[i]), "f"(b[(i + u) & 7]));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(d[i]) : "f"(s), "f"(t));
            }
    }
    float r = 0; for (int i = 0; i < 8; i++) r += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int DISTINCT> __global__ void __launch_bounds__(256) k_ffma2(float *out, int iters, float s, float t) {
    p64 d[8], a[8], b[8];
    for (int i = 0; i < 8; i++) { d[i] = mk(threadIdx.x * 0.001f + i, i + 0.5f); a[i] = mk(s + i * 1e-6f, s - i * 1e-6f); b[i] = mk(t + i * 1e-6f, t - i * 1e-6f); }
    const p64 s2 = mk(s, s), t2 = mk(t, t);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (DISTINCT) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d[i]) : "l"(a[i]), "l"(b[(i + u) & 7]));
                else asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(s2), "l"(t2));
            }
    }
    float r = 0; for (int i = 0; i < 8; i++) r += __uint_as_float((unsigned)d[i]) + __uint_as_float((unsigned)(d[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 10000;
    const char *names[4] = {"FFMA  shared-ops ", "FFMA2 shared-ops ", "FFMA  distinct-ops", "FFMA2 distinct-ops"};
    for (int warps_per_sm = 4; warps_per_sm <= 32; warps_per_sm *= 2) {
        const int blocks = 148 * warps_per_sm / 8;
        for (int mode = 0; mode < 4; mode++) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k_ffma<0><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
                else if (mode == 1) k_ffma2<0><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
                else if (mode == 2) k_ffma<1><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
                else k_ffma2<1><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double inst = (double)blocks * 8 * iters * 64; // warp instructions
            const double per_sm_clk = inst / 148 / (best * 1e-3 * 1.965e9);
            printf("%s warps/SM %2d: %.3f ms, %.2f warp-inst/clk/SM, %.1f TFLOP/s\n", names[mode], warps_per_sm, best, per_sm_clk,
                   inst * 32 * 2 * ((mode & 1) ? 2 : 1) / (best * 1e-3) / 1e12);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
