// Stand-alone probe of the TMA tile load used by k_sor_tiled (debug aid, not part of the library).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float *out, int *status, int x0, int y0, int z, int boxw, int boxh) {
    extern __shared__ unsigned char raw[];
    unsigned char *base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    float *stage = (float *)base;
    uint64_t *bar = (uint64_t *)(base + boxw * boxh * 4);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(boxw * boxh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(stage)), "l"(&tmap), "r"(smem_u32(bar)), "r"(x0), "r"(y0), "r"(z) : "memory");
    }
    uint32_t ok = 0, spins = 0;
    while (!ok && spins < (1u << 20)) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
        spins++;
    }
    if (threadIdx.x == 0) { status[0] = ok; status[1] = spins; }
    __syncthreads();
    for (int i = threadIdx.x; i < boxw * boxh; i += blockDim.x) out[i] = ok ? stage[i] : -1.0f;
}
int main() {
    const int W = 64, H = 64, S = 64, NP = 11, BW = 64, BH = 64;
    size_t P = (size_t)S * H;
    std::vector<float> h(P * NP);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100000) * 0.001f + 1.0f;
    float *d, *out; int *status;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, BW * BH * 4); cudaMalloc(&status, 8);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    PFN_encodeTiled enc = (PFN_encodeTiled)p;
    CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, NP}; cuuint64_t strides[2] = {(cuuint64_t)S * 4, (cuuint64_t)P * 4};
    cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    int smem = BW * BH * 4 + 64 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int trial = 0; trial < 3; trial++) {
        int x0 = trial == 0 ? 0 : -2, y0 = trial == 0 ? 0 : -2, z = trial == 2 ? 7 : 0;
        probe<<<1, 128, smem>>>(tm, out, status, x0, y0, z, BW, BH);
        e = cudaDeviceSynchronize();
        int st[2]; std::vector<float> o(BW * BH);
        cudaMemcpy(st, status, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), out, BW * BH * 4, cudaMemcpyDeviceToHost);
        printf("trial %d (x0=%d,y0=%d,z=%d): sync=%s ok=%d spins=%d  out[0]=%g out[2*64+2]=%g expect h[z*P]=%g out[65]=%g\n", trial, x0, y0, z,
               cudaGetErrorString(e), st[0], st[1], o[0], o[2 * 64 + 2], h[z * P], o[65]);
    }
    return 0;
}
