// Micro-benchmark: how fast can one persistent CTA per SM pull haloed tiles of a 9-plane arena through TMA?
// (debug aid for k_sor_tiled; not part of the library)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
// nstage buffers of (planes x BH x BW floats); each stage has its own mbarrier; consumer = whole CTA reads one float per thread
__global__ void __launch_bounds__(512, 1) pull(const __grid_constant__ CUtensorMap tmap, int planes, int BW, int BH, int IW, int IH, int tiles_x,
                                               int ntiles, int nstage, float *sink) {
    extern __shared__ unsigned char raw[];
    unsigned char *base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const int stage_bytes = planes * BW * BH * 4;
    uint64_t *bars = (uint64_t *)(base + nstage * stage_bytes);
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int tile, int s) {
        const int tx = tile % tiles_x, ty = tile / tiles_x;
        const int x0 = tx * IW - (BW - IW) / 2, y0 = ty * IH - (BH - IH) / 2;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(stage_bytes) : "memory");
        for (int p = 0; p < planes; p++)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(base + s * stage_bytes + p * BW * BH * 4)), "l"(&tmap), "r"(smem_u32(bars + s)), "r"(x0), "r"(y0), "r"(p) : "memory");
    };
    int tile = blockIdx.x, k = 0;
    if (threadIdx.x == 0)
        for (int s = 0; s < nstage; s++) { int t = tile + s * gridDim.x; if (t < ntiles) issue(t, s); }
    float acc = 0.f;
    for (; tile < ntiles; tile += gridDim.x, k++) {
        const int s = k % nstage;
        const uint32_t parity = (k / nstage) & 1;
        while (!try_wait(bars + s, parity)) {}
        acc += ((float *)(base + s * stage_bytes))[threadIdx.x];
        __syncthreads();
        const int nxt = tile + nstage * gridDim.x;
        if (threadIdx.x == 0 && nxt < ntiles) issue(nxt, s);
    }
    if (acc == 12345.678f) sink[0] = acc;
}
int main(int argc, char **argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 2560, H = argc > 2 ? atoi(argv[2]) : 1440, S = W, NP = 11;
    size_t P = (size_t)S * H;
    float *d, *sink;
    cudaMalloc(&d, P * NP * 4); cudaMemset(d, 0, P * NP * 4); cudaMalloc(&sink, 4);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)p;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int BW, BH, halo_x, halo_y, nstage, planes, promo; };
    Cfg cfgs[] = {{64, 64, 8, 8, 1, 9, 1}, {64, 64, 8, 8, 1, 9, 2}, {64, 32, 8, 8, 2, 9, 1}, {64, 32, 8, 8, 3, 9, 1}, {128, 32, 8, 8, 1, 9, 1},
                  {128, 16, 8, 4, 2, 9, 1}, {256, 16, 8, 4, 1, 9, 1}, {64, 16, 8, 4, 5, 9, 1}, {64, 64, 0, 0, 1, 9, 1}, {64, 64, 8, 8, 1, 4, 1}, {64, 64, 8, 8, 1, 4, 1}};
    for (auto &c : cfgs) {
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, NP}; cuuint64_t strides[2] = {(cuuint64_t)S * 4, (cuuint64_t)P * 4};
        cuuint32_t box[3] = {(cuuint32_t)c.BW, (cuuint32_t)c.BH, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         c.promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (c.promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode fail %d\n", r); continue; }
        const int IW = c.BW - 2 * c.halo_x, IH = c.BH - 2 * c.halo_y;
        const int tiles_x = (W + IW - 1) / IW, tiles_y = (H + IH - 1) / IH, ntiles = tiles_x * tiles_y;
        const int smem = c.nstage * c.planes * c.BW * c.BH * 4 + 64 + 1024;
        cudaFuncSetAttribute(pull, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            pull<<<148, 512, smem>>>(tm, c.planes, c.BW, c.BH, IW, IH, tiles_x, ntiles, c.nstage, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double bytes = (double)ntiles * c.planes * c.BW * c.BH * 4;
        printf("box %3dx%2d halo %d,%d stages %d planes %d promo %d: %5d tiles, %.1f us, tile bytes moved %.0f MB -> %.2f TB/s (%s) useful px rate %.1f Gpx/s\n", c.BW, c.BH, c.halo_x, c.halo_y,
               c.nstage, c.planes, c.promo, ntiles, best * 1e3, bytes / 1e6, bytes / best / 1e9, cudaGetErrorString(cudaGetLastError()), (double)W * H / best / 1e6);
    }
    return 0;
}
