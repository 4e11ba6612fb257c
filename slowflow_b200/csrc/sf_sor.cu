// sf_sor.cu -- K4: red-black SOR for the coupled 2x2-block 5-point system of sor_coupled (solver.c:63-399).
//
// Per pixel and sweep (solver.c:132-138, with the pre-inverted blocks of :101-106):
//     B1 = psi_l*du_l + (psi_r*du_r + psi_t*du_t + psi_b*du_b + b1)      (same for B2 with dv)
//     du += omega*(a11'*B1 + a12'*B2 - du);   dv += omega*(a12'*B1 + a22'*B2 - dv)
// The reference visits pixels row-major (Gauss-Seidel, serial in i).  Here colour (i+j)&1 == 0 is
// updated first, then colour 1, per sweep -- the ordering of the oracle's CPU red-black mode.
//
// Variant 0 (production): temporally blocked.  A persistent CTA (16 warps) owns a 64 x 64 pixel tile.
//   * TMA (cp.async.bulk.tensor, 3-D map over the 11-plane SOR arena) stages the 9 input planes of the
//     NEXT tile in shared memory while the current tile is being relaxed; out-of-image texels are
//     zero-filled by the TMA unit, which gives psi = 0 / a' = 0 boundary handling for free.
//   * The 7 coefficient planes and du,dv of the current tile live in REGISTERS: lane l of warp w owns
//     pixel columns {2l, 2l+1} of rows 4w..4w+3.  Horizontal neighbours come from warp shuffles,
//     vertical neighbours across warps through a 16 KB double-buffered exchange area.
//   * T sweeps (2T half sweeps) run per HBM round trip; the outer 2T pixels of the tile are halo that
//     is recomputed by the neighbouring tile.  HBM traffic per sweep drops from 44 B/px to
//     (36/f + 8)/T B/px with f = the interior fraction of the tile (e.g. 56x56/64x64 for T = 2).
//   * du,dv ping-pong between two arena buffers (a tile's halo is another tile's interior).
// Variant 1 (validation, tiny images): one launch per half sweep straight from global memory.
#include "sf_internal.cuh"

namespace sf {

// ------------------------------------------------------------------------------------------ variant 1
template <int COLOUR>
__global__ void __launch_bounds__(256) k_sor_half_global(Geom g, const float *__restrict__ a11, const float *__restrict__ a12,
                                                         const float *__restrict__ a22, const float *__restrict__ b1,
                                                         const float *__restrict__ b2, const float *__restrict__ ph,
                                                         const float *__restrict__ pv, float *du, float *dv, float omega) {
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= g.H) return;
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + ((j + COLOUR) & 1);
    if (i >= g.W) return;
    const size_t o = (size_t)j * g.S + i;
    const float psr = ph[o], psb = pv[o];
    const float psl = (i > 0) ? ph[o - 1] : 0.0f, pst = (j > 0) ? pv[o - g.S] : 0.0f;
    const size_t orr = (i < g.W - 1) ? o + 1 : o, ol = (i > 0) ? o - 1 : o;
    const size_t ot = (j > 0) ? o - g.S : o, ob = (j < g.H - 1) ? o + g.S : o;
    const float s1 = psr * du[orr] + pst * du[ot] + psb * du[ob] + b1[o];
    const float s2 = psr * dv[orr] + pst * dv[ot] + psb * dv[ob] + b2[o];
    const float B1 = psl * du[ol] + s1, B2 = psl * dv[ol] + s2;
    const float u = du[o], v = dv[o];
    du[o] = u + omega * (a11[o] * B1 + a12[o] * B2 - u);
    dv[o] = v + omega * (a12[o] * B1 + a22[o] * B2 - v);
}

// ------------------------------------------------------------------------------------------ variant 0
constexpr int SOR_R = 4;                    // rows per warp (must be even: row parity == r parity)
constexpr int SOR_NW = 16;                  // warps per CTA
constexpr int SOR_TW = 64;                  // tile width: 2 pixels per lane
constexpr int SOR_TH = SOR_R * SOR_NW;      // tile height
constexpr int SOR_STAGED = 9;               // a11' a12' a22' b1 b2 psi_h psi_v du dv
constexpr int SOR_PLANE_FLOATS = SOR_TW * SOR_TH;
constexpr int SOR_STAGE_BYTES = SOR_STAGED * SOR_PLANE_FLOATS * 4;
constexpr int SOR_EXCH_BYTES = 2 /*buffers*/ * 2 /*top,bottom*/ * SOR_NW * 32 * 8;
constexpr int SOR_SMEM_BYTES = SOR_STAGE_BYTES + SOR_EXCH_BYTES + 64 + 1024 /*alignment slack*/;

struct SorTiledArgs {
    Geom g;
    float *out_du, *out_dv; // destination planes (interior of every tile)
    int in_du_plane, in_dv_plane;
    int T;          // sweeps fused in this launch
    int tiles_x, tiles_y;
    float omega;
    int zero_init;  // initial iterate is 0: du,dv are not loaded
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a TMA that never completes (bad descriptor) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

struct SorRegs {
    float a11[SOR_R][2], a12[SOR_R][2], a22[SOR_R][2], b1[SOR_R][2], b2[SOR_R][2];
    float phl[SOR_R], phm[SOR_R], phr[SOR_R]; // psi_h at columns 2l-1, 2l, 2l+1
    float pv[SOR_R][2], pvt[2];               // psi_v of the rows, and of the row above the strip
    float du[SOR_R][2], dv[SOR_R][2];
};

template <int C>
__device__ __forceinline__ void sor_half_sweep(SorRegs &q, const float2 up, const float2 dn, const float omega) {
    constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
    for (int r = 0; r < SOR_R; r++) {
        const int e = (C + r) & 1; // column of the pair that has colour C in this row (tile origin is even/even)
        float ul, vl, ur, vr, psl, psr;
        if (e == 0) {
            ul = __shfl_up_sync(FULL, q.du[r][1], 1);
            vl = __shfl_up_sync(FULL, q.dv[r][1], 1);
            ur = q.du[r][1];
            vr = q.dv[r][1];
            psl = q.phl[r];
            psr = q.phm[r];
        } else {
            ul = q.du[r][0];
            vl = q.dv[r][0];
            ur = __shfl_down_sync(FULL, q.du[r][0], 1);
            vr = __shfl_down_sync(FULL, q.dv[r][0], 1);
            psl = q.phm[r];
            psr = q.phr[r];
        }
        float ut, vt, pst, ub, vb;
        if (r == 0) { ut = up.x; vt = up.y; pst = q.pvt[e]; }
        else { ut = q.du[r - 1][e]; vt = q.dv[r - 1][e]; pst = q.pv[r - 1][e]; }
        if (r == SOR_R - 1) { ub = dn.x; vb = dn.y; }
        else { ub = q.du[r + 1][e]; vb = q.dv[r + 1][e]; }
        const float psb = q.pv[r][e];
        // B = b + sum_nb psi_nb*d_nb as one FMA chain (bottom, top, right, left); the reference's
        // left-to-right sum differs from this only in rounding (tests gate at 2e-4 after 30 sweeps)
        const float B1 = fmaf(psl, ul, fmaf(psr, ur, fmaf(pst, ut, fmaf(psb, ub, q.b1[r][e]))));
        const float B2 = fmaf(psl, vl, fmaf(psr, vr, fmaf(pst, vt, fmaf(psb, vb, q.b2[r][e]))));
        const float ru = fmaf(q.a11[r][e], B1, fmaf(q.a12[r][e], B2, -q.du[r][e]));
        const float rv = fmaf(q.a12[r][e], B1, fmaf(q.a22[r][e], B2, -q.dv[r][e]));
        q.du[r][e] = fmaf(omega, ru, q.du[r][e]);
        q.dv[r][e] = fmaf(omega, rv, q.dv[r][e]);
    }
}

__global__ void __launch_bounds__(SOR_NW * 32, 1) k_sor_tiled(const __grid_constant__ CUtensorMap tmap, SorTiledArgs a) {
    extern __shared__ unsigned char smem_raw[];
    // TMA destinations need 128-byte alignment; round the dynamic base up to 1 KB
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float *stage = reinterpret_cast<float *>(base);
    float2 *exch = reinterpret_cast<float2 *>(base + SOR_STAGE_BYTES);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(base + SOR_STAGE_BYTES + SOR_EXCH_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // halo: 2T pixels per side.  TMA needs the box origin 16-byte aligned along x, so the horizontal
    // halo is rounded up to a multiple of 4 floats (tile origins stay multiples of 4).
    const int hy = 2 * a.T, hx = (2 * a.T + 3) & ~3;
    const int IW = SOR_TW - 2 * hx, IH = SOR_TH - 2 * hy;
    const int ntiles = a.tiles_x * a.tiles_y;
    const int nload = a.zero_init ? 7 : 9;

    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int tile) {
        const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
        const int x0 = tx * IW - hx, y0 = ty * IH - hy;
        mbar_expect_tx(mbar, (uint32_t)(nload * SOR_PLANE_FLOATS * 4));
#pragma unroll 1
        for (int pl = 0; pl < nload; pl++) {
            const int z = (pl < 7) ? pl : (pl == 7 ? a.in_du_plane : a.in_dv_plane);
            tma_load_3d(stage + pl * SOR_PLANE_FLOATS, &tmap, mbar, x0, y0, z);
        }
    };

    int tile = blockIdx.x;
    uint32_t phase = 0;
    if (threadIdx.x == 0 && tile < ntiles) issue(tile);

    // exchange area: exch[((buf*2 + side)*SOR_NW + warp)*32 + lane], side 0 = top row, 1 = bottom row
    auto ex = [&](int buf, int side, int w) -> float2 * { return exch + ((buf * 2 + side) * SOR_NW + w) * 32 + lane; };

    while (tile < ntiles) {
        const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
        const int x0 = tx * IW - hx, y0 = ty * IH - hy;

        mbar_wait(mbar, phase);
        phase ^= 1u;

        // ---- shared -> registers
        SorRegs q;
#pragma unroll
        for (int r = 0; r < SOR_R; r++) {
            const int tr = warp * SOR_R + r;
            const float2 *row = reinterpret_cast<const float2 *>(stage + tr * SOR_TW) + lane;
            float2 v;
            v = row[(SP_A11 * SOR_PLANE_FLOATS) / 2]; q.a11[r][0] = v.x; q.a11[r][1] = v.y;
            v = row[(SP_A12 * SOR_PLANE_FLOATS) / 2]; q.a12[r][0] = v.x; q.a12[r][1] = v.y;
            v = row[(SP_A22 * SOR_PLANE_FLOATS) / 2]; q.a22[r][0] = v.x; q.a22[r][1] = v.y;
            v = row[(SP_B1 * SOR_PLANE_FLOATS) / 2];  q.b1[r][0] = v.x;  q.b1[r][1] = v.y;
            v = row[(SP_B2 * SOR_PLANE_FLOATS) / 2];  q.b2[r][0] = v.x;  q.b2[r][1] = v.y;
            v = row[(SP_PH * SOR_PLANE_FLOATS) / 2];  q.phm[r] = v.x;    q.phr[r] = v.y;
            const float left = __shfl_up_sync(0xffffffffu, v.y, 1);
            q.phl[r] = (lane == 0) ? 0.0f : left; // column -1 of the tile is never needed for a valid pixel
            v = row[(SP_PV * SOR_PLANE_FLOATS) / 2];  q.pv[r][0] = v.x;  q.pv[r][1] = v.y;
            if (a.zero_init) {
                q.du[r][0] = q.du[r][1] = q.dv[r][0] = q.dv[r][1] = 0.0f;
            } else {
                v = row[(7 * SOR_PLANE_FLOATS) / 2]; q.du[r][0] = v.x; q.du[r][1] = v.y;
                v = row[(8 * SOR_PLANE_FLOATS) / 2]; q.dv[r][0] = v.x; q.dv[r][1] = v.y;
            }
        }
        if (warp > 0) {
            const float2 v = (reinterpret_cast<const float2 *>(stage + SP_PV * SOR_PLANE_FLOATS + (warp * SOR_R - 1) * SOR_TW))[lane];
            q.pvt[0] = v.x; q.pvt[1] = v.y;
        } else {
            q.pvt[0] = q.pvt[1] = 0.0f; // row -1 of the tile: halo or outside the image
        }
        // initial publication "as if colour 1 had just been relaxed": top row column 1, bottom row column 0
        *ex(1, 0, warp) = make_float2(q.du[0][1], q.dv[0][1]);
        *ex(1, 1, warp) = make_float2(q.du[SOR_R - 1][0], q.dv[SOR_R - 1][0]);
        __syncthreads(); // staging buffer fully consumed + exchange visible

        const int next = tile + gridDim.x;
        if (threadIdx.x == 0 && next < ntiles) issue(next); // overlaps with the relaxation below

        // ---- 2T half sweeps in registers
        const float2 zero2 = make_float2(0.0f, 0.0f);
        // Half sweep k (0-based) only has to be right for pixels at depth >= k+1 from the tile edge (depth
        // d becomes garbage-tolerant once k >= d).  A warp whose innermost row has depth <= k therefore
        // idles in half sweep k (it still joins the barrier):
        const int wdepth = (warp < SOR_NW / 2) ? (warp * SOR_R + SOR_R - 1) : ((SOR_NW - 1 - warp) * SOR_R + SOR_R - 1);
#pragma unroll 1
        for (int t = 0; t < a.T; t++) {
            if (2 * t < wdepth) {
                const float2 up = (warp > 0) ? *ex(1, 1, warp - 1) : zero2;
                const float2 dn = (warp < SOR_NW - 1) ? *ex(1, 0, warp + 1) : zero2;
                sor_half_sweep<0>(q, up, dn, a.omega);
                *ex(0, 0, warp) = make_float2(q.du[0][0], q.dv[0][0]);
                *ex(0, 1, warp) = make_float2(q.du[SOR_R - 1][1], q.dv[SOR_R - 1][1]);
            }
            __syncthreads();
            if (2 * t + 1 < wdepth) {
                const float2 up = (warp > 0) ? *ex(0, 1, warp - 1) : zero2;
                const float2 dn = (warp < SOR_NW - 1) ? *ex(0, 0, warp + 1) : zero2;
                sor_half_sweep<1>(q, up, dn, a.omega);
                *ex(1, 0, warp) = make_float2(q.du[0][1], q.dv[0][1]);
                *ex(1, 1, warp) = make_float2(q.du[SOR_R - 1][0], q.dv[SOR_R - 1][0]);
            }
            __syncthreads();
        }

        // ---- interior of the tile -> global (float2 per lane and row: 256 B per warp row)
        const int cx = 2 * lane, gx = x0 + cx;
        if (cx >= hx && cx < SOR_TW - hx && gx < a.g.W) {
#pragma unroll
            for (int r = 0; r < SOR_R; r++) {
                const int tr = warp * SOR_R + r, gy = y0 + tr;
                if (tr >= hy && tr < SOR_TH - hy && gy < a.g.H) {
                    const size_t o = (size_t)gy * a.g.S + gx;
                    *reinterpret_cast<float2 *>(a.out_du + o) = make_float2(q.du[r][0], q.du[r][1]);
                    *reinterpret_cast<float2 *>(a.out_dv + o) = make_float2(q.dv[r][0], q.dv[r][1]);
                }
            }
        }
        tile = next;
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

bool sor_plan_init(SorPlan &plan, Geom g, float *arena, int num_sms) {
    plan.g = g;
    plan.arena = arena;
    plan.num_sms = num_sms;
    plan.tmap_valid = false;
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
        return false;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)SP_COUNT};
    const cuuint64_t strides[2] = {(cuuint64_t)g.S * 4, (cuuint64_t)g.plane() * 4};
    const cuuint32_t box[3] = {SOR_TW, SOR_TH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&plan.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, arena, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d) for %dx%d stride %d", (int)r, g.W, g.H, g.S);
        set_error(buf);
        return false;
    }
    plan.tmap_valid = true;
    return true;
}

int launch_sor(cudaStream_t st, SorPlan &plan, int iterations, float omega, int variant, int fuse, int *cur,
               bool zero_init) {
    const Geom g = plan.g;
    float *A = plan.arena;
    const size_t P = g.plane();
    int launches = 0;
    if (iterations <= 0) return 0;
    if (variant == 1 || !plan.tmap_valid) {
        float *du = A + (size_t)(*cur ? SP_DUB : SP_DUA) * P, *dv = A + (size_t)(*cur ? SP_DVB : SP_DVA) * P;
        if (zero_init) {
            cudaMemsetAsync(du, 0, P * sizeof(float), st);
            cudaMemsetAsync(dv, 0, P * sizeof(float), st);
        }
        dim3 b(32, 8), grid(((g.W + 1) / 2 + 31) / 32, (g.H + 7) / 8);
        for (int it = 0; it < iterations; it++) {
            k_sor_half_global<0><<<grid, b, 0, st>>>(g, A + SP_A11 * P, A + SP_A12 * P, A + SP_A22 * P, A + SP_B1 * P,
                                                     A + SP_B2 * P, A + SP_PH * P, A + SP_PV * P, du, dv, omega);
            k_sor_half_global<1><<<grid, b, 0, st>>>(g, A + SP_A11 * P, A + SP_A12 * P, A + SP_A22 * P, A + SP_B1 * P,
                                                     A + SP_B2 * P, A + SP_PH * P, A + SP_PV * P, du, dv, omega);
            launches += 2;
        }
        return launches;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_sor_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, SOR_SMEM_BYTES) != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_sor_tiled) failed");
            return -1;
        }
        attr_set = true;
    }
    if (fuse < 1) fuse = 1;
    if (fuse > 7) fuse = 7; // halo 2*fuse per side must leave an interior: 64 - 4*fuse >= 36
    int done = 0;
    while (done < iterations) {
        const int T = (iterations - done < fuse) ? (iterations - done) : fuse;
        SorTiledArgs a;
        a.g = g;
        a.T = T;
        const int IW = SOR_TW - 2 * ((2 * T + 3) & ~3), IH = SOR_TH - 4 * T;
        a.tiles_x = (g.W + IW - 1) / IW;
        a.tiles_y = (g.H + IH - 1) / IH;
        a.omega = omega;
        a.zero_init = (zero_init && done == 0) ? 1 : 0;
        a.in_du_plane = *cur ? SP_DUB : SP_DUA;
        a.in_dv_plane = *cur ? SP_DVB : SP_DVA;
        a.out_du = A + (size_t)(*cur ? SP_DUA : SP_DUB) * P;
        a.out_dv = A + (size_t)(*cur ? SP_DVA : SP_DVB) * P;
        const int ntiles = a.tiles_x * a.tiles_y;
        const int grid = ntiles < plan.num_sms ? ntiles : plan.num_sms;
        k_sor_tiled<<<grid, SOR_NW * 32, SOR_SMEM_BYTES, st>>>(plan.tmap, a);
        *cur ^= 1;
        done += T;
        launches++;
    }
    return launches;
}

} // namespace sf
