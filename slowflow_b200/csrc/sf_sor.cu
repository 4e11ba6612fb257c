// sf_sor.cu -- K4: red-black SOR for the coupled 2x2-block 5-point system of sor_coupled (solver.c:63-399).
//
// Per pixel and sweep (solver.c:132-138, with the pre-inverted blocks of :101-106):
//     B1 = psi_l*du_l + (psi_r*du_r + psi_t*du_t + psi_b*du_b + b1)      (same for B2 with dv)
//     du += omega*(a11'*B1 + a12'*B2 - du);   dv += omega*(a12'*B1 + a22'*B2 - dv)
// The reference visits pixels row-major (Gauss-Seidel, serial in i).  Here colour (i+j)&1 == 0 is
// updated first, then colour 1, per sweep -- the ordering of the oracle's CPU red-black mode.
//
// Variant 0 (production): temporally blocked, ONE launch per sor_coupled call.  A persistent CTA (8 compute warps + a
// helper warp group) owns one 64 x 64 pixel tile at a time.
//   * T sweeps (2T half sweeps) run per pass over the image; the outer 2T pixels of a tile are halo that is recomputed
//     by the neighbouring tile.  HBM traffic per sweep drops from 44 B/px to (36/f + 8)/T B/px with f = the interior
//     fraction of the tile (48x48/64x64 for T = 4).  du,dv ping-pong between two arena buffers from pass to pass.
//   * All passes of the call run inside the launch: tiles are tickets (pass-major) from an atomic counter and tile (tx,ty)
//     of pass p waits for the 3x3 tile neighbourhood of pass p-1 through per-tile flags -- see k_sor_tiled.
//   * TMA (cp.async.bulk.tensor, 3-D maps over the 11-plane SOR arena) stages the 9 input planes of the coming tiles in
//     shared memory while the current tile is being relaxed; out-of-image texels are zero-filled by the TMA unit, which
//     gives psi = 0 / a' = 0 boundary handling for free.
//   * The 7 coefficient planes and du,dv of the current tile live in REGISTERS: lane l of warp w owns pixel columns
//     {2l, 2l+1} of rows 8w..8w+7.  Horizontal neighbours come from warp shuffles, vertical neighbours across warps
//     through a double-buffered exchange area in shared memory.
// Variant 1 (validation, tiny images): one launch per half sweep straight from global memory.
#include "sf_internal.cuh"
#include "sf_pack.cuh"
#include "sf_tma.cuh"

#ifndef SF_SOR_NW
#define SF_SOR_NW 8
#endif

namespace sf {

// ------------------------------------------------------------------------------------------ variant 1
template <int COLOUR>
__global__ void __launch_bounds__(256) k_sor_half_global(Geom g, const float *__restrict__ a11, const float *__restrict__ a12,
                                                         const float *__restrict__ a22, const float *__restrict__ b1,
                                                         const float *__restrict__ b2, const float *__restrict__ ph,
                                                         const float *__restrict__ pv, float *du, float *dv, float omega) {
    if (g.cancelled()) return;
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
#ifndef SF_SOR_R
#define SF_SOR_R 8
#endif
#ifndef SF_SOR_SKIP
#define SF_SOR_SKIP 0
#endif
#ifndef SF_SOR_PDL
#define SF_SOR_PDL 1 // chain consecutive launches with programmatic dependent launch
#endif
#ifndef SF_SOR_EDGE_TILES
#define SF_SOR_EDGE_TILES 1 // 1: tiles whose side lies on the image border keep that side's halo as interior (fewer tiles)
#endif
#define SOR_EDGE (SF_SOR_EDGE_TILES ? 0 : 1)
#ifndef SF_SOR_A_AFTER_B
#define SF_SOR_A_AFTER_B 1 // 1: group A of the round after next is requested when group B of the next round has landed
#endif
#ifndef SF_SOR_MULTIPASS
#define SF_SOR_MULTIPASS 1 // 1: all passes of `fuse` sweeps of a call in one launch; 0: one launch per pass (A/B timing)
#endif
#ifndef SF_SOR_SYNC
#define SF_SOR_SYNC 0 // 0: one CTA barrier per half sweep; 1: neighbour-to-neighbour publication counters
#endif
constexpr int SOR_R = SF_SOR_R;             // rows per lane (4 or 8); rows p and p + SOR_R/2 form one packed-fp32 pair
constexpr int SOR_HP = SOR_R / 2;           // row pairs per lane and column
constexpr int SOR_NW = SF_SOR_NW;           // warps per CTA
constexpr int SOR_TW = 64;                  // tile width: 2 pixels per lane
constexpr int SOR_TH = SOR_R * SOR_NW;      // tile height
// Staging area of a tile: the 9 planes a11' a12' a22' b1 | b2 psi_h psi_v du dv, each [SOR_TH rows][64].  The first group
// (SOR_GA planes, no dependencies: coefficients) is double-buffered and pulled TWO tiles ahead, the second group (+ the
// psi_v row above the tile) has one buffer and is pulled one tile ahead, after the compute warps have read the previous
// tile out of it.  (A single buffer for all nine planes left the bulk copies ~5000 clk per tile, and they take ~5400
// while the half sweeps' shuffles share the shared-memory data path: 1000-1400 clk of every tile were spent waiting.)
constexpr int SOR_GA = 4;                         // planes of group A: a11' a12' a22' b1
constexpr int SOR_GB = 5;                         // planes of group B: b2 psi_h psi_v du dv
constexpr int SOR_PLANE_FLOATS = SOR_TH * SOR_TW;
constexpr int SOR_STAGE_A_BYTES = SOR_GA * SOR_PLANE_FLOATS * 4;            // one of the two buffers of group A
constexpr int SOR_STAGE_B_BYTES = (SOR_GB * SOR_PLANE_FLOATS + SOR_TW) * 4; // group B + the row above
constexpr int SOR_STAGE_BYTES = 2 * SOR_STAGE_A_BYTES + ((SOR_STAGE_B_BYTES + 1023) & ~1023);
constexpr int SOR_EXCH_BYTES = 2 /*slots*/ * 2 /*top,bottom*/ * SOR_NW * 32 * 8;
constexpr int SOR_SMEM_BYTES = SOR_STAGE_BYTES + SOR_EXCH_BYTES + 128 /*pub, tickets*/ + 8 * 8 /*mbarriers*/ + 1024 /*alignment slack*/;
static_assert(SOR_SMEM_BYTES <= 227 * 1024, "the staging area does not fit");

#ifdef SF_SOR_CLOCKS
__device__ unsigned long long g_sor_hclk[8]; // loader warp: free-wait, issue B, fetch + wait for B, issue A + deps, -, -, -, rounds
__device__ unsigned long long g_sor_clk[8]; // tma-wait, load, sweeps, barrier-wait, store, -, -, tiles (warp 0 of every CTA)
#define SOR_CLK(var) const long long var = clock64()
#ifndef SF_SOR_CLOCK_THREAD
#define SF_SOR_CLOCK_THREAD 0
#endif
#else
#define SOR_CLK(var)
#endif

// Horizontal halo of a tile for T fused sweeps.  TMA needs the box origin 16-byte aligned along x.  With tile origins at
// multiples of the interior width (SF_SOR_EDGE_TILES) that holds for the exact halo 2T (interior 64 - 4T is a multiple
// of 4); with origins shifted by the halo it has to be rounded up to a multiple of 4 floats.
__host__ __device__ __forceinline__ int sor_halo_x(int T) { return SF_SOR_EDGE_TILES ? 2 * T : ((2 * T + 3) & ~3); }

struct SorTiledArgs {
    Geom g;
    float *arena;   // plane 0 of the SOR arena (the iterate planes are addressed by plane index)
    int plane_a;    // du plane of the iterate buffer pass 0 reads (dv = du + 1); pass p reads buffer (p & 1), writes the other
    int plane_b;
    int T;          // sweeps fused per pass (fixes the tiling: halo 2T)
    int passes;     // passes chained in this launch (tile-level dependencies instead of kernel boundaries)
    int T_last;     // sweeps of the last pass (<= T)
    int tiles_x, tiles_y;
    float omega;
    int zero_init;  // the initial iterate is 0: pass 0 does not load du,dv
    float one;      // 1.0f, opaque to the compiler (see ldp in k_sor_tiled)
    unsigned *sync; // [0] ticket counter, [1] CTAs that have left, [2] generation (see SorPlan::sync)
    unsigned *done; // per tile: generation + number of passes the tile has completed
};

// gpu-scope flag accesses for the tile dependencies
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Per-lane register tile: pixel columns {2l, 2l+1} of the warp's 8 rows.  Every quantity is held as a
// packed pair (row p, row p+4) of one column in ONE 64-bit register so that the whole relaxation runs on
// packed-fp32 FFMA2 (fma.rn.f32x2, sm_100): both rows of a pair have the same colour in a given column.
// (64-bit inline-asm operands make ptxas keep the pairs in aligned register pairs; with float2 values it
// re-assembled every operand with MOVs.)
struct SorRegs {
    p64 na11[2][SOR_HP], na12[2][SOR_HP], na22[2][SOR_HP]; // NEGATED inverse blocks, [column][pair]
    p64 b1[2][SOR_HP], b2[2][SOR_HP];
    p64 du[2][SOR_HP], dv[2][SOR_HP];
    p64 phl[SOR_HP], phm[SOR_HP], phr[SOR_HP]; // psi_h at columns 2l-1, 2l, 2l+1
    p64 pv[2][SOR_HP];                         // psi_v of the rows
    p64 pvt[2];                                // psi_v of (row -1, row 3): the rows above pair 0
};

// Relax the colour-C pixels of row pair P (rows P and P+4).  up = (du,dv) of the row above the warp's strip
// (used by P == 0), dn = of the row below (P == SOR_HP-1).  Same FMA chain per component as the scalar form
//   B = psl*dl + (psr*dr + (pst*dt + (psb*db + b)));  d += omega*(a1*B1 + (a2*B2 - d))
// written with negated a and omega (fma(-a,B,d) = -fma(a,B,-d) exactly), which needs no negation op:
//   n = (-a1)*B1 + ((-a2)*B2 + d);  d += (-omega)*n
template <int C, int P>
__device__ __forceinline__ void sor_relax_pair(SorRegs &q, const float2 up, const float2 dn, const p64 nomega2) {
    constexpr int e = (C + P) & 1; // column of the lane's pair that has colour C in these rows (tile origin is even/even)
    p64 ul, vl, ur, vr, psl, psr;
    if (e == 0) {
        ul = shfl_up2(q.du[1][P]);
        vl = shfl_up2(q.dv[1][P]);
        ur = q.du[1][P];
        vr = q.dv[1][P];
        psl = q.phl[P];
        psr = q.phm[P];
    } else {
        ul = q.du[0][P];
        vl = q.dv[0][P];
        ur = shfl_down2(q.du[0][P]);
        vr = shfl_down2(q.dv[0][P]);
        psl = q.phm[P];
        psr = q.phr[P];
    }
    p64 ut, vt, pst, ub, vb;
    if (P == 0) {
        ut = pk(up.x, lo_of(q.du[e][SOR_HP - 1]));
        vt = pk(up.y, lo_of(q.dv[e][SOR_HP - 1]));
        pst = q.pvt[e];
    } else {
        ut = q.du[e][P - 1];
        vt = q.dv[e][P - 1];
        pst = q.pv[e][P - 1];
    }
    if (P == SOR_HP - 1) {
        ub = pk(hi_of(q.du[e][0]), dn.x);
        vb = pk(hi_of(q.dv[e][0]), dn.y);
    } else {
        ub = q.du[e][P + 1];
        vb = q.dv[e][P + 1];
    }
    const p64 psb = q.pv[e][P];
    const p64 B1 = fma2(psl, ul, fma2(psr, ur, fma2(pst, ut, fma2(psb, ub, q.b1[e][P]))));
    const p64 B2 = fma2(psl, vl, fma2(psr, vr, fma2(pst, vt, fma2(psb, vb, q.b2[e][P]))));
    const p64 nu = fma2(q.na11[e][P], B1, fma2(q.na12[e][P], B2, q.du[e][P]));
    const p64 nv = fma2(q.na12[e][P], B1, fma2(q.na22[e][P], B2, q.dv[e][P]));
    q.du[e][P] = fma2(nomega2, nu, q.du[e][P]);
    q.dv[e][P] = fma2(nomega2, nv, q.dv[e][P]);
}

// relax pairs [P, PEND) of colour C; with SF_SOR_SKIP a pair whose deeper row has depth <= k is skipped
// (half sweep k only has to be right for pixels at depth >= k+1 from the tile edge; warp-uniform)
template <int C, int P, int PEND>
__device__ __forceinline__ void sor_relax_range(SorRegs &q, const float2 up, const float2 dn, const p64 nomega2, int k,
                                                const int (&pdepth)[SOR_HP]) {
    if constexpr (P < PEND) {
        if (!SF_SOR_SKIP || k < pdepth[P]) sor_relax_pair<C, P>(q, up, dn, nomega2);
        sor_relax_range<C, P + 1, PEND>(q, up, dn, nomega2, k, pdepth);
    }
}

// CTA = SOR_NW compute warps + ONE helper warp.  The compute warps only ever touch shared memory and their own output
// rows; everything that needs a gpu-scope fence or an L2 round trip -- fetching tickets, watching the dependency flags,
// requesting the TMA copies of the next tile for all strips, publishing finished tiles -- runs on the helper warp, where
// a stalled fence does not hold up the relaxation (a release store issued by a compute warp waited for the bulk copies
// in flight: +5000 clk per tile).
//
// A launch runs `passes` passes of T sweeps over the image.  Work items ("tickets") are numbered pass-major, tile-minor
// and handed out by an atomic counter, so a CTA that holds ticket t only ever waits for tickets < t, which are held by
// CTAs that are running: no deadlock whatever the residency of the grid.  Tile (tx,ty) of pass p reads the iterate
// buffer pass p-1 wrote (tile + halo) and overwrites the interior of the buffer pass p-1 read, so it has to wait for the
// 3x3 tile neighbourhood of pass p-1: `done[tile]` holds generation + passes completed.  The passes of a launch
// therefore overlap: no drained pipeline, no cold first wave and no partial last wave between them.
//
// Shared-memory hand-shakes of a round (= one tile of this CTA), all mbarriers:
//   fullB      (count 1)       helper: arrive.expect_tx + the bulk copies of group B  ->  compute warps: group B staged;
//                              the ticket of round r is written to tickq[r & 3] before the arrive
//   fullA[r&1] (count 1)       the same for group A of round r (requested one round earlier)
//   freeb      (count SOR_NW)  warp w has read its strip into registers  ->  helper: both buffers round r used are free
//   stored     (count SOR_NW)  warp w has issued the tile's global stores  ->  helper: release-store the tile's flag
// The register file is split between the four schedulers (512 registers per lane each), so a ninth warp would cap every
// warp of the CTA at 168 registers.  The helper is therefore a whole warp group (4 warps, 3 of which leave at once)
// that hands its registers back (setmaxnreg), and the compute warp groups take them: (8 x 232 + 4 x 40) x 32 = the 384 x 168 registers the CTA is launched with.
constexpr int SOR_THREADS = (SOR_NW + 4) * 32;
static_assert(SOR_NW % 4 == 0, "setmaxnreg works on warp groups of four warps");

__global__ void __launch_bounds__(SOR_THREADS, 1)
k_sor_tiled(const __grid_constant__ CUtensorMap tmap_ga, const __grid_constant__ CUtensorMap tmap_gb,
            const __grid_constant__ CUtensorMap tmap_iter, const __grid_constant__ CUtensorMap tmap_row, SorTiledArgs a) {
    extern __shared__ unsigned char smem_raw[];
    // TMA destinations need 128-byte alignment; round the dynamic base up to 1 KB
    unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *stageA = reinterpret_cast<float *>(base);                           // [2][SOR_GA planes][SOR_TH][64]
    float *stageB = reinterpret_cast<float *>(base + 2 * SOR_STAGE_A_BYTES);   // [SOR_GB planes][SOR_TH][64] + row above [64]
    float2 *exch = reinterpret_cast<float2 *>(base + SOR_STAGE_BYTES);
    int *pub = reinterpret_cast<int *>(base + SOR_STAGE_BYTES + SOR_EXCH_BYTES);  // publications per warp
    volatile unsigned *tickq = reinterpret_cast<volatile unsigned *>(pub + 16);  // [4] ticket of round r & 3
    uint64_t *fullA = reinterpret_cast<uint64_t *>(pub + 32);                     // [2]
    uint64_t *fullB = fullA + 2, *freeb = fullA + 3, *stored = fullA + 4;

    // halo: 2T pixels per side (sor_halo_x: exact with tile origins at multiples of the interior width)
    const int hy = 2 * a.T, hx = sor_halo_x(a.T);
    const int IW = SOR_TW - 2 * hx, IH = SOR_TH - 2 * hy;
    const int ntiles = a.tiles_x * a.tiles_y;
    const unsigned total = (unsigned)(ntiles * a.passes);

    if (threadIdx.x == 0) {
        for (int w = 0; w < SOR_NW; w++) pub[w] = 0;
        pub[24] = 0;  // rounds published
        pub[25] = -1; // rounds in all: not known yet
        mbar_init(fullA, 1);
        mbar_init(fullA + 1, 1);
        mbar_init(fullB, 1);
        mbar_init(freeb, SOR_NW);
        mbar_init(stored, SOR_NW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

#if SF_SOR_PDL
    // the next launch of the chain may become resident while this grid drains (sf_internal.cuh: pdl_enter); the
    // barrier set-up above touches no global memory
    pdl_enter();
#endif
    if (a.g.cancelled()) return; // (nothing is in flight yet)

    // ================================================================================== helper warps
    if (warp >= SOR_NW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp > SOR_NW + 1) return;
        const unsigned gen = *reinterpret_cast<volatile unsigned *>(a.sync + 2);
        volatile int *hstate = reinterpret_cast<volatile int *>(pub + 24); // [0] rounds published, [1] rounds in all (-1: open)

        // ------------------------------------------------------------------------------ publisher (warp SOR_NW + 1)
        // Release-stores the flag of every finished tile.  A warp of its own because the store's gpu-scope fence takes
        // 1300-2300 clk, which the loader does not have between two tiles.
        if (warp == SOR_NW + 1) {
            int k = 0;
            for (;;) {
                uint32_t spins = 0;
                bool got;
                while (!(got = mbar_try_wait(stored, (uint32_t)(k & 1)))) { // (suspends for a bounded time)
                    const int n = hstate[1];
                    if (n >= 0 && k >= n) break;
                    if (++spins > (1u << 24)) __trap();
                }
                if (!got) break;
                if (lane == 0) {
                    const unsigned pt = tickq[k & 3], pass = pt / (unsigned)ntiles;
                    st_release_gpu(a.done + (pt - pass * (unsigned)ntiles), gen + pass + 1u);
                    hstate[0] = k + 1;
                }
                k++;
            }
            // the last CTA to leave re-arms the counters for the next launch on this plan (by then every loader has
            // fetched its last ticket and every helper has read `gen`)
            if (lane == 0) {
                if (atomicAdd(a.sync + 1, 1u) == gridDim.x - 1) {
                    a.sync[0] = 0;
                    a.sync[1] = 0;
                    a.sync[2] = gen + (unsigned)a.passes;
                    __threadfence();
                }
            }
            return;
        }

        // ------------------------------------------------------------------------------ loader (warp SOR_NW)
        auto fetch = [&]() -> unsigned { // next ticket (warp-uniform)
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(a.sync, 1u);
            return __shfl_sync(0xffffffffu, t, 0);
        };
        // the 3x3 neighbourhood of the previous pass: lanes 0..8 look at one flag each (acquire)
        auto deps_wait = [&](unsigned t) {
            const int pass = (int)(t / (unsigned)ntiles), tl = (int)(t - (unsigned)(pass * ntiles));
            const int nx = tl % a.tiles_x + lane % 3 - 1, ny = tl / a.tiles_x + (lane / 3) - 1;
            const bool check = pass > 0 && lane < 9 && nx >= 0 && nx < a.tiles_x && ny >= 0 && ny < a.tiles_y;
            uint32_t spins = 0;
            for (;;) {
                bool ok = true;
                if (check) ok = (int)(ld_acquire_gpu(a.done + ny * a.tiles_x + nx) - (gen + (unsigned)pass)) >= 0;
                if (__all_sync(0xffffffffu, ok)) break;
                __nanosleep(200);
                if (++spins > (1u << 22)) __trap();
            }
            // the neighbours' stores went through the generic proxy of other SMs; the bulk copies read through the async proxy
            if (lane == 0) asm volatile("fence.proxy.async.global;" ::: "memory");
        };
        auto origin = [&](unsigned t, int &x0, int &y0, int &pass) {
            pass = (int)(t / (unsigned)ntiles);
            const int tl = (int)(t - (unsigned)(pass * ntiles));
            x0 = (tl % a.tiles_x) * IW - SOR_EDGE * hx;
            y0 = (tl / a.tiles_x) * IH - SOR_EDGE * hy;
        };
        // group A of a ticket (coefficient planes: no dependencies) into buffer `buf`
        auto issue_a = [&](unsigned t, int buf) { // lane 0
            int x0, y0, pass;
            origin(t, x0, y0, pass);
            mbar_expect_tx(fullA + buf, (uint32_t)SOR_STAGE_A_BYTES);
            tma_load_3d(stageA + buf * (SOR_GA * SOR_PLANE_FLOATS), &tmap_ga, fullA + buf, x0, y0, 0);
        };
        // group B: b2 psi_h psi_v, the iterate of the pass (not for the zero iterate of pass 0), the psi_v row above the tile
        auto issue_b = [&](unsigned t) { // lane 0
            int x0, y0, pass;
            origin(t, x0, y0, pass);
            const bool zero_it = a.zero_init && pass == 0;
            mbar_expect_tx(fullB, (uint32_t)(((zero_it ? 3 : 5) * SOR_PLANE_FLOATS + SOR_TW) * 4));
            tma_load_3d(stageB, &tmap_gb, fullB, x0, y0, SOR_GA);
            if (!zero_it) tma_load_3d(stageB + 3 * SOR_PLANE_FLOATS, &tmap_iter, fullB, x0, y0, (pass & 1) ? a.plane_b : a.plane_a);
            tma_load_3d(stageB + SOR_GB * SOR_PLANE_FLOATS, &tmap_row, fullB, x0, y0 - 1, SP_PV);
        };
        // Round k: group B is requested when round k-1 has left the staging area (freeb(k-1)), group A -- into the buffer
        // round k-2 used -- one round earlier, but only after group B of round k-1 has landed: the two would share the
        // copy bandwidth and group B is the one the compute warps wait for.  The ticket of round k+1 and the dependency
        // flags of round k are looked at during the sweeps of round k-1.  A wait for a tile of another CTA cannot dead-
        // lock: rounds handed out before are computed and published without this warp.
        unsigned tk[2]; // tickets of rounds k, k+1
        tk[0] = fetch();
        tk[1] = tk[0] < total ? fetch() : total;
        int k = 0;
        if (tk[0] < total) {
            if (lane == 0) issue_a(tk[0], 0);
            deps_wait(tk[0]);
        }
        for (;;) {
            SOR_CLK(h0);
            const unsigned t = tk[k & 1];
            if (k > 0) mbar_poll_wait(freeb, (uint32_t)((k - 1) & 1)); // round k-1 has left the staging area
            SOR_CLK(h1);
#ifdef SF_SOR_CLOCKS
            if (lane == 0 && k > 0) atomicAdd(&g_sor_hclk[4], (unsigned long long)((unsigned)h1 - (unsigned)pub[28]));
#endif
            {   // tickq[k & 3] was read by the publisher for round k-4
                uint32_t spins = 0;
                while (hstate[0] < k - 3)
                    if (++spins > (1u << 26)) __trap();
            }
            if (lane == 0) {
                if (t < total) {
                    tickq[k & 3] = t;
                    issue_b(t);
                } else { // end marker
                    tickq[k & 3] = total;
                    hstate[1] = k;
                    mbar_arrive(fullB);
                }
            }
            SOR_CLK(h2);
            if (t >= total) break;
            // round k+1: its group A goes into the buffer round k-1 has just left, once group B of round k is in
            const unsigned tn = tk[(k + 1) & 1];
            tk[k & 1] = tn < total ? fetch() : total; // (ticket of round k+2)
            if (tn < total) {
#if SF_SOR_A_AFTER_B
                mbar_poll_wait(fullB, (uint32_t)(k & 1));
#endif
                SOR_CLK(h3);
                if (lane == 0) issue_a(tn, (k + 1) & 1);
                deps_wait(tn);
#ifdef SF_SOR_CLOCKS
                if (lane == 0) {
                    const long long h4 = clock64();
                    atomicAdd(&g_sor_hclk[0], (unsigned long long)(h1 - h0));
                    atomicAdd(&g_sor_hclk[1], (unsigned long long)(h2 - h1));
                    atomicAdd(&g_sor_hclk[2], (unsigned long long)(h3 - h2));
                    atomicAdd(&g_sor_hclk[3], (unsigned long long)(h4 - h3));
                    atomicAdd(&g_sor_hclk[7], 1ull);
                }
#endif
            }
            k++;
        }
        return;
    }

    // ================================================================================== compute warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    uint32_t phase = 0;
    int npub = 0; // publications of this warp so far (= of every warp it is in step with)
    int round = 0;

    // Boundary rows travel between neighbouring warps through a 2-slot ring per warp and side:
    // exch[((slot*2 + side)*SOR_NW + warp)*32 + lane], side 0 = top row, 1 = bottom row.
    auto ex = [&](int slot, int side, int w) -> float2 * { return exch + ((slot * 2 + side) * SOR_NW + w) * 32 + lane; };
#ifdef SF_SOR_CLOCKS
    long long wait_clk = 0;
#endif
    auto publish = [&](float2 top, float2 bottom) {
        *ex(npub & 1, 0, warp) = top;
        *ex(npub & 1, 1, warp) = bottom;
        npub++;
#if SF_SOR_SYNC == 1
        __syncwarp();
        if (lane == 0) st_release_shared(pub + warp, npub);
#endif
    };
    // wait until both neighbours have made publication number `need` (1-based)
    auto wait_neighbours = [&](int need) {
        SOR_CLK(w0);
#if SF_SOR_SYNC == 0
        (void)need;
        asm volatile("bar.sync 1, %0;" ::"n"(SOR_NW * 32) : "memory"); // the compute warps only
#else
        uint32_t spins = 0;
        if (warp > 0)
            while (ld_acquire_shared(pub + warp - 1) < need)
                if (++spins > (1u << 24)) __trap();
        if (warp < SOR_NW - 1)
            while (ld_acquire_shared(pub + warp + 1) < need)
                if (++spins > (1u << 24)) __trap();
        __syncwarp();
#endif
#ifdef SF_SOR_CLOCKS
        wait_clk += clock64() - w0;
#endif
    };
    const p64 nomega2 = pk(-a.omega, -a.omega);
    const float2 zero2 = make_float2(0.0f, 0.0f);
    // Half sweep k only has to be right for pixels at depth >= k+1 from the tile edge, so (with SF_SOR_SKIP) a
    // row pair whose deeper row has depth <= k is skipped in half sweep k (warp-uniform).
    int pdepth[SOR_HP];
#pragma unroll
    for (int p = 0; p < SOR_HP; p++) {
        const int r0 = warp * SOR_R + p, r1 = r0 + SOR_HP;
        const int d0 = r0 < SOR_TH - 1 - r0 ? r0 : SOR_TH - 1 - r0, d1 = r1 < SOR_TH - 1 - r1 ? r1 : SOR_TH - 1 - r1;
        pdepth[p] = d0 > d1 ? d0 : d1;
    }

    for (;;) {
        SOR_CLK(c0);
        mbar_wait(fullB, phase);
        phase ^= 1u;
        const unsigned cur_t = tickq[round & 3];
        if (cur_t >= total) break;
        SOR_CLK(c0b);
        mbar_wait(fullA + (round & 1), (uint32_t)((round >> 1) & 1));
        const int pass = (int)(cur_t / (unsigned)ntiles), tile = (int)(cur_t - (unsigned)(pass * ntiles));
        const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
        const int x0 = tx * IW - SOR_EDGE * hx, y0 = ty * IH - SOR_EDGE * hy;
        const bool zero_it = a.zero_init && pass == 0;
        const int nhalf = 2 * (pass == a.passes - 1 ? a.T_last : a.T);
        SOR_CLK(c1);

        // ---- shared -> registers
        SorRegs q;
        {
            // LDS.64 delivers the two COLUMNS of one row; a pair is (row p, row p+4) of ONE column, so every two
            // loads are transposed into two pairs.  Each pair is then passed through one FFMA2 (x*1 + 0, or x*(-1) + 0
            // for the negated blocks) so that it is DEFINED by a 64-bit instruction: ptxas then keeps it in an
            // aligned register pair instead of re-assembling it from two scalars with MOVs in front of every use.
            const p64 one2 = pk(a.one, a.one), mone2 = pk(-a.one, -a.one), z2 = pk(0.0f, 0.0f);
            // the warp's strip within group A (planes 0..3, this round's buffer) and group B (planes 4..8)
            const float2 *rowA = reinterpret_cast<const float2 *>(stageA + (round & 1) * (SOR_GA * SOR_PLANE_FLOATS)) + (warp * SOR_R * SOR_TW) / 2 + lane;
            const float2 *rowB = reinterpret_cast<const float2 *>(stageB) + (warp * SOR_R * SOR_TW) / 2 + lane;
            auto ld2 = [&](int plane, int p, p64 &c0, p64 &c1, p64 scale) {
                const float2 *row0 = plane < SOR_GA ? rowA + (plane * SOR_PLANE_FLOATS) / 2 : rowB + ((plane - SOR_GA) * SOR_PLANE_FLOATS) / 2;
                const float2 lo = row0[(p * SOR_TW) / 2];
                const float2 hi = row0[((p + SOR_HP) * SOR_TW) / 2];
                c0 = fma2(pk(lo.x, hi.x), scale, z2);
                c1 = fma2(pk(lo.y, hi.y), scale, z2);
            };
#pragma unroll
            for (int p = 0; p < SOR_HP; p++) {
                ld2(SP_A11, p, q.na11[0][p], q.na11[1][p], mone2);
                ld2(SP_A12, p, q.na12[0][p], q.na12[1][p], mone2);
                ld2(SP_A22, p, q.na22[0][p], q.na22[1][p], mone2);
                ld2(SP_B1, p, q.b1[0][p], q.b1[1][p], one2);
                ld2(SP_B2, p, q.b2[0][p], q.b2[1][p], one2);
                ld2(SP_PV, p, q.pv[0][p], q.pv[1][p], one2);
                ld2(SP_PH, p, q.phm[p], q.phr[p], one2);
                const p64 left = shfl_up2(q.phr[p]);
                q.phl[p] = (lane == 0) ? z2 : left; // column -1 of the tile is never needed for a valid pixel
                if (zero_it) {
                    q.du[0][p] = q.du[1][p] = q.dv[0][p] = q.dv[1][p] = z2;
                } else {
                    ld2(7, p, q.du[0][p], q.du[1][p], one2);
                    ld2(8, p, q.dv[0][p], q.dv[1][p], one2);
                }
            }
            // psi_v of the row above the strip: the last row of the strip above, or the extra row staged above the tile
            // (zero-filled by the TMA unit above the image)
            const float2 above = (warp > 0) ? rowB[((SP_PV - SOR_GA) * SOR_PLANE_FLOATS - SOR_TW) / 2]
                                            : reinterpret_cast<const float2 *>(stageB)[(SOR_GB * SOR_PLANE_FLOATS) / 2 + lane];
            q.pvt[0] = fma2(pk(above.x, lo_of(q.pv[0][SOR_HP - 1])), one2, z2);
            q.pvt[1] = fma2(pk(above.y, lo_of(q.pv[1][SOR_HP - 1])), one2, z2);
        }
        __syncwarp(); // every lane has left the staging block
        if (lane == 0) mbar_arrive(freeb); // the helper may pull the coming tiles into the buffers this round used
#ifdef SF_SOR_CLOCKS
        if (threadIdx.x == 0) pub[28] = (int)(unsigned)clock64();
#endif

        // initial publication "as if colour 1 had just been relaxed": top row column 1, bottom row column 0
        const int pub0 = npub;
        publish(make_float2(lo_of(q.du[1][0]), lo_of(q.dv[1][0])), make_float2(hi_of(q.du[0][SOR_HP - 1]), hi_of(q.dv[0][SOR_HP - 1])));

        // ---- 2T half sweeps in registers.  Per half sweep: the interior pairs (own rows only), then -- once the
        // neighbours' boundary rows of the previous half sweep are visible -- the two boundary pairs, then publish.
        SOR_CLK(c2);
#pragma unroll 1
        for (int k = 0; k < nhalf; k += 2) {
            sor_relax_range<0, 1, SOR_HP - 1>(q, zero2, zero2, nomega2, k, pdepth);
            wait_neighbours(pub0 + k + 1);
            {
                const int slot = (pub0 + k) & 1;
                const float2 up = (warp > 0) ? *ex(slot, 1, warp - 1) : zero2;
                const float2 dn = (warp < SOR_NW - 1) ? *ex(slot, 0, warp + 1) : zero2;
                sor_relax_range<0, 0, 1>(q, up, dn, nomega2, k, pdepth);
                sor_relax_range<0, SOR_HP - 1, SOR_HP>(q, up, dn, nomega2, k, pdepth);
                publish(make_float2(lo_of(q.du[0][0]), lo_of(q.dv[0][0])), make_float2(hi_of(q.du[1][SOR_HP - 1]), hi_of(q.dv[1][SOR_HP - 1])));
            }
            sor_relax_range<1, 1, SOR_HP - 1>(q, zero2, zero2, nomega2, k + 1, pdepth);
            wait_neighbours(pub0 + k + 2);
            {
                const int slot = (pub0 + k + 1) & 1;
                const float2 up = (warp > 0) ? *ex(slot, 1, warp - 1) : zero2;
                const float2 dn = (warp < SOR_NW - 1) ? *ex(slot, 0, warp + 1) : zero2;
                sor_relax_range<1, 0, 1>(q, up, dn, nomega2, k + 1, pdepth);
                sor_relax_range<1, SOR_HP - 1, SOR_HP>(q, up, dn, nomega2, k + 1, pdepth);
                if (k + 2 < nhalf)
                    publish(make_float2(lo_of(q.du[1][0]), lo_of(q.dv[1][0])), make_float2(hi_of(q.du[0][SOR_HP - 1]), hi_of(q.dv[0][SOR_HP - 1])));
            }
        }

        // ---- interior of the tile -> global (float2 per lane and row: 256 B per warp row)
        SOR_CLK(c3);
        const int cx = 2 * lane, gx = x0 + cx;
#if SF_SOR_EDGE_TILES
        // a tile side that lies on (or beyond) the image border needs no halo: the boundary condition there is exact,
        // so the first tile of a row / column keeps its first hx columns / hy rows and the last one its last
        const int cx_lo = (tx == 0) ? 0 : hx, cx_hi = (x0 + SOR_TW >= a.g.W) ? SOR_TW : SOR_TW - hx;
        const int tr_lo = (ty == 0) ? 0 : hy, tr_hi = (y0 + SOR_TH >= a.g.H) ? SOR_TH : SOR_TH - hy;
#else
        const int cx_lo = hx, cx_hi = SOR_TW - hx, tr_lo = hy, tr_hi = SOR_TH - hy;
#endif
        const size_t Pl = a.g.plane();
        float *const out_du = a.arena + (size_t)((pass & 1) ? a.plane_a : a.plane_b) * Pl, *const out_dv = out_du + Pl;
        if (cx >= cx_lo && cx < cx_hi && gx < a.g.W) {
#pragma unroll
            for (int p = 0; p < SOR_HP; p++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int tr = warp * SOR_R + p + h * SOR_HP, gy = y0 + tr;
                    if (tr >= tr_lo && tr < tr_hi && gy < a.g.H) {
                        const size_t o = (size_t)gy * a.g.S + gx;
                        *reinterpret_cast<float2 *>(out_du + o) = h ? make_float2(hi_of(q.du[0][p]), hi_of(q.du[1][p]))
                                                                    : make_float2(lo_of(q.du[0][p]), lo_of(q.du[1][p]));
                        *reinterpret_cast<float2 *>(out_dv + o) = h ? make_float2(hi_of(q.dv[0][p]), hi_of(q.dv[1][p]))
                                                                    : make_float2(lo_of(q.dv[0][p]), lo_of(q.dv[1][p]));
                    }
                }
            }
        }
        // the strip's stores are ordered before the helper's release store of the tile flag by this arrive (release.cta)
        __syncwarp();
        if (lane == 0) mbar_arrive(stored);
        round++;
#ifdef SF_SOR_CLOCKS
        if (threadIdx.x == SF_SOR_CLOCK_THREAD) {
            const long long c4 = clock64();
            atomicAdd(&g_sor_clk[0], (unsigned long long)(c1 - c0));
            atomicAdd(&g_sor_clk[1], (unsigned long long)(c2 - c1));
            atomicAdd(&g_sor_clk[2], (unsigned long long)(c3 - c2));
            atomicAdd(&g_sor_clk[3], (unsigned long long)wait_clk);
            atomicAdd(&g_sor_clk[4], (unsigned long long)(c4 - c3));
            atomicAdd(&g_sor_clk[5], (unsigned long long)(c0b - c0));
            if (c0b - c0 > 500) atomicAdd(&g_sor_clk[6], 1ull);
            atomicAdd(&g_sor_clk[7], 1ull);
            wait_clk = 0;
        }
#endif
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled lookup_encode_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        return reinterpret_cast<PFN_encodeTiled>(p);
    return nullptr;
}
static PFN_encodeTiled get_encode_fn() {
    static const PFN_encodeTiled fn = lookup_encode_fn(); // initialised once, thread-safe (one host thread per device)
    return fn;
}

// Function attributes are per device: every context sets them on its own device when it is created (a process-wide
// "done once" flag left the second GPU of a multi-threaded driver without its shared-memory opt-in).
bool sor_device_init() {
    if (cudaFuncSetAttribute(k_sor_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, SOR_SMEM_BYTES) != cudaSuccess) {
        set_error("cudaFuncSetAttribute(k_sor_tiled) failed");
        return false;
    }
    return sor_stream_device_init();
}

void sor_plan_release(SorPlan &plan) {
    if (plan.sync) cudaFree(plan.sync);
    plan.sync = nullptr;
    plan.sync_tiles = 0;
    plan.tmap_valid = false;
}

bool sor_plan_init(SorPlan &plan, Geom g, float *arena, int num_sms) {
    plan.g = g;
    plan.arena = arena;
    plan.num_sms = num_sms;
    plan.tmap_valid = false;
    if (!plan.sync) {
        // ticket counter, exit counter, generation, pad + one completion word per tile (zeroed once; the kernel keeps them
        // consistent from launch to launch)
        const size_t words = 4 + SorPlan::SYNC_TILES;
        if (cudaMalloc(&plan.sync, words * sizeof(unsigned)) != cudaSuccess || cudaMemset(plan.sync, 0, words * sizeof(unsigned)) != cudaSuccess) {
            plan.sync = nullptr;
            set_error("cudaMalloc of the SOR tile flags failed");
            return false;
        }
        plan.sync_tiles = SorPlan::SYNC_TILES;
    }
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
        return false;
    }
    // views of the same (x, y, plane) arena that differ in the box: a tile of the first four coefficient planes, of the
    // other three, of a du,dv plane pair, and a single row (psi_v above the tile)
    const cuuint64_t dims[3] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)SP_COUNT};
    const cuuint64_t strides[2] = {(cuuint64_t)g.S * 4, (cuuint64_t)g.plane() * 4};
    const cuuint32_t estr[3] = {1, 1, 1};
    unsigned sb[3][3];
    sor_stream_boxes(sb);
    const cuuint32_t boxes[7][3] = {{SOR_TW, SOR_TH, SOR_GA}, {SOR_TW, SOR_TH, 2}, {SOR_TW, 1, 1},
                                    {sb[0][0], sb[0][1], sb[0][2]}, {sb[1][0], sb[1][1], sb[1][2]}, {sb[2][0], sb[2][1], sb[2][2]},
                                    {SOR_TW, SOR_TH, SOR_GB - 2}};
    CUtensorMap *maps[7] = {&plan.tmap, &plan.tmap_iter, &plan.tmap_row, &plan.tmap_s_coef, &plan.tmap_s_iter, &plan.tmap_s_row,
                            &plan.tmap_gb};
    for (int m = 0; m < 7; m++) {
        const CUresult r = enc(maps[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, arena, dims, strides, boxes[m], estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char buf[128];
            snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (CUresult %d) for %dx%d stride %d", (int)r, g.W, g.H, g.S);
            set_error(buf);
            return false;
        }
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
    if (iterations <= 0) {
        // the reference erases du,dv and then runs a 0-iteration sor_coupled (variational.c:44-45, solver.c:20): the
        // increment the callers add to the flow must be a defined 0, not whatever the arena held
        if (zero_init) {
            cudaMemsetAsync(A + (size_t)(*cur ? SP_DUB : SP_DUA) * P, 0, P * sizeof(float), st);
            cudaMemsetAsync(A + (size_t)(*cur ? SP_DVB : SP_DVA) * P, 0, P * sizeof(float), st);
        }
        return 0;
    }
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
    if (variant == 2) return launch_sor_stream(st, plan, iterations, omega, fuse, cur, zero_init);
    if (fuse < 1) fuse = 1;
    const int max_fuse = (SOR_TH - 4) / 4 < 7 ? (SOR_TH - 4) / 4 : 7; // the halo (2*fuse per side) must leave an interior
    if (fuse > max_fuse) fuse = max_fuse;
    int done = 0;
    while (done < iterations) {
        // all passes go into ONE launch (tile-level dependencies between the passes).  A remainder of fewer than `fuse`
        // sweeps is the last pass of that launch: it keeps the tiling (the halo is then deeper than it needs) and runs
        // fewer half sweeps -- a launch of its own with the tighter tiling costs more than the bytes it saves (2560x1440,
        // 7 x 4 + 2 sweeps: 43 us for the separate 2-sweep launch).
        const int T = (iterations - done < fuse) ? (iterations - done) : fuse;
        int passes = (iterations - done + T - 1) / T;
        int T_last = (iterations - done) - (passes - 1) * T;
        SorTiledArgs a;
        a.g = g;
        a.T = T;
        const int IW = SOR_TW - 2 * sor_halo_x(T), IH = SOR_TH - 4 * T;
#if SF_SOR_EDGE_TILES
        // tile origins at multiples of the interior size: the first and the last tile of a row / column keep the halo
        // side that lies on the image border (exact there), so n tiles cover n * I + 2 * halo pixels
        a.tiles_x = (g.W - (SOR_TW - IW) + IW - 1) / IW;
        a.tiles_y = (g.H - (SOR_TH - IH) + IH - 1) / IH;
        if (a.tiles_x < 1) a.tiles_x = 1;
        if (a.tiles_y < 1) a.tiles_y = 1;
#else
        a.tiles_x = (g.W + IW - 1) / IW;
        a.tiles_y = (g.H + IH - 1) / IH;
#endif
        const int ntiles = a.tiles_x * a.tiles_y;
        if (SF_SOR_MULTIPASS == 0 || (size_t)ntiles > plan.sync_tiles) { // (no dependency flags for that many tiles)
            passes = 1;
            T_last = T;
        }
        a.passes = passes;
        a.T_last = T_last;
        a.omega = omega;
        a.one = 1.0f;
        a.zero_init = (zero_init && done == 0) ? 1 : 0;
        a.arena = A;
        a.plane_a = *cur ? SP_DUB : SP_DUA;
        a.plane_b = *cur ? SP_DUA : SP_DUB;
        a.sync = plan.sync;
        a.done = plan.sync + 4;
        const long long work = (long long)ntiles * passes;
        const int grid = work < plan.num_sms ? (int)work : plan.num_sms;
#if SF_SOR_PDL
        if (launch_pdl(k_sor_tiled, dim3(grid), dim3(SOR_THREADS), (size_t)SOR_SMEM_BYTES, st, plan.tmap, plan.tmap_gb, plan.tmap_iter, plan.tmap_row, a) !=
            cudaSuccess) {
            set_error("cudaLaunchKernelEx(k_sor_tiled) failed");
            return -1;
        }
#else
        k_sor_tiled<<<grid, SOR_THREADS, SOR_SMEM_BYTES, st>>>(plan.tmap, plan.tmap_gb, plan.tmap_iter, plan.tmap_row, a);
#endif
        if (passes & 1) *cur ^= 1;
        done += T * (passes - 1) + T_last;
        launches++;
    }
    return launches;
}

} // namespace sf
