// sf_wderivs.cu -- multi-frame data pass, first half: bilinear warp of ONE frame with its time factor + the frame's five
// spatial derivative images, fused, marching form (sm_100a).
//
// Replaces, per warped frame and outer iteration of the multi-frame path (variational_mt.cpp:96-166):
//     image_warp (time factor)      variational_aux_mt.cpp:722-760   -> warped frame (3 planes) + raw in-bounds mask
//     get_derivatives, per frame    variational_mt.cpp:111-166       -> Ix Iy Ixx Ixy Iyy, 15 planes [derivative][channel]
// which were two launches (k_warp, then k_data_term<DK_DERIVS> re-reading the warped frame through a 40x40 shared-memory
// tile: 15.6 + 40.9 us per 1280x1024 frame).  Here the warped row is formed once, stored, and fed straight into the
// rolling windows of the derivative stencils: the traversal, the window layout and the border rules are those of
// k_prep_two_frame (sf_prep.cu, sf_march.cuh) without the temporal difference, the data term and the linear system.
// HBM traffic per pixel: read wx,wy (8 B) + the frame through L1/L2 gathers (12); write warped (12) + mask (4) + 15
// derivative planes (60) = 96 B.
//
//   row r is warped (and stored) at step r; windows hold rows r-4..r-1 of the warped image I and rows r-6..r-3 of Ix, Iy;
//   the output row of step r is o = r - 4:  Ix(o), Iy(o) from the windows, Ixx = d/dx Ix(o), Ixy = d/dy Ix, Iyy = d/dy Iy.
#include <atomic>

#include "sf_internal.cuh"
#include "sf_march.cuh"

namespace sf {

constexpr int WD_WARPS = 2;          // warps per CTA (independent)
constexpr int SF_WD_MAX_FRAMES = 8;  // frames per launch (2S warped frames of a window; more go into further launches)
constexpr int WD_RING = 36 * 32;     // packed values of one warp's windows: I, Ix, Iy x 3 channels x 4 rows
#ifndef SF_WD_MINB
#define SF_WD_MINB 8
#endif

// tuning knobs (tools/wderivs_bench.cu): minimum rows per segment, resident-warp override
int g_wd_min_rows = 4, g_wd_warps_per_sm = 0; // (short segments: the kernel is parallelism-bound at 1 Mpx, tools/wderivs_bench.cu)

// one frame of a launch
struct WdFrame {
    const float *src;     // the frame (3 planes)
    float factor;         // time factor of the frame
    float *warped, *mask; // 3 planes, 1 plane
    float *derivs;        // 15 planes [Ix Iy Ixx Ixy Iyy][channel]
};
// All warped frames of a window go into ONE launch: the (strip, segment) items of every frame share the one wave of
// resident warps, so the segments are F times longer and the 8 window-filling rows of a segment weigh F times less
// (1280x1024, 4 frames: 44-row instead of 12-row segments).
struct WdArgs {
    Geom g;
    const float *wx, *wy; // flow of the reference frame
    WdFrame f[SF_WD_MAX_FRAMES];
    int nframes;
    int strips, seg_rows, segs, nwork;
};

template <bool EDGE>
__device__ __forceinline__ void wd_march(const WdArgs &wa, const WdFrame &a, const int strip, const int seg, const int lane, p64 *ring_sm) {
    const Geom g = wa.g;
    const int W = g.W, H = g.H, H1 = H - 1, S = g.S;
    const size_t P = g.plane();

    Lane L;
    const int X0 = strip * PR_OUT_W - 2 * PR_OUT_LO;
    L.x0 = X0 + 2 * lane;
    L.xc0 = clampi(L.x0, 0, W - 1);
    L.xc1 = clampi(L.x0 + 1, 0, W - 1);
    L.lane_l = clampi((0 - X0) >> 1, 0, 31);
    L.lane_r = clampi((W - 1 - X0) >> 1, 0, 31);
    L.comp_r = (W - 1 - X0) & 1;

    const int Y0 = seg * wa.seg_rows;
    const int Yend = min(Y0 + wa.seg_rows, H);
    const int Rbase = Y0 - 4;

    const bool out_lane = (lane >= PR_OUT_LO) && (lane <= PR_OUT_HI) && (L.x0 < S);
    const bool v0 = L.x0 < W, v1 = L.x0 + 1 < W; // valid (non-padding) columns of the pair
    const float fxc0 = (float)L.xc0, fxc1 = (float)L.xc1;
    const float Wm1 = (float)(W - 1), Hm1 = (float)H1;
    const p64 zero2 = splat2(0.0f);

    enum { RG_I = 0, RG_IX = 1, RG_IY = 2 };
    p64 *const sm = ring_sm + lane;
    auto rd = [&](int arr, int c, int s) -> p64 { return sm[((arr * 3 + c) * 4 + s) * 32]; };
    auto wr = [&](int arr, int c, int s, p64 v) { sm[((arr * 3 + c) * 4 + s) * 32] = v; };
#pragma unroll
    for (int k = 0; k < 36; k++) sm[k * 32] = zero2;

    // the flow of the NEXT row to be warped is loaded one step ahead
    p64 fx, fy;
    {
        const int ro = clampi(Rbase, 0, H1) * S;
        fx = load_pair<EDGE>(wa.wx, ro, L);
        fy = load_pair<EDGE>(wa.wy, ro, L);
    }

    // warp row r (clamped), store it when it is a row of this segment, return the three channel pairs
    auto warp_row = [&](const int r, p64 (&B)[3]) {
        const int rr = clampi(r, 0, H1);
        const float x0f = fmaf(a.factor, lo_of(fx), fxc0), y0f = fmaf(a.factor, lo_of(fy), (float)rr);
        const float x1f = fmaf(a.factor, hi_of(fx), fxc1), y1f = fmaf(a.factor, hi_of(fy), (float)rr);
        const Taps t0 = warp_taps(g, x0f, y0f), t1 = warp_taps(g, x1f, y1f);
#pragma unroll
        for (int c = 0; c < 3; c++) B[c] = pk(warp_fetch(a.src + c * P, t0), warp_fetch(a.src + c * P, t1));
        {
            const int rn = clampi(r + 1, 0, H1) * S;
            fx = load_pair<EDGE>(wa.wx, rn, L);
            fy = load_pair<EDGE>(wa.wy, rn, L);
        }
        if (r >= Y0 && r < Yend && out_lane) { // (rows of the segment are never clamped)
            const size_t off = (size_t)r * S + L.x0;
#pragma unroll
            for (int c = 0; c < 3; c++)
                *reinterpret_cast<float2 *>(a.warped + c * P + off) = make_float2(v0 ? lo_of(B[c]) : 0.0f, v1 ? hi_of(B[c]) : 0.0f);
            const float mk0 = (x0f >= 0.0f && x0f <= Wm1 && y0f >= 0.0f && y0f <= Hm1) ? 1.0f : 0.0f;
            const float mk1 = (x1f >= 0.0f && x1f <= Wm1 && y1f >= 0.0f && y1f <= Hm1) ? 1.0f : 0.0f;
            *reinterpret_cast<float2 *>(a.mask + off) = make_float2(v0 ? mk0 : 0.0f, v1 ? mk1 : 0.0f);
        }
    };

    // one marching step; U = (r - Rbase) & 3 is the ring slot of row r.  FULL: also the second-stage derivatives and the
    // stores of output row o = r - 4 (the first 8 steps of a segment only fill the windows)
    auto step = [&](auto full_tag, const int U, const int r) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int U1 = (U + 1) & 3, U2 = (U + 2) & 3, U3 = (U + 3) & 3;
        const int o = r - 4;
        p64 B[3];
        warp_row(r, B);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            p64 ix_new = hconv_pair(rd(RG_I, c, U2)); // Ix of row r-2
            if (EDGE) ix_new = xedge_fix(ix_new, L, W);
            p64 iy_new = vconv_pair(rd(RG_I, c, U), rd(RG_I, c, U1), rd(RG_I, c, U3), B[c]); // Iy of row r-2
            const p64 iy_p1 = rd(RG_IY, c, U1);
            if (r - 2 > H1) iy_new = iy_p1; // second-stage replicate border in y: Iy beyond the last row is Iy(H-1)
            if (FULL) {
                const p64 ix_o = rd(RG_IX, c, U);
                const p64 ixx = hconv_pair(ix_o); // (columns outside the image were patched in Ix itself)
                const p64 ixy = vconv_pair(rd(RG_IX, c, U2), rd(RG_IX, c, U3), rd(RG_IX, c, U1), ix_new);
                const p64 iy_c = rd(RG_IY, c, U), iy_b = rd(RG_IY, c, U3);
                p64 iy_m2 = rd(RG_IY, c, U2), iy_m1 = iy_b;
                if (o == 0) iy_m2 = iy_m1 = iy_c; // rows -2, -1 take Iy(0)
                else if (o == 1) iy_m2 = iy_b;    // row -1 takes Iy(0)
                const p64 iyy = vconv_pair(iy_m2, iy_m1, iy_p1, iy_new);
                if (o >= Y0 && o < Yend && out_lane) {
                    float *d = a.derivs + (size_t)c * P + (size_t)o * S + L.x0;
                    // padding columns: defined zeros (the reference leaves garbage there)
                    *reinterpret_cast<float2 *>(d) = make_float2(v0 ? lo_of(ix_o) : 0.0f, v1 ? hi_of(ix_o) : 0.0f);
                    *reinterpret_cast<float2 *>(d + 3 * P) = make_float2(v0 ? lo_of(iy_c) : 0.0f, v1 ? hi_of(iy_c) : 0.0f);
                    *reinterpret_cast<float2 *>(d + 6 * P) = make_float2(v0 ? lo_of(ixx) : 0.0f, v1 ? hi_of(ixx) : 0.0f);
                    *reinterpret_cast<float2 *>(d + 9 * P) = make_float2(v0 ? lo_of(ixy) : 0.0f, v1 ? hi_of(ixy) : 0.0f);
                    *reinterpret_cast<float2 *>(d + 12 * P) = make_float2(v0 ? lo_of(iyy) : 0.0f, v1 ? hi_of(iyy) : 0.0f);
                }
            }
            // rotate this channel's windows (every value of the slots being overwritten has been consumed above)
            wr(RG_IX, c, U2, ix_new);
            wr(RG_I, c, U, B[c]);
            wr(RG_IY, c, U2, iy_new);
        }
    };
    struct Lean { enum { value = 0 }; };
    struct Full { enum { value = 1 }; };

    const int Rend = Yend + 4; // last warped row is Yend + 3
    // output row o = r - 4 reaches the segment at r = Y0 + 4: steps Rbase .. Y0 + 3 are warm-up
#pragma unroll 1
    for (int r = Rbase; r < Y0 + 4; r++) step(Lean{}, (r - Rbase) & 3, r);
#pragma unroll 1
    for (int r = Y0 + 4; r < Rend; r++) step(Full{}, (r - Rbase) & 3, r);
}

__global__ void __launch_bounds__(WD_WARPS * 32, SF_WD_MINB) k_warp_derivs(WdArgs a) {
    pdl_enter();
    if (a.g.cancelled()) return;
    __shared__ p64 ring[WD_WARPS * WD_RING];
    const int lane = threadIdx.x & 31;
    const int work = blockIdx.x * WD_WARPS + (threadIdx.x >> 5);
    if (work >= a.nwork) return; // whole warp
    const int strip = work % a.strips, rest = work / a.strips, seg = rest % a.segs;
    const WdFrame fr = a.f[rest / a.segs];
    const int X0 = strip * PR_OUT_W - 2 * PR_OUT_LO;
    p64 *ring_sm = ring + (threadIdx.x >> 5) * WD_RING;
    if ((X0 < 0) || (X0 + 63 > a.g.W - 1)) wd_march<true>(a, fr, strip, seg, lane, ring_sm);
    else wd_march<false>(a, fr, strip, seg, lane, ring_sm);
}

void launch_warp_derivs_batch(cudaStream_t st, Geom g, int num_sms, const float *wx, const float *wy, int nframes,
                              const float *const *src3, const int *factor, float *const *warped3, float *const *mask,
                              float *const *derivs15) {
    // resident warps per SM (the same on every device of the box; relaxed atomics: one host thread per device may get here
    // at the same time)
    static std::atomic<int> resident{0};
    int res = resident.load(std::memory_order_relaxed);
    if (!res) {
        int blocks_per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_warp_derivs, WD_WARPS * 32, 0);
        res = (blocks_per_sm > 0 ? blocks_per_sm : 1) * WD_WARPS;
        resident.store(res, std::memory_order_relaxed);
    }
    if (g_wd_warps_per_sm > 0) res = g_wd_warps_per_sm;
    for (int f0 = 0; f0 < nframes; f0 += SF_WD_MAX_FRAMES) {
        WdArgs a;
        a.g = g;
        a.wx = wx; a.wy = wy;
        a.nframes = nframes - f0 < SF_WD_MAX_FRAMES ? nframes - f0 : SF_WD_MAX_FRAMES;
        for (int k = 0; k < SF_WD_MAX_FRAMES; k++) {
            const int f = f0 + (k < a.nframes ? k : 0);
            a.f[k].src = src3[f]; a.f[k].factor = (float)factor[f];
            a.f[k].warped = warped3[f]; a.f[k].mask = mask[f]; a.f[k].derivs = derivs15[f];
        }
        a.strips = (g.S + PR_OUT_W - 1) / PR_OUT_W;
        // rows per segment: as many (frame, strip, segment) items as fit ONE wave of resident warps
        int segs = (num_sms * res) / (a.strips * a.nframes);
        if (segs < 1) segs = 1;
        int rows = (g.H + segs - 1) / segs;
        if (rows < g_wd_min_rows) rows = g_wd_min_rows;
        a.seg_rows = (rows + 3) & ~3;
        a.segs = (g.H + a.seg_rows - 1) / a.seg_rows;
        a.nwork = a.strips * a.segs * a.nframes;
        launch_pdl(k_warp_derivs, dim3((a.nwork + WD_WARPS - 1) / WD_WARPS), dim3(WD_WARPS * 32), 0, st, a);
    }
}

void launch_warp_derivs(cudaStream_t st, Geom g, int num_sms, const float *src3, const float *wx, const float *wy, int factor,
                        float *warped3, float *mask, float *derivs15) {
    launch_warp_derivs_batch(st, g, num_sms, wx, wy, 1, &src3, &factor, &warped3, &mask, &derivs15);
}

} // namespace sf
