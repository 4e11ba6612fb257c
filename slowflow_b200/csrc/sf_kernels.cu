// sf_kernels.cu -- per-pixel operators of the two-frame path (sm_100a).
//
//   K0 dpsis_weight   variational_aux.c:183-209 / variational_aux_mt.cpp:673-719
//   K1 image_warp     variational_aux.c:18-52   / variational_aux_mt.cpp:722-756
//   K3 smoothness     variational_aux.c:84-149  / variational_aux_mt.cpp:18-92 (modes 0,1)
//   (K2, the fused derivative + data-term kernel, lives in sf_data.cu)
//   K5 add / fill     variational.c:44-48, 60-65
//
// All kernels are HBM-bound stencils: coalesced row-major access, halos staged in shared memory,
// every intermediate (the 24 derivative planes of the reference) kept on chip.  Border semantics
// follow image.c:400-526: horizontal taps clamp the column index, vertical taps fold the missing
// coefficients into the nearest row.
#include "sf_internal.cuh"
#include "sf_penalty.cuh"
#include "sf_stencil.cuh"

namespace sf {

// ------------------------------------------------------------------------------------------ K0
// CTA = 32 x 8 pixels; the luminance of the tile + 2 px halo is formed once in shared memory (1.7 evaluations per pixel
// instead of 9: the kernel was instruction-bound on the nine de-normalise / weight / divide chains per pixel).
__global__ void __launch_bounds__(256) k_dpsis_weight(Geom g, const float *__restrict__ im, float *__restrict__ out,
                                                      float coef, float a1, float a2, float a3, float s1, float s2,
                                                      float s3, float divisor) {
    pdl_enter();
    constexpr int TW = 32, TH = 8, LW = TW + 4, LH = TH + 4;
    __shared__ float lum[LH][LW + 1];
    const int i0 = blockIdx.x * TW, j0 = blockIdx.y * TH;
    const int tid = threadIdx.y * TW + threadIdx.x;
    const size_t P = g.plane();
    const int W1 = g.W - 1, H1 = g.H - 1;
    // luminance / divisor as a product with the correctly rounded reciprocal (<= 1 ulp from the reference's quotient,
    // variational_aux.c:190-196)
    const float inv_div = __frcp_rn(divisor);
    for (int idx = tid; idx < LH * LW; idx += TW * TH) {
        const int ly = idx / LW, lx = idx - ly * LW;
        const size_t o = (size_t)clampi(j0 - 2 + ly, 0, H1) * g.S + clampi(i0 - 2 + lx, 0, W1);
        lum[ly][lx] = (0.299f * (im[o] * s1 + a1) + 0.587f * (im[o + P] * s2 + a2) + 0.114f * (im[o + 2 * P] * s3 + a3)) * inv_div;
    }
    __syncthreads();
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = i0 + tx, j = j0 + ty;
    if (i >= g.S || j >= g.H) return;
    float v = 0.0f;
    if (i < g.W) {
        const float lx = hconv5(lum[ty + 2][tx], lum[ty + 2][tx + 1], lum[ty + 2][tx + 2], lum[ty + 2][tx + 3], lum[ty + 2][tx + 4]);
        const float ly = vconv5(lum[ty][tx + 2], lum[ty + 1][tx + 2], lum[ty + 2][tx + 2], lum[ty + 3][tx + 2], lum[ty + 4][tx + 2], j, g.H);
        v = 0.5f * expf(-coef * sqrtf(lx * lx + ly * ly));
    }
    out[(size_t)j * g.S + i] = v;
}

void launch_dpsis_weight(cudaStream_t st, Geom g, const float *im3, float *out, float coef, const float avg[3],
                         const float stdv[3], float divisor) {
    dim3 b(32, 8), grid((g.S + 31) / 32, (g.H + 7) / 8);
    launch_pdl(k_dpsis_weight, grid, b, 0, st, g, im3, out, coef, avg[0], avg[1], avg[2], stdv[0], stdv[1], stdv[2], divisor);
}

// ------------------------------------------------------------------------------------------ K1
// One thread warps 4 consecutive pixels: float4 loads of the flow, float4 stores of the three
// warped channels and the mask.  The 4x3 bilinear gathers hit L1/L2 (flow is smooth).
__global__ void __launch_bounds__(256) k_warp(Geom g, const float *__restrict__ src, const float *__restrict__ wx,
                                              const float *__restrict__ wy, float factor, float *__restrict__ dst,
                                              float *__restrict__ mask) {
    pdl_enter();
    if (g.cancelled()) return;
    const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i4 >= g.S || j >= g.H) return;
    const size_t P = g.plane();
    const size_t o = (size_t)j * g.S + i4;
    const float4 fx = *reinterpret_cast<const float4 *>(wx + o);
    const float4 fy = *reinterpret_cast<const float4 *>(wy + o);
    const float fxs[4] = {fx.x, fx.y, fx.z, fx.w}, fys[4] = {fy.x, fy.y, fy.z, fy.w};
    float r0[4], r1[4], r2[4], mk[4];
    const float Wm1 = (float)(g.W - 1), Hm1 = (float)(g.H - 1);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = i4 + k;
        r0[k] = r1[k] = r2[k] = 0.0f;
        mk[k] = 0.0f;
        if (i < g.W) {
            const float xx = (float)i + factor * fxs[k];
            const float yy = (float)j + factor * fys[k];
            const float xf = floorf(xx), yf = floorf(yy);
            const float dx = xx - xf, dy = yy - yf;
            mk[k] = (xx >= 0.0f && xx <= Wm1 && yy >= 0.0f && yy <= Hm1) ? 1.0f : 0.0f;
            // clamp in float first so that huge flows cannot overflow the int conversion
            const int x = (int)fminf(fmaxf(xf, -2.0f), Wm1 + 2.0f), y = (int)fminf(fmaxf(yf, -2.0f), Hm1 + 2.0f);
            const int x1 = clampi(x, 0, g.W - 1), x2 = clampi(x + 1, 0, g.W - 1);
            const int y1 = clampi(y, 0, g.H - 1), y2 = clampi(y + 1, 0, g.H - 1);
            const size_t o11 = (size_t)y1 * g.S + x1, o12 = (size_t)y1 * g.S + x2, o21 = (size_t)y2 * g.S + x1,
                         o22 = (size_t)y2 * g.S + x2;
            // reference order: s11*(1-dx)*(1-dy) + s12*dx*(1-dy) + s21*(1-dx)*dy + s22*dx*dy
            r0[k] = __ldg(src + o11) * (1.0f - dx) * (1.0f - dy) + __ldg(src + o12) * dx * (1.0f - dy) +
                    __ldg(src + o21) * (1.0f - dx) * dy + __ldg(src + o22) * dx * dy;
            r1[k] = __ldg(src + P + o11) * (1.0f - dx) * (1.0f - dy) + __ldg(src + P + o12) * dx * (1.0f - dy) +
                    __ldg(src + P + o21) * (1.0f - dx) * dy + __ldg(src + P + o22) * dx * dy;
            r2[k] = __ldg(src + 2 * P + o11) * (1.0f - dx) * (1.0f - dy) + __ldg(src + 2 * P + o12) * dx * (1.0f - dy) +
                    __ldg(src + 2 * P + o21) * (1.0f - dx) * dy + __ldg(src + 2 * P + o22) * dx * dy;
        }
    }
    *reinterpret_cast<float4 *>(dst + o) = make_float4(r0[0], r0[1], r0[2], r0[3]);
    *reinterpret_cast<float4 *>(dst + P + o) = make_float4(r1[0], r1[1], r1[2], r1[3]);
    *reinterpret_cast<float4 *>(dst + 2 * P + o) = make_float4(r2[0], r2[1], r2[2], r2[3]);
    if (mask) *reinterpret_cast<float4 *>(mask + o) = make_float4(mk[0], mk[1], mk[2], mk[3]);
}

void launch_warp(cudaStream_t st, Geom g, const float *src3, const float *wx, const float *wy, int factor,
                 float *dst3, float *mask) {
    dim3 b(32, 8), grid((g.S / 4 + 31) / 32, (g.H + 7) / 8);
    launch_pdl(k_warp, grid, b, 0, st, g, src3, wx, wy, (float)factor, dst3, mask);
}

// ------------------------------------------------------------------------------------------ K3
// psi_h(i,j): edge (i,j)-(i+1,j); psi_v(i,j): edge (i,j)-(i,j+1).  3-tap central differences
// [-0.5, 0, 0.5] with image.c:400-423 vertical border folding and clamped horizontal reads.
//
// One thread owns 4 consecutive pixels of one row: the 3x6 neighbourhood of the flow comes from float4 loads of rows
// j-1, j, j+1 plus the neighbouring lanes' edge columns (warp shuffle; the two outer lanes of a warp load theirs).
// UPDATE: the flow is uu = wx + du, vv = wy + dv (variational.c:60-61 / :68-69 for niter_inner == 1) -- the sum is
// formed on the fly, stored to (wx_out, wy_out) (a different buffer: neighbouring rows are read by other CTAs) and the
// separate flow-update pass disappears.
struct Row6 {
    float c[6]; // columns i4-1 .. i4+4, clamped to the image
};
template <bool UPDATE>
__device__ __forceinline__ float4 load_flow4(const float *__restrict__ w, const float *__restrict__ d, size_t o) {
    float4 a = *reinterpret_cast<const float4 *>(w + o);
    if (UPDATE) {
        const float4 b = *reinterpret_cast<const float4 *>(d + o);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    return a;
}
template <bool UPDATE>
__device__ __forceinline__ float load_flow1(const float *__restrict__ w, const float *__restrict__ d, size_t o) {
    return UPDATE ? w[o] + d[o] : w[o];
}
// Row `r` of the (updated) flow around columns i4..i4+3, in two phases so that EVERY global load of the thread is in
// flight before the first shuffle waits for one of them (with load and shuffle interleaved per row, ptxas serialised
// seven memory round trips per thread and the kernel ran at a quarter of the HBM rate):
//   load_raw : the float4 (wx [+ du]) plus the edge neighbours that only the outer lanes of a warp need
//   finish_row6 : replicate-border patching and the neighbour columns from the adjacent lanes
struct RawRow {
    float4 a;   // columns i4 .. i4+3 (unclamped: what the update stores)
    float l, r; // columns i4-1 / i4+4 (meaningful in lane 0 / lane 31 only)
};
template <bool UPDATE>
__device__ __forceinline__ RawRow load_raw(const Geom &g, const float *__restrict__ w, const float *__restrict__ d, int r, int i4,
                                           int lane) {
    const size_t ro = (size_t)r * g.S;
    RawRow q;
    q.a = load_flow4<UPDATE>(w, d, ro + i4);
    q.l = (lane == 0 && i4 > 0) ? load_flow1<UPDATE>(w, d, ro + i4 - 1) : 0.0f;
    q.r = (lane == 31 && i4 + 4 <= g.W - 1) ? load_flow1<UPDATE>(w, d, ro + i4 + 4) : 0.0f;
    return q;
}
__device__ __forceinline__ Row6 finish_row6(const Geom &g, const RawRow &raw, int i4, int lane) {
    const int W1 = g.W - 1;
    const float4 a = raw.a;
    Row6 q;
    q.c[1] = a.x; q.c[2] = a.y; q.c[3] = a.z; q.c[4] = a.w;
    // columns beyond W-1 (stride padding) replicate column W-1, which lies in this float4 (S - W <= 3)
    if (i4 + 3 > W1) {
        const float last = (W1 - i4 == 0) ? a.x : (W1 - i4 == 1) ? a.y : a.z;
        if (i4 + 1 > W1) q.c[2] = last;
        if (i4 + 2 > W1) q.c[3] = last;
        q.c[4] = last;
    }
    const float from_l = __shfl_up_sync(0xffffffffu, q.c[4], 1), from_r = __shfl_down_sync(0xffffffffu, q.c[1], 1);
    if (i4 == 0) q.c[0] = q.c[1];
    else q.c[0] = (lane > 0) ? from_l : raw.l;
    if (i4 + 4 > W1) q.c[5] = q.c[4];
    else q.c[5] = (lane < 31) ? from_r : raw.r;
    return q;
}

// Central differences: with the rows / columns clamped at load time, -0.5*m + (-0)*c + 0.5*p (and its border-folded
// forms, image.c:400-423) equals 0.5*(p - m) exactly (scaling by 0.5 is exact), and the average of two of them is
// 0.25*((p - m) + (p' - m')).
// TWO_FRAME: (w + w')*half_alpha / sqrt(s^2 + 1e-6) (variational_aux.c:124,139) as a product with MUFU.RSQ (2 ulp;
// the reference rounds a double quotient once) -- otherwise the pluggable penalty of variational_aux_mt.cpp:65,90.
template <bool TWO_FRAME>
__device__ __forceinline__ float edge_weight(const Penalty &reg, float ww, float alpha_factor, float ssq) {
    if (TWO_FRAME) return ww * alpha_factor * rsqrtf(ssq + 0.001f * 0.001f);
    return smooth_weight(reg, ww, alpha_factor, ssq);
}

template <bool UPDATE, bool TWO_FRAME>
__global__ void __launch_bounds__(256) k_flow_smooth(Geom g, const float *__restrict__ wx, const float *__restrict__ wy,
                                                     const float *__restrict__ du, const float *__restrict__ dv,
                                                     const float *__restrict__ w, float alpha_factor, Penalty reg, int mode,
                                                     float *__restrict__ wx_out, float *__restrict__ wy_out,
                                                     float *__restrict__ ph, float *__restrict__ pv) {
    pdl_enter();
    if (g.cancelled()) return;
    const int lane = threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (j >= g.H) return; // warp-uniform (a warp is one row segment)
    int i4 = (blockIdx.x * 32 + lane) * 4;
    const bool live = i4 < g.S;
    if (!live) i4 = g.S - 4; // keep the lane in the shuffles with a valid address
    const int H1 = g.H - 1, W1 = g.W - 1;
    const int jm = j > 0 ? j - 1 : 0, jp = j < H1 ? j + 1 : H1;
    const size_t o = (size_t)j * g.S + i4;
    // ---- phase 1: every load of this thread
    const RawRow rUm = load_raw<UPDATE>(g, wx, du, jm, i4, lane), rU0 = load_raw<UPDATE>(g, wx, du, j, i4, lane),
                 rUp = load_raw<UPDATE>(g, wx, du, jp, i4, lane);
    const RawRow rVm = load_raw<UPDATE>(g, wy, dv, jm, i4, lane), rV0 = load_raw<UPDATE>(g, wy, dv, j, i4, lane),
                 rVp = load_raw<UPDATE>(g, wy, dv, jp, i4, lane);
    const float4 wa = *reinterpret_cast<const float4 *>(w + o);
    const float4 wb = *reinterpret_cast<const float4 *>(w + (size_t)jp * g.S + i4);
    const float we = (lane == 31 && i4 + 4 <= W1) ? w[o + 4] : 0.0f;
    // ---- phase 2: neighbour columns
    const float4 ru = rU0.a, rv = rV0.a;
    const Row6 Um = finish_row6(g, rUm, i4, lane), U0 = finish_row6(g, rU0, i4, lane), Up = finish_row6(g, rUp, i4, lane);
    const Row6 Vm = finish_row6(g, rVm, i4, lane), V0 = finish_row6(g, rV0, i4, lane), Vp = finish_row6(g, rVp, i4, lane);
    // smoothness weights: row j columns i4..i4+4, row j+1 columns i4..i4+3
    float w0[5], w1[4];
    {
        w0[0] = wa.x; w0[1] = wa.y; w0[2] = wa.z; w0[3] = wa.w;
        const float from_r = __shfl_down_sync(0xffffffffu, wa.x, 1);
        w0[4] = (lane < 31) ? from_r : we;
        w1[0] = wb.x; w1[1] = wb.y; w1[2] = wb.z; w1[3] = wb.w;
    }
    if (!live) return;
    const float cross = (mode != 0) ? 0.25f : 0.0f; // mode 0 (variational_aux_mt.cpp:30-60): forward differences only
    float h[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int i = i4 + k;
        const float u00 = U0.c[k + 1], v00 = V0.c[k + 1];
        {
            const float ux1 = U0.c[k + 2] - u00, vx1 = V0.c[k + 2] - v00;
            const float tu = cross * ((Up.c[k + 1] - Um.c[k + 1]) + (Up.c[k + 2] - Um.c[k + 2]));
            const float tv = cross * ((Vp.c[k + 1] - Vm.c[k + 1]) + (Vp.c[k + 2] - Vm.c[k + 2]));
            const float uxsq = ux1 * ux1 + tu * tu;
            const float vxsq = vx1 * vx1 + tv * tv;
            h[k] = (i < W1) ? edge_weight<TWO_FRAME>(reg, w0[k] + w0[k + 1], alpha_factor, uxsq + vxsq) : 0.0f;
        }
        {
            const float uy1 = Up.c[k + 1] - u00, vy1 = Vp.c[k + 1] - v00;
            const float tu = cross * ((U0.c[k + 2] - U0.c[k]) + (Up.c[k + 2] - Up.c[k]));
            const float tv = cross * ((V0.c[k + 2] - V0.c[k]) + (Vp.c[k + 2] - Vp.c[k]));
            const float uysq = uy1 * uy1 + tu * tu;
            const float vysq = vy1 * vy1 + tv * tv;
            v[k] = (i < g.W && j < H1) ? edge_weight<TWO_FRAME>(reg, w0[k] + w1[k], alpha_factor, uysq + vysq) : 0.0f;
        }
    }
    *reinterpret_cast<float4 *>(ph + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4 *>(pv + o) = make_float4(v[0], v[1], v[2], v[3]);
    if (UPDATE) {
        *reinterpret_cast<float4 *>(wx_out + o) = ru;
        *reinterpret_cast<float4 *>(wy_out + o) = rv;
    }
}

void launch_smoothness(cudaStream_t st, Geom g, const float *uu, const float *vv, const float *w, float alpha_factor,
                       Penalty reg, int mode, float *ph, float *pv) {
    dim3 b(32, 8), grid((g.S / 4 + 31) / 32, (g.H + 7) / 8);
    if (reg.type < 0)
        launch_pdl(k_flow_smooth<false, true>, grid, b, 0, st, g, uu, vv, nullptr, nullptr, w, alpha_factor, reg, mode, nullptr, nullptr, ph, pv);
    else
        launch_pdl(k_flow_smooth<false, false>, grid, b, 0, st, g, uu, vv, nullptr, nullptr, w, alpha_factor, reg, mode, nullptr, nullptr, ph, pv);
}

void launch_update_smoothness(cudaStream_t st, Geom g, const float *wx, const float *wy, const float *du, const float *dv,
                              const float *w, float alpha_factor, Penalty reg, int mode, float *wx_out, float *wy_out,
                              float *ph, float *pv) {
    dim3 b(32, 8), grid((g.S / 4 + 31) / 32, (g.H + 7) / 8);
    if (reg.type < 0)
        launch_pdl(k_flow_smooth<true, true>, grid, b, 0, st, g, wx, wy, du, dv, w, alpha_factor, reg, mode, wx_out, wy_out, ph, pv);
    else
        launch_pdl(k_flow_smooth<true, false>, grid, b, 0, st, g, wx, wy, du, dv, w, alpha_factor, reg, mode, wx_out, wy_out, ph, pv);
}

// ------------------------------------------------------------------------------------------ separable filters
// convolve_horiz / convolve_vert (image.c:400-645) with 3 or 5 taps over `planes` consecutive planes, all stride
// columns like the reference: horizontal taps clamp the column to [0, W-1] (which is also what the reference's in-place
// padding of src amounts to), vertical taps fold the coefficients of missing rows into the nearest row.  Operator
// twins only -- the hot path never materialises a derivative plane.
struct Taps5 {
    float c[5];
};
template <bool VERT>
__global__ void __launch_bounds__(256) k_convolve(Geom g, const float *__restrict__ src, float *__restrict__ dst, int order,
                                                  Taps5 t, int planes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.S || j >= g.H) return;
    const size_t P = g.plane();
    const int W1 = g.W - 1, H = g.H;
    const float *c = t.c;
    for (int p = 0; p < planes; p++) {
        const float *s = src + p * P;
        const size_t o = (size_t)j * g.S + i;
        float r;
        if (!VERT) {
            const float *row = s + (size_t)j * g.S;
            if (order == 1) r = c[0] * row[clampi(i - 1, 0, W1)] + c[1] * row[clampi(i, 0, W1)] + c[2] * row[clampi(i + 1, 0, W1)];
            else r = c[0] * row[clampi(i - 2, 0, W1)] + c[1] * row[clampi(i - 1, 0, W1)] + c[2] * row[clampi(i, 0, W1)] +
                     c[3] * row[clampi(i + 1, 0, W1)] + c[4] * row[clampi(i + 2, 0, W1)];
        } else {
            auto at = [&](int dj) { return s[(size_t)clampi(j + dj, 0, H - 1) * g.S + i]; }; // only in-image rows are weighted
            if (order == 1) {
                if (j == 0) r = (c[0] + c[1]) * at(0) + c[2] * at(1);
                else if (j == H - 1) r = c[0] * at(-1) + (c[1] + c[2]) * at(0);
                else r = c[0] * at(-1) + c[1] * at(0) + c[2] * at(1);
            } else {
                if (j == 0) r = (c[0] + c[1] + c[2]) * at(0) + c[3] * at(1) + c[4] * at(2);
                else if (j == 1) r = (c[0] + c[1]) * at(-1) + c[2] * at(0) + c[3] * at(1) + c[4] * at(2);
                else if (j == H - 2) r = c[0] * at(-2) + c[1] * at(-1) + c[2] * at(0) + (c[3] + c[4]) * at(1);
                else if (j == H - 1) r = c[0] * at(-2) + c[1] * at(-1) + (c[2] + c[3] + c[4]) * at(0);
                else r = c[0] * at(-2) + c[1] * at(-1) + c[2] * at(0) + c[3] * at(1) + c[4] * at(2);
            }
        }
        dst[p * P + o] = r;
    }
}
void launch_convolve(cudaStream_t st, Geom g, const float *src, float *dst, bool vertical, int order, const float *coeffs,
                     int planes) {
    Taps5 t;
    for (int k = 0; k < 5; k++) t.c[k] = (k < 2 * order + 1) ? coeffs[k] : 0.0f;
    dim3 b(32, 8), grid((g.S + 31) / 32, (g.H + 7) / 8);
    if (vertical) k_convolve<true><<<grid, b, 0, st>>>(g, src, dst, order, t, planes);
    else k_convolve<false><<<grid, b, 0, st>>>(g, src, dst, order, t, planes);
}

// m = 0.5*(im2 + im1), dt = im2 - im1 over 3 planes (variational_aux.c:63-69)
__global__ void __launch_bounds__(256) k_mean_diff(size_t n, const float *__restrict__ im1, const float *__restrict__ im2,
                                                   float *__restrict__ mean, float *__restrict__ dt) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const float a = im1[k], b = im2[k];
        mean[k] = 0.5f * (b + a);
        dt[k] = b - a;
    }
}
void launch_mean_diff(cudaStream_t st, size_t n, const float *im1, const float *im2, float *mean, float *dt) {
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_mean_diff<<<blocks, 256, 0, st>>>(n, im1, im2, mean, dt);
}

// ------------------------------------------------------------------------------------------ small operators
__global__ void __launch_bounds__(256) k_sub_laplacian(Geom g, float *__restrict__ dst, const float *__restrict__ src,
                                                       const float *__restrict__ ph, const float *__restrict__ pv) {
    if (g.cancelled()) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.W || j >= g.H) return;
    const size_t o = (size_t)j * g.S + i;
    const float wc = src[o];
    float b = dst[o];
    if (i > 0) b -= ph[o - 1] * (wc - src[o - 1]);
    if (i < g.W - 1) b += ph[o] * (src[o + 1] - wc);
    if (j > 0) b -= pv[o - g.S] * (wc - src[o - g.S]);
    if (j < g.H - 1) b += pv[o] * (src[o + g.S] - wc);
    dst[o] = b;
}
void launch_sub_laplacian(cudaStream_t st, Geom g, float *dst, const float *src, const float *ph, const float *pv) {
    dim3 b(32, 8), grid((g.W + 31) / 32, (g.H + 7) / 8);
    k_sub_laplacian<<<grid, b, 0, st>>>(g, dst, src, ph, pv);
}

__global__ void __launch_bounds__(256) k_invert_blocks(Geom g, float *__restrict__ a11, float *__restrict__ a12,
                                                       float *__restrict__ a22, const float *__restrict__ ph,
                                                       const float *__restrict__ pv) {
    if (g.cancelled()) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.W || j >= g.H) return;
    const size_t o = (size_t)j * g.S + i;
    const float hl = (i > 0) ? ph[o - 1] : 0.0f, hr = ph[o];
    const float vt = (j > 0) ? pv[o - g.S] : 0.0f, vb = (j < g.H - 1) ? pv[o] : 0.0f;
    const float sp = ((hl + hr) + vt) + vb;
    const float D11 = a22[o] + sp, D22 = a11[o] + sp, q = a12[o];
    const float det = D11 * D22 - q * q;
    a11[o] = D11 / det;
    a22[o] = D22 / det;
    a12[o] = q / -det;
}
void launch_invert_blocks(cudaStream_t st, Geom g, float *a11, float *a12, float *a22, const float *ph, const float *pv) {
    dim3 b(32, 8), grid((g.W + 31) / 32, (g.H + 7) / 8);
    k_invert_blocks<<<grid, b, 0, st>>>(g, a11, a12, a22, ph, pv);
}

// dst may alias a (the callers update a flow plane in place), so neither carries __restrict__
__global__ void __launch_bounds__(256) k_add4(size_t n4, float4 *dst, const float4 *a, const float4 *__restrict__ b) {
    pdl_enter();
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (size_t)gridDim.x * blockDim.x) {
        const float4 x = a[k], y = b[k];
        dst[k] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
}
void launch_add(cudaStream_t st, Geom g, float *dst, const float *a, const float *b) {
    const size_t n4 = g.plane() / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    launch_pdl(k_add4, dim3(blocks), dim3(256), 0, st, n4, reinterpret_cast<float4 *>(dst), reinterpret_cast<const float4 *>(a),
               reinterpret_cast<const float4 *>(b));
}

__global__ void __launch_bounds__(256) k_fill(size_t n, float *__restrict__ dst, float v) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) dst[k] = v;
}
void launch_fill(cudaStream_t st, float *dst, size_t n, float v) {
    if (v == 0.0f) {
        cudaMemsetAsync(dst, 0, n * sizeof(float), st);
        return;
    }
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_fill<<<blocks, 256, 0, st>>>(n, dst, v);
}

} // namespace sf
