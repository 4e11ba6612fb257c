// sf_kernels.cu -- per-pixel operators of the two-frame path (sm_100a).
//
//   K0 dpsis_weight   variational_aux.c:183-209 / variational_aux_mt.cpp:673-719
//   K1 image_warp     variational_aux.c:18-52   / variational_aux_mt.cpp:722-756
//   K3 smoothness     variational_aux.c:84-149  / variational_aux_mt.cpp:18-92 (modes 0,1)
//   K2 data term      variational_aux.c:55-78 (derivatives) + :215-302 (data term)
//                     [+ :153-180 laplacian + solver.c:101-106 block inverse when fused]
//   K5 add / fill     variational.c:44-48, 60-65
//
// All kernels are HBM-bound stencils: coalesced row-major access, halos staged in shared memory,
// every intermediate (the 24 derivative planes of the reference) kept on chip.  Border semantics
// follow image.c:400-526: horizontal taps clamp the column index, vertical taps fold the missing
// coefficients into the nearest row.
#include "sf_internal.cuh"
#include "sf_penalty.cuh"

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

// ------------------------------------------------------------------------------------------ K0
__global__ void __launch_bounds__(256) k_dpsis_weight(Geom g, const float *__restrict__ im, float *__restrict__ out,
                                                      float coef, float a1, float a2, float a3, float s1, float s2,
                                                      float s3, float divisor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.S || j >= g.H) return;
    const size_t P = g.plane();
    auto lum = [&](int x, int y) -> float {
        const size_t o = (size_t)y * g.S + x;
        return (0.299f * (im[o] * s1 + a1) + 0.587f * (im[o + P] * s2 + a2) + 0.114f * (im[o + 2 * P] * s3 + a3)) / divisor;
    };
    float v = 0.0f;
    if (i < g.W) {
        const int W1 = g.W - 1, H1 = g.H - 1;
        const float lx = hconv5(lum(clampi(i - 2, 0, W1), j), lum(clampi(i - 1, 0, W1), j), lum(i, j),
                                lum(clampi(i + 1, 0, W1), j), lum(clampi(i + 2, 0, W1), j));
        const float ly = vconv5(lum(i, clampi(j - 2, 0, H1)), lum(i, clampi(j - 1, 0, H1)), lum(i, j),
                                lum(i, clampi(j + 1, 0, H1)), lum(i, clampi(j + 2, 0, H1)), j, g.H);
        v = 0.5f * expf(-coef * sqrtf(lx * lx + ly * ly));
    }
    out[(size_t)j * g.S + i] = v;
}

void launch_dpsis_weight(cudaStream_t st, Geom g, const float *im3, float *out, float coef, const float avg[3],
                         const float stdv[3], float divisor) {
    dim3 b(32, 8), grid((g.S + 31) / 32, (g.H + 7) / 8);
    k_dpsis_weight<<<grid, b, 0, st>>>(g, im3, out, coef, avg[0], avg[1], avg[2], stdv[0], stdv[1], stdv[2], divisor);
}

// ------------------------------------------------------------------------------------------ K1
// One thread warps 4 consecutive pixels: float4 loads of the flow, float4 stores of the three
// warped channels and the mask.  The 4x3 bilinear gathers hit L1/L2 (flow is smooth).
__global__ void __launch_bounds__(256) k_warp(Geom g, const float *__restrict__ src, const float *__restrict__ wx,
                                              const float *__restrict__ wy, float factor, float *__restrict__ dst,
                                              float *__restrict__ mask) {
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
            const float w11 = (1.0f - dx) * (1.0f - dy), w12 = dx * (1.0f - dy), w21 = (1.0f - dx) * dy, w22 = dx * dy;
            // reference order: s11*(1-dx)*(1-dy) + s12*dx*(1-dy) + s21*(1-dx)*dy + s22*dx*dy
            r0[k] = __ldg(src + o11) * (1.0f - dx) * (1.0f - dy) + __ldg(src + o12) * dx * (1.0f - dy) +
                    __ldg(src + o21) * (1.0f - dx) * dy + __ldg(src + o22) * dx * dy;
            r1[k] = __ldg(src + P + o11) * (1.0f - dx) * (1.0f - dy) + __ldg(src + P + o12) * dx * (1.0f - dy) +
                    __ldg(src + P + o21) * (1.0f - dx) * dy + __ldg(src + P + o22) * dx * dy;
            r2[k] = __ldg(src + 2 * P + o11) * (1.0f - dx) * (1.0f - dy) + __ldg(src + 2 * P + o12) * dx * (1.0f - dy) +
                    __ldg(src + 2 * P + o21) * (1.0f - dx) * dy + __ldg(src + 2 * P + o22) * dx * dy;
            (void)w11; (void)w12; (void)w21; (void)w22;
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
    k_warp<<<grid, b, 0, st>>>(g, src3, wx, wy, (float)factor, dst3, mask);
}

// ------------------------------------------------------------------------------------------ K3
// psi_h(i,j): edge (i,j)-(i+1,j); psi_v(i,j): edge (i,j)-(i,j+1).  3-tap central differences
// [-0.5, 0, 0.5] with image.c:400-423 vertical border folding and clamped horizontal reads.
__device__ __forceinline__ float hdiff3(float m1, float s0, float p1) { return -0.5f * m1 + (-0.0f) * s0 + 0.5f * p1; }
__device__ __forceinline__ float vdiff3(float m1, float s0, float p1, int j, int H) {
    if (j == 0) return (-0.5f + -0.0f) * s0 + 0.5f * p1;
    if (j == H - 1) return -0.5f * m1 + (-0.0f + 0.5f) * s0;
    return -0.5f * m1 + (-0.0f) * s0 + 0.5f * p1;
}

__global__ void __launch_bounds__(256) k_smoothness(Geom g, const float *__restrict__ uu, const float *__restrict__ vv,
                                                    const float *__restrict__ w, float alpha_factor, Penalty reg,
                                                    int mode, float *__restrict__ ph, float *__restrict__ pv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.S || j >= g.H) return;
    const size_t o = (size_t)j * g.S + i;
    float h = 0.0f, v = 0.0f;
    if (i < g.W) {
        const int W1 = g.W - 1, H1 = g.H - 1;
        auto U = [&](int x, int y) { return uu[(size_t)clampi(y, 0, H1) * g.S + clampi(x, 0, W1)]; };
        auto V = [&](int x, int y) { return vv[(size_t)clampi(y, 0, H1) * g.S + clampi(x, 0, W1)]; };
        const float u00 = U(i, j), v00 = V(i, j);
        if (i < g.W - 1) {
            const float ux1 = U(i + 1, j) - u00, vx1 = V(i + 1, j) - v00;
            float tu = 0.0f, tv = 0.0f;
            if (mode != 0) {
                const float uy2a = vdiff3(U(i, j - 1), u00, U(i, j + 1), j, g.H);
                const float uy2b = vdiff3(U(i + 1, j - 1), U(i + 1, j), U(i + 1, j + 1), j, g.H);
                const float vy2a = vdiff3(V(i, j - 1), v00, V(i, j + 1), j, g.H);
                const float vy2b = vdiff3(V(i + 1, j - 1), V(i + 1, j), V(i + 1, j + 1), j, g.H);
                tu = 0.5f * (uy2a + uy2b);
                tv = 0.5f * (vy2a + vy2b);
            }
            const float uxsq = ux1 * ux1 + tu * tu;
            const float vxsq = vx1 * vx1 + tv * tv;
            const float ssq = uxsq + vxsq;
            const float ww = w[o] + w[o + 1];
            h = smooth_weight(reg, ww, alpha_factor, ssq);
        }
        if (j < g.H - 1) {
            const float uy1 = U(i, j + 1) - u00, vy1 = V(i, j + 1) - v00;
            float tu = 0.0f, tv = 0.0f;
            if (mode != 0) {
                const float ux2a = hdiff3(U(i - 1, j), u00, U(i + 1, j));
                const float ux2b = hdiff3(U(i - 1, j + 1), U(i, j + 1), U(i + 1, j + 1));
                const float vx2a = hdiff3(V(i - 1, j), v00, V(i + 1, j));
                const float vx2b = hdiff3(V(i - 1, j + 1), V(i, j + 1), V(i + 1, j + 1));
                tu = 0.5f * (ux2a + ux2b);
                tv = 0.5f * (vx2a + vx2b);
            }
            const float uysq = uy1 * uy1 + tu * tu;
            const float vysq = vy1 * vy1 + tv * tv;
            const float ssq = uysq + vysq;
            const float ww = w[o] + w[o + g.S];
            v = smooth_weight(reg, ww, alpha_factor, ssq);
        }
    }
    ph[o] = h;
    pv[o] = v;
}

void launch_smoothness(cudaStream_t st, Geom g, const float *uu, const float *vv, const float *w, float alpha_factor,
                       Penalty reg, int mode, float *ph, float *pv) {
    dim3 b(32, 8), grid((g.S + 31) / 32, (g.H + 7) / 8);
    k_smoothness<<<grid, b, 0, st>>>(g, uu, vv, w, alpha_factor, reg, mode, ph, pv);
}

// ------------------------------------------------------------------------------------------ K2
// Output tile 32 x 32, 256 threads (32 x 8), 4 pixels per thread.  Shared memory per channel:
//   m  = 0.5*(im2w + im1), z = im2w - im1 on the tile + 4-pixel halo (clamped coordinates)
//   Ix = d/dx m on (tile + 2) and Iy = d/dy m on tile columns x (tile rows + 2)
// All three channels are staged at once; each thread then evaluates the second derivatives and
// the robust data term of one pixel at a time entirely in registers.
constexpr int DT_TW = 32, DT_TH = 32, DT_HALO = 4;
constexpr int DT_MW = DT_TW + 2 * DT_HALO; // 40 columns of m / z
constexpr int DT_MH = DT_TH + 2 * DT_HALO; // 40 rows
constexpr int DT_MP = DT_MW + 1;           // pitch (odd: conflict-free column walks)
constexpr int DT_XW = DT_TW + 4;           // Ix columns (x-2 .. x+2)
constexpr int DT_XH = DT_TH + 4;           // Ix / Iy rows (y-2 .. y+2)
constexpr int DT_XP = DT_XW + 1;
constexpr int DT_YP = DT_TW + 1;
constexpr size_t DT_SMEM_FLOATS = 3 * (2 * DT_MH * DT_MP + DT_XH * DT_XP + DT_XH * DT_YP);

struct DataTermArgs {
    const float *im1, *im2w, *mask, *du, *dv;
    float half_delta_over3, half_gamma_over3;
    const float *ph, *pv, *lap_u, *lap_v;
    float *a11, *a12, *a22, *b1, *b2;
};

template <bool FUSE_SYSTEM>
__global__ void __launch_bounds__(256) k_data_two_frame(Geom g, DataTermArgs a) {
    extern __shared__ float smem[];
    float *sm_m = smem;                           // [3][DT_MH][DT_MP]
    float *sm_z = sm_m + 3 * DT_MH * DT_MP;       // [3][DT_MH][DT_MP]
    float *sm_ix = sm_z + 3 * DT_MH * DT_MP;      // [3][DT_XH][DT_XP]
    float *sm_iy = sm_ix + 3 * DT_XH * DT_XP;     // [3][DT_XH][DT_YP]

    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int x0 = blockIdx.x * DT_TW, y0 = blockIdx.y * DT_TH;
    const int W1 = g.W - 1, H1 = g.H - 1;
    const size_t P = g.plane();

    // stage 1: mean and temporal difference on the haloed tile, coordinates clamped to the image
    for (int idx = tid; idx < 3 * DT_MH * DT_MW; idx += 256) {
        const int c = idx / (DT_MH * DT_MW);
        const int rem = idx - c * (DT_MH * DT_MW);
        const int ry = rem / DT_MW, rx = rem - ry * DT_MW;
        const int gx = clampi(x0 - DT_HALO + rx, 0, W1), gy = clampi(y0 - DT_HALO + ry, 0, H1);
        const size_t o = (size_t)c * P + (size_t)gy * g.S + gx;
        const float p1 = __ldg(a.im1 + o), p2 = __ldg(a.im2w + o);
        sm_m[(c * DT_MH + ry) * DT_MP + rx] = 0.5f * (p2 + p1);
        sm_z[(c * DT_MH + ry) * DT_MP + rx] = p2 - p1;
    }
    __syncthreads();

    // stage 2: first derivatives of the mean image.
    // Ix at tile-local (lx in [-2, TW+2), ly in [-2, TH+2)); columns outside the image take the value
    // at the clamped column (the reference convolves Ix with replicate borders, image.c:475-516).
    for (int idx = tid; idx < 3 * DT_XH * DT_XW; idx += 256) {
        const int c = idx / (DT_XH * DT_XW);
        const int rem = idx - c * (DT_XH * DT_XW);
        const int ry = rem / DT_XW, rx = rem - ry * DT_XW;
        const int gx = clampi(x0 - 2 + rx, 0, W1);         // clamped global column
        const int mx = gx - (x0 - DT_HALO);                // column inside sm_m
        const float *row = sm_m + (c * DT_MH + (ry + 2)) * DT_MP;
        sm_ix[(c * DT_XH + ry) * DT_XP + rx] = hconv5(row[mx - 2], row[mx - 1], row[mx], row[mx + 1], row[mx + 2]);
    }
    for (int idx = tid; idx < 3 * DT_XH * DT_TW; idx += 256) {
        const int c = idx / (DT_XH * DT_TW);
        const int rem = idx - c * (DT_XH * DT_TW);
        const int ry = rem / DT_TW, rx = rem - ry * DT_TW;
        const int gy = y0 - 2 + ry; // rows outside the image are never consumed (vconv5 folding)
        const float *col = sm_m + (c * DT_MH + (ry + 2)) * DT_MP + (rx + DT_HALO);
        sm_iy[(c * DT_XH + ry) * DT_YP + rx] =
            vconv5(col[-2 * DT_MP], col[-DT_MP], col[0], col[DT_MP], col[2 * DT_MP], gy, g.H);
    }
    __syncthreads();

    // stage 3: per pixel second derivatives + robust data term
    const float dnorm = 0.1f * 0.1f, eps_color = 0.001f * 0.001f, eps_grad = 0.001f * 0.001f;
    const int lx = threadIdx.x;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int ly = threadIdx.y + 8 * k;
        const int i = x0 + lx, j = y0 + ly;
        if (i >= g.S || j >= g.H) continue;
        const size_t o = (size_t)j * g.S + i;
        if (i >= g.W) { // padding columns: defined zeros (the reference leaves garbage there, Q1)
            a.a11[o] = 0.0f; a.a12[o] = 0.0f; a.a22[o] = 0.0f; a.b1[o] = 0.0f; a.b2[o] = 0.0f;
            continue;
        }
        const float u = a.du ? a.du[o] : 0.0f, v = a.dv ? a.dv[o] : 0.0f, m = a.mask[o];
        float ixx[3], ixy[3], iyy[3], ixz[3], iyz[3], ix[3], iy[3], iz[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float *pix = sm_ix + (c * DT_XH + (ly + 2)) * DT_XP + (lx + 2);
            const float *piy = sm_iy + (c * DT_XH + (ly + 2)) * DT_YP + lx;
            const float *pz = sm_z + (c * DT_MH + (ly + DT_HALO)) * DT_MP + (lx + DT_HALO);
            ix[c] = pix[0];
            iy[c] = piy[0];
            iz[c] = pz[0];
            ixx[c] = hconv5(pix[-2], pix[-1], pix[0], pix[1], pix[2]);
            ixy[c] = vconv5(pix[-2 * DT_XP], pix[-DT_XP], pix[0], pix[DT_XP], pix[2 * DT_XP], j, g.H);
            iyy[c] = vconv5(piy[-2 * DT_YP], piy[-DT_YP], piy[0], piy[DT_YP], piy[2 * DT_YP], j, g.H);
            ixz[c] = hconv5(pz[-2], pz[-1], pz[0], pz[1], pz[2]);
            iyz[c] = vconv5(pz[-2 * DT_MP], pz[-DT_MP], pz[0], pz[DT_MP], pz[2 * DT_MP], j, g.H);
        }
        float A11 = 0.0f, A12 = 0.0f, A22 = 0.0f, B1 = 0.0f, B2 = 0.0f;
        if (a.half_delta_over3 != 0.0f) { // colour constancy (variational_aux.c:242-265)
            float r[3], nn[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                r[c] = iz[c] + ix[c] * u + iy[c] * v;
                nn[c] = ix[c] * ix[c] + iy[c] * iy[c] + dnorm;
            }
            const float t = m * a.half_delta_over3 /
                            sqrtf(r[0] * r[0] / nn[0] + r[1] * r[1] / nn[1] + r[2] * r[2] / nn[2] + eps_color);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float gc = t / nn[c];
                A11 += gc * ix[c] * ix[c];
                A12 += gc * ix[c] * iy[c];
                A22 += gc * iy[c] * iy[c];
                B1 -= gc * iz[c] * ix[c];
                B2 -= gc * iz[c] * iy[c];
            }
        }
        { // gradient constancy (variational_aux.c:267-296)
            float rx[3], ry[3], nx[3], ny[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                nx[c] = ixx[c] * ixx[c] + ixy[c] * ixy[c] + dnorm;
                ny[c] = iyy[c] * iyy[c] + ixy[c] * ixy[c] + dnorm;
                rx[c] = ixz[c] + ixx[c] * u + ixy[c] * v;
                ry[c] = iyz[c] + ixy[c] * u + iyy[c] * v;
            }
            const float t = m * a.half_gamma_over3 /
                            sqrtf(rx[0] * rx[0] / nx[0] + ry[0] * ry[0] / ny[0] + rx[1] * rx[1] / nx[1] +
                                  ry[1] * ry[1] / ny[1] + rx[2] * rx[2] / nx[2] + ry[2] * ry[2] / ny[2] + eps_grad);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float gx = t / nx[c], gy = t / ny[c];
                A11 += gx * ixx[c] * ixx[c] + gy * ixy[c] * ixy[c];
                A12 += gx * ixx[c] * ixy[c] + gy * ixy[c] * iyy[c];
                A22 += gy * iyy[c] * iyy[c] + gx * ixy[c] * ixy[c];
                B1 -= gx * ixx[c] * ixz[c] + gy * ixy[c] * iyz[c];
                B2 -= gy * iyy[c] * iyz[c] + gx * ixy[c] * ixz[c];
            }
        }
        if (FUSE_SYSTEM) {
            // b += div(psi grad w): gather form of sub_laplacian with its accumulation order
            // (left edge, right edge, upper edge, lower edge; variational_aux.c:158-179)
            const float hl = (i > 0) ? a.ph[o - 1] : 0.0f, hr = a.ph[o];
            const float vt = (j > 0) ? a.pv[o - g.S] : 0.0f, vb = a.pv[o];
            const size_t ol = (i > 0) ? o - 1 : o, orr = (i < W1) ? o + 1 : o;
            const size_t ot = (j > 0) ? o - g.S : o, ob = (j < H1) ? o + g.S : o;
            {
                const float wc = a.lap_u[o];
                B1 -= hl * (wc - a.lap_u[ol]);
                B1 += hr * (a.lap_u[orr] - wc);
                B1 -= vt * (wc - a.lap_u[ot]);
                B1 += vb * (a.lap_u[ob] - wc);
            }
            {
                const float wc = a.lap_v[o];
                B2 -= hl * (wc - a.lap_v[ol]);
                B2 += hr * (a.lap_v[orr] - wc);
                B2 -= vt * (wc - a.lap_v[ot]);
                B2 += vb * (a.lap_v[ob] - wc);
            }
            // inverse of [[a11+sum psi, a12],[a12, a22+sum psi]] (solver.c:101-106)
            const float sp = ((hl + hr) + vt) + vb;
            const float D11 = A22 + sp, D22 = A11 + sp;
            const float det = D11 * D22 - A12 * A12;
            A11 = D11 / det;
            A22 = D22 / det;
            A12 = A12 / -det;
        }
        a.a11[o] = A11; a.a12[o] = A12; a.a22[o] = A22; a.b1[o] = B1; a.b2[o] = B2;
    }
}

void launch_data_two_frame(cudaStream_t st, Geom g, const float *im1, const float *im2w, const float *mask,
                           const float *du, const float *dv, float half_delta_over3, float half_gamma_over3,
                           bool fuse_system, const float *ph, const float *pv, const float *lap_u,
                           const float *lap_v, float *a11, float *a12, float *a22, float *b1, float *b2) {
    DataTermArgs a{im1, im2w, mask, du, dv, half_delta_over3, half_gamma_over3, ph, pv, lap_u, lap_v, a11, a12, a22, b1, b2};
    dim3 b(32, 8), grid((g.S + DT_TW - 1) / DT_TW, (g.H + DT_TH - 1) / DT_TH);
    const size_t smem = DT_SMEM_FLOATS * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_data_two_frame<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_data_two_frame<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    if (fuse_system) k_data_two_frame<true><<<grid, b, smem, st>>>(g, a);
    else k_data_two_frame<false><<<grid, b, smem, st>>>(g, a);
}

// ------------------------------------------------------------------------------------------ small operators
__global__ void __launch_bounds__(256) k_sub_laplacian(Geom g, float *__restrict__ dst, const float *__restrict__ src,
                                                       const float *__restrict__ ph, const float *__restrict__ pv) {
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

__global__ void __launch_bounds__(256) k_add4(size_t n4, float4 *__restrict__ dst, const float4 *__restrict__ a,
                                              const float4 *__restrict__ b) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (size_t)gridDim.x * blockDim.x) {
        const float4 x = a[k], y = b[k];
        dst[k] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
}
void launch_add(cudaStream_t st, Geom g, float *dst, const float *a, const float *b) {
    const size_t n4 = g.plane() / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    k_add4<<<blocks, 256, 0, st>>>(n4, reinterpret_cast<float4 *>(dst), reinterpret_cast<const float4 *>(a),
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
