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
