// sf_data.cu -- K2: image derivatives + robust data term in one fused pass (sm_100a).
//
// Replaces, per data term, the reference's chain
//     mean / temporal difference  (variational_aux.c:63-69 ; variational_mt.cpp:118-161)
//     7 separable 5-tap convolutions per colour image -> 24 derivative planes (variational_aux.c:71-77)
//     compute_data_and_match      (variational_aux.c:215-302)            kind DK_TWO_FRAME
//     add_data_and_match          (variational_aux_mt.cpp:166-403)       kind DK_MT_SUCC
//     add_data_and_match_ref      (variational_aux_mt.cpp:408-634)       kind DK_MT_REF
//   [+ sub_laplacian (variational_aux.c:153-180) and the 2x2 block inverse of sor_coupled's first
//      sweep (solver.c:101-106) when fuse_system is set on the last term]
// Nothing but the two input images, the mask and the 5 system planes touches HBM: the derivative
// planes live in shared memory / registers.
//
// CTA = 256 threads (32 x 8), output tile 32 x 32, 4 pixels per thread.
//   stage 1  m = 0.5*(B + A), z = +-(B - A) on the tile + 4 px halo   (float4 loads on interior tiles)
//   stage 2  Ix = d/dx m on tile+2, Iy = d/dy m on tile columns x (tile rows + 2)
//   stage 3  per pixel: Ixx Ixy Iyy Ixz Iyz (+ Ix Iy Iz), robust weights, 2x2 system contribution
// Border semantics follow image.c:400-526 (clamped columns, folded vertical taps); tiles that do not
// touch the image border take a branch-free path.
#include "sf_internal.cuh"
#include "sf_pack.cuh"
#include "sf_penalty.cuh"
#include "sf_stencil.cuh"

namespace sf {

constexpr int DT_TW = 32, DT_TH = 32, DT_HALO = 4;
constexpr int DT_MW = DT_TW + 2 * DT_HALO; // 40 columns of m / z (multiple of 4: float4 staging)
constexpr int DT_MH = DT_TH + 2 * DT_HALO; // 40 rows
constexpr int DT_XW = DT_TW + 4;           // Ix columns (x-2 .. x+2)
constexpr int DT_XH = DT_TH + 4;           // Ix / Iy rows (y-2 .. y+2)
constexpr size_t DT_SMEM_FLOATS = 3 * (2 * DT_MH * DT_MW + DT_XH * DT_XW + DT_XH * DT_TW);

// 1/x: MUFU.RCP (1 ulp) + one Newton step (~0.5 ulp); the reference divides (divps).  4 instructions instead of the ~10 of
// an IEEE-rounded reciprocal -- the term arithmetic, not the stencils, is the bulk of the multi-frame data pass.
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

struct Derivs {
    float ix[3], iy[3], iz[3], ixx[3], ixy[3], iyy[3], ixz[3], iyz[3];
};
struct Acc {
    float a11, a12, a22, b1, b2;
};

// ---- two-frame data term (variational_aux.c:241-296).  Quotients x/n are evaluated as x * rcp(n) with a
// correctly rounded reciprocal and 1/sqrt through rsqrtf (<= 2 ulp from the reference's divps / sqrtps).
__device__ __forceinline__ void term_two_frame(const Derivs &d, float u, float v, float m, float hd, float hg, Acc &acc) {
    const float dnorm = 0.1f * 0.1f, eps_color = 0.001f * 0.001f, eps_grad = 0.001f * 0.001f;
    if (hd != 0.0f) {
        float r[3], inv[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            r[c] = d.iz[c] + d.ix[c] * u + d.iy[c] * v;
            inv[c] = fast_rcp(d.ix[c] * d.ix[c] + d.iy[c] * d.iy[c] + dnorm);
        }
        const float t = m * hd * rsqrtf(r[0] * r[0] * inv[0] + r[1] * r[1] * inv[1] + r[2] * r[2] * inv[2] + eps_color);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float gc = t * inv[c];
            acc.a11 += gc * d.ix[c] * d.ix[c];
            acc.a12 += gc * d.ix[c] * d.iy[c];
            acc.a22 += gc * d.iy[c] * d.iy[c];
            acc.b1 -= gc * d.iz[c] * d.ix[c];
            acc.b2 -= gc * d.iz[c] * d.iy[c];
        }
    }
    float rx[3], ry[3], ivx[3], ivy[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        ivx[c] = fast_rcp(d.ixx[c] * d.ixx[c] + d.ixy[c] * d.ixy[c] + dnorm);
        ivy[c] = fast_rcp(d.iyy[c] * d.iyy[c] + d.ixy[c] * d.ixy[c] + dnorm);
        rx[c] = d.ixz[c] + d.ixx[c] * u + d.ixy[c] * v;
        ry[c] = d.iyz[c] + d.ixy[c] * u + d.iyy[c] * v;
    }
    const float t = m * hg * rsqrtf(rx[0] * rx[0] * ivx[0] + ry[0] * ry[0] * ivy[0] + rx[1] * rx[1] * ivx[1] +
                                    ry[1] * ry[1] * ivy[1] + rx[2] * rx[2] * ivx[2] + ry[2] * ry[2] * ivy[2] + eps_grad);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float gx = t * ivx[c], gy = t * ivy[c];
        acc.a11 += gx * d.ixx[c] * d.ixx[c] + gy * d.ixy[c] * d.ixy[c];
        acc.a12 += gx * d.ixx[c] * d.ixy[c] + gy * d.ixy[c] * d.iyy[c];
        acc.a22 += gy * d.iyy[c] * d.iyy[c] + gx * d.ixy[c] * d.ixy[c];
        acc.b1 -= gx * d.ixx[c] * d.ixz[c] + gy * d.ixy[c] * d.iyz[c];
        acc.b2 -= gy * d.iyy[c] * d.iyz[c] + gx * d.ixy[c] * d.ixz[c];
    }
}

// ---- multi-frame successive term (variational_aux_mt.cpp:186-363): warped frames s and s+1, effective
// gradient s*I - (s+1)*I evaluated exactly as written; wc = channel weights; psi' pluggable.
template <int PC = -1, int PG = -1>
__device__ __forceinline__ void term_mt_succ(const Derivs &d, float u, float v, float m, float wd, float wg, float s,
                                             const float wc[3], int dt_norm, const Penalty &pc, const Penalty &pg, Acc &acc) {
    const float dnorm = 0.1f * 0.1f;
    const float f = s, f1 = s + 1.0f;
    const float g = f - f1;
    if (wd != 0.0f) {
        // The reference writes the effective gradient of the pair as s*I - (s+1)*I and expands the residual term by term
        // (variational_aux_mt.cpp:196-215); s - (s+1) is exactly -1 in float for every time factor the path uses, so the
        // gradient is g*I with g = f - f1 and the residual folds to Iz + gx*u + gy*v (same value up to the rounding order).
        float r[3], gx[3], gy[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            gx[c] = g * d.ix[c];
            gy[c] = g * d.iy[c];
            r[c] = wc[c] * (d.iz[c] + gx[c] * u + gy[c] * v);
        }
        if (!dt_norm) {
            const float t = m * wd * penalty_deriv_vt<PC>(pc, r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float g = t * wc[c];
                acc.a11 += g * gx[c] * gx[c];
                acc.a12 += g * gx[c] * gy[c];
                acc.a22 += g * gy[c] * gy[c];
                acc.b1 -= g * d.iz[c] * gx[c];
                acc.b2 -= g * d.iz[c] * gy[c];
            }
        } else {
            float inv[3];
#pragma unroll
            for (int c = 0; c < 3; c++) inv[c] = fast_rcp(gx[c] * gx[c] + gy[c] * gy[c] + dnorm);
            const float t = m * wd * penalty_deriv_vt<PC>(pc, r[0] * r[0] * inv[0] + r[1] * r[1] * inv[1] + r[2] * r[2] * inv[2]);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float g = t * inv[c] * wc[c];
                acc.a11 += g * gx[c] * gx[c];
                acc.a12 += g * gx[c] * gy[c];
                acc.a22 += g * gy[c] * gy[c];
                acc.b1 -= g * d.iz[c] * gx[c];
                acc.b2 -= g * d.iz[c] * gy[c];
            }
        }
    }
    float rx[3], ry[3], gxx[3], gyy[3], gxy[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        gxx[c] = g * d.ixx[c];
        gyy[c] = g * d.iyy[c];
        gxy[c] = g * d.ixy[c];
        rx[c] = wc[c] * (d.ixz[c] + gxx[c] * u + gxy[c] * v);
        ry[c] = wc[c] * (d.iyz[c] + gxy[c] * u + gyy[c] * v);
    }
    if (!dt_norm) {
        const float t = m * wg * penalty_deriv_vt<PG>(pg, rx[0] * rx[0] + ry[0] * ry[0] + rx[1] * rx[1] + ry[1] * ry[1] +
                                                         rx[2] * rx[2] + ry[2] * ry[2]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float g = t * wc[c];
            acc.a11 += g * gxx[c] * gxx[c] + g * gxy[c] * gxy[c];
            acc.a12 += g * gxx[c] * gxy[c] + g * gxy[c] * gyy[c];
            acc.a22 += g * gyy[c] * gyy[c] + g * gxy[c] * gxy[c];
            acc.b1 -= g * d.ixz[c] * gxx[c] + g * d.iyz[c] * gxy[c];
            acc.b2 -= g * d.iyz[c] * gyy[c] + g * d.ixz[c] * gxy[c];
        }
    } else {
        float ivx[3], ivy[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            ivx[c] = fast_rcp(gxx[c] * gxx[c] + gxy[c] * gxy[c] + dnorm);
            ivy[c] = fast_rcp(gyy[c] * gyy[c] + gxy[c] * gxy[c] + dnorm);
        }
        const float t = m * wg * penalty_deriv_vt<PG>(pg, rx[0] * rx[0] * ivx[0] + ry[0] * ry[0] * ivy[0] + rx[1] * rx[1] * ivx[1] +
                                                         ry[1] * ry[1] * ivy[1] + rx[2] * rx[2] * ivx[2] + ry[2] * ry[2] * ivy[2]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float g1 = t * ivx[c] * wc[c], g2 = t * ivy[c] * wc[c];
            acc.a11 += g1 * gxx[c] * gxx[c] + g2 * gxy[c] * gxy[c];
            acc.a12 += g1 * gxx[c] * gxy[c] + g2 * gxy[c] * gyy[c];
            acc.a22 += g2 * gyy[c] * gyy[c] + g1 * gxy[c] * gxy[c];
            acc.b1 -= g1 * d.ixz[c] * gxx[c] + g2 * d.iyz[c] * gxy[c];
            acc.b2 -= g2 * d.iyz[c] * gyy[c] + g1 * d.ixz[c] * gxy[c];
        }
    }
}

// ---- multi-frame reference term (variational_aux_mt.cpp:416-592): frame vs. unwarped reference frame with
// time factor s (sign flipped for s >= 0, :424-425).  The un-normalised branch reproduces the reference
// literally, including its copy-paste slips (:458-471 channel 3, :527-530 channel 1).
template <int PC = -1, int PG = -1>
__device__ __forceinline__ void term_mt_ref(const Derivs &d, float u, float v, float m, float wd, float wg, float s,
                                            const float wc[3], int dt_norm, const Penalty &pc, const Penalty &pg, Acc &acc) {
    const float dnorm = 0.1f * 0.1f;
    const float fsq = s * s;
    const float f = (s >= 0.0f) ? -s : s;
    if (wd != 0.0f) {
        float r[3];
#pragma unroll
        for (int c = 0; c < 3; c++) r[c] = wc[c] * (d.iz[c] + d.ix[c] * f * u + d.iy[c] * f * v);
        if (!dt_norm) {
            float t = m * wd * penalty_deriv_vt<PC>(pc, r[0] * r[0] / fsq + r[1] * r[1] / fsq + r[2] * r[2] / fsq);
            t /= fsq;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float g = (c == 0) ? t * wc[0] * f : t * f * wc[c];
                acc.b1 -= g * d.iz[c] * d.ix[c];
                acc.b2 -= g * d.iz[c] * d.iy[c];
                g = (c == 2) ? t * f : g * f; // channel 3 drops the weight and one factor (as written, :469)
                acc.a11 += g * d.ix[c] * d.ix[c];
                acc.a12 += g * d.ix[c] * d.iy[c];
                acc.a22 += g * d.iy[c] * d.iy[c];
            }
        } else {
            float inv[3];
#pragma unroll
            for (int c = 0; c < 3; c++) inv[c] = fast_rcp(fsq * d.ix[c] * d.ix[c] + fsq * d.iy[c] * d.iy[c] + dnorm);
            const float t = m * wd * penalty_deriv_vt<PC>(pc, r[0] * r[0] * inv[0] + r[1] * r[1] * inv[1] + r[2] * r[2] * inv[2]);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float g = t * inv[c] * wc[c] * f;
                acc.b1 -= g * d.iz[c] * d.ix[c];
                acc.b2 -= g * d.iz[c] * d.iy[c];
                g = g * f;
                acc.a11 += g * d.ix[c] * d.ix[c];
                acc.a12 += g * d.ix[c] * d.iy[c];
                acc.a22 += g * d.iy[c] * d.iy[c];
            }
        }
    }
    float rx[3], ry[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        rx[c] = wc[c] * (d.ixz[c] + d.ixx[c] * f * u + d.ixy[c] * f * v);
        ry[c] = wc[c] * (d.iyz[c] + d.ixy[c] * f * u + d.iyy[c] * f * v);
    }
    if (!dt_norm) {
        float t = m * wg * penalty_deriv_vt<PG>(pg, rx[0] * rx[0] / fsq + ry[0] * ry[0] / fsq + rx[1] * rx[1] / fsq +
                                                   ry[1] * ry[1] / fsq + rx[2] * rx[2] / fsq + ry[2] * ry[2] / fsq);
        t /= fsq;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float g = t * wc[c] * f;
            acc.b1 -= g * d.ixx[c] * d.ixz[c] + g * d.ixy[c] * d.iyz[c];
            acc.b2 -= g * d.iyy[c] * d.iyz[c] + g * d.ixy[c] * d.ixz[c];
            g = g * f;
            if (c == 0) g = g * fsq; // channel 1 carries an extra factorsq (as written, :528-530)
            acc.a11 += g * d.ixx[c] * d.ixx[c] + g * d.ixy[c] * d.ixy[c];
            acc.a12 += g * d.ixx[c] * d.ixy[c] + g * d.ixy[c] * d.iyy[c];
            acc.a22 += g * d.iyy[c] * d.iyy[c] + g * d.ixy[c] * d.ixy[c];
        }
    } else {
        float ivx[3], ivy[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            ivx[c] = fast_rcp(fsq * d.ixx[c] * d.ixx[c] + fsq * d.ixy[c] * d.ixy[c] + dnorm);
            ivy[c] = fast_rcp(fsq * d.iyy[c] * d.iyy[c] + fsq * d.ixy[c] * d.ixy[c] + dnorm);
        }
        const float t = m * wg * penalty_deriv_vt<PG>(pg, rx[0] * rx[0] * ivx[0] + ry[0] * ry[0] * ivy[0] + rx[1] * rx[1] * ivx[1] +
                                                         ry[1] * ry[1] * ivy[1] + rx[2] * rx[2] * ivx[2] + ry[2] * ry[2] * ivy[2]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float g1 = t * ivx[c] * wc[c] * f, g2 = t * ivy[c] * wc[c] * f;
            acc.b1 -= g1 * d.ixx[c] * d.ixz[c] + g2 * d.ixy[c] * d.iyz[c];
            acc.b2 -= g2 * d.iyy[c] * d.iyz[c] + g1 * d.ixy[c] * d.ixz[c];
            g1 = g1 * f;
            g2 = g2 * f;
            acc.a11 += g1 * d.ixx[c] * d.ixx[c] + g2 * d.ixy[c] * d.ixy[c];
            acc.a12 += g1 * d.ixx[c] * d.ixy[c] + g2 * d.ixy[c] * d.iyy[c];
            acc.a22 += g2 * d.iyy[c] * d.iyy[c] + g1 * d.ixy[c] * d.ixy[c];
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) k_data_term(Geom g, DataTermDesc t, DataCommon cm) {
    pdl_enter();
    if (g.cancelled()) return;
    extern __shared__ float smem[];
    float *sm_m = smem;                       // [3][DT_MH][DT_MW]
    float *sm_z = sm_m + 3 * DT_MH * DT_MW;   // [3][DT_MH][DT_MW]
    float *sm_ix = sm_z + 3 * DT_MH * DT_MW;  // [3][DT_XH][DT_XW]
    float *sm_iy = sm_ix + 3 * DT_XH * DT_XW; // [3][DT_XH][DT_TW]

    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int x0 = blockIdx.x * DT_TW, y0 = blockIdx.y * DT_TH;
    const int W1 = g.W - 1, H1 = g.H - 1;
    const size_t P = g.plane();
    // tile (with its halo) strictly inside the image: no clamping, no folded vertical taps anywhere
    const bool interior = (x0 - DT_HALO >= 0) && (x0 + DT_TW + DT_HALO <= g.W) && (y0 - DT_HALO >= 0) &&
                          (y0 + DT_TH + DT_HALO <= g.H);
    const float zs = (t.zsign > 0) ? 1.0f : -1.0f;

    // ---- stage 1 (interior tiles: 2-D thread mapping, no per-element index division: lane = channel * 10 + float4
    // column, warp ty takes rows ty, ty+8, ...)
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (interior) {
        constexpr int Q = DT_MW / 4; // float4 per row
        if (tx < 3 * Q) {
            const int c = tx / Q, q = tx - c * Q;
            const size_t base = (size_t)c * P + (size_t)(y0 - DT_HALO) * g.S + (x0 - DT_HALO + 4 * q);
#pragma unroll
            for (int ry = ty; ry < DT_MH; ry += 8) {
                const size_t o = base + (size_t)ry * g.S;
                const float4 a = __ldg(reinterpret_cast<const float4 *>(t.A + o));
                const float4 b = __ldg(reinterpret_cast<const float4 *>(t.B + o));
                const int so = (c * DT_MH + ry) * DT_MW + 4 * q;
                *reinterpret_cast<float4 *>(sm_m + so) =
                    make_float4(0.5f * (b.x + a.x), 0.5f * (b.y + a.y), 0.5f * (b.z + a.z), 0.5f * (b.w + a.w));
                *reinterpret_cast<float4 *>(sm_z + so) =
                    make_float4(zs * (b.x - a.x), zs * (b.y - a.y), zs * (b.z - a.z), zs * (b.w - a.w));
            }
        }
    } else {
        for (int idx = tid; idx < 3 * DT_MH * DT_MW; idx += 256) {
            const int c = idx / (DT_MH * DT_MW);
            const int rem = idx - c * (DT_MH * DT_MW);
            const int ry = rem / DT_MW, rx = rem - ry * DT_MW;
            const int gx = clampi(x0 - DT_HALO + rx, 0, W1), gy = clampi(y0 - DT_HALO + ry, 0, H1);
            const size_t o = (size_t)c * P + (size_t)gy * g.S + gx;
            const float a = __ldg(t.A + o), b = __ldg(t.B + o);
            sm_m[(c * DT_MH + ry) * DT_MW + rx] = 0.5f * (b + a);
            sm_z[(c * DT_MH + ry) * DT_MW + rx] = zs * (b - a);
        }
    }
    __syncthreads();

    // ---- stage 2: Ix on (tile+2)^2 -- columns outside the image take the value at the clamped column
    // (the reference convolves Ix itself with replicate borders, image.c:475-516) -- and Iy on tile columns.
    // Columns 0..31 of a row go to the 32 lanes of a warp, rows ty, ty+8, ...; the 4 remaining Ix columns of 8 rows
    // make one more warp-wide pass per channel.
    auto ix_at = [&](int c, int ry, int rx) {
        const int gx = interior ? (x0 - 2 + rx) : clampi(x0 - 2 + rx, 0, W1);
        const int mx = gx - (x0 - DT_HALO);
        const float *row = sm_m + (c * DT_MH + (ry + 2)) * DT_MW;
        sm_ix[(c * DT_XH + ry) * DT_XW + rx] = hconv5(row[mx - 2], row[mx - 1], row[mx], row[mx + 1], row[mx + 2]);
    };
#pragma unroll
    for (int c = 0; c < 3; c++) {
        for (int ry = ty; ry < DT_XH; ry += 8) {
            ix_at(c, ry, tx);
            const float *col = sm_m + (c * DT_MH + (ry + 2)) * DT_MW + (tx + DT_HALO);
            sm_iy[(c * DT_XH + ry) * DT_TW + tx] =
                interior ? hconv5(col[-2 * DT_MW], col[-DT_MW], col[0], col[DT_MW], col[2 * DT_MW])
                         : vconv5(col[-2 * DT_MW], col[-DT_MW], col[0], col[DT_MW], col[2 * DT_MW], y0 - 2 + ry, g.H);
        }
        // Ix columns 32..35: lane -> (row within a group of 8, column); warp ty takes row groups ty, ty+8, ...
        for (int ry = ty * 8 + (tx >> 2); ry < DT_XH; ry += 64) ix_at(c, ry, DT_TW + (tx & 3));
    }
    __syncthreads();

    // ---- stage 3
    const int lx = tx;
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
        const int ly = threadIdx.y + 8 * k;
        const int i = x0 + lx, j = y0 + ly;
        if (i >= g.S || j >= g.H) continue;
        const size_t o = (size_t)j * g.S + i;
        if (i >= g.W) { // padding columns: defined zeros (the reference leaves garbage there)
            if (KIND == DK_DERIVS) {
#pragma unroll
                for (int k = 0; k < 15; k++) cm.a11[(size_t)k * P + o] = 0.0f;
            } else {
                cm.a11[o] = 0.0f; cm.a12[o] = 0.0f; cm.a22[o] = 0.0f; cm.b1[o] = 0.0f; cm.b2[o] = 0.0f;
            }
            continue;
        }
        Derivs d;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float *pix = sm_ix + (c * DT_XH + (ly + 2)) * DT_XW + (lx + 2);
            const float *piy = sm_iy + (c * DT_XH + (ly + 2)) * DT_TW + lx;
            const float *pz = sm_z + (c * DT_MH + (ly + DT_HALO)) * DT_MW + (lx + DT_HALO);
            d.ix[c] = pix[0];
            d.iy[c] = piy[0];
            d.iz[c] = pz[0];
            d.ixx[c] = hconv5(pix[-2], pix[-1], pix[0], pix[1], pix[2]);
            d.ixz[c] = hconv5(pz[-2], pz[-1], pz[0], pz[1], pz[2]);
            if (interior) {
                d.ixy[c] = hconv5(pix[-2 * DT_XW], pix[-DT_XW], pix[0], pix[DT_XW], pix[2 * DT_XW]);
                d.iyy[c] = hconv5(piy[-2 * DT_TW], piy[-DT_TW], piy[0], piy[DT_TW], piy[2 * DT_TW]);
                d.iyz[c] = hconv5(pz[-2 * DT_MW], pz[-DT_MW], pz[0], pz[DT_MW], pz[2 * DT_MW]);
            } else {
                d.ixy[c] = vconv5(pix[-2 * DT_XW], pix[-DT_XW], pix[0], pix[DT_XW], pix[2 * DT_XW], j, g.H);
                d.iyy[c] = vconv5(piy[-2 * DT_TW], piy[-DT_TW], piy[0], piy[DT_TW], piy[2 * DT_TW], j, g.H);
                d.iyz[c] = vconv5(pz[-2 * DT_MW], pz[-DT_MW], pz[0], pz[DT_MW], pz[2 * DT_MW], j, g.H);
            }
        }
        if (KIND == DK_DERIVS) { // per-frame derivative planes [Ix Iy Ixx Ixy Iyy][channel] at cm.a11 (A == B: m is the frame)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                cm.a11[(size_t)(0 + c) * P + o] = d.ix[c];
                cm.a11[(size_t)(3 + c) * P + o] = d.iy[c];
                cm.a11[(size_t)(6 + c) * P + o] = d.ixx[c];
                cm.a11[(size_t)(9 + c) * P + o] = d.ixy[c];
                cm.a11[(size_t)(12 + c) * P + o] = d.iyy[c];
            }
            continue;
        }
        const float u = cm.du ? cm.du[o] : 0.0f, v = cm.dv ? cm.dv[o] : 0.0f;
        float m = t.mask[o];
        Acc acc;
        if (cm.accumulate) {
            acc.a11 = cm.a11[o]; acc.a12 = cm.a12[o]; acc.a22 = cm.a22[o]; acc.b1 = cm.b1[o]; acc.b2 = cm.b2[o];
        } else {
            acc.a11 = acc.a12 = acc.a22 = acc.b1 = acc.b2 = 0.0f;
        }
        if (KIND == DK_TWO_FRAME) {
            term_two_frame(d, u, v, m, t.wd, t.wg, acc);
        } else {
            // occlusion / window factor of variational_mt.cpp:293-320, applied on the fly to the raw mask
            if (t.dir >= 0 && cm.occ) {
                const float oc = cm.occ[o];
                const float fac = (1.0f + ((oc == 0.0f) ? 1.0f : 0.0f)) * cm.data_norm;
                const float sel = (t.dir == 0) ? ((oc >= 0.0f) ? 1.0f : 0.0f) : ((oc <= 0.0f) ? 1.0f : 0.0f);
                m = (1.0f * (sel / fac)) * m;
            }
            float wc[3] = {1.0f, 1.0f, 1.0f};
            if (cm.chw) {
                wc[0] = cm.chw[o]; wc[1] = cm.chw[o + cm.chw_pstride]; wc[2] = cm.chw[o + 2 * cm.chw_pstride];
            }
            if (KIND == DK_MT_SUCC) term_mt_succ(d, u, v, m, t.wd, t.wg, t.s, wc, cm.dt_norm, cm.pc, cm.pg, acc);
            else term_mt_ref(d, u, v, m, t.wd, t.wg, t.s, wc, cm.dt_norm, cm.pc, cm.pg, acc);
        }
        if (cm.fuse_system) {
            // b += div(psi grad w): gather form of sub_laplacian with its accumulation order (left edge, right
            // edge, upper edge, lower edge; variational_aux.c:158-179)
            const float hl = (i > 0) ? cm.ph[o - 1] : 0.0f, hr = cm.ph[o];
            const float vt = (j > 0) ? cm.pv[o - g.S] : 0.0f, vb = cm.pv[o];
            const size_t ol = (i > 0) ? o - 1 : o, orr = (i < W1) ? o + 1 : o;
            const size_t ot = (j > 0) ? o - g.S : o, ob = (j < H1) ? o + g.S : o;
            {
                const float wcn = cm.lap_u[o];
                acc.b1 -= hl * (wcn - cm.lap_u[ol]);
                acc.b1 += hr * (cm.lap_u[orr] - wcn);
                acc.b1 -= vt * (wcn - cm.lap_u[ot]);
                acc.b1 += vb * (cm.lap_u[ob] - wcn);
            }
            {
                const float wcn = cm.lap_v[o];
                acc.b2 -= hl * (wcn - cm.lap_v[ol]);
                acc.b2 += hr * (cm.lap_v[orr] - wcn);
                acc.b2 -= vt * (wcn - cm.lap_v[ot]);
                acc.b2 += vb * (cm.lap_v[ob] - wcn);
            }
            // inverse of [[a11 + sum psi, a12], [a12, a22 + sum psi]] (solver.c:101-106)
            const float sp = ((hl + hr) + vt) + vb;
            const float D11 = acc.a22 + sp, D22 = acc.a11 + sp;
            const float det = D11 * D22 - acc.a12 * acc.a12;
            acc.a11 = D11 / det;
            acc.a22 = D22 / det;
            acc.a12 = acc.a12 / -det;
        }
        cm.a11[o] = acc.a11; cm.a12[o] = acc.a12; cm.a22[o] = acc.a22; cm.b1[o] = acc.b1; cm.b2[o] = acc.b2;
    }
}

// ---- all multi-frame terms of one outer / inner iteration in one pointwise pass (see sf_internal.cuh)
#ifndef SF_MT_TERMS_MINB
#define SF_MT_TERMS_MINB 4 // resident CTAs per SM the register allocation aims at: 64 registers, 32 warps per SM (2: 101 registers, 12.5 ms per config-3 window; 3: 11.9; 4: 11.8; 5: 12.2)
#endif
// PC, PG: the colour / gradient penalties as compile-time functor ids (select_robust_function,
// variational_aux_mt.cpp:889-926): one instantiation per pair instead of a switch per evaluation
template <int PC, int PG>
__global__ void __launch_bounds__(256, SF_MT_TERMS_MINB) k_mt_terms(Geom g, MtTermsArgs ta, DataCommon cm) {
    pdl_enter();
    if (g.cancelled()) return;
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    if (i >= g.S || j >= g.H) return;
    const size_t P = g.plane(), o = (size_t)j * g.S + i;
    if (i >= g.W) {
        cm.a11[o] = 0.0f; cm.a12[o] = 0.0f; cm.a22[o] = 0.0f; cm.b1[o] = 0.0f; cm.b2[o] = 0.0f;
        return;
    }
    const int W1 = g.W - 1, H1 = g.H - 1;
    const float u = cm.du ? cm.du[o] : 0.0f, v = cm.dv ? cm.dv[o] : 0.0f;
    float wc[3] = {1.0f, 1.0f, 1.0f};
    if (cm.chw) {
        wc[0] = cm.chw[o]; wc[1] = cm.chw[o + cm.chw_pstride]; wc[2] = cm.chw[o + 2 * cm.chw_pstride];
    }
    const float oc = cm.occ ? cm.occ[o] : 0.0f;
    Acc acc;
    acc.a11 = acc.a12 = acc.a22 = acc.b1 = acc.b2 = 0.0f;
#pragma unroll 1
    for (int k = 0; k < ta.nterms; k++) {
        const MtTerm t = ta.term[k];
        const float *IA = ta.I[t.fa] + o, *IB = ta.I[t.fb] + o, *DA = ta.D[t.fa] + o, *DB = ta.D[t.fb] + o;
        Derivs d;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            // the reference forms m = (A + B)/2 and z = A - B first and differentiates those (variational_mt.cpp:118-161);
            // the filters are linear, so the same values (up to rounding) come from the frames' own derivatives
            d.iz[c] = __ldg(IA + c * P) - __ldg(IB + c * P);
            const float xa = __ldg(DA + (0 + c) * P), xb = __ldg(DB + (0 + c) * P);
            const float ya = __ldg(DA + (3 + c) * P), yb = __ldg(DB + (3 + c) * P);
            d.ix[c] = 0.5f * (xb + xa);
            d.iy[c] = 0.5f * (yb + ya);
            d.ixz[c] = xa - xb;
            d.iyz[c] = ya - yb;
            d.ixx[c] = 0.5f * (__ldg(DB + (6 + c) * P) + __ldg(DA + (6 + c) * P));
            d.ixy[c] = 0.5f * (__ldg(DB + (9 + c) * P) + __ldg(DA + (9 + c) * P));
            d.iyy[c] = 0.5f * (__ldg(DB + (12 + c) * P) + __ldg(DA + (12 + c) * P));
        }
        float m = __ldg(ta.mask[t.mask_frame] + o);
        if (cm.occ) { // occlusion / window factor of variational_mt.cpp:293-320, applied on the fly to the raw mask
            const float fac = (1.0f + ((oc == 0.0f) ? 1.0f : 0.0f)) * cm.data_norm;
            const float sel = (t.dir == 0) ? ((oc >= 0.0f) ? 1.0f : 0.0f) : ((oc <= 0.0f) ? 1.0f : 0.0f);
            m = (1.0f * (sel / fac)) * m;
        }
        if (t.kind == DK_MT_SUCC) term_mt_succ<PC, PG>(d, u, v, m, t.wd, t.wg, t.s, wc, cm.dt_norm, cm.pc, cm.pg, acc);
        else term_mt_ref<PC, PG>(d, u, v, m, t.wd, t.wg, t.s, wc, cm.dt_norm, cm.pc, cm.pg, acc);
    }
    {
        // b += div(psi grad w) and the 2x2 block inverse, as in k_data_term's fuse_system (variational_aux.c:158-179,
        // solver.c:101-106)
        const float hl = (i > 0) ? cm.ph[o - 1] : 0.0f, hr = cm.ph[o];
        const float vt = (j > 0) ? cm.pv[o - g.S] : 0.0f, vb = cm.pv[o];
        const size_t ol = (i > 0) ? o - 1 : o, orr = (i < W1) ? o + 1 : o;
        const size_t ot = (j > 0) ? o - g.S : o, ob = (j < H1) ? o + g.S : o;
        {
            const float wcn = cm.lap_u[o];
            acc.b1 -= hl * (wcn - cm.lap_u[ol]);
            acc.b1 += hr * (cm.lap_u[orr] - wcn);
            acc.b1 -= vt * (wcn - cm.lap_u[ot]);
            acc.b1 += vb * (cm.lap_u[ob] - wcn);
        }
        {
            const float wcn = cm.lap_v[o];
            acc.b2 -= hl * (wcn - cm.lap_v[ol]);
            acc.b2 += hr * (cm.lap_v[orr] - wcn);
            acc.b2 -= vt * (wcn - cm.lap_v[ot]);
            acc.b2 += vb * (cm.lap_v[ob] - wcn);
        }
        const float sp = ((hl + hr) + vt) + vb;
        const float D11 = acc.a22 + sp, D22 = acc.a11 + sp;
        const float det = D11 * D22 - acc.a12 * acc.a12;
        acc.a11 = D11 / det;
        acc.a22 = D22 / det;
        acc.a12 = acc.a12 / -det;
    }
    cm.a11[o] = acc.a11; cm.a12[o] = acc.a12; cm.a22[o] = acc.a22; cm.b1[o] = acc.b1; cm.b2[o] = acc.b2;
}

// ---- the same pass with TWO columns per thread on the packed-fp32 pipe (sf_pack.cuh).  k_mt_terms is instruction bound
// (ncu: ~3200 thread instructions per pixel for the six terms of config 3, issue slots 59 % busy at 33 % of the DRAM
// peak): with a (column 2l, column 2l+1) pair in every 64-bit register the loads become LDG.64 and the term arithmetic
// FFMA2 / FMUL2, i.e. half the instructions per pixel.  The terms below are term_mt_succ / term_mt_ref with the products
// grouped so that every update of A,b is one packed fma (h = g * I_a once, then fma(h, I_b, acc)); the results agree with
// the scalar forms -- which stay the operator twins (k_data_term<DK_MT_*>) -- to rounding.
struct Derivs2 {
    p64 ix[3], iy[3], iz[3], ixx[3], ixy[3], iyy[3], ixz[3], iyz[3];
};
struct Acc2 {
    p64 a11, a12, a22, b1, b2;
};

// colour part shared by both terms: acc += g * (gx, gy) (x) (gx, gy), b -= g * iz * (gx, gy)
__device__ __forceinline__ void acc_color2(Acc2 &acc, p64 ga, p64 gb, p64 gx, p64 gy, p64 niz) {
    const p64 kx = mul2(ga, gx), ky = mul2(ga, gy);
    acc.a11 = fma2(kx, gx, acc.a11);
    acc.a12 = fma2(kx, gy, acc.a12);
    acc.a22 = fma2(ky, gy, acc.a22);
    acc.b1 = fma2(mul2(gb, gx), niz, acc.b1);
    acc.b2 = fma2(mul2(gb, gy), niz, acc.b2);
}
// gradient part: ga1/ga2 weight the A updates of the x / y residual, gb1/gb2 the b updates
__device__ __forceinline__ void acc_grad2(Acc2 &acc, p64 ga1, p64 ga2, p64 gb1, p64 gb2, p64 gxx, p64 gxy, p64 gyy, p64 nixz,
                                          p64 niyz) {
    const p64 k1x = mul2(ga1, gxx), k1y = mul2(ga1, gxy), k2x = mul2(ga2, gxy), k2y = mul2(ga2, gyy);
    acc.a11 = fma2(k1x, gxx, fma2(k2x, gxy, acc.a11));
    acc.a12 = fma2(k1x, gxy, fma2(k2x, gyy, acc.a12));
    acc.a22 = fma2(k2y, gyy, fma2(k1y, gxy, acc.a22));
    acc.b1 = fma2(mul2(gb1, gxx), nixz, fma2(mul2(gb2, gxy), niyz, acc.b1));
    acc.b2 = fma2(mul2(gb2, gyy), niyz, fma2(mul2(gb1, gxy), nixz, acc.b2));
}

template <int PC, int PG>
__device__ __forceinline__ void term_mt_succ2(const Derivs2 &d, p64 u, p64 v, p64 m, float wd, float wg, float s, const p64 wc[3],
                                              int dt_norm, const Penalty &pc, const Penalty &pg, Acc2 &acc) {
    const p64 dnorm = splat2(0.1f * 0.1f), mone = splat2(-1.0f);
    const float f = s, f1 = s + 1.0f;
    const p64 g = splat2(f - f1); // -1 for every time factor of the path (see term_mt_succ)
    if (wd != 0.0f) {
        p64 r[3], gx[3], gy[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            gx[c] = mul2(g, d.ix[c]);
            gy[c] = mul2(g, d.iy[c]);
            r[c] = mul2(wc[c], fma2(gy[c], v, fma2(gx[c], u, d.iz[c])));
        }
        const p64 mw = mul2(m, splat2(wd));
        if (!dt_norm) {
            const p64 t = mul2(mw, penalty_deriv_v2<PC>(pc, fma2(r[2], r[2], fma2(r[1], r[1], mul2(r[0], r[0])))));
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const p64 gg = mul2(t, wc[c]);
                acc_color2(acc, gg, gg, gx[c], gy[c], mul2(d.iz[c], mone));
            }
        } else {
            p64 inv[3];
#pragma unroll
            for (int c = 0; c < 3; c++) inv[c] = rcp2(fma2(gy[c], gy[c], fma2(gx[c], gx[c], dnorm)));
            const p64 x = fma2(mul2(r[2], r[2]), inv[2], fma2(mul2(r[1], r[1]), inv[1], mul2(mul2(r[0], r[0]), inv[0])));
            const p64 t = mul2(mw, penalty_deriv_v2<PC>(pc, x));
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const p64 gg = mul2(mul2(t, inv[c]), wc[c]);
                acc_color2(acc, gg, gg, gx[c], gy[c], mul2(d.iz[c], mone));
            }
        }
    }
    p64 rx[3], ry[3], gxx[3], gyy[3], gxy[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        gxx[c] = mul2(g, d.ixx[c]);
        gyy[c] = mul2(g, d.iyy[c]);
        gxy[c] = mul2(g, d.ixy[c]);
        rx[c] = mul2(wc[c], fma2(gxy[c], v, fma2(gxx[c], u, d.ixz[c])));
        ry[c] = mul2(wc[c], fma2(gyy[c], v, fma2(gxy[c], u, d.iyz[c])));
    }
    const p64 mw = mul2(m, splat2(wg));
    if (!dt_norm) {
        p64 x = mul2(rx[0], rx[0]);
        x = fma2(ry[0], ry[0], x);
        x = fma2(rx[1], rx[1], x);
        x = fma2(ry[1], ry[1], x);
        x = fma2(rx[2], rx[2], x);
        x = fma2(ry[2], ry[2], x);
        const p64 t = mul2(mw, penalty_deriv_v2<PG>(pg, x));
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const p64 gg = mul2(t, wc[c]);
            acc_grad2(acc, gg, gg, gg, gg, gxx[c], gxy[c], gyy[c], mul2(d.ixz[c], mone), mul2(d.iyz[c], mone));
        }
    } else {
        p64 ivx[3], ivy[3];
        p64 x = splat2(0.0f);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const p64 xy = fma2(gxy[c], gxy[c], dnorm);
            ivx[c] = rcp2(fma2(gxx[c], gxx[c], xy));
            ivy[c] = rcp2(fma2(gyy[c], gyy[c], xy));
            x = fma2(mul2(rx[c], rx[c]), ivx[c], x);
            x = fma2(mul2(ry[c], ry[c]), ivy[c], x);
        }
        const p64 t = mul2(mw, penalty_deriv_v2<PG>(pg, x));
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const p64 tw = mul2(t, wc[c]);
            const p64 g1 = mul2(tw, ivx[c]), g2 = mul2(tw, ivy[c]);
            acc_grad2(acc, g1, g2, g1, g2, gxx[c], gxy[c], gyy[c], mul2(d.ixz[c], mone), mul2(d.iyz[c], mone));
        }
    }
}

template <int PC, int PG>
__device__ __forceinline__ void term_mt_ref2(const Derivs2 &d, p64 u, p64 v, p64 m, float wd, float wg, float s, const p64 wc[3],
                                             int dt_norm, const Penalty &pc, const Penalty &pg, Acc2 &acc) {
    const p64 dnorm = splat2(0.1f * 0.1f), mone = splat2(-1.0f);
    const float fsq = s * s;
    const float f = (s >= 0.0f) ? -s : s;
    const p64 f2 = splat2(f), fsq2 = splat2(fsq);
    const p64 fu = mul2(f2, u), fv = mul2(f2, v);
    const p64 ifsq = splat2(1.0f / fsq); // the un-normalised branches divide by factorsq (:452-471, :518-530)
    if (wd != 0.0f) {
        p64 r[3];
#pragma unroll
        for (int c = 0; c < 3; c++) r[c] = mul2(wc[c], fma2(d.iy[c], fv, fma2(d.ix[c], fu, d.iz[c])));
        const p64 mw = mul2(m, splat2(wd));
        if (!dt_norm) {
            const p64 x = mul2(fma2(r[2], r[2], fma2(r[1], r[1], mul2(r[0], r[0]))), ifsq);
            const p64 t = mul2(mul2(mw, penalty_deriv_v2<PC>(pc, x)), ifsq);
            const p64 tf = mul2(t, f2);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const p64 gb = mul2(tf, wc[c]);
                const p64 ga = (c == 2) ? tf : mul2(gb, f2); // channel 3 drops the weight and one factor (as written, :469)
                acc_color2(acc, ga, gb, d.ix[c], d.iy[c], mul2(d.iz[c], mone));
            }
        } else {
            p64 inv[3];
#pragma unroll
            for (int c = 0; c < 3; c++)
                inv[c] = rcp2(fma2(mul2(fsq2, d.iy[c]), d.iy[c], fma2(mul2(fsq2, d.ix[c]), d.ix[c], dnorm)));
            const p64 x = fma2(mul2(r[2], r[2]), inv[2], fma2(mul2(r[1], r[1]), inv[1], mul2(mul2(r[0], r[0]), inv[0])));
            const p64 tf = mul2(mul2(mw, penalty_deriv_v2<PC>(pc, x)), f2);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const p64 gb = mul2(mul2(tf, inv[c]), wc[c]);
                acc_color2(acc, mul2(gb, f2), gb, d.ix[c], d.iy[c], mul2(d.iz[c], mone));
            }
        }
    }
    p64 rx[3], ry[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        rx[c] = mul2(wc[c], fma2(d.ixy[c], fv, fma2(d.ixx[c], fu, d.ixz[c])));
        ry[c] = mul2(wc[c], fma2(d.iyy[c], fv, fma2(d.ixy[c], fu, d.iyz[c])));
    }
    const p64 mw = mul2(m, splat2(wg));
    if (!dt_norm) {
        p64 x = mul2(rx[0], rx[0]);
        x = fma2(ry[0], ry[0], x);
        x = fma2(rx[1], rx[1], x);
        x = fma2(ry[1], ry[1], x);
        x = fma2(rx[2], rx[2], x);
        x = fma2(ry[2], ry[2], x);
        const p64 t = mul2(mul2(mw, penalty_deriv_v2<PG>(pg, mul2(x, ifsq))), ifsq);
        const p64 tf = mul2(t, f2);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const p64 gb = mul2(tf, wc[c]);
            p64 ga = mul2(gb, f2);
            if (c == 0) ga = mul2(ga, fsq2); // channel 1 carries an extra factorsq (as written, :528-530)
            acc_grad2(acc, ga, ga, gb, gb, d.ixx[c], d.ixy[c], d.iyy[c], mul2(d.ixz[c], mone), mul2(d.iyz[c], mone));
        }
    } else {
        p64 ivx[3], ivy[3];
        p64 x = splat2(0.0f);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const p64 xy = fma2(mul2(fsq2, d.ixy[c]), d.ixy[c], dnorm);
            ivx[c] = rcp2(fma2(mul2(fsq2, d.ixx[c]), d.ixx[c], xy));
            ivy[c] = rcp2(fma2(mul2(fsq2, d.iyy[c]), d.iyy[c], xy));
            x = fma2(mul2(rx[c], rx[c]), ivx[c], x);
            x = fma2(mul2(ry[c], ry[c]), ivy[c], x);
        }
        const p64 tf = mul2(mul2(mw, penalty_deriv_v2<PG>(pg, x)), f2);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const p64 tw = mul2(tf, wc[c]);
            const p64 gb1 = mul2(tw, ivx[c]), gb2 = mul2(tw, ivy[c]);
            acc_grad2(acc, mul2(gb1, f2), mul2(gb2, f2), gb1, gb2, d.ixx[c], d.ixy[c], d.iyy[c], mul2(d.ixz[c], mone),
                      mul2(d.iyz[c], mone));
        }
    }
}

#ifndef SF_MT_TERMS2_MINB
#define SF_MT_TERMS2_MINB 2 // resident CTAs per SM the register allocation aims at (128 registers, 16 warps = 1024 pixels per SM; 3: 80 registers + 300 B of spills, config-3 window 11.5 instead of 11.1 ms, config 4 65.9 instead of 60.7)
#endif
template <int PC, int PG>
__global__ void __launch_bounds__(256, SF_MT_TERMS2_MINB) k_mt_terms2(Geom g, MtTermsArgs ta, DataCommon cm) {
    pdl_enter();
    if (g.cancelled()) return;
    const int i = (blockIdx.x * 32 + threadIdx.x) * 2, j = blockIdx.y * 8 + threadIdx.y; // columns i, i + 1
    if (i >= g.S || j >= g.H) return;
    const size_t P = g.plane(), o = (size_t)j * g.S + i;
    auto store2 = [](float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); };
    if (i >= g.W) { // padding columns: defined zeros
        store2(cm.a11 + o, 0.f, 0.f); store2(cm.a12 + o, 0.f, 0.f); store2(cm.a22 + o, 0.f, 0.f);
        store2(cm.b1 + o, 0.f, 0.f); store2(cm.b2 + o, 0.f, 0.f);
        return;
    }
    const int W1 = g.W - 1, H1 = g.H - 1;
    const bool v1 = i + 1 <= W1; // the pair's second column is a pixel (else padding: computed on whatever is there, stored as 0)
    const p64 zero2 = splat2(0.0f), one2 = splat2(1.0f), half2 = splat2(0.5f), mone = splat2(-1.0f);
    const p64 u = cm.du ? ldg2(cm.du + o) : zero2, v = cm.dv ? ldg2(cm.dv + o) : zero2;
    p64 wc[3] = {one2, one2, one2};
    if (cm.chw) {
        wc[0] = ldg2(cm.chw + o); wc[1] = ldg2(cm.chw + o + cm.chw_pstride); wc[2] = ldg2(cm.chw + o + 2 * cm.chw_pstride);
    }
    // occlusion / window factor of variational_mt.cpp:293-320 per direction: (1 * (sel / fac)) with sel in {0, 1}
    p64 occ_past = one2, occ_future = one2;
    if (cm.occ) {
        const p64 oc = ldg2(cm.occ + o);
        const float oc0 = lo_of(oc), oc1 = hi_of(oc);
        const float q0 = 1.0f / ((1.0f + ((oc0 == 0.0f) ? 1.0f : 0.0f)) * cm.data_norm);
        const float q1 = 1.0f / ((1.0f + ((oc1 == 0.0f) ? 1.0f : 0.0f)) * cm.data_norm);
        occ_past = pk(oc0 >= 0.0f ? q0 : 0.0f, oc1 >= 0.0f ? q1 : 0.0f);
        occ_future = pk(oc0 <= 0.0f ? q0 : 0.0f, oc1 <= 0.0f ? q1 : 0.0f);
    }
    Acc2 acc;
    acc.a11 = acc.a12 = acc.a22 = acc.b1 = acc.b2 = zero2;
#pragma unroll 1
    for (int k = 0; k < ta.nterms; k++) {
        const MtTerm t = ta.term[k];
        const float *IA = ta.I[t.fa] + o, *IB = ta.I[t.fb] + o, *DA = ta.D[t.fa] + o, *DB = ta.D[t.fb] + o;
        Derivs2 d;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            d.iz[c] = sub2(ldg2(IA + c * P), ldg2(IB + c * P));
            const p64 xa = ldg2(DA + (0 + c) * P), xb = ldg2(DB + (0 + c) * P);
            const p64 ya = ldg2(DA + (3 + c) * P), yb = ldg2(DB + (3 + c) * P);
            d.ix[c] = mul2(add2(xb, xa), half2);
            d.iy[c] = mul2(add2(yb, ya), half2);
            d.ixz[c] = sub2(xa, xb);
            d.iyz[c] = sub2(ya, yb);
            d.ixx[c] = mul2(add2(ldg2(DB + (6 + c) * P), ldg2(DA + (6 + c) * P)), half2);
            d.ixy[c] = mul2(add2(ldg2(DB + (9 + c) * P), ldg2(DA + (9 + c) * P)), half2);
            d.iyy[c] = mul2(add2(ldg2(DB + (12 + c) * P), ldg2(DA + (12 + c) * P)), half2);
        }
        p64 m = ldg2(ta.mask[t.mask_frame] + o);
        if (cm.occ) m = mul2((t.dir == 0) ? occ_past : occ_future, m);
        if (t.kind == DK_MT_SUCC) term_mt_succ2<PC, PG>(d, u, v, m, t.wd, t.wg, t.s, wc, cm.dt_norm, cm.pc, cm.pg, acc);
        else term_mt_ref2<PC, PG>(d, u, v, m, t.wd, t.wg, t.s, wc, cm.dt_norm, cm.pc, cm.pg, acc);
    }
    {
        // b += div(psi grad w) and the 2x2 block inverse (variational_aux.c:158-179, solver.c:101-106): left edge, right
        // edge, upper edge, lower edge; neighbours outside the image are the pixel itself, their diffusivities 0
        const p64 hr = ldg2(cm.ph + o), vb = ldg2(cm.pv + o);
        const p64 hl = pk((i > 0) ? __ldg(cm.ph + o - 1) : 0.0f, lo_of(hr));
        const p64 vt = (j > 0) ? ldg2(cm.pv + o - g.S) : zero2;
        const size_t ot = (j > 0) ? o - g.S : o, ob = (j < H1) ? o + g.S : o;
        const p64 nhl = mul2(hl, mone), nvt = mul2(vt, mone);
        auto lap = [&](const float *w, p64 b) {
            const p64 wcn = ldg2(w + o), wt = ldg2(w + ot), wb = ldg2(w + ob);
            const float c0 = lo_of(wcn), c1 = hi_of(wcn);
            const p64 wl = pk((i > 0) ? __ldg(w + o - 1) : c0, c0);
            const p64 wr = pk((i < W1) ? c1 : c0, (i + 1 < W1) ? __ldg(w + o + 2) : c1);
            b = fma2(nhl, sub2(wcn, wl), b);
            b = fma2(hr, sub2(wr, wcn), b);
            b = fma2(nvt, sub2(wcn, wt), b);
            return fma2(vb, sub2(wb, wcn), b);
        };
        acc.b1 = lap(cm.lap_u, acc.b1);
        acc.b2 = lap(cm.lap_v, acc.b2);
        const p64 sp = add2(add2(add2(hl, hr), vt), vb);
        const p64 D11 = add2(acc.a22, sp), D22 = add2(acc.a11, sp);
        const p64 det = fma2(mul2(acc.a12, acc.a12), mone, mul2(D11, D22));
        const p64 rdet = rcp2(det);
        acc.a11 = mul2(D11, rdet);
        acc.a22 = mul2(D22, rdet);
        acc.a12 = mul2(mul2(acc.a12, mone), rdet);
    }
    store2(cm.a11 + o, lo_of(acc.a11), v1 ? hi_of(acc.a11) : 0.0f);
    store2(cm.a12 + o, lo_of(acc.a12), v1 ? hi_of(acc.a12) : 0.0f);
    store2(cm.a22 + o, lo_of(acc.a22), v1 ? hi_of(acc.a22) : 0.0f);
    store2(cm.b1 + o, lo_of(acc.b1), v1 ? hi_of(acc.b1) : 0.0f);
    store2(cm.b2 + o, lo_of(acc.b2), v1 ? hi_of(acc.b2) : 0.0f);
}

bool data_term_device_init() { // per device, called by sfgpu_create
    const int smem = (int)(DT_SMEM_FLOATS * sizeof(float));
    return cudaFuncSetAttribute(k_data_term<DK_TWO_FRAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
           cudaFuncSetAttribute(k_data_term<DK_MT_SUCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
           cudaFuncSetAttribute(k_data_term<DK_MT_REF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
           cudaFuncSetAttribute(k_data_term<DK_DERIVS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
}

void launch_data_term(cudaStream_t st, Geom g, const DataTermDesc &t, const DataCommon &cm) {
    dim3 b(32, 8), grid((g.S + DT_TW - 1) / DT_TW, (g.H + DT_TH - 1) / DT_TH);
    const size_t smem = DT_SMEM_FLOATS * sizeof(float);
    switch (t.kind) {
    case DK_MT_SUCC: launch_pdl(k_data_term<DK_MT_SUCC>, grid, b, smem, st, g, t, cm); break;
    case DK_MT_REF: launch_pdl(k_data_term<DK_MT_REF>, grid, b, smem, st, g, t, cm); break;
    default: launch_pdl(k_data_term<DK_TWO_FRAME>, grid, b, smem, st, g, t, cm); break;
    }
}

void launch_frame_derivs(cudaStream_t st, Geom g, const float *image3, float *derivs15) {
    dim3 b(32, 8), grid((g.S + DT_TW - 1) / DT_TW, (g.H + DT_TH - 1) / DT_TH);
    DataTermDesc t{image3, image3, +1, nullptr, DK_DERIVS, 0.0f, 0.0f, 1.0f, -1};
    DataCommon cm{};
    cm.a11 = derivs15;
    launch_pdl(k_data_term<DK_DERIVS>, grid, b, DT_SMEM_FLOATS * sizeof(float), st, g, t, cm);
}

template <int PC>
static void launch_mt_terms_pg(cudaStream_t st, dim3 grid, dim3 b, Geom g, const MtTermsArgs &ta, const DataCommon &cm) {
    switch (cm.pg.type) {
    case SF_ROBUST_QUADRATIC: launch_pdl(k_mt_terms2<PC, SF_ROBUST_QUADRATIC>, grid, b, 0, st, g, ta, cm); break;
    case SF_ROBUST_MODL1: launch_pdl(k_mt_terms2<PC, SF_ROBUST_MODL1>, grid, b, 0, st, g, ta, cm); break;
    case SF_ROBUST_LORENTZIAN: launch_pdl(k_mt_terms2<PC, SF_ROBUST_LORENTZIAN>, grid, b, 0, st, g, ta, cm); break;
    case SF_ROBUST_TRUNC_MODL1: launch_pdl(k_mt_terms2<PC, SF_ROBUST_TRUNC_MODL1>, grid, b, 0, st, g, ta, cm); break;
    case SF_ROBUST_GEMAN_MCCLURE: launch_pdl(k_mt_terms2<PC, SF_ROBUST_GEMAN_MCCLURE>, grid, b, 0, st, g, ta, cm); break;
    default: launch_pdl(k_mt_terms2<PC, -1>, grid, b, 0, st, g, ta, cm); break;
    }
}

// every plane the packed kernel touches must start on an 8-byte boundary (rows do: S is a multiple of 4 floats)
static bool mt_terms_pairs_aligned(const MtTermsArgs &ta, const DataCommon &cm) {
    uintptr_t bits = 0;
    auto add = [&](const void *p) { bits |= reinterpret_cast<uintptr_t>(p); };
    for (int k = 0; k < ta.nterms; k++) {
        const MtTerm &t = ta.term[k];
        add(ta.I[t.fa]); add(ta.I[t.fb]); add(ta.D[t.fa]); add(ta.D[t.fb]); add(ta.mask[t.mask_frame]);
    }
    add(cm.du); add(cm.dv); add(cm.chw); add(cm.occ); add(cm.ph); add(cm.pv); add(cm.lap_u); add(cm.lap_v);
    add(cm.a11); add(cm.a12); add(cm.a22); add(cm.b1); add(cm.b2);
    return (bits & 7) == 0 && (cm.chw_pstride & 1) == 0;
}

void launch_mt_terms(cudaStream_t st, Geom g, const MtTermsArgs &ta, const DataCommon &cm, bool force_scalar) {
    if (force_scalar || !mt_terms_pairs_aligned(ta, cm)) { // one column per thread, penalties by run-time switch
        launch_pdl(k_mt_terms<-1, -1>, dim3((g.S + 31) / 32, (g.H + 7) / 8), dim3(32, 8), 0, st, g, ta, cm);
        return;
    }
    dim3 b(32, 8), grid((g.S + 63) / 64, (g.H + 7) / 8);
    switch (cm.pc.type) {
    case SF_ROBUST_QUADRATIC: launch_mt_terms_pg<SF_ROBUST_QUADRATIC>(st, grid, b, g, ta, cm); break;
    case SF_ROBUST_MODL1: launch_mt_terms_pg<SF_ROBUST_MODL1>(st, grid, b, g, ta, cm); break;
    case SF_ROBUST_LORENTZIAN: launch_mt_terms_pg<SF_ROBUST_LORENTZIAN>(st, grid, b, g, ta, cm); break;
    case SF_ROBUST_TRUNC_MODL1: launch_mt_terms_pg<SF_ROBUST_TRUNC_MODL1>(st, grid, b, g, ta, cm); break;
    case SF_ROBUST_GEMAN_MCCLURE: launch_mt_terms_pg<SF_ROBUST_GEMAN_MCCLURE>(st, grid, b, g, ta, cm); break;
    default: launch_mt_terms_pg<-1>(st, grid, b, g, ta, cm); break;
    }
}

} // namespace sf
