/*
 * sf_oracle.c -- CPU restatement of the two-frame variational refinement (see sf_oracle.h).
 * TEST INFRASTRUCTURE ONLY; never linked into or loaded by the product library.
 *
 * Written as plain index arithmetic over (column i, row j) with o = j*stride + i.  Each block
 * keeps the reference's fp32 evaluation order so results are bit-identical to the reference
 * objects (checked by tests/test_oracle_pin.py).  Build: gcc -O2 -ffp-contract=off.
 */
#include "sf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ containers (image.c:17-89) */

static void *sfo_alloc(size_t bytes) {
    void *p = NULL;
    if (bytes == 0) bytes = 16;
    if (posix_memalign(&p, 64, bytes) != 0 || !p) {
        fprintf(stderr, "sf_oracle: out of memory\n");
        exit(1);
    }
    memset(p, 0, bytes);
    return p;
}

sfo_image_t *sfo_image_new(int width, int height) {
    sfo_image_t *im = (sfo_image_t *)malloc(sizeof(*im));
    im->width = width;
    im->height = height;
    im->stride = ((width + 3) / 4) * 4;
    im->data = (float *)sfo_alloc(sizeof(float) * (size_t)im->stride * height);
    return im;
}

void sfo_image_delete(sfo_image_t *im) {
    if (!im) return;
    free(im->data);
    free(im);
}

sfo_color_image_t *sfo_color_image_new(int width, int height) {
    sfo_color_image_t *im = (sfo_color_image_t *)malloc(sizeof(*im));
    im->width = width;
    im->height = height;
    im->stride = ((width + 3) / 4) * 4;
    im->c1 = (float *)sfo_alloc(sizeof(float) * 3 * (size_t)im->stride * height);
    im->c2 = im->c1 + (size_t)im->stride * height;
    im->c3 = im->c2 + (size_t)im->stride * height;
    return im;
}

void sfo_color_image_delete(sfo_color_image_t *im) {
    if (!im) return;
    free(im->c1);
    free(im);
}

/* ------------------------------------------------------------------ filters (image.c:351-373)
 * convolution_new(order, half, even=0): taps[order-k] = +half[k], taps[order+k] = -half[k]; the
 * k=0 assignment happens twice, leaving the centre tap at -half[0] (= -0.0f for both filters). */
static void sfo_deriv5(float t[5]) { /* variational.c:118-119 */
    const float h1 = -8.0f / 12.0f, h2 = 1.0f / 12.0f;
    t[0] = h2; t[1] = h1; t[2] = -0.0f; t[3] = -h1; t[4] = -h2;
}
static void sfo_deriv3(float t[3]) { /* variational.c:120-121 */
    t[0] = -0.5f; t[1] = -0.0f; t[2] = 0.5f;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* image.c:425-447 (3 taps) and :449-526 wait -- horizontal: :460-526.  The reference builds
 * shifted row copies whose out-of-range entries replicate the first / last VALID pixel (it first
 * overwrites the padding columns of src with src[width-1]); that is a clamp of the column index
 * to [0, width-1] for every output column of the stride. */
void sfo_convolve_horiz(sfo_image_t *dst, const sfo_image_t *src, int order, const float *c) {
    const int W = src->width, H = src->height, S = src->stride;
    for (int j = 0; j < H; j++) {
        const float *s = src->data + (size_t)j * S;
        float *d = dst->data + (size_t)j * S;
        for (int i = 0; i < S; i++) {
            if (order == 1) {
                d[i] = c[0] * s[clampi(i - 1, 0, W - 1)] + c[1] * s[clampi(i, 0, W - 1)] +
                       c[2] * s[clampi(i + 1, 0, W - 1)];
            } else {
                d[i] = c[0] * s[clampi(i - 2, 0, W - 1)] + c[1] * s[clampi(i - 1, 0, W - 1)] +
                       c[2] * s[clampi(i, 0, W - 1)] + c[3] * s[clampi(i + 1, 0, W - 1)] +
                       c[4] * s[clampi(i + 2, 0, W - 1)];
            }
        }
    }
}

/* image.c:400-458.  Border rows fold the taps that fall outside into the nearest row's tap
 * (sum of coefficients first, then one multiply) -- kept because it rounds differently from a
 * replicate read.  All stride columns are processed. */
void sfo_convolve_vert(sfo_image_t *dst, const sfo_image_t *src, int order, const float *c) {
    const int H = src->height, S = src->stride;
    for (int j = 0; j < H; j++) {
        float *d = dst->data + (size_t)j * S;
        const float *s0 = src->data + (size_t)j * S;
        if (order == 1) {
            const float *sm = s0 - S, *sp = s0 + S;
            for (int i = 0; i < S; i++) {
                if (j == 0) d[i] = (c[0] + c[1]) * s0[i] + c[2] * sp[i];
                else if (j == H - 1) d[i] = c[0] * sm[i] + (c[1] + c[2]) * s0[i];
                else d[i] = c[0] * sm[i] + c[1] * s0[i] + c[2] * sp[i];
            }
        } else {
            const float *sm2 = s0 - 2 * S, *sm1 = s0 - S, *sp1 = s0 + S, *sp2 = s0 + 2 * S;
            for (int i = 0; i < S; i++) {
                if (j == 0) d[i] = (c[0] + c[1] + c[2]) * s0[i] + c[3] * sp1[i] + c[4] * sp2[i];
                else if (j == 1) d[i] = (c[0] + c[1]) * sm1[i] + c[2] * s0[i] + c[3] * sp1[i] + c[4] * sp2[i];
                else if (j == H - 2) d[i] = c[0] * sm2[i] + c[1] * sm1[i] + c[2] * s0[i] + (c[3] + c[4]) * sp1[i];
                else if (j == H - 1) d[i] = c[0] * sm2[i] + c[1] * sm1[i] + (c[2] + c[3] + c[4]) * s0[i];
                else d[i] = c[0] * sm2[i] + c[1] * sm1[i] + c[2] * s0[i] + c[3] * sp1[i] + c[4] * sp2[i];
            }
        }
    }
}

static void color_convolve(sfo_color_image_t *dst, const sfo_color_image_t *src, int horiz, const float *c5) {
    /* image.c:658-688 with exactly one of horiz_conv / vert_conv given */
    float *dp[3] = {dst->c1, dst->c2, dst->c3};
    float *sp[3] = {src->c1, src->c2, src->c3};
    for (int ch = 0; ch < 3; ch++) {
        sfo_image_t d = {src->width, src->height, src->stride, dp[ch]};
        sfo_image_t s = {src->width, src->height, src->stride, sp[ch]};
        if (horiz) sfo_convolve_horiz(&d, &s, 2, c5);
        else sfo_convolve_vert(&d, &s, 2, c5);
    }
}

/* ------------------------------------------------------------------ warp
 * variational_aux.c:18-52; integer time factor as in variational_aux_mt.cpp:722-756. */
void sfo_image_warp(sfo_color_image_t *dst, sfo_image_t *mask, const sfo_color_image_t *src,
                    const sfo_image_t *wx, const sfo_image_t *wy, int factor) {
    const int W = src->width, H = src->height, S = src->stride;
    const float *sc[3] = {src->c1, src->c2, src->c3};
    float *dc[3] = {dst->c1, dst->c2, dst->c3};
    for (int j = 0; j < H; j++) {
        for (int i = 0; i < W; i++) {
            const size_t o = (size_t)j * S + i;
            const float xx = i + factor * wx->data[o];
            const float yy = j + factor * wy->data[o];
            const int x = (int)floor(xx), y = (int)floor(yy);
            const float dx = xx - x, dy = yy - y;
            if (mask) mask->data[o] = (xx >= 0 && xx <= W - 1 && yy >= 0 && yy <= H - 1) ? 1.0f : 0.0f;
            const int x1 = clampi(x, 0, W - 1), x2 = clampi(x + 1, 0, W - 1);
            const int y1 = clampi(y, 0, H - 1), y2 = clampi(y + 1, 0, H - 1);
            for (int ch = 0; ch < 3; ch++) {
                const float *s = sc[ch];
                dc[ch][o] = s[(size_t)y1 * S + x1] * (1.0f - dx) * (1.0f - dy) + s[(size_t)y1 * S + x2] * dx * (1.0f - dy) +
                            s[(size_t)y2 * S + x1] * (1.0f - dx) * dy + s[(size_t)y2 * S + x2] * dx * dy;
            }
        }
    }
}

/* ------------------------------------------------------------------ derivatives (variational_aux.c:55-78) */
void sfo_get_derivatives(const sfo_color_image_t *im1, const sfo_color_image_t *im2,
                         sfo_color_image_t *dx, sfo_color_image_t *dy, sfo_color_image_t *dt,
                         sfo_color_image_t *dxx, sfo_color_image_t *dxy, sfo_color_image_t *dyy,
                         sfo_color_image_t *dxt, sfo_color_image_t *dyt) {
    float c5[5];
    sfo_deriv5(c5);
    sfo_color_image_t *mean = sfo_color_image_new(im2->width, im2->height);
    const size_t n = 3 * (size_t)im1->height * im1->stride;
    for (size_t k = 0; k < n; k++) {
        mean->c1[k] = 0.5f * (im2->c1[k] + im1->c1[k]);
        dt->c1[k] = im2->c1[k] - im1->c1[k];
    }
    color_convolve(dx, mean, 1, c5);
    color_convolve(dy, mean, 0, c5);
    color_convolve(dxx, dx, 1, c5);
    color_convolve(dxy, dx, 0, c5);
    color_convolve(dyy, dy, 0, c5);
    color_convolve(dxt, dt, 1, c5);
    color_convolve(dyt, dt, 0, c5);
    sfo_color_image_delete(mean);
}

/* ------------------------------------------------------------------ smoothness (variational_aux.c:84-149) */
void sfo_compute_smoothness(sfo_image_t *dst_horiz, sfo_image_t *dst_vert, const sfo_image_t *uu,
                            const sfo_image_t *vv, const sfo_image_t *w, float half_alpha) {
    const int W = uu->width, H = uu->height, S = uu->stride;
    const float eps_smooth = 0.001f * 0.001f; /* variational_aux.c:13 */
    float c3[3];
    sfo_deriv3(c3);
    sfo_image_t *ux2 = sfo_image_new(W, H), *uy2 = sfo_image_new(W, H), *vx2 = sfo_image_new(W, H),
                *vy2 = sfo_image_new(W, H);
    sfo_convolve_horiz(ux2, uu, 1, c3);
    sfo_convolve_horiz(vx2, vv, 1, c3);
    sfo_convolve_vert(uy2, uu, 1, c3);
    sfo_convolve_vert(vy2, vv, 1, c3);
    for (int j = 0; j < H; j++) {
        for (int i = 0; i < W - 1; i++) {
            const size_t o = (size_t)j * S + i;
            const float ux1 = uu->data[o + 1] - uu->data[o];
            const float vx1 = vv->data[o + 1] - vv->data[o];
            float t = 0.5f * (uy2->data[o] + uy2->data[o + 1]);
            const float uxsq = ux1 * ux1 + t * t;
            t = 0.5f * (vy2->data[o] + vy2->data[o + 1]);
            const float vxsq = vx1 * vx1 + t * t;
            t = uxsq + vxsq;
            /* float product / double sqrt, rounded to float on store (:124) */
            dst_horiz->data[o] = (float)((w->data[o] + w->data[o + 1]) * half_alpha / sqrt(t + eps_smooth));
        }
        for (int i = W - 1; i < S; i++) dst_horiz->data[(size_t)j * S + i] = 0.0f; /* :126 */
    }
    for (int j = 0; j < H - 1; j++) {
        for (int i = 0; i < W; i++) {
            const size_t o = (size_t)j * S + i;
            const float uy1 = uu->data[o + S] - uu->data[o];
            const float vy1 = vv->data[o + S] - vv->data[o];
            float t = 0.5f * (ux2->data[o] + ux2->data[o + S]);
            const float uysq = uy1 * uy1 + t * t;
            t = 0.5f * (vx2->data[o] + vx2->data[o + S]);
            const float vysq = vy1 * vy1 + t * t;
            t = uysq + vysq;
            dst_vert->data[o] = (float)((w->data[o] + w->data[o + S]) * half_alpha / sqrt(t + eps_smooth));
        }
    }
    for (int i = 0; i < S; i++) dst_vert->data[(size_t)(H - 1) * S + i] = 0.0f; /* :146 */
    sfo_image_delete(ux2); sfo_image_delete(uy2); sfo_image_delete(vx2); sfo_image_delete(vy2);
}

/* ------------------------------------------------------------------ laplacian (variational_aux.c:153-180)
 * Scatter form kept (the accumulation order into dst decides the rounding). */
void sfo_sub_laplacian(sfo_image_t *dst, const sfo_image_t *src, const sfo_image_t *wh, const sfo_image_t *wv) {
    const int W = src->width, H = src->height, S = src->stride;
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W - 1; i++) {
            const size_t o = (size_t)j * S + i;
            const float t = wh->data[o] * (src->data[o + 1] - src->data[o]);
            dst->data[o] += t;
            dst->data[o + 1] -= t;
        }
    for (size_t o = 0; o < (size_t)(H - 1) * S; o++) { /* every stride column, rows 0..H-2 */
        const float t = wv->data[o] * (src->data[o + S] - src->data[o]);
        dst->data[o] += t;
        dst->data[o + S] -= t;
    }
}

/* ------------------------------------------------------------------ smoothness weight (variational_aux.c:183-209) */
sfo_image_t *sfo_compute_dpsis_weight(const sfo_color_image_t *im, float coef) {
    const int W = im->width, H = im->height;
    float c5[5];
    sfo_deriv5(c5);
    sfo_image_t *lum = sfo_image_new(W, H), *lx = sfo_image_new(W, H), *ly = sfo_image_new(W, H);
    const size_t n = (size_t)im->height * im->stride;
    for (size_t k = 0; k < n; k++)
        lum->data[k] = (0.299f * im->c1[k] + 0.587f * im->c2[k] + 0.114f * im->c3[k]) / 255.0f;
    sfo_convolve_horiz(lx, lum, 2, c5);
    sfo_convolve_vert(ly, lum, 2, c5);
    for (size_t k = 0; k < n; k++) {
        const float g = -coef * sqrtf(lx->data[k] * lx->data[k] + ly->data[k] * ly->data[k]);
        lum->data[k] = 0.5f * expf(g);
    }
    sfo_image_delete(lx);
    sfo_image_delete(ly);
    return lum;
}

/* ------------------------------------------------------------------ data term (variational_aux.c:215-302) */
void sfo_compute_data_and_match(sfo_image_t *a11, sfo_image_t *a12, sfo_image_t *a22, sfo_image_t *b1,
                                sfo_image_t *b2, const sfo_image_t *mask, const sfo_image_t *du,
                                const sfo_image_t *dv, const sfo_color_image_t *Ix, const sfo_color_image_t *Iy,
                                const sfo_color_image_t *Iz, const sfo_color_image_t *Ixx,
                                const sfo_color_image_t *Ixy, const sfo_color_image_t *Iyy,
                                const sfo_color_image_t *Ixz, const sfo_color_image_t *Iyz,
                                float half_delta_over3, float half_gamma_over3) {
    const float dnorm = 0.1f * 0.1f, eps_color = 0.001f * 0.001f, eps_grad = 0.001f * 0.001f; /* :10-12 */
    const size_t n = (size_t)du->height * du->stride; /* runs over the padding as well (:239) */
    const size_t P = n;                               /* plane size: c2 = c1 + P, c3 = c2 + P */
    for (size_t k = 0; k < n; k++) {
        float A11 = 0.0f, A12 = 0.0f, A22 = 0.0f, B1 = 0.0f, B2 = 0.0f;
        const float u = du->data[k], v = dv->data[k], m = mask->data[k];
        if (half_delta_over3) { /* colour constancy (:242-265) */
            float r[3], nn[3], g[3];
            for (int c = 0; c < 3; c++) {
                const float ix = Ix->c1[k + c * P], iy = Iy->c1[k + c * P], iz = Iz->c1[k + c * P];
                r[c] = iz + ix * u + iy * v;
                nn[c] = ix * ix + iy * iy + dnorm;
            }
            const float t = m * half_delta_over3 /
                            sqrtf(r[0] * r[0] / nn[0] + r[1] * r[1] / nn[1] + r[2] * r[2] / nn[2] + eps_color);
            g[2] = t / nn[2]; g[1] = t / nn[1]; g[0] = t / nn[0];
            for (int c = 0; c < 3; c++) {
                const float ix = Ix->c1[k + c * P], iy = Iy->c1[k + c * P], iz = Iz->c1[k + c * P];
                A11 += g[c] * ix * ix;
                A12 += g[c] * ix * iy;
                A22 += g[c] * iy * iy;
                B1 -= g[c] * iz * ix;
                B2 -= g[c] * iz * iy;
            }
        }
        { /* gradient constancy (:267-296) */
            float rx[3], ry[3], nx[3], ny[3], gx[3], gy[3];
            for (int c = 0; c < 3; c++) {
                const float ixx = Ixx->c1[k + c * P], ixy = Ixy->c1[k + c * P], iyy = Iyy->c1[k + c * P];
                nx[c] = ixx * ixx + ixy * ixy + dnorm;
                ny[c] = iyy * iyy + ixy * ixy + dnorm;
                rx[c] = Ixz->c1[k + c * P] + ixx * u + ixy * v;
                ry[c] = Iyz->c1[k + c * P] + ixy * u + iyy * v;
            }
            const float t = m * half_gamma_over3 /
                            sqrtf(rx[0] * rx[0] / nx[0] + ry[0] * ry[0] / ny[0] + rx[1] * rx[1] / nx[1] +
                                  ry[1] * ry[1] / ny[1] + rx[2] * rx[2] / nx[2] + ry[2] * ry[2] / ny[2] + eps_grad);
            gy[2] = t / ny[2]; gx[2] = t / nx[2]; gy[1] = t / ny[1]; gx[1] = t / nx[1]; gy[0] = t / ny[0]; gx[0] = t / nx[0];
            for (int c = 0; c < 3; c++) {
                const float ixx = Ixx->c1[k + c * P], ixy = Ixy->c1[k + c * P], iyy = Iyy->c1[k + c * P];
                const float ixz = Ixz->c1[k + c * P], iyz = Iyz->c1[k + c * P];
                A11 += gx[c] * ixx * ixx + gy[c] * ixy * ixy;
                A12 += gx[c] * ixx * ixy + gy[c] * ixy * iyy;
                A22 += gy[c] * iyy * iyy + gx[c] * ixy * ixy;
                B1 -= gx[c] * ixx * ixz + gy[c] * ixy * iyz;
                B2 -= gy[c] * iyy * iyz + gx[c] * ixy * ixz;
            }
        }
        a11->data[k] = A11; a12->data[k] = A12; a22->data[k] = A22; b1->data[k] = B1; b2->data[k] = B2;
    }
}

/* ------------------------------------------------------------------ SOR
 * solver.c:17-57: the readable specification. */
void sfo_sor_coupled_readable(sfo_image_t *du, sfo_image_t *dv, const sfo_image_t *a11, const sfo_image_t *a12,
                              const sfo_image_t *a22, const sfo_image_t *b1, const sfo_image_t *b2,
                              const sfo_image_t *ph, const sfo_image_t *pv, int iterations, float omega) {
    const int W = du->width, H = du->height, S = du->stride;
    for (int it = 0; it < iterations; it++)
        for (int j = 0; j < H; j++)
            for (int i = 0; i < W; i++) {
                const size_t o = (size_t)j * S + i;
                float su = 0.0f, sv = 0.0f, sp = 0.0f;
                if (j > 0) { su -= pv->data[o - S] * du->data[o - S]; sv -= pv->data[o - S] * dv->data[o - S]; sp += pv->data[o - S]; }
                if (i > 0) { su -= ph->data[o - 1] * du->data[o - 1]; sv -= ph->data[o - 1] * dv->data[o - 1]; sp += ph->data[o - 1]; }
                if (j < H - 1) { su -= pv->data[o] * du->data[o + S]; sv -= pv->data[o] * dv->data[o + S]; sp += pv->data[o]; }
                if (i < W - 1) { su -= ph->data[o] * du->data[o + 1]; sv -= ph->data[o] * dv->data[o + 1]; sp += ph->data[o]; }
                const float A11 = a11->data[o] + sp, A12 = a12->data[o], A22 = a22->data[o] + sp;
                const float det = A11 * A22 - A12 * A12;
                const float B1 = b1->data[o] - su, B2 = b2->data[o] - sv;
                du->data[o] = (1.0f - omega) * du->data[o] + omega * (A22 * B1 - A12 * B2) / det;
                dv->data[o] = (1.0f - omega) * dv->data[o] + omega * (-A12 * B1 + A11 * B2) / det;
            }
}

/* In-place inverse of the 2x2 diagonal blocks, solver.c:101-106 (and :158-163, :206-211).
 * The reference reads psi_h(i-1) from a zero-prefixed shifted row, psi_h(W-1..) = 0 and psi_v(H-1) = 0
 * by construction of compute_smoothness; missing top/bottom terms are omitted, not added as 0. */
static void sfo_invert_blocks(sfo_image_t *a11, sfo_image_t *a12, sfo_image_t *a22, const sfo_image_t *ph,
                              const sfo_image_t *pv) {
    const int W = a11->width, H = a11->height, S = a11->stride;
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            const size_t o = (size_t)j * S + i;
            const float hl = (i > 0) ? ph->data[o - 1] : 0.0f;
            float sp = hl + ph->data[o];
            if (j > 0) sp = sp + pv->data[o - S];
            if (j < H - 1) sp = sp + pv->data[o];
            const float A11 = a22->data[o] + sp, A22 = a11->data[o] + sp; /* adjugate: swapped on purpose */
            const float det = A11 * A22 - a12->data[o] * a12->data[o];
            a11->data[o] = A11 / det;
            a22->data[o] = A22 / det;
            a12->data[o] = a12->data[o] / -det;
        }
}

/* One pixel update, solver.c:132-138 (same expression in every row class). */
static inline void sfo_sor_pixel(float *du, float *dv, const float *a11, const float *a12, const float *a22,
                                 const float *b1, const float *b2, const float *ph, const float *pv, int i, int j,
                                 int W, int H, int S, float omega) {
    const size_t o = (size_t)j * S + i;
    /* right neighbour: the reference multiplies psi_h(i) by a shifted copy that holds 0 beyond W-2 */
    const float dur = (i < W - 1) ? du[o + 1] : 0.0f, dvr = (i < W - 1) ? dv[o + 1] : 0.0f;
    float s1 = ph[o] * dur, s2 = ph[o] * dvr;
    if (j > 0) { s1 = s1 + pv[o - S] * du[o - S]; s2 = s2 + pv[o - S] * dv[o - S]; }
    if (j < H - 1) { s1 = s1 + pv[o] * du[o + S]; s2 = s2 + pv[o] * dv[o + S]; }
    s1 = s1 + b1[o];
    s2 = s2 + b2[o];
    float B1 = s1, B2 = s2;
    if (i > 0) { B1 = ph[o - 1] * du[o - 1] + s1; B2 = ph[o - 1] * dv[o - 1] + s2; }
    du[o] += omega * (a11[o] * B1 + a12[o] * B2 - du[o]);
    dv[o] += omega * (a12[o] * B1 + a22[o] * B2 - dv[o]);
}

void sfo_sor_coupled(sfo_image_t *du, sfo_image_t *dv, sfo_image_t *a11, sfo_image_t *a12, sfo_image_t *a22,
                     const sfo_image_t *b1, const sfo_image_t *b2, const sfo_image_t *ph, const sfo_image_t *pv,
                     int iterations, float omega, int mode) {
    const int W = du->width, H = du->height, S = du->stride;
    if (W < 2 || H < 2 || iterations < 1) { /* solver.c:66-69 */
        sfo_sor_coupled_readable(du, dv, a11, a12, a22, b1, b2, ph, pv, iterations, omega);
        return;
    }
    sfo_invert_blocks(a11, a12, a22, ph, pv);
    for (int it = 0; it < iterations; it++) {
        if (mode == SFO_SOR_LEX) {
            for (int j = 0; j < H; j++)
                for (int i = 0; i < W; i++)
                    sfo_sor_pixel(du->data, dv->data, a11->data, a12->data, a22->data, b1->data, b2->data,
                                  ph->data, pv->data, i, j, W, H, S, omega);
        } else {
            for (int colour = 0; colour < 2; colour++)
                for (int j = 0; j < H; j++)
                    for (int i = (j + colour) & 1; i < W; i += 2)
                        sfo_sor_pixel(du->data, dv->data, a11->data, a12->data, a22->data, b1->data, b2->data,
                                      ph->data, pv->data, i, j, W, H, S, omega);
        }
    }
}

/* ------------------------------------------------------------------ driver (variational.c:19-143) */
void sfo_variational_params_default(sfo_variational_params_t *p) {
    if (!p) {
        fprintf(stderr, "Error optical_flow_params_default: argument is null\n");
        exit(1);
    }
    p->alpha = 1.0f; p->gamma = 0.71f; p->delta = 0.0f; p->sigma = 1.00f;
    p->niter_outer = 5; p->niter_inner = 1; p->niter_solver = 30; p->sor_omega = 1.9f;
}

void sfo_variational(sfo_image_t *wx, sfo_image_t *wy, const sfo_color_image_t *im1,
                     const sfo_color_image_t *im2, const sfo_variational_params_t *params, int sor_mode) {
    sfo_variational_params_t defaults;
    if (!params) {
        sfo_variational_params_default(&defaults);
        params = &defaults;
    }
    const float half_alpha = 0.5f * params->alpha;                 /* variational.c:114-116 */
    const float half_gamma_over3 = params->gamma * 0.5f / 3.0f;
    const float half_delta_over3 = params->delta * 0.5f / 3.0f;
    const int W = wx->width, H = wx->height, S = wx->stride;
    const size_t n = (size_t)S * H;

    sfo_image_t *du = sfo_image_new(W, H), *dv = sfo_image_new(W, H), *mask = sfo_image_new(W, H),
                *sh = sfo_image_new(W, H), *sv = sfo_image_new(W, H), *uu = sfo_image_new(W, H),
                *vv = sfo_image_new(W, H), *a11 = sfo_image_new(W, H), *a12 = sfo_image_new(W, H),
                *a22 = sfo_image_new(W, H), *b1 = sfo_image_new(W, H), *b2 = sfo_image_new(W, H);
    sfo_color_image_t *w_im2 = sfo_color_image_new(W, H), *Ix = sfo_color_image_new(W, H),
                      *Iy = sfo_color_image_new(W, H), *Iz = sfo_color_image_new(W, H),
                      *Ixx = sfo_color_image_new(W, H), *Ixy = sfo_color_image_new(W, H),
                      *Iyy = sfo_color_image_new(W, H), *Ixz = sfo_color_image_new(W, H),
                      *Iyz = sfo_color_image_new(W, H);
    sfo_image_t *dpsis_weight = sfo_compute_dpsis_weight(im1, 5.0f); /* variational.c:34 */

    for (int outer = 0; outer < params->niter_outer; outer++) {
        sfo_image_warp(w_im2, mask, im2, wx, wy, 1);
        sfo_get_derivatives(im1, w_im2, Ix, Iy, Iz, Ixx, Ixy, Iyy, Ixz, Iyz);
        memset(du->data, 0, n * sizeof(float));
        memset(dv->data, 0, n * sizeof(float));
        memcpy(uu->data, wx->data, n * sizeof(float));
        memcpy(vv->data, wy->data, n * sizeof(float));
        for (int inner = 0; inner < params->niter_inner; inner++) {
            sfo_compute_smoothness(sh, sv, uu, vv, dpsis_weight, half_alpha);
            sfo_compute_data_and_match(a11, a12, a22, b1, b2, mask, du, dv, Ix, Iy, Iz, Ixx, Ixy, Iyy, Ixz, Iyz,
                                       half_delta_over3, half_gamma_over3);
            sfo_sub_laplacian(b1, wx, sh, sv);
            sfo_sub_laplacian(b2, wy, sh, sv);
            sfo_sor_coupled(du, dv, a11, a12, a22, b1, b2, sh, sv, params->niter_solver, params->sor_omega, sor_mode);
            for (size_t k = 0; k < n; k++) {
                uu->data[k] = wx->data[k] + du->data[k];
                vv->data[k] = wy->data[k] + dv->data[k];
            }
        }
        memcpy(wx->data, uu->data, n * sizeof(float));
        memcpy(wy->data, vv->data, n * sizeof(float));
    }
    sfo_image_delete(du); sfo_image_delete(dv); sfo_image_delete(mask); sfo_image_delete(sh); sfo_image_delete(sv);
    sfo_image_delete(uu); sfo_image_delete(vv); sfo_image_delete(a11); sfo_image_delete(a12); sfo_image_delete(a22);
    sfo_image_delete(b1); sfo_image_delete(b2); sfo_image_delete(dpsis_weight);
    sfo_color_image_delete(w_im2); sfo_color_image_delete(Ix); sfo_color_image_delete(Iy); sfo_color_image_delete(Iz);
    sfo_color_image_delete(Ixx); sfo_color_image_delete(Ixy); sfo_color_image_delete(Iyy);
    sfo_color_image_delete(Ixz); sfo_color_image_delete(Iyz);
}


/* ---------------------------------------------------------------- on-disk outputs (checkers of sfgpu_write_flo / _pbm) */
/* writeFlowFile, epic_flow_extended/io.c:78-96: element-wise fwrite exactly as written there */
int sfo_write_flo(const char *filename, const sfo_image_t *flowx, const sfo_image_t *flowy) {
    FILE *stream = fopen(filename, "wb");
    if (stream == 0) return 1;
    const float help = 202021.25;
    fwrite(&help, sizeof(float), 1, stream);
    const int aXSize = flowx->width, aYSize = flowx->height;
    fwrite(&aXSize, sizeof(int), 1, stream);
    fwrite(&aYSize, sizeof(int), 1, stream);
    int y, x;
    for (y = 0; y < aYSize; y++)
        for (x = 0; x < aXSize; x++) {
            fwrite(&flowx->data[y * flowx->stride + x], sizeof(float), 1, stream);
            fwrite(&flowy->data[y * flowy->stride + x], sizeof(float), 1, stream);
        }
    fclose(stream);
    return 0;
}

/* slow_flow.cpp:893-905: occ_mat = 0.5*(occ + 1); convertTo(CV_8UC1, 255) (saturate_cast: round half to even);
 * imwrite(".pbm", PXM_BINARY=1) -- returns the 8-bit image; the P4 packing of cv::imwrite is pinned against python cv2
 * in tests/test_io.py */
int sfo_occlusion_to_u8(const sfo_image_t *occ, unsigned char *dst /* width*height, dense */) {
    int y, x;
    for (y = 0; y < occ->height; y++)
        for (x = 0; x < occ->width; x++) {
            const double v = nearbyint(0.5 * ((double)occ->data[y * occ->stride + x] + 1.0) * 255.0);
            dst[y * occ->width + x] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    return 0;
}
