/*
 * sf_oracle.h -- CPU restatement of Slow Flow's variational refinement hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the reported CPU baseline.  The product (libslowflow_gpu.so) never
 * links, loads or falls back to this code.
 *
 * Every function restates one function of the reference (JJanai/slowflow, paths relative to the
 * reference root) and cites it.  The arithmetic is written operation by operation in the
 * reference's evaluation order (fp32, no FMA contraction: build with -ffp-contract=off), so the
 * two-frame path is bit-identical to the reference's own objects built with its flags
 * (-O3 -msse4).  Pinning: tests/test_oracle_pin.py compares against oracle/_ref/libsf_ref.so
 * (the reference's unmodified sources compiled in place) where that library exists, and against
 * the golden vectors under tests/golden/ (generated from oracle/_ref by
 * tests/golden/make_golden.py) everywhere.
 *
 * One deliberate addition: SFO_SOR_REDBLACK, the "CPU red-black reference mode" that isolates the
 * SOR-ordering change of the GPU solver (same per-pixel update as solver.c, colour (i+j)&1 == 0
 * first, then colour 1, per sweep).
 */
#ifndef SF_ORACLE_H_
#define SF_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as the reference's image_t / color_image_t (epic_flow_extended/image.h:17-34). */
typedef struct sfo_image_s {
    int width, height, stride; /* stride = ceil4(width) floats */
    float *data;
} sfo_image_t;

typedef struct sfo_color_image_s {
    int width, height, stride;
    float *c1, *c2, *c3; /* planar, c2 = c1 + stride*height, c3 = c2 + stride*height */
} sfo_color_image_t;

/* Same layout as variational_params_t (epic_flow_extended/variational.h:15-24). */
typedef struct sfo_variational_params_s {
    float alpha, gamma, delta, sigma;
    int niter_outer, niter_inner, niter_solver;
    float sor_omega;
} sfo_variational_params_t;

enum { SFO_SOR_LEX = 0, SFO_SOR_REDBLACK = 1 };

/* image.c:17-33, 71-89 (uninitialised in the reference; zero-filled here so that padding is defined) */
sfo_image_t *sfo_image_new(int width, int height);
void sfo_image_delete(sfo_image_t *im);
sfo_color_image_t *sfo_color_image_new(int width, int height);
void sfo_color_image_delete(sfo_color_image_t *im);

/* image.c:400-526 -- order-1 and order-2 separable convolutions, replicate borders */
void sfo_convolve_horiz(sfo_image_t *dst, const sfo_image_t *src, int order, const float *coeffs);
void sfo_convolve_vert(sfo_image_t *dst, const sfo_image_t *src, int order, const float *coeffs);

/* variational_aux.c:18-52 ; variational_aux_mt.cpp:722-756 (factor) */
void sfo_image_warp(sfo_color_image_t *dst, sfo_image_t *mask, const sfo_color_image_t *src,
                    const sfo_image_t *wx, const sfo_image_t *wy, int factor);
/* variational_aux.c:55-78 */
void sfo_get_derivatives(const sfo_color_image_t *im1, const sfo_color_image_t *im2,
                         sfo_color_image_t *dx, sfo_color_image_t *dy, sfo_color_image_t *dt,
                         sfo_color_image_t *dxx, sfo_color_image_t *dxy, sfo_color_image_t *dyy,
                         sfo_color_image_t *dxt, sfo_color_image_t *dyt);
/* variational_aux.c:84-149 */
void sfo_compute_smoothness(sfo_image_t *dst_horiz, sfo_image_t *dst_vert, const sfo_image_t *uu,
                            const sfo_image_t *vv, const sfo_image_t *dpsis_weight, float half_alpha);
/* variational_aux.c:153-180 */
void sfo_sub_laplacian(sfo_image_t *dst, const sfo_image_t *src, const sfo_image_t *weight_horiz,
                       const sfo_image_t *weight_vert);
/* variational_aux.c:183-209 */
sfo_image_t *sfo_compute_dpsis_weight(const sfo_color_image_t *im, float coef);
/* variational_aux.c:215-302 */
void sfo_compute_data_and_match(sfo_image_t *a11, sfo_image_t *a12, sfo_image_t *a22, sfo_image_t *b1,
                                sfo_image_t *b2, const sfo_image_t *mask, const sfo_image_t *du,
                                const sfo_image_t *dv, const sfo_color_image_t *Ix, const sfo_color_image_t *Iy,
                                const sfo_color_image_t *Iz, const sfo_color_image_t *Ixx,
                                const sfo_color_image_t *Ixy, const sfo_color_image_t *Iyy,
                                const sfo_color_image_t *Ixz, const sfo_color_image_t *Iyz,
                                float half_delta_over3, float half_gamma_over3);
/* solver.c:63-399 (mode SFO_SOR_LEX) and its red-black re-ordering (SFO_SOR_REDBLACK).
 * Like the reference, overwrites a11,a12,a22 with the inverted 2x2 blocks. */
void sfo_sor_coupled(sfo_image_t *du, sfo_image_t *dv, sfo_image_t *a11, sfo_image_t *a12, sfo_image_t *a22,
                     const sfo_image_t *b1, const sfo_image_t *b2, const sfo_image_t *dpsis_horiz,
                     const sfo_image_t *dpsis_vert, int iterations, float omega, int mode);
/* solver.c:17-57 */
void sfo_sor_coupled_readable(sfo_image_t *du, sfo_image_t *dv, const sfo_image_t *a11, const sfo_image_t *a12,
                              const sfo_image_t *a22, const sfo_image_t *b1, const sfo_image_t *b2,
                              const sfo_image_t *dpsis_horiz, const sfo_image_t *dpsis_vert, int iterations,
                              float omega);

/* variational.c:85-98 */
void sfo_variational_params_default(sfo_variational_params_t *params);
/* variational.c:101-143 + compute_one_level :19-82.  params may be NULL (defaults). */
void sfo_variational(sfo_image_t *wx, sfo_image_t *wy, const sfo_color_image_t *im1,
                     const sfo_color_image_t *im2, const sfo_variational_params_t *params, int sor_mode);

#ifdef __cplusplus
}
#endif
#endif
