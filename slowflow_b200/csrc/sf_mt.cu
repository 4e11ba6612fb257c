// sf_mt.cu -- multi-frame driver (Variational_MT).  PLACEHOLDER until the MT kernels land.
#include "sf_context.cuh"
#include <string.h>
using namespace sf;
extern "C" {
void sf_mt_params_default(sf_mt_params_t *p) { // slow_flow.cpp:64-128 (setDefault)
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->S = 2; p->layers = 1; p->p_scale = 0.9f;
    p->alpha = 4.0f; p->gamma = 6.0f; p->delta = 1.0f;
    p->dataterm = 1; p->smoothing = 1; p->one_direction = 0;
    for (int a = 0; a < SF_MT_MAX_REF; a++) { p->rho[a] = 1.0f; p->omega[a] = 1.0f; }
    p->rho[0] = 1.0f; p->rho[1] = 1.0f; p->omega[0] = 0.0f; p->omega[1] = 2.0f;
    p->robust_color = 1; p->robust_color_eps = 0.001f; p->robust_color_truncation = 0.5f;
    p->robust_grad = -1; p->robust_grad_eps = 0.001f; p->robust_grad_truncation = 0.5f;
    p->robust_reg = 1; p->robust_reg_eps = 0.001f; p->robust_reg_truncation = 0.5f;
    p->niter_alter = 10; p->niter_outer = 10; p->niter_inner = 1; p->niter_solver = 30; p->niter_graphc = 10;
    p->thres_outer = 1e-5f; p->thres_inner = 1e-5f; p->sor_omega = 1.9f;
    p->occlusion_reasoning = 1; p->occlusion_penalty = 0.1f; p->occlusion_alpha = 0.1f;
    p->graphcut_int_terms = 0; p->hbit = 1;
    for (int k = 0; k < 3; k++) { p->img_norm_avg[k] = 0.0f; p->img_norm_std[k] = 1.0f; }
}
int sfgpu_variational_mt(sfgpu_ctx *, image_t *, image_t *, const color_image_t *const *, const sf_mt_params_t *,
                         const color_image_t *, image_t *, float *) {
    set_error("sfgpu_variational_mt: not built yet");
    return SFGPU_ERR_UNSUPPORTED;
}
int sfgpu_normalize(sfgpu_ctx *, color_image_t *const *, int, sf_mt_params_t *) {
    set_error("sfgpu_normalize: not built yet");
    return SFGPU_ERR_UNSUPPORTED;
}
int sfgpu_get_mt_stats(sfgpu_ctx *c, sfgpu_mt_stats_t *out) {
    if (!c || !out) return SFGPU_ERR_ARG;
    *out = c->mt_stats;
    return SFGPU_OK;
}
}
