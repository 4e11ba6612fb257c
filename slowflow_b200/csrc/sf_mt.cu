// sf_mt.cu -- multi-frame driver (Variational_MT).  PLACEHOLDER until the MT kernels land.
#include "sf_context.cuh"
using namespace sf;
extern "C" {
void sf_mt_params_default(sf_mt_params_t *p) { (void)p; }
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
