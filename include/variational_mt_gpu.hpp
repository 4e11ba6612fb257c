// variational_mt_gpu.hpp -- header-only C++ shim with the shape of the reference class Variational_MT
// (epic_flow_extended/variational_mt.h:23-71) over the C ABI of slowflow_gpu.h, so that the call sites
// slow_flow.cpp:875-888 and :1018-1023 compile unchanged apart from the include:
//
//     Variational_MT minimzer_f;
//     minimzer_f.setChannelWeights(channel_weights);
//     minimzer_f.variational(wx, wy, im, thread_params);     // thread_params: the reference's ParameterList
//
// `Params` is any type with the ParameterList lookups the reference uses (utils/parameter_list.h:20-143):
//     bool exists(std::string);  std::string parameter(const char*);
//     template<class T> T parameter(std::string key, std::string default_value);
//     void insert(std::string key, std::string value, bool overwrite);
// cv::Point2f is replaced by a 2-float POD (OpenCV is not a dependency of this library); it converts implicitly
// to anything constructible from (float, float).
#ifndef VARIATIONAL_MT_GPU_HPP_
#define VARIATIONAL_MT_GPU_HPP_

#include <stdio.h>
#include <stdlib.h>

#include <sstream>
#include <string>

#include "slowflow_gpu.h"

struct sf_point2f {
    float x, y;
    template <class P> operator P() const { return P(x, y); }
};

// the keys of variational_mt.cpp:173-192, 250-251, 533-568 -> POD
template <class Params> inline void sf_mt_params_from_list(Params &p, sf_mt_params_t *m) {
    sf_mt_params_default(m);
    m->S = p.template parameter<int>("slow_flow_S", "2");
    m->layers = p.template parameter<int>("slow_flow_layers", "1");
    m->p_scale = p.template parameter<float>("slow_flow_p_scale", "0.9");
    m->alpha = p.template parameter<float>("slow_flow_alpha", "4");
    m->gamma = p.template parameter<float>("slow_flow_gamma", "6");
    m->delta = p.template parameter<float>("slow_flow_delta", "1");
    m->dataterm = p.template parameter<bool>("slow_flow_dataterm", "1");
    m->smoothing = p.template parameter<int>("slow_flow_smoothing", "0");
    m->one_direction = (p.exists("slow_flow_method") && std::string(p.parameter("slow_flow_method")) == "forward") ? 1 : 0;
    for (int a = 0; a < m->S - 1 && a < SF_MT_MAX_REF; a++) {
        std::ostringstream ko, kr;
        ko << "slow_flow_omega_" << a;
        kr << "slow_flow_rho_" << a;
        m->omega[a] = p.template parameter<float>(ko.str(), "1.0");
        m->rho[a] = p.template parameter<float>(kr.str(), "1.0");
    }
    m->robust_color = p.template parameter<int>("slow_flow_robust_color", "1");
    m->robust_color_eps = p.template parameter<float>("slow_flow_robust_color_eps", "0.001");
    m->robust_color_truncation = p.template parameter<float>("slow_flow_robust_color_truncation", "0.5");
    if (p.exists("slow_flow_robust_grad")) { // else: the colour settings are reused (variational_mt.cpp:556-557)
        m->robust_grad = p.template parameter<int>("slow_flow_robust_grad", "1");
        m->robust_grad_eps = p.template parameter<float>("slow_flow_robust_grad_eps", "0.001");
        m->robust_grad_truncation = p.template parameter<float>("slow_flow_robust_grad_truncation", "0.5");
    } else {
        m->robust_grad = -1;
    }
    m->robust_reg = p.template parameter<int>("slow_flow_robust_reg", "1");
    m->robust_reg_eps = p.template parameter<float>("slow_flow_robust_reg_eps", "0.001");
    m->robust_reg_truncation = p.template parameter<float>("slow_flow_robust_reg_truncation", "0.5");
    m->niter_alter = p.template parameter<int>("slow_flow_niter_alter", "1");
    m->niter_outer = p.template parameter<int>("slow_flow_niter_outer", "10");
    m->niter_inner = p.template parameter<int>("slow_flow_niter_inner", "1");
    m->niter_solver = p.template parameter<int>("slow_flow_niter_solver", "30");
    m->niter_graphc = p.template parameter<int>("slow_flow_niter_graphc", "10");
    m->thres_outer = p.template parameter<float>("slow_flow_thres_outer", "1e-5");
    m->thres_inner = p.template parameter<float>("slow_flow_thres_inner", "1e-5");
    m->sor_omega = p.template parameter<float>("slow_flow_sor_omega", "1.9");
    m->occlusion_reasoning = p.template parameter<bool>("slow_flow_occlusion_reasoning", "0");
    m->occlusion_penalty = p.template parameter<float>("slow_flow_occlusion_penalty", "1.0");
    m->occlusion_alpha = p.template parameter<float>("slow_flow_occlusion_alpha", "0.5");
    m->hbit = p.template parameter<bool>("16bit", "0");
    const char *avg[3] = {"slow_flow_img_norm_avg_1", "slow_flow_img_norm_avg_2", "slow_flow_img_norm_avg_3"};
    const char *sd[3] = {"slow_flow_img_norm_std_1", "slow_flow_img_norm_std_2", "slow_flow_img_norm_std_3"};
    for (int k = 0; k < 3; k++) {
        m->img_norm_avg[k] = (float)p.template parameter<double>(avg[k], "0");
        m->img_norm_std[k] = (float)p.template parameter<double>(sd[k], "1");
    }
}

// One context per host thread and device, shared by every Variational_MT object and normalize() call of that thread:
// the unmodified call sites construct a minimiser per jet (slow_flow.cpp:875, :1018) and call normalize() per window
// (:673); with a context per object each window would pay context creation and a ~1 GB workspace allocation.  The
// contexts are destroyed when the thread exits.
inline sfgpu_ctx *sf_shim_thread_context() {
    struct Cache {
        enum { MAXDEV = 64 };
        sfgpu_ctx *ctx[MAXDEV];
        Cache() { for (int k = 0; k < MAXDEV; k++) ctx[k] = NULL; }
        ~Cache() { for (int k = 0; k < MAXDEV; k++) if (ctx[k]) sfgpu_destroy(ctx[k]); }
    };
    static thread_local Cache cache;
    int dev = sfgpu_get_device();
    if (dev < 0 || dev >= Cache::MAXDEV) dev = 0;
    if (!cache.ctx[dev] && sfgpu_create(dev, NULL, &cache.ctx[dev]) != SFGPU_OK) return NULL;
    return cache.ctx[dev];
}

class Variational_MT {
public:
    Variational_MT() : one_direction(false), channel_w_(NULL) { occ_.data = NULL; }
    ~Variational_MT() {
        if (occ_.data) free(occ_.data);
    }
    void setChannelWeights(color_image_t *weights) { channel_w_ = weights; }
    image_t *getOcclusions() { return occ_.data ? &occ_ : NULL; }

    // Variational_MT::variational (variational_mt.cpp:526): refines wx, wy in place, returns the last outer
    // iteration's mean absolute change; like the reference it writes the key "final" (:527, :764)
    template <class Params> sf_point2f variational(image_t *wx, image_t *wy, color_image_t *const *im, Params &params) {
        params.insert("final", "0", true);
        sf_mt_params_t m;
        sf_mt_params_from_list(params, &m);
        // the reference latches the member once the key says "forward" (variational_mt.cpp:548-549): it stays set for
        // later calls on the same object even if their parameters lack the key
        if (m.one_direction) one_direction = true;
        if (one_direction) m.one_direction = 1;
        // on the calling thread's current device: one host thread per device (slow_flow.cpp:706)
        sfgpu_ctx *ctx = sf_shim_thread_context();
        if (!ctx) fail();
        if (occ_.data) free(occ_.data);
        occ_.width = wx->width; occ_.height = wx->height; occ_.stride = wx->stride;
        occ_.data = (float *)calloc((size_t)wx->stride * wx->height, sizeof(float));
        float avg[2] = {0.f, 0.f};
        if (sfgpu_variational_mt(ctx, wx, wy, im, &m, channel_w_, &occ_, avg) != SFGPU_OK) fail();
        sf_point2f r = {avg[0], avg[1]};
        return r;
    }

    bool one_direction;

private:
    static void fail() { // the reference's error convention: message + exit(1) (image.c:19-30)
        fprintf(stderr, "error in Variational_MT::variational(): %s\n", sfgpu_last_error());
        exit(1);
    }
    color_image_t *channel_w_;
    image_t occ_;
};

// normalize() of variational_mt.cpp:17-85 with the reference's signature shape
template <class Params> inline void normalize(color_image_t **seq, unsigned F, Params &params) {
    sfgpu_ctx *ctx = sf_shim_thread_context();
    sf_mt_params_t m;
    sf_mt_params_from_list(params, &m);
    if (!ctx || sfgpu_normalize(ctx, seq, (int)F, &m) != SFGPU_OK) {
        fprintf(stderr, "error in normalize(): %s\n", sfgpu_last_error());
        exit(1);
    }
    const char *avg[3] = {"slow_flow_img_norm_avg_1", "slow_flow_img_norm_avg_2", "slow_flow_img_norm_avg_3"};
    const char *sd[3] = {"slow_flow_img_norm_std_1", "slow_flow_img_norm_std_2", "slow_flow_img_norm_std_3"};
    for (int k = 0; k < 3; k++) {
        std::ostringstream a, s;
        a << m.img_norm_avg[k];
        s << m.img_norm_std[k];
        params.insert(avg[k], a.str(), true);
        params.insert(sd[k], s.str(), true);
    }
}

#endif // VARIATIONAL_MT_GPU_HPP_
