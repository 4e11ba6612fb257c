// oracle_api.cpp -- C entry points over the reference's UNMODIFIED multi-frame driver
// (epic_flow_extended/variational_mt.cpp + variational_aux_mt.cpp + utils/parameter_list.cpp, compiled in
// place into oracle/_ref/libsf_ref.so).  TEST INFRASTRUCTURE ONLY.
//
// Fills a ParameterList from the POD of include/slowflow_gpu.h (the same keys the reference reads,
// SURVEY Appendix B) and calls normalize() / Variational_MT::variational() exactly like
// slow_flow.cpp:673,875-888 does.
#include <sstream>
#include <string>

#include "variational_mt.h" // the reference's own header (pulls image.h, parameter_list.h)

#include "variational.h"    // reference's variational_params_t (so the ABI header reuses it)
#include "../../include/slowflow_gpu.h" // sf_mt_params_t POD

extern "C" {
int sf_gco_int_terms = 0;
int sf_gco_calls = 0;
void sf_ref_set_sor_mode(int mode);
}

template <typename T> static std::string str(T v) {
    std::ostringstream s;
    s.precision(9);
    s << v;
    return s.str();
}

static void fill(ParameterList &p, const sf_mt_params_t *m) {
    const int ref = m->S - 1;
    p.insert("verbose", "0", true);
    p.verbose = "0000000000";
    p.insert("slow_flow_S", str(m->S), true);
    p.insert("slow_flow_layers", str(m->layers), true);
    p.insert("slow_flow_p_scale", str(m->p_scale), true);
    p.insert("slow_flow_alpha", str(m->alpha), true);
    p.insert("slow_flow_gamma", str(m->gamma), true);
    p.insert("slow_flow_delta", str(m->delta), true);
    p.insert("slow_flow_dataterm", str(m->dataterm), true);
    p.insert("slow_flow_smoothing", str(m->smoothing), true);
    p.insert("slow_flow_method", m->one_direction ? "forward" : "symmetric", true);
    for (int a = 0; a < ref; a++) {
        p.insert("slow_flow_rho_" + str(a), str(m->rho[a]), true);
        p.insert("slow_flow_omega_" + str(a), str(m->omega[a]), true);
    }
    p.insert("slow_flow_robust_color", str(m->robust_color), true);
    p.insert("slow_flow_robust_color_eps", str(m->robust_color_eps), true);
    p.insert("slow_flow_robust_color_truncation", str(m->robust_color_truncation), true);
    if (m->robust_grad >= 0) {
        p.insert("slow_flow_robust_grad", str(m->robust_grad), true);
        p.insert("slow_flow_robust_grad_eps", str(m->robust_grad_eps), true);
        p.insert("slow_flow_robust_grad_truncation", str(m->robust_grad_truncation), true);
    }
    p.insert("slow_flow_robust_reg", str(m->robust_reg), true);
    p.insert("slow_flow_robust_reg_eps", str(m->robust_reg_eps), true);
    p.insert("slow_flow_robust_reg_truncation", str(m->robust_reg_truncation), true);
    p.insert("slow_flow_niter_alter", str(m->niter_alter), true);
    p.insert("slow_flow_niter_outer", str(m->niter_outer), true);
    p.insert("slow_flow_niter_inner", str(m->niter_inner), true);
    p.insert("slow_flow_niter_solver", str(m->niter_solver), true);
    p.insert("slow_flow_niter_graphc", str(m->niter_graphc), true);
    p.insert("slow_flow_thres_outer", str(m->thres_outer), true);
    p.insert("slow_flow_thres_inner", str(m->thres_inner), true);
    p.insert("slow_flow_sor_omega", str(m->sor_omega), true);
    p.insert("slow_flow_occlusion_reasoning", str(m->occlusion_reasoning), true);
    p.insert("slow_flow_occlusion_penalty", str(m->occlusion_penalty), true);
    p.insert("slow_flow_occlusion_alpha", str(m->occlusion_alpha), true);
    p.insert("16bit", str(m->hbit), true);
    p.insert("sigma", "0", true);
    for (int k = 0; k < 3; k++) {
        p.insert("slow_flow_img_norm_avg_" + str(k + 1), str(m->img_norm_avg[k]), true);
        p.insert("slow_flow_img_norm_std_" + str(k + 1), str(m->img_norm_std[k]), true);
    }
}

extern "C" {

// normalize() of variational_mt.cpp:17-85, in place; publishes avg/std into the POD
int sf_ref_normalize(color_image_t *const *seq, int F, sf_mt_params_t *m) {
    ParameterList p;
    fill(p, m);
    normalize(const_cast<color_image_t **>(seq), (u_int32_t)F, p);
    for (int k = 0; k < 3; k++) {
        m->img_norm_avg[k] = p.parameter<float>("slow_flow_img_norm_avg_" + str(k + 1), "0");
        m->img_norm_std[k] = p.parameter<float>("slow_flow_img_norm_std_" + str(k + 1), "1");
    }
    return 0;
}

// Variational_MT::variational (variational_mt.cpp:526); stats = {sor calls unavailable here: 0, graph-cut calls}
int sf_ref_variational_mt(image_t *wx, image_t *wy, color_image_t *const *im, const sf_mt_params_t *m,
                          const color_image_t *channel_w, image_t *occlusions_out, float avg_change[2], int sor_mode,
                          int *stats) {
    ParameterList p;
    fill(p, m);
    sf_gco_int_terms = m->graphcut_int_terms;
    sf_gco_calls = 0;
    sf_ref_set_sor_mode(sor_mode);
    Variational_MT solver;
    if (channel_w) solver.setChannelWeights(const_cast<color_image_t *>(channel_w));
    if (m->one_direction) solver.one_direction = true;
    const Point2f r = solver.variational(wx, wy, im, p);
    sf_ref_set_sor_mode(0);
    if (avg_change) { avg_change[0] = r.x; avg_change[1] = r.y; }
    if (occlusions_out && solver.getOcclusions()) {
        image_t *o = solver.getOcclusions();
        if (o->width == occlusions_out->width && o->height == occlusions_out->height)
            memcpy(occlusions_out->data, o->data, sizeof(float) * o->stride * o->height);
    }
    if (stats) { stats[0] = 0; stats[1] = sf_gco_calls; }
    return 0;
}

// ---- test hooks for the restated third-party pieces (pinned by tests/test_oracle_pin.py)
// interleaved float image (rows x cols x cn) -> optional GaussianBlur(sigma) -> optional resize(dst_rows x dst_cols)
int sf_ref_blur_resize(const float *src, int rows, int cols, int cn, double sigma, int dst_rows, int dst_cols, float *dst) {
    Mat m(rows, cols, CV_MAKETYPE(CV_32F, cn));
    memcpy(m.data, src, sizeof(float) * rows * cols * cn);
    if (sigma > 0) GaussianBlur(m, m, Size(0, 0), sigma, sigma, BORDER_REPLICATE);
    if (dst_rows != rows || dst_cols != cols) resize(m, m, Size(dst_cols, dst_rows), 0, 0, INTER_LINEAR);
    memcpy(dst, m.data, sizeof(float) * dst_rows * dst_cols * cn);
    return 0;
}
// binary labelling through the gco stand-in: data costs d0/d1 per site, Potts weight alpha
int sf_ref_mincut(int w, int h, const float *d0, const float *d1, float alpha, int int_terms, int *labels) {
    sf_gco_int_terms = int_terms;
    GCoptimizationGridGraph gc(w, h, 2);
    for (int p = 0; p < w * h; p++) { gc.setDataCost(p, 0, d0[p]); gc.setDataCost(p, 1, d1[p]); }
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) gc.setSmoothCost(a, b, a != b ? alpha : 0.0f);
    gc.expansion(10);
    for (int p = 0; p < w * h; p++) labels[p] = gc.whatLabel(p);
    return 0;
}

} // extern "C"
