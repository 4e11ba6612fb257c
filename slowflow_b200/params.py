"""ctypes mirrors of ``variational_params_t`` (variational.h:15-24) and ``sf_mt_params_t``
(include/slowflow_gpu.h; the ParameterList keys of SURVEY Appendix B)."""
import ctypes as C

SF_MT_MAX_REF = 8
ROBUST = {"quadratic": 0, "modl1": 1, "lorentzian": 2, "trunc_modl1": 3, "geman_mcclure": 4}


class VariationalParams(C.Structure):
    _fields_ = [("alpha", C.c_float), ("gamma", C.c_float), ("delta", C.c_float), ("sigma", C.c_float),
                ("niter_outer", C.c_int), ("niter_inner", C.c_int), ("niter_solver", C.c_int),
                ("sor_omega", C.c_float)]


def variational_params_default():
    """variational.c:85-98"""
    return VariationalParams(1.0, 0.71, 0.0, 1.0, 5, 1, 30, 1.9)


class MTParams(C.Structure):
    _fields_ = [
        ("S", C.c_int), ("layers", C.c_int), ("p_scale", C.c_float),
        ("alpha", C.c_float), ("gamma", C.c_float), ("delta", C.c_float),
        ("dataterm", C.c_int), ("smoothing", C.c_int), ("one_direction", C.c_int),
        ("rho", C.c_float * SF_MT_MAX_REF), ("omega", C.c_float * SF_MT_MAX_REF),
        ("robust_color", C.c_int), ("robust_color_eps", C.c_float), ("robust_color_truncation", C.c_float),
        ("robust_grad", C.c_int), ("robust_grad_eps", C.c_float), ("robust_grad_truncation", C.c_float),
        ("robust_reg", C.c_int), ("robust_reg_eps", C.c_float), ("robust_reg_truncation", C.c_float),
        ("niter_alter", C.c_int), ("niter_outer", C.c_int), ("niter_inner", C.c_int), ("niter_solver", C.c_int),
        ("niter_graphc", C.c_int),
        ("thres_outer", C.c_float), ("thres_inner", C.c_float), ("sor_omega", C.c_float),
        ("occlusion_reasoning", C.c_int), ("occlusion_penalty", C.c_float), ("occlusion_alpha", C.c_float),
        ("graphcut_int_terms", C.c_int), ("hbit", C.c_int),
        ("img_norm_avg", C.c_float * 3), ("img_norm_std", C.c_float * 3),
    ]


def mt_params_default():
    """slow_flow.cpp:64-128 (setDefault) -- same values as sf_mt_params_default() in the library."""
    p = MTParams()
    p.S, p.layers, p.p_scale = 2, 1, 0.9
    p.alpha, p.gamma, p.delta = 4.0, 6.0, 1.0
    p.dataterm, p.smoothing, p.one_direction = 1, 1, 0
    for a in range(SF_MT_MAX_REF):
        p.rho[a], p.omega[a] = 1.0, 1.0
    p.rho[0], p.rho[1], p.omega[0], p.omega[1] = 1.0, 1.0, 0.0, 2.0
    p.robust_color, p.robust_color_eps, p.robust_color_truncation = 1, 0.001, 0.5
    p.robust_grad, p.robust_grad_eps, p.robust_grad_truncation = -1, 0.001, 0.5
    p.robust_reg, p.robust_reg_eps, p.robust_reg_truncation = 1, 0.001, 0.5
    p.niter_alter, p.niter_outer, p.niter_inner, p.niter_solver, p.niter_graphc = 10, 10, 1, 30, 10
    p.thres_outer, p.thres_inner, p.sor_omega = 1e-5, 1e-5, 1.9
    p.occlusion_reasoning, p.occlusion_penalty, p.occlusion_alpha = 1, 0.1, 0.1
    p.graphcut_int_terms, p.hbit = 0, 1
    for k in range(3):
        p.img_norm_avg[k], p.img_norm_std[k] = 0.0, 1.0
    return p
