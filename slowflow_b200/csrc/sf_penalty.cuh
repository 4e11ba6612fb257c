// sf_penalty.cuh -- the robust penalties of penalty_functions/*.h as device functions.
//
// The reference has two code paths per penalty with different rounding: the v4sf overloads (fp32,
// used by the data terms, variational_aux_mt.cpp:166-634, and the occlusion costs :815-824) and the
// scalar overloads (computed in double where epsilon_sq is a double member, used by
// compute_smoothness :65,90).  Both are reproduced as written, including the unmatched
// Geman-McClure psi / psi' pair (geman_mcclure.h:20-38) and the truncated-L1 compare of
// sqrt(s^2) against tau (trunc_modified_l1_norm.h:20-54).
#pragma once
#include "sf_internal.cuh"
#include "sf_pack.cuh"

namespace sf {

// psi'(s^2), v4sf overloads (fp32).  Quotients go through MUFU.RCP / MUFU.RSQ plus one Newton step (within 1 ulp of the
// reference's divps / sqrtps chain) instead of IEEE division sequences: this sits in the innermost per-term arithmetic.
__device__ __forceinline__ float penalty_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ float penalty_half_rsqrt(float x) { // 1 / (2 sqrt(x)), x > 0
    const float r = rsqrtf(x);
    const float h = 0.5f * r;
    return fmaf(0.5f * h, fmaf(-x * r, r, 1.0f), h); // one Newton step on rsqrt, r' = r + r/2 (1 - x r^2), halved
}
// TYPE: the penalty as a compile-time functor id (SF_ROBUST_*; the multi-frame terms kernel is instantiated per pair
// of ids), or -1 for the run-time switch below
template <int TYPE>
__device__ __forceinline__ float penalty_deriv_vt(const Penalty &p, float xsq);
__device__ __forceinline__ float penalty_deriv_v(const Penalty &p, float xsq) {
    switch (p.type) {
    case SF_ROBUST_QUADRATIC: return 1.0f;
    case SF_ROBUST_LORENTZIAN: return penalty_rcp(2.0f * p.eps_sq_f + xsq);
    case SF_ROBUST_GEMAN_MCCLURE: {
        float t = p.eps_sq_f + xsq;
        t = t * t;
        return (p.eps_sq_f + 2.0f * xsq) * penalty_rcp(t);
    }
    case SF_ROBUST_TRUNC_MODL1: {
        if (sqrtf(xsq) > p.trunc) return 0.0f;
        return penalty_half_rsqrt(xsq + p.eps_sq_f);
    }
    default: return penalty_half_rsqrt(xsq + p.eps_sq_f); // ModifiedL1Norm
    }
}

template <int TYPE>
__device__ __forceinline__ float penalty_deriv_vt(const Penalty &p, float xsq) {
    if constexpr (TYPE == SF_ROBUST_QUADRATIC) {
        return 1.0f;
    } else if constexpr (TYPE == SF_ROBUST_LORENTZIAN) {
        return penalty_rcp(2.0f * p.eps_sq_f + xsq);
    } else if constexpr (TYPE == SF_ROBUST_GEMAN_MCCLURE) {
        float t = p.eps_sq_f + xsq;
        t = t * t;
        return (p.eps_sq_f + 2.0f * xsq) * penalty_rcp(t);
    } else if constexpr (TYPE == SF_ROBUST_TRUNC_MODL1) {
        if (sqrtf(xsq) > p.trunc) return 0.0f;
        return penalty_half_rsqrt(xsq + p.eps_sq_f);
    } else if constexpr (TYPE == SF_ROBUST_MODL1) {
        return penalty_half_rsqrt(xsq + p.eps_sq_f);
    } else {
        return penalty_deriv_v(p, xsq);
    }
}

// psi'(s^2) of TWO pixels on the packed pipe (multi-frame terms kernel, two columns per thread): the same formulas with
// the MUFU seeds taken per half and the Newton steps / products as FFMA2 / FMUL2
template <int TYPE>
__device__ __forceinline__ p64 penalty_deriv_v2(const Penalty &p, p64 xsq) {
    if constexpr (TYPE == SF_ROBUST_QUADRATIC) {
        return splat2(1.0f);
    } else if constexpr (TYPE == SF_ROBUST_LORENTZIAN) {
        return rcp2(add2(xsq, splat2(2.0f * p.eps_sq_f)));
    } else if constexpr (TYPE == SF_ROBUST_GEMAN_MCCLURE) {
        const p64 e = splat2(p.eps_sq_f);
        p64 t = add2(e, xsq);
        t = mul2(t, t);
        return mul2(fma2(splat2(2.0f), xsq, e), rcp2(t));
    } else if constexpr (TYPE == SF_ROBUST_MODL1) {
        const p64 x = add2(xsq, splat2(p.eps_sq_f));
        const p64 r = pk(rsqrtf(lo_of(x)), rsqrtf(hi_of(x)));
        const p64 h = mul2(r, splat2(0.5f));
        // one Newton step on rsqrt, halved: h + h/2 (1 - x r^2)
        const p64 e = fma2(mul2(mul2(x, splat2(-1.0f)), r), r, splat2(1.0f));
        return fma2(mul2(h, splat2(0.5f)), e, h);
    } else { // truncated L1 (a compare per pixel) and the run-time switch: per half
        return pk(penalty_deriv_vt<TYPE>(p, lo_of(xsq)), penalty_deriv_vt<TYPE>(p, hi_of(xsq)));
    }
}

// psi'(s^2), scalar overloads.  The reference evaluates these in double where epsilon_sq is a double member and
// rounds the result to float; the fp32 evaluation here is within 2 ulp of that (kept off the FP64 pipe).
__device__ __forceinline__ float penalty_deriv_s(const Penalty &p, float xsq) {
    switch (p.type) {
    case SF_ROBUST_QUADRATIC: return 1.0f;
    case SF_ROBUST_LORENTZIAN: return __fdiv_rn(1.0f, 2.0f * p.eps_sq_f + xsq);
    case SF_ROBUST_GEMAN_MCCLURE: {
        float t = p.eps_sq_f + xsq;
        t = t * t;
        return __fdiv_rn(p.eps_sq_f + 2.0f * xsq, t);
    }
    case SF_ROBUST_TRUNC_MODL1: {
        if (sqrtf(xsq) > p.trunc) return 0.0f;
        return 1.0f / (2.0f * sqrtf(xsq + p.eps_sq_f));
    }
    default: return __fdiv_rn(1.0f, 2.0f * sqrtf(xsq + p.eps_sq_f));
    }
}

// psi(s^2), v4sf overloads (occlusion data costs)
__device__ __forceinline__ float penalty_apply_v(const Penalty &p, float xsq) {
    switch (p.type) {
    case SF_ROBUST_QUADRATIC: return xsq;
    case SF_ROBUST_LORENTZIAN: return (float)log(1.0 + 0.5 * (double)xsq / p.eps_sq_d);
    case SF_ROBUST_GEMAN_MCCLURE: return xsq / ((xsq + 1.0f) * (xsq + 1.0f));
    case SF_ROBUST_TRUNC_MODL1: {
        if (sqrtf(xsq) > p.trunc) return sqrtf(p.trunc + p.eps_sq_f);
        return sqrtf(xsq + p.eps_sq_f);
    }
    default: return sqrtf(xsq + p.eps_sq_f);
    }
}

// smoothness diffusivity of one edge.
//   type < 0 : two-frame form (w+w')*half_alpha / sqrt(s^2 + 1e-6) (variational_aux.c:124).  The reference
//              takes the sqrt and the quotient in double and rounds once; here both are IEEE fp32, which
//              differs by at most 1-2 ulp (2e-7 relative) and keeps the kernel off the FP64 pipe.
//   else     : (w+w')*alpha * psi'_reg(s^2) through the scalar overload (variational_aux_mt.cpp:65)
__device__ __forceinline__ float smooth_weight(const Penalty &reg, float ww, float alpha_factor, float ssq) {
    if (reg.type < 0) {
        const float eps_smooth = 0.001f * 0.001f;
        return __fdiv_rn(ww * alpha_factor, sqrtf(ssq + eps_smooth));
    }
    return ww * alpha_factor * penalty_deriv_s(reg, ssq);
}

} // namespace sf
