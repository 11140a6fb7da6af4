// Device functions of the truncated-normal negative log-likelihood (shared by loss.cu and train.cu).
//
// Reference: safe_log_erf (/root/reference/spock_reg_model.py:323-335) and VarModel._lossfnc (:547-577).
#pragma once
#include "common.cuh"

namespace bnn {

// f_under(0) of safe_log_erf evaluated in fp32 (0.643278438654541f - 0.643250926022749f): the
// where-masked sum at :335 adds it to every x >= -1.
constexpr float kSleOffset = 2.7477741241455078e-05f;

__device__ __forceinline__ float sle_dev(float x) {
    if (x < -1.0f) {
        // left-to-right fp32 evaluation of :330-331, no FMA contraction
        float t = __fmul_rn(0.485660082730562f, x);
        t = __fadd_rn(t, __fmul_rn(0.643278438654541f, expf(x)));
        t = __fadd_rn(t, __fmul_rn(0.00200084619923262f, __fmul_rn(__fmul_rn(x, x), x)));
        t = __fsub_rn(t, 0.643250926022749f);
        t = __fsub_rn(t, __fmul_rn(0.955350621183745f, __fmul_rn(x, x)));
        return t;  // + f_over(0) = log(1 + erf(0)) = 0
    }
    return __fadd_rn(kSleOffset, logf(__fadd_rn(1.0f, erff(x))));
}

__device__ __forceinline__ float sle_grad_dev(float x) {
    if (x < -1.0f)
        return 0.485660082730562f + 0.643278438654541f * expf(x) + 3.0f * 0.00200084619923262f * x * x -
               2.0f * 0.955350621183745f * x;
    return 1.1283791670955126f * expf(-x * x) / (1.0f + erff(x));
}

__device__ __forceinline__ void nll_terms(float mu, float sd, float y, float& l, float& dmu, float& dsd) {
    const float var = __fmul_rn(sd, sd);
    const float s2 = sqrtf(__fmul_rn(2.0f, var));  // torch.sqrt(2*var)
    if (y >= 9.0f) {
        const float b = __fdiv_rn(__fsub_rn(mu, 9.0f), s2);
        l = sle_dev(b);
        const float g = sle_grad_dev(b);
        dmu = g / s2;
        dsd = -g * b / sd;
    } else {
        const float a = __fdiv_rn(__fsub_rn(mu, 4.0f), s2);
        const float r = __fsub_rn(y, mu);
        l = __fdiv_rn(-__fmul_rn(r, r), __fmul_rn(2.0f, var));
        l = __fadd_rn(l, -logf(sd));
        l = __fadd_rn(l, -sle_dev(a));
        const float g = sle_grad_dev(a);
        dmu = r / var - g / s2;
        dsd = r * r / (var * sd) - 1.0f / sd + g * a / sd;
    }
    if (!isfinite(l)) {
        l = -100.0f;
        dmu = 0.f;
        dsd = 0.f;
    }
}

}  // namespace bnn
