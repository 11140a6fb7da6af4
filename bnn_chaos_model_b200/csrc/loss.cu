// K3: truncated-normal negative log-likelihood, forward + analytic backward, fused.
//
// Reference: safe_log_erf (/root/reference/spock_reg_model.py:323-335) and
// VarModel._lossfnc (:547-577).  Per label j in {0,1} of a system with prediction (mu, sd):
//   y_j <  9:  l = -(y-mu)^2/(2 sd^2) - log(sd) - sle((mu-4)/(sqrt2 sd))
//   y_j >= 9:  l =  sle((mu-9)/(sqrt2 sd))
//   non-finite l -> -100;   loss = -(l_0 + l_1)
#include "common.cuh"

namespace bnn {

// f_under(0) of safe_log_erf evaluated in fp32 (0.643278438654541f - 0.643250926022749f): the
// where-masked sum at :335 adds it to every x >= -1.
__device__ __constant__ const float kSleOffset = 2.7477741241455078e-05f;

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

__global__ void nll_kernel(const float2* __restrict__ mu_sd, const float2* __restrict__ y, int64_t B,
                           float* __restrict__ loss, float2* __restrict__ grad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float2 o = mu_sd[i], yy = y[i];
    float l0, l1, a0, a1, b0, b1;
    nll_terms(o.x, o.y, yy.x, l0, a0, b0);
    nll_terms(o.x, o.y, yy.y, l1, a1, b1);
    if (loss) loss[i] = -(l0 + l1);
    if (grad) grad[i] = make_float2(-(a0 + a1), -(b0 + b1));
}

// Deterministic sum of n floats by ONE block: fixed strided order, then a fixed tree.
__global__ void __launch_bounds__(1024) det_sum_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
    __shared__ float sh[1024];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 1024) acc += v[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

int launch_nll(const float* mu_sd, const float* y, int64_t B, float* loss, float* loss_sum, float* grad,
               cudaStream_t st) {
    const int threads = 256;
    const int64_t blocks = (B + threads - 1) / threads;
    nll_kernel<<<(unsigned)blocks, threads, 0, st>>>((const float2*)mu_sd, (const float2*)y, B, loss, (float2*)grad);
    BNN_CUDA(cudaGetLastError());
    if (loss_sum) {
        det_sum_kernel<<<1, 1024, 0, st>>>(loss, B, loss_sum);
        BNN_CUDA(cudaGetLastError());
    }
    return BNN_OK;
}

}  // namespace bnn

extern "C" int bnn_nll_fwd_bwd(const float* d_mu_sd, const float* d_y, int64_t B, float* d_loss_per_system,
                               float* d_loss_sum, float* d_grad, void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_mu_sd && d_y && B > 0, BNN_E_ARG, "bnn_nll_fwd_bwd: null pointer or B<=0");
    BNN_REQUIRE(!d_loss_sum || d_loss_per_system, BNN_E_ARG,
                "bnn_nll_fwd_bwd: d_loss_sum needs d_loss_per_system (the sum is a second, deterministic pass)");
    BNN_REQUIRE(B < (1ll << 38), BNN_E_ARG, "bnn_nll_fwd_bwd: B too large");
    return launch_nll(d_mu_sd, d_y, B, d_loss_per_system, d_loss_sum, d_grad, (cudaStream_t)stream);
}
