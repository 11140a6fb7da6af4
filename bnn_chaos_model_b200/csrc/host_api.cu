// Host-buffer convenience entry point and the FFMA-peak micro-kernel.
#include "common.cuh"

namespace bnn {

// 16 independent accumulators per thread; a, b stay in registers.  packed: 8 fma.rn.f32x2
// per iteration, scalar: 16 fma.rn.f32.  Measures the FP32 roofline denominator live.
template <bool PACKED>
__global__ void __launch_bounds__(256) ffma_peak_kernel(int64_t iters, float seed_a, float seed_b, float* sink) {
    float a = seed_a + threadIdx.x * 1e-9f, b = seed_b;
    if (PACKED) {
        u64 acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = pack2(0.1f * i, 0.2f * i);
        const u64 aa = pack2(a, a * 0.999f), bb = pack2(b, b * 1.001f);
        for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma2(acc[i], aa, bb);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            unpack2(acc[i], lo, hi);
            s += lo + hi;
        }
        if (s == 123.456f) sink[0] = s;
    } else {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.1f * i;
        for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
        if (s == 123.456f) sink[0] = s;
    }
}

}  // namespace bnn

extern "C" {

int bnn_ffma_peak(int32_t packed, int64_t iters, float* d_sink, int64_t* flops_out, void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_sink && iters > 0, BNN_E_ARG, "bnn_ffma_peak: null sink or iters<=0");
    int dev = 0, sms = 0;
    BNN_CUDA(cudaGetDevice(&dev));
    BNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256;
    if (packed)
        ffma_peak_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, 0.999f, 1e-3f, d_sink);
    else
        ffma_peak_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, 0.999f, 1e-3f, d_sink);
    BNN_CUDA(cudaGetLastError());
    if (flops_out) *flops_out = 2ll * blocks * threads * iters * 16;
    return BNN_OK;
}

static size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t bnn_multiswag_host_scratch_bytes(const bnn_model_config* cfg, int64_t n_systems, int64_t n_units) {
    if (bnn::validate_config(cfg) != BNN_OK || n_systems <= 0 || n_units <= 0) return 0;
    const size_t d = bnn::FlatLayout(cfg->n_features).d, P = bnn::PackedLayout(bnn::live_columns(cfg).n, cfg->n_features).P;
    return up256((size_t)n_systems * cfg->n_times * cfg->n_features * 4) + up256((size_t)n_units * d * 4) +
           up256((size_t)n_units * P * 4) + up256((size_t)n_units * n_systems * 2 * 4);
}

int bnn_multiswag_predict_host(const bnn_model_config* cfg, const float* h_x, int64_t n_systems, const float* d_w_avg,
                               const float* d_w2_avg, const float* d_pre_D, int32_t n_models, int32_t K,
                               int32_t samples_per_model, float scale, uint64_t seed, float* h_out, void* d_scratch,
                               void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(h_x && h_out && d_scratch, BNN_E_ARG, "bnn_multiswag_predict_host: null pointer");
    BNN_REQUIRE(n_models >= 1 && samples_per_model >= 1 && n_systems >= 1, BNN_E_ARG,
                "bnn_multiswag_predict_host: empty problem");
    const int64_t U = (int64_t)n_models * samples_per_model;
    const size_t d = FlatLayout(cfg->n_features).d, P = PackedLayout(live_columns(cfg).n, cfg->n_features).P;
    const size_t xb = (size_t)n_systems * cfg->n_times * cfg->n_features * 4, ob = (size_t)U * n_systems * 2 * 4;
    char* base = (char*)d_scratch;
    float* dx = (float*)base;
    float* dth = (float*)(base + up256(xb));
    float* dthp = (float*)((char*)dth + up256(U * d * 4));
    float* dout = (float*)((char*)dthp + up256(U * P * 4));
    cudaStream_t st = (cudaStream_t)stream;
    BNN_CUDA(cudaMemcpyAsync(dx, h_x, xb, cudaMemcpyHostToDevice, st));
    rc = bnn_swag_sample(cfg, d_w_avg, d_w2_avg, d_pre_D, n_models, K, nullptr, U, 0, samples_per_model, scale, seed,
                         nullptr, nullptr, dth, dthp, stream);
    if (rc != BNN_OK) return rc;
    rc = bnn_predict(cfg, dx, n_systems, dthp, U, nullptr, nullptr, seed, 0, 0, 0, dout, nullptr, nullptr, stream);
    if (rc != BNN_OK) return rc;
    BNN_CUDA(cudaMemcpyAsync(h_out, dout, ob, cudaMemcpyDeviceToHost, st));
    BNN_CUDA(cudaStreamSynchronize(st));
    return BNN_OK;
}

}  // extern "C"
