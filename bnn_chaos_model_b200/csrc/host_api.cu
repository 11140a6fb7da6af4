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
    const size_t P = bnn::PackedLayout(bnn::live_columns(cfg).n, cfg->n_features).P;
    return up256((size_t)n_systems * cfg->n_times * cfg->n_features * 4) + up256((size_t)n_units * P * 4) +
           up256((size_t)n_units * n_systems * 2 * 4);
}

namespace {
// two side streams + events of one host-entry call, released on every return path
struct HostPipe {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~HostPipe() {
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (in) cudaStreamDestroy(in);
        if (out) cudaStreamDestroy(out);
    }
};
}  // namespace

// Copies, sampling and prediction are pipelined over the systems: the batch is cut into up to four chunks at multiples of
// the kernel's system granule (4 % / 47 % / 47 % / 2 %: the first upload runs under the weight sampler, only 2 % of the
// download is left when the last kernel ends), the upload of chunk k + 1 and the download of chunk k - 1 run on two side
// streams under the predictive kernel of chunk k.  Philox draws are keyed on global (unit, system) indices, so the result
// equals the one-launch result bit for bit.
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
    const int64_t U = (int64_t)n_models * samples_per_model, N = n_systems;
    const size_t P = PackedLayout(live_columns(cfg).n, cfg->n_features).P;
    const size_t row_floats = (size_t)cfg->n_times * cfg->n_features;
    const size_t xb = (size_t)N * row_floats * 4;
    char* base = (char*)d_scratch;
    float* dx = (float*)base;
    float* dthp = (float*)(base + up256(xb));
    float* dout = (float*)((char*)dthp + up256(U * P * 4));
    cudaStream_t st = (cudaStream_t)stream;

    const int64_t g = bnn_predict_system_granule(cfg);
    BNN_REQUIRE(g >= 1, (int)g, "bnn_multiswag_predict_host: no kernel for this configuration");
    int64_t cut[5] = {0, 0, 0, 0, N};
    int n_chunks = 1;
    if (N >= 64 * g) {   // worth pipelining
        cut[1] = (int64_t)(0.04 * N / g + 0.5) * g;
        cut[2] = (int64_t)(0.51 * N / g + 0.5) * g;
        cut[3] = (int64_t)(0.98 * N / g + 0.5) * g;
        if (cut[1] < g) cut[1] = g;
        if (cut[2] <= cut[1]) cut[2] = cut[1] + g;
        if (cut[3] <= cut[2]) cut[3] = cut[2] + g;
        if (cut[3] >= N) cut[3] = N - g;     // N >= 64 g: still behind cut[2]
        n_chunks = 4;
    } else {
        cut[1] = N;
    }
    HostPipe hp;
    BNN_CUDA(cudaStreamCreateWithFlags(&hp.in, cudaStreamNonBlocking));
    BNN_CUDA(cudaStreamCreateWithFlags(&hp.out, cudaStreamNonBlocking));
    for (auto& e : hp.ev) BNN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaEvent_t ev_start = hp.ev[0], ev_done = hp.ev[1];
    cudaEvent_t* ev_in = &hp.ev[2];   // [4]
    cudaEvent_t* ev_k = &hp.ev[6];    // [4]
    // work already queued on the caller's stream (it may still use the scratch) comes first
    BNN_CUDA(cudaEventRecord(ev_start, st));
    BNN_CUDA(cudaStreamWaitEvent(hp.in, ev_start, 0));
    BNN_CUDA(cudaStreamWaitEvent(hp.out, ev_start, 0));
    for (int k = 0; k < n_chunks; ++k) {
        const int64_t lo = cut[k], hi = k + 1 == n_chunks ? N : cut[k + 1];
        BNN_CUDA(cudaMemcpyAsync(dx + lo * row_floats, h_x + lo * row_floats, (size_t)(hi - lo) * row_floats * 4,
                                 cudaMemcpyHostToDevice, hp.in));
        BNN_CUDA(cudaEventRecord(ev_in[k], hp.in));
    }
    rc = bnn_swag_sample(cfg, d_w_avg, d_w2_avg, d_pre_D, n_models, K, nullptr, U, 0, samples_per_model, scale, seed,
                         nullptr, nullptr, nullptr, dthp, stream);
    if (rc != BNN_OK) return rc;
    for (int k = 0; k < n_chunks; ++k) {
        const int64_t lo = cut[k], hi = k + 1 == n_chunks ? N : cut[k + 1];
        BNN_CUDA(cudaStreamWaitEvent(st, ev_in[k], 0));
        // [U, N, 2] output: unit stride 2 N floats, this chunk's systems start at column lo
        rc = bnn_predict_strided(cfg, dx + lo * row_floats, hi - lo, dthp, U, nullptr, nullptr, seed, 0, lo, 2 * N, 2,
                                 dout + 2 * lo, nullptr, nullptr, stream);
        if (rc != BNN_OK) return rc;
        BNN_CUDA(cudaEventRecord(ev_k[k], st));
        BNN_CUDA(cudaStreamWaitEvent(hp.out, ev_k[k], 0));
        BNN_CUDA(cudaMemcpy2DAsync(h_out + 2 * lo, (size_t)N * 8, dout + 2 * lo, (size_t)N * 8, (size_t)(hi - lo) * 8, (size_t)U,
                                   cudaMemcpyDeviceToHost, hp.out));
    }
    BNN_CUDA(cudaEventRecord(ev_done, hp.out));
    BNN_CUDA(cudaStreamWaitEvent(st, ev_done, 0));
    BNN_CUDA(cudaStreamSynchronize(st));
    BNN_CUDA(cudaStreamSynchronize(hp.out));
    return BNN_OK;
}

}  // extern "C"
