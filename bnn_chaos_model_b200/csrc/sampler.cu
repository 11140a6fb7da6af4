// K1: SWAG weight sampling for all (model, weight-sample) units at once, and the
// flatten()-order -> packed-layout gather.
//
// Reference: SWAGModel.sample_weights, /root/reference/spock_reg_model.py:815-838
//   D = pre_D - w_avg[:,None]; z1 ~ N(0,I_d); z2 ~ N(0,I_K)
//   w  = w_avg + (scale/sqrt2) * z1 * sqrt|w2_avg - w_avg^2|     (dense diag matmul in the
//        reference, :832-834; every off-diagonal term is an exact +0)
//   w += scale * (D z2) / sqrt(2 (K-1))                           (:835)
// and SWAGModel.load (:748-761), which installs the flat vector into the module.
//
// HBM-bound: writes d floats per unit; pre_D rows are re-read from L2.
#include "common.cuh"

namespace bnn {

constexpr int SAMPLER_THREADS = 128;
constexpr int MAXK = 64;

// One block = 512 consecutive flat indices (4 per thread = one Philox block) of one unit.
__global__ void __launch_bounds__(SAMPLER_THREADS)
swag_sample_kernel(const float* __restrict__ w_avg, const float* __restrict__ w2_avg,
                   const float* __restrict__ pre_D, int d, int K, const int32_t* __restrict__ unit_model,
                   int64_t unit_offset, int samples_per_model, int n_models, float c1, float scale, float c2div,
                   uint64_t seed, const float* __restrict__ z1, const float* __restrict__ z2,
                   float* __restrict__ theta, int blocks_per_unit) {
    __shared__ float z2s[MAXK];
    const int64_t u = blockIdx.x / blocks_per_unit;
    const int jb = (blockIdx.x % blocks_per_unit) * (SAMPLER_THREADS * 4);
    const int64_t gu = unit_offset + u;
    int m = unit_model ? unit_model[u] : (int)(gu / samples_per_model);
    m = min(max(m, 0), n_models - 1);

    if (threadIdx.x < (K + 3) / 4) {
        float4 n4;
        if (z2) {
            const float* p = z2 + u * K + threadIdx.x * 4;
            int rem = K - threadIdx.x * 4;
            n4.x = p[0];
            n4.y = rem > 1 ? p[1] : 0.f;
            n4.z = rem > 2 ? p[2] : 0.f;
            n4.w = rem > 3 ? p[3] : 0.f;
        } else {
            n4 = philox_normal4(seed, STREAM_Z2, (uint32_t)gu, 0u, threadIdx.x);
        }
        z2s[threadIdx.x * 4 + 0] = n4.x;
        z2s[threadIdx.x * 4 + 1] = n4.y;
        z2s[threadIdx.x * 4 + 2] = n4.z;
        z2s[threadIdx.x * 4 + 3] = n4.w;
    }
    __syncthreads();

    const int j0 = jb + threadIdx.x * 4;
    if (j0 >= d) return;
    float zz[4];
    if (z1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) zz[i] = (j0 + i < d) ? z1[u * (int64_t)d + j0 + i] : 0.f;
    } else {
        float4 n4 = philox_normal4(seed, STREAM_Z1, (uint32_t)gu, 0u, (uint32_t)(j0 >> 2));
        zz[0] = n4.x; zz[1] = n4.y; zz[2] = n4.z; zz[3] = n4.w;
    }
    const float* wa = w_avg + (int64_t)m * d;
    const float* w2a = w2_avg + (int64_t)m * d;
    const float* pd = pre_D + (int64_t)m * d * K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j0 + i;
        if (j >= d) break;
        const float w = wa[j];
        // explicit _rn intrinsics: keep the reference's separate mul/sub/add roundings (no FMA contraction)
        const float sig = fabsf(__fsub_rn(w2a[j], __fmul_rn(w, w)));
        float th = __fadd_rn(w, __fmul_rn(__fmul_rn(c1, zz[i]), sqrtf(sig)));
        const float* row = pd + (int64_t)j * K;
        float dot = 0.f;
        for (int k = 0; k < K; ++k) dot = fmaf(__fsub_rn(__ldg(row + k), w), z2s[k], dot);
        th = __fadd_rn(th, __fdiv_rn(__fmul_rn(scale, dot), c2div));
        theta[u * (int64_t)d + j] = th;
    }
}

// theta [U,d] (flatten order) -> theta_packed [U,P].  One thread per packed float.
__global__ void pack_theta_kernel(const float* __restrict__ theta, int64_t n_units, FlatLayout fl, PackedLayout pl,
                                  LiveCols lc, float* __restrict__ packed) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n_units * pl.P) return;
    const int64_t u = idx / pl.P;
    const int i = (int)(idx % pl.P);
    const float* th = theta + u * fl.d;
    int src = -1;
    auto grp = [](int c, int& q, int& ii) { q = c / GC; ii = c % GC; };
    int q, ii;
    if (i < pl.b0p) {  // W0p[kin][48]
        int k = i / HP;
        grp(i % HP, q, ii);
        if (ii < 10) src = fl.W0 + (q * 10 + ii) * fl.F + lc.col[k];
    } else if (i < pl.W1p) {
        grp(i - pl.b0p, q, ii);
        if (ii < 10) src = fl.b0 + q * 10 + ii;
    } else if (i < pl.b1p) {
        int r = i - pl.W1p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.W1 + (q * 10 + ii) * H + k;
    } else if (i < pl.W2p) {
        grp(i - pl.b1p, q, ii);
        if (ii < 10) src = fl.b1 + q * 10 + ii;
    } else if (i < pl.b2p) {  // W2p duplicated pairs
        int r = i - pl.W2p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.W2 + (q * 5 + ii / 2) * H + k;
    } else if (i < pl.V0p) {
        grp(i - pl.b2p, q, ii);
        if (ii < 10) src = fl.b2 + q * 5 + ii / 2;
    } else if (i < pl.c0p) {
        int r = i - pl.V0p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.V0 + (q * 10 + ii) * S2 + k;
    } else if (i < pl.V1p) {
        grp(i - pl.c0p, q, ii);
        if (ii < 10) src = fl.c0 + q * 10 + ii;
    } else if (i < pl.c1p) {
        int r = i - pl.V1p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.V1 + (q * 10 + ii) * H + k;
    } else if (i < pl.V2) {
        grp(i - pl.c1p, q, ii);
        if (ii < 10) src = fl.c1 + q * 10 + ii;
    } else if (i < pl.c2) {
        src = fl.V2 + (i - pl.V2);
    } else if (i < pl.lv_sum) {
        int r = i - pl.c2;
        if (r < 2) src = fl.c2 + r;
    } else if (i < pl.lv_in) {
        src = fl.lv_sum + (i - pl.lv_sum);
    } else if (i < pl.B1h) {
        int r = i - pl.lv_in;
        if (r < fl.F) src = fl.lv_in + r;
    } else {
        if (i >= pl.Bb) {  // fp32 bias block b0[48] | b1[48] | b2[32]
            const int r = i - pl.Bb;
            float b = 0.f;
            if (r < TC_N) { if (r < H) b = th[fl.b0 + r]; }
            else if (r < 2 * TC_N) { if (r - TC_N < H) b = th[fl.b1 + r - TC_N]; }
            else if (r - 2 * TC_N < L) b = th[fl.b2 + r - 2 * TC_N];
            packed[idx] = b;
            return;
        }
        // tensor-core B operands: element (n, k) of a [N][K] matrix sits at ((k/4)*N + n)*4 + k%4
        int r, N, nreal, wsrc, kreal, ld;
        bool lo;
        if (i < pl.B2h) {
            r = i - pl.B1h; N = TC_N; lo = r >= pl.tc_k1 * TC_N; r -= lo ? pl.tc_k1 * TC_N : 0;
            nreal = H; kreal = pl.kin; wsrc = fl.W0; ld = fl.F;
        } else if (i < pl.B3h) {
            r = i - pl.B2h; N = TC_N; lo = r >= TC_K2 * TC_N; r -= lo ? TC_K2 * TC_N : 0;
            nreal = H; kreal = H; wsrc = fl.W1; ld = H;
        } else {
            r = i - pl.B3h; N = TC_N3; lo = r >= TC_K2 * TC_N3; r -= lo ? TC_K2 * TC_N3 : 0;
            nreal = L; kreal = H; wsrc = fl.W2; ld = H;
        }
        const int chunk = r / (N * 4), n = (r / 4) % N, k = chunk * 4 + (r & 3);
        float w = 0.f;
        if (n < nreal && k < kreal) w = th[wsrc + n * ld + ((wsrc == fl.W0) ? (int)lc.col[k] : k)];
        const float hi = tf32_rna(w);
        packed[idx] = lo ? tf32_rna(w - hi) : hi;  // lo pre-rounded: the tensor core would truncate it
        return;
    }
    packed[idx] = src >= 0 ? th[src] : 0.f;
}

int launch_pack_theta(const bnn_model_config* cfg, const float* d_theta, int64_t n_units, float* d_packed,
                      cudaStream_t st) {
    FlatLayout fl(cfg->n_features);
    LiveCols lc = live_columns(cfg);
    PackedLayout pl(lc.n, cfg->n_features);
    const int64_t total = n_units * pl.P;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "too many units for one pack launch");
    pack_theta_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_theta, n_units, fl, pl, lc, d_packed);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

}  // namespace bnn

extern "C" {

int bnn_pack_theta(const bnn_model_config* cfg, const float* d_theta, int64_t n_units, float* d_theta_packed,
                   void* stream) {
    int rc = bnn::validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = bnn::check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_theta && d_theta_packed && n_units > 0, BNN_E_ARG, "bnn_pack_theta: null pointer or n_units<=0");
    return bnn::launch_pack_theta(cfg, d_theta, n_units, d_theta_packed, (cudaStream_t)stream);
}

int bnn_swag_sample(const bnn_model_config* cfg, const float* d_w_avg, const float* d_w2_avg, const float* d_pre_D,
                    int32_t n_models, int32_t K, const int32_t* d_unit_model, int64_t n_units, int64_t unit_offset,
                    int32_t samples_per_model, float scale, uint64_t seed, const float* d_z1, const float* d_z2,
                    float* d_theta, float* d_theta_packed, void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_w_avg && d_w2_avg && d_pre_D, BNN_E_ARG, "bnn_swag_sample: SWAG statistics pointer is NULL");
    BNN_REQUIRE(d_theta, BNN_E_ARG, "bnn_swag_sample: d_theta is required (it also feeds the packed layout)");
    BNN_REQUIRE(n_models >= 1 && n_units >= 1, BNN_E_ARG, "bnn_swag_sample: n_models/n_units must be >= 1");
    // the reference fails in D @ z2 when pre_D has fewer than K columns (:835); K>=2 for sqrt(2(K-1))
    BNN_REQUIRE(K >= 2 && K <= MAXK, BNN_E_ARG, "bnn_swag_sample: K=%d out of [2,%d]", K, MAXK);
    BNN_REQUIRE((d_z1 == nullptr) == (d_z2 == nullptr), BNN_E_ARG, "bnn_swag_sample: give both z1 and z2 or neither");
    BNN_REQUIRE(d_unit_model || samples_per_model >= 1, BNN_E_ARG, "bnn_swag_sample: samples_per_model must be >= 1");
    const int d = FlatLayout(cfg->n_features).d;
    const int bpu = (d + SAMPLER_THREADS * 4 - 1) / (SAMPLER_THREADS * 4);
    const int64_t blocks = n_units * bpu;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "bnn_swag_sample: too many units for one launch");
    // scale * (1/np.sqrt(2.0)) is a double that torch casts to fp32 when it meets the fp32 tensor (:834)
    const float c1 = (float)((double)scale * (1.0 / sqrt(2.0)));
    const float c2div = (float)sqrt(2.0 * (K - 1));
    cudaStream_t st = (cudaStream_t)stream;
    swag_sample_kernel<<<(unsigned)blocks, SAMPLER_THREADS, 0, st>>>(
        d_w_avg, d_w2_avg, d_pre_D, d, K, d_unit_model, unit_offset, samples_per_model, n_models, c1, scale, c2div,
        seed, d_z1, d_z2, d_theta, bpu);
    BNN_CUDA(cudaGetLastError());
    if (d_theta_packed) return launch_pack_theta(cfg, d_theta, n_units, d_theta_packed, st);
    return BNN_OK;
}

}  // extern "C"
