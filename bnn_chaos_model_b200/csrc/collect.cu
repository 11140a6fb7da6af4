// K5: SWAG moment collection for several seed models at once.
//
// Reference: SWAGModel.aggregate_model (/root/reference/spock_reg_model.py:763-785):
//   first call:  w_avg = w, w2_avg = w^2, pre_D = w[:,None]
//   later:       w_avg = (w_avg*n + w)/(n+1), w2_avg likewise;
//                if current_epoch % c == 0: append w as newest column, keep the last K
//   n += 1
// pre_D is kept as a fixed [d,K] row-major buffer with n_cols valid columns (oldest first),
// so a full buffer is exactly the tensor save_swag writes (:917).  HBM-bound, tiny.
#include "common.cuh"

namespace bnn {

__global__ void swag_collect_kernel(const float* __restrict__ w, int64_t d, int K, float* __restrict__ w_avg,
                                    float* __restrict__ w2_avg, float* __restrict__ pre_D,
                                    const int32_t* __restrict__ n_models, const int32_t* __restrict__ n_cols,
                                    int current_epoch, int c) {
    const int s = blockIdx.y;
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= d) return;
    const int n = n_models[s];
    const int nc = n_cols[s];
    const int64_t o = (int64_t)s * d + j;
    const float cw = w[o];
    const float cw2 = __fmul_rn(cw, cw);
    if (n == 0) {
        w_avg[o] = cw;
        w2_avg[o] = cw2;
    } else {
        const float nf = (float)n, n1 = (float)(n + 1);
        w_avg[o] = __fdiv_rn(__fadd_rn(__fmul_rn(w_avg[o], nf), cw), n1);
        w2_avg[o] = __fdiv_rn(__fadd_rn(__fmul_rn(w2_avg[o], nf), cw2), n1);
    }
    const bool record = (nc == 0) || (current_epoch % c == 0);
    if (record) {
        float* row = pre_D + o * K;
        if (nc < K) {
            row[nc] = cw;
        } else {
            for (int k = 0; k + 1 < K; ++k) row[k] = row[k + 1];
            row[K - 1] = cw;
        }
    }
}

__global__ void swag_bump_kernel(int32_t* n_models, int32_t* n_cols, int n_seeds, int K, int current_epoch, int c) {
    const int s = threadIdx.x + blockIdx.x * blockDim.x;
    if (s >= n_seeds) return;
    const int nc = n_cols[s];
    if ((nc == 0) || (current_epoch % c == 0)) n_cols[s] = min(nc + 1, K);
    n_models[s] += 1;
}

}  // namespace bnn

extern "C" int bnn_swag_collect(const float* d_w, int64_t d, int32_t n_seeds, int32_t K, float* d_w_avg,
                                float* d_w2_avg, float* d_pre_D, int32_t* d_n_models, int32_t* d_n_cols,
                                int32_t current_epoch, int32_t c, void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_w && d_w_avg && d_w2_avg && d_pre_D && d_n_models && d_n_cols, BNN_E_ARG,
                "bnn_swag_collect: null pointer");
    BNN_REQUIRE(d > 0 && n_seeds > 0 && n_seeds < 65536 && K >= 1 && c >= 1, BNN_E_ARG,
                "bnn_swag_collect: bad sizes (d=%lld seeds=%d K=%d c=%d)", (long long)d, n_seeds, K, c);
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    dim3 grid((unsigned)((d + threads - 1) / threads), (unsigned)n_seeds);
    swag_collect_kernel<<<grid, threads, 0, st>>>(d_w, d, K, d_w_avg, d_w2_avg, d_pre_D, d_n_models, d_n_cols,
                                                  current_epoch, c);
    BNN_CUDA(cudaGetLastError());
    swag_bump_kernel<<<(n_seeds + 127) / 128, 128, 0, st>>>(d_n_models, d_n_cols, n_seeds, K, current_epoch, c);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

// ---------------------------------------------------------------------------------------
// K6: input packing -- the step immediately upstream of the predictive kernel.
//
// Reference: data_setup_kernel (/root/reference/figures/spock/regression.py:183-213) followed by
// ssX.transform (:144, float64 sklearn StandardScaler) and torch.tensor(X).float() (:145):
//   raw row = [26 time-series columns | 3 masses | isnotfinite(col 3), (col 6), (col 7)]   (32 columns)
//   nan_to_num(posinf=0, neginf=0);  angle columns {11,12,13,17,18,19,23,24,25} -> (cos, sin)    (41 columns)
//   x = float32((X - mean) / scale)        (all arithmetic in float64, like numpy)
// HBM-bound: 232 B read + 164 B written per time-step row.
// ---------------------------------------------------------------------------------------
namespace bnn {

__device__ __forceinline__ bool is_angle_col(int j) { return j >= 11 && j <= 25 && ((j - 11) % 6) < 3; }

__global__ void __launch_bounds__(256) pack_inputs_kernel(const double* __restrict__ ts, const double* __restrict__ mass,
                                                          const double* __restrict__ mean, const double* __restrict__ scale,
                                                          int64_t n_rows, int T, float* __restrict__ x) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // (row, raw column j of 32)
    if (idx >= n_rows * 32) return;
    const int64_t row = idx >> 5;
    const int j = (int)(idx & 31);
    double v;
    if (j < 26) {
        v = ts[row * 26 + j];
    } else if (j < 29) {
        v = mass[(row / T) * 3 + (j - 26)];
    } else {
        const int src = (j == 29) ? 3 : (j == 30 ? 6 : 7);
        v = isfinite(ts[row * 26 + src]) ? 0.0 : 1.0;  // flags are taken before nan_to_num
    }
    if (!isfinite(v)) v = 0.0;  // np.nan_to_num(posinf=0.0, neginf=0.0): NaN -> 0 too
    // output column: every angle column before j adds one
    int oc = j;
    if (j > 11) oc += min(j - 11, 3);
    if (j > 17) oc += min(j - 17, 3);
    if (j > 23) oc += min(j - 23, 3);
    float* o = x + row * 41;
    if (is_angle_col(j)) {
        o[oc] = (float)((cos(v) - mean[oc]) / scale[oc]);
        o[oc + 1] = (float)((sin(v) - mean[oc + 1]) / scale[oc + 1]);
    } else {
        o[oc] = (float)((v - mean[oc]) / scale[oc]);
    }
}

}  // namespace bnn

extern "C" int bnn_pack_inputs(const double* d_tseries, const double* d_masses, const double* d_ss_mean,
                               const double* d_ss_scale, int64_t n_systems, int32_t n_times, float* d_x, void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_tseries && d_masses && d_ss_mean && d_ss_scale && d_x, BNN_E_ARG, "bnn_pack_inputs: null pointer");
    BNN_REQUIRE(n_systems > 0 && n_times > 0, BNN_E_ARG, "bnn_pack_inputs: empty input");
    const int64_t total = n_systems * n_times * 32;
    const int64_t blocks = (total + 255) / 256;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "bnn_pack_inputs: too many rows for one launch");
    pack_inputs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_tseries, d_masses, d_ss_mean, d_ss_scale,
                                                                          n_systems * n_times, n_times, d_x);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}
