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
