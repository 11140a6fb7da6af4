// K4: fused SWAG training step -- placeholder until the kernels land.
#include "common.cuh"

extern "C" {

size_t bnn_train_workspace_bytes(const bnn_model_config*, int64_t, int32_t) { return 0; }

int bnn_train_step(const bnn_model_config*, const bnn_train_hparams*, int32_t, float*, float*, const float*,
                   const float*, const int32_t*, int64_t, const float*, const float*, const float*, uint64_t,
                   uint64_t, float*, float*, void*, void*) {
    bnn::set_error("bnn_train_step: not built yet");
    return BNN_E_CONFIG;
}

int bnn_eval_loss(const bnn_model_config* cfg, const float* d_x, const float* d_y, int64_t B,
                  const float* d_theta_packed, int64_t n_units, const float* d_eps, uint64_t seed,
                  float* d_out_mu_sd, float* d_loss_sum, void* d_workspace, void* stream) {
    bnn::set_error("bnn_eval_loss: not built yet");
    return BNN_E_CONFIG;
}

}  // extern "C"
