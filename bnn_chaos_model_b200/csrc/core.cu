// Error plumbing, configuration checks and layout queries of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace bnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice: %s (no CUDA device; this library has no CPU fallback)", cudaGetErrorString(e));
        return (int)e;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
    }
    if (major != 10) {
        set_error("device compute capability %d.x is not sm_100 (B200); kernels are sm_100a-only", major);
        return BNN_E_ARCH;
    }
    return BNN_OK;
}

int validate_config(const bnn_model_config* cfg) {
    BNN_REQUIRE(cfg != nullptr, BNN_E_ARG, "cfg is NULL");
    BNN_REQUIRE(cfg->hidden == H && cfg->latent == L, BNN_E_CONFIG,
                "kernels are compiled for hidden=%d latent=%d (got %d, %d)", H, L, cfg->hidden, cfg->latent);
    BNN_REQUIRE(cfg->n_in_layers == 1 && cfg->n_out_layers == 1, BNN_E_CONFIG,
                "kernels are compiled for in=1 out=1 (got %d, %d)", cfg->n_in_layers, cfg->n_out_layers);
    BNN_REQUIRE(cfg->n_features >= 1 && cfg->n_features <= MAXF, BNN_E_CONFIG, "n_features=%d out of [1,%d]",
                cfg->n_features, MAXF);
    BNN_REQUIRE(cfg->n_times >= 4 && cfg->n_times % 4 == 0 && cfg->n_times <= 128, BNN_E_CONFIG,
                "n_times=%d must be a multiple of 4 in [4,128]", cfg->n_times);
    LiveCols lc = live_columns(cfg);
    BNN_REQUIRE(lc.n >= 1, BNN_E_CONFIG, "zero_mask removes every input column");
    return BNN_OK;
}

}  // namespace bnn

extern "C" {

int bnn_abi_version(void) { return BNN_ABI_VERSION; }

const char* bnn_last_error_string(void) { return bnn::g_err; }

int64_t bnn_param_count(const bnn_model_config* cfg) {
    int rc = bnn::validate_config(cfg);
    if (rc != BNN_OK) return rc;
    return bnn::FlatLayout(cfg->n_features).d;
}

int64_t bnn_packed_param_count(const bnn_model_config* cfg) {
    int rc = bnn::validate_config(cfg);
    if (rc != BNN_OK) return rc;
    return bnn::PackedLayout(bnn::live_columns(cfg).n, cfg->n_features).P;
}

}  // extern "C"
