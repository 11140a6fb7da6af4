// K2: fused MultiSWAG posterior predictive.  See predict_device.cuh for the decomposition.
//
// Replaces, for every (unit, system) pair in one launch, the reference's per-sample Python
// loop body SWAGModel.forward_swag_fast (/root/reference/spock_reg_model.py:878-908; callers
// figures/main_figures.py:127-156, figures/spock/regression.py:74-92,
// figures/multiswag_5_planet.py:295-298).
#include <stdlib.h>
#include <string.h>

#include "predict_device.cuh"
#include "pipeline.cuh"

namespace bnn {

struct PredictParams {
    const float* X;       // [N,T,F]
    const float* thp;     // [U,P]
    const float* eps;     // [U,N,2L] or null
    const float* eps_sum; // [U,N,2L] or null (noisy forward)
    float* summary;       // [U,N,2L] or null
    float* out;           // [U,N,2] or [N,U,2]
#ifdef BNN_TC_TIMELINE
    long long* dbg;       // timeline buffer of the diagnostic build (tools/tc_timeline.py)
#endif
    int64_t N, U;
    int64_t unit_offset, system_offset;
    int64_t out_unit_stride, out_sys_stride;  // in floats
    uint64_t seed;
    int F, kin;
    int8_t wide_col[16];  // tensor-core kernel, wide variant: X column of the live inputs 32..47 (-1: none)
    int units_per_cta;
    HeadConsts hc;
    ColMap cm;
};

// ---------------------------------------------------------------------------------------
// v1: synchronous variant.  grid = (tiles, unit chunks).  Per unit: stage the feature
// weights in shared memory, warps take tasks round-robin, warp 0 runs the tail of the
// previous unit's records while the others already work on the next unit.
// ---------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) predict_v1_kernel(const PredictParams prm, const int T) {
    extern __shared__ __align__(16) float smem[];
    const TileGeom g(T);
    const PackedLayout pl(prm.kin, prm.F);
    float* xT = smem;
    float* wbuf = xT + prm.kin * g.RP;
    float* hbuf = wbuf + pl.feat_floats;
    float* rec = hbuf + NW * HT_FLOATS;
    float* scratch = rec + 2 * rec_floats(g);

    const int warp = threadIdx.x >> 5;
    const int64_t n0 = (int64_t)blockIdx.x * SYS_TILE;
    const int n_valid = (int)min((int64_t)SYS_TILE, prm.N - n0);
    const int64_t u_begin = (int64_t)blockIdx.y * prm.units_per_cta;
    const int64_t u_end = min(prm.U, u_begin + prm.units_per_cta);

    load_x_tile(prm.X, n0, n_valid, prm.F, g, prm.kin, prm.cm, xT, reinterpret_cast<int*>(hbuf));

    for (int64_t u = u_begin; u < u_end; ++u) {
        const float* thp = prm.thp + u * pl.P;
        // stage feature weights (all threads), previous unit's tasks are complete (barrier below)
        for (int i = threadIdx.x; i < pl.feat_floats / 4; i += NW * 32)
            reinterpret_cast<float4*>(wbuf)[i] = __ldg(reinterpret_cast<const float4*>(thp) + i);
        __syncthreads();
        float* rec_u = rec + (int)((u - u_begin) & 1) * rec_floats(g);
        for (int t = warp; t < g.n_tasks; t += NW)
            mlp_task(xT, g, prm.kin, wbuf, pl, hbuf + warp * HT_FLOATS, t, rec_u);
        __syncthreads();
        // tail by the last warp (it had the fewest tasks when n_tasks % NW != 0)
        if (warp == NW - 1) {
            const float* eps_u = prm.eps ? prm.eps + u * prm.N * S2 : nullptr;
            const float* eps_sum_u = prm.eps_sum ? prm.eps_sum + u * prm.N * S2 : nullptr;
            float* summary_u = prm.summary ? prm.summary + u * prm.N * S2 : nullptr;
            tail_unit<true>(rec_u, g, thp, pl, eps_u, eps_sum_u, summary_u, prm.seed, (uint32_t)(prm.unit_offset + u),
                      prm.system_offset + n0, n0, n_valid, prm.hc, scratch, prm.out + u * prm.out_unit_stride,
                      prm.out_sys_stride);
        }
        // the other warps run ahead to the next unit: they touch wbuf (safe: all tasks done) and the
        // other rec slot; the barrier after the next staging orders the tail before slot reuse.
    }
}

static size_t v1_smem_bytes(int kin, int T, int NW) {
    TileGeom g(T);
    PackedLayout pl(kin);
    size_t fl = (size_t)kin * g.RP + pl.feat_floats + (size_t)NW * HT_FLOATS + 2 * rec_floats(g) + 1024;
    return fl * sizeof(float);
}

template <int NW>
static int launch_v1(const PredictParams& prm, int T, cudaStream_t st) {
    const size_t smem = v1_smem_bytes(prm.kin, T, NW);
    BNN_REQUIRE(smem <= 227 * 1024, BNN_E_CONFIG, "predict tile needs %zu bytes of shared memory (> 227 KB)", smem);
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        BNN_CUDA(cudaFuncSetAttribute(predict_v1_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    const int64_t tiles = (prm.N + SYS_TILE - 1) / SYS_TILE;
    const int64_t chunks = (prm.U + prm.units_per_cta - 1) / prm.units_per_cta;
    BNN_REQUIRE(tiles < (1ll << 31) && chunks < 65536, BNN_E_ARG, "grid too large (%lld tiles, %lld chunks)",
                (long long)tiles, (long long)chunks);
    dim3 grid((unsigned)tiles, (unsigned)chunks);
    predict_v1_kernel<NW><<<grid, NW * 32, smem, st>>>(prm, T);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}


// ---------------------------------------------------------------------------------------
// v2: warp-specialised persistent kernel.  grid = #SMs; every CTA walks a static list of
// work items (tile of 8 systems, contiguous range of units).  Per item:
//   producer warp : streams each unit's packed weights (38 KB) into a 2-slot shared-memory
//                   ring with cp.async.bulk (TMA 1-D) + mbarrier complete_tx
//   NC consumers  : fetch (unit, 32-row task) pairs from a shared counter, run feature_nn on
//                   FFMA2, write pooled (mean, M2) records, arrive on done[slot]
//   tail warp     : when the 25 tasks of a unit are done: merge records, sample summary
//                   statistics, regress_nn from the shared-memory copy of the head weights,
//                   store (mu, std), release the slot to the producer
// No CTA-wide barrier inside an item; x stays resident in shared memory for all units.
// ---------------------------------------------------------------------------------------
struct V2Smem {
    uint64_t full[2], done[2], tail_done[2];
    int task_ctr;
    int pad;
};

template <int NC>
__global__ void __launch_bounds__((NC + 2) * 32, 1)
predict_v2_kernel(const PredictParams prm, const int T, const int n_tiles, const int chunks) {
    extern __shared__ __align__(128) float smem_v2[];
    float* smem = smem_v2;
    const TileGeom g(T);
    const PackedLayout pl(prm.kin, prm.F);
    const int P4 = pl.B1h;  // floats staged per unit: feature + head (+ logvars); the tensor-core section is skipped
    float* xT = smem;
    float* ring = xT + prm.kin * g.RP;
    float* hq = ring + 2 * P4;
    float* rec = hq + NC * HQ_FLOATS;
    float* scratch = rec + 2 * rec_floats(g);
    V2Smem* sh = reinterpret_cast<V2Smem*>(scratch + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_items = n_tiles * chunks;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / chunks, chunk = item % chunks;
        const int64_t n0 = (int64_t)tile * SYS_TILE;
        const int n_valid = (int)min((int64_t)SYS_TILE, prm.N - n0);
        // units of this item: even split of [0,U) into `chunks` ranges
        const int64_t u_begin = prm.U * chunk / chunks, u_end = prm.U * (chunk + 1) / chunks;
        const int n_units = (int)(u_end - u_begin);

        __syncthreads();  // previous item fully drained by every role
        if (threadIdx.x == 0) {
            if (item != (int)blockIdx.x) {
                for (int s = 0; s < 2; ++s) { mbar_inval(&sh->full[s]); mbar_inval(&sh->done[s]); mbar_inval(&sh->tail_done[s]); }
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(&sh->full[s], 1);
                mbar_init(&sh->done[s], g.n_tasks);
                mbar_init(&sh->tail_done[s], 1);
            }
            sh->task_ctr = 0;
            mbar_init_fence();
        }
        // hq doubles as the poison scratch of the tile load (NC*640 floats >= RP ints for NC >= 2)
        load_x_tile(prm.X, n0, n_valid, prm.F, g, prm.kin, prm.cm, xT, reinterpret_cast<int*>(hq));
        __syncthreads();

        if (warp == NC) {
            // ---------------- producer ----------------
            if (lane == 0) {
                for (int i = 0; i < n_units; ++i) {
                    const int s = i & 1;
                    if (i >= 2) mbar_wait(&sh->tail_done[s], ((i >> 1) - 1) & 1);
                    mbar_arrive_expect_tx(&sh->full[s], (uint32_t)(P4 * sizeof(float)));
                    bulk_g2s(ring + s * P4, prm.thp + (u_begin + i) * pl.P, (uint32_t)(P4 * sizeof(float)), &sh->full[s]);
                }
            }
        } else if (warp == NC + 1) {
            // ---------------- tail ----------------
            for (int i = 0; i < n_units; ++i) {
                const int s = i & 1;
                const uint32_t par = (i >> 1) & 1;
                mbar_wait(&sh->full[s], par);  // visibility of the bulk-copied head weights
                mbar_wait(&sh->done[s], par);  // all tasks of the unit have written their records
                const int64_t u = u_begin + i;
                const float* eps_u = prm.eps ? prm.eps + u * prm.N * S2 : nullptr;
                const float* eps_sum_u = prm.eps_sum ? prm.eps_sum + u * prm.N * S2 : nullptr;
                float* summary_u = prm.summary ? prm.summary + u * prm.N * S2 : nullptr;
                tail_unit<false>(rec + s * rec_floats(g), g, ring + s * P4, pl, eps_u, eps_sum_u, summary_u, prm.seed,
                                 (uint32_t)(prm.unit_offset + u), prm.system_offset + n0, n0, n_valid, prm.hc,
                                 scratch, prm.out + u * prm.out_unit_stride, prm.out_sys_stride);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh->tail_done[s]);
            }
        } else {
            // ---------------- consumers ----------------
            float* my_hq = hq + warp * HQ_FLOATS;
            const int total = n_units * g.n_tasks;
            while (true) {
                int id = 0;
                if (lane == 0) id = atomicAdd(&sh->task_ctr, 1);
                id = __shfl_sync(0xffffffffu, id, 0);
                if (id >= total) break;
                const int i = id / g.n_tasks, t = id - i * g.n_tasks;
                const int s = i & 1;
                mbar_wait(&sh->full[s], (i >> 1) & 1);
                mlp_task_v2(xT, g, prm.kin, ring + s * P4, pl, my_hq, t, rec + s * rec_floats(g));
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh->done[s]);
            }
        }
    }
}

static size_t v2_smem_bytes(int kin, int F, int T, int NC) {
    TileGeom g(T);
    PackedLayout pl(kin, F);
    size_t fl = (size_t)kin * g.RP + 2 * (size_t)pl.B1h + (size_t)NC * HQ_FLOATS + 2 * rec_floats(g) + 1024;
    return fl * sizeof(float) + sizeof(V2Smem);
}

template <int NC>
static int launch_v2(const PredictParams& prm, int T, cudaStream_t st) {
    const size_t smem = v2_smem_bytes(prm.kin, prm.F, T, NC);
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        BNN_CUDA(cudaFuncSetAttribute(predict_v2_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    int n_sms = 0;
    {
        int dev = 0;
        BNN_CUDA(cudaGetDevice(&dev));
        BNN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t tiles = (prm.N + SYS_TILE - 1) / SYS_TILE;
    BNN_REQUIRE(tiles < (1ll << 24), BNN_E_ARG, "too many system tiles for one launch (%lld)", (long long)tiles);
    // Split each tile's units into `chunks` items so that the item count is close to a multiple of the
    // SM count (static round-robin over persistent CTAs) while items stay long enough to amortise the
    // x-tile load and the pipeline fill (>= 32 units per item when U allows).
    int64_t max_chunks = prm.U >= 64 ? prm.U / 32 : 1;
    if (max_chunks > 64) max_chunks = 64;
    int best = 1;
    double best_eff = 0.0;
    for (int c = 1; c <= max_chunks; ++c) {
        const int64_t items = tiles * c;
        const int64_t rounds = (items + n_sms - 1) / n_sms;
        const double eff = (double)items / (double)(rounds * n_sms);
        if (eff > best_eff + 0.005) { best_eff = eff; best = c; }
    }
    const int64_t items = tiles * best;
    const int grid = (int)(items < n_sms ? items : n_sms);
    predict_v2_kernel<NC><<<grid, (NC + 2) * 32, smem, st>>>(prm, T, (int)tiles, best);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

}  // namespace bnn

#include "predict_tc_kernel.cuh"

extern "C" {

size_t bnn_predict_workspace_bytes(const bnn_model_config*, int64_t, int64_t) { return 0; }

// Kernel selection.  Default (0): the tensor-core kernel (tcgen05, 3xTF32) when T = 100 (at most 32 live input columns: one
// layer-1 pass; 33..48, e.g. all 41 columns of the noisy forward: K = 40 + a second 8-column pass),
// else the FP32 FFMA2 kernels (v2: warp-specialised with a TMA weight ring, when its tile fits in shared memory; else
// v1).  A process-wide override for tests / tools: bnn_set_predict_variant() (diagnostic header), initialised ONCE from
// the environment variable BNN_PREDICT_VARIANT = tc | v2 | v1; likewise the unit chunk (BNN_PREDICT_UNIT_CHUNK).
enum { PV_AUTO = 0, PV_TC = 1, PV_V2 = 2, PV_V1 = 3 };
static int g_predict_variant = -1;
static long long g_unit_chunk = -1;
static int predict_variant() {
    if (g_predict_variant < 0) {
        const char* f = getenv("BNN_PREDICT_VARIANT");
        g_predict_variant = !f ? PV_AUTO : (!strncmp(f, "tc", 2) ? PV_TC : (!strncmp(f, "v2", 2) ? PV_V2 : (!strcmp(f, "v1") ? PV_V1 : PV_AUTO)));
    }
    return g_predict_variant;
}
static long long predict_unit_chunk() {
    if (g_unit_chunk < 0) {
        const char* uc = getenv("BNN_PREDICT_UNIT_CHUNK");
        g_unit_chunk = uc ? strtoll(uc, nullptr, 10) : 0;
        if (g_unit_chunk < 0) g_unit_chunk = 0;
    }
    return g_unit_chunk;
}
int bnn_set_predict_variant(int32_t variant) {
    BNN_REQUIRE(variant >= 0 && variant <= 3, BNN_E_ARG, "bnn_set_predict_variant: 0 = auto, 1 = tensor-core, 2 = FFMA v2, 3 = FFMA v1");
    g_predict_variant = variant;
    return BNN_OK;
}
int bnn_set_predict_unit_chunk(int64_t units) {
    BNN_REQUIRE(units >= 0, BNN_E_ARG, "bnn_set_predict_unit_chunk: units per launch, 0 = default (1024)");
    g_unit_chunk = units;
    return BNN_OK;
}

// true when bnn_predict runs the tensor-core kernel for this config
static bool tc_selected(const bnn_model_config* cfg, int kin) {
    const int v = predict_variant();
    if (v != PV_AUTO && v != PV_TC) return false;
    return cfg->n_times == bnn::tc::T_FIXED && kin <= bnn::TC_K1W;
}

int32_t bnn_predict_system_granule(const bnn_model_config* cfg) {
    int rc = bnn::validate_config(cfg);
    if (rc != BNN_OK) return rc;
    return tc_selected(cfg, bnn::live_columns(cfg).n) ? bnn::tc::SYS : 1;
}

int bnn_predict(const bnn_model_config* cfg, const float* d_x, int64_t n_systems, const float* d_theta_packed,
                int64_t n_units, const float* d_eps, const float* d_eps_sum, uint64_t seed, int64_t unit_offset,
                int64_t system_offset, int32_t out_system_major, float* d_out, float* d_summary_out,
                void* d_workspace, void* stream) {
    return bnn_predict_strided(cfg, d_x, n_systems, d_theta_packed, n_units, d_eps, d_eps_sum, seed, unit_offset,
                               system_offset, out_system_major ? 2 : n_systems * 2, out_system_major ? n_units * 2 : 2,
                               d_out, d_summary_out, d_workspace, stream);
}

int bnn_predict_strided(const bnn_model_config* cfg, const float* d_x, int64_t n_systems, const float* d_theta_packed,
                        int64_t n_units, const float* d_eps, const float* d_eps_sum, uint64_t seed, int64_t unit_offset,
                        int64_t system_offset, int64_t out_unit_stride, int64_t out_system_stride, float* d_out,
                        float* d_summary_out, void* d_workspace, void* stream) {
    using namespace bnn;
    (void)d_workspace;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_x && d_theta_packed && d_out, BNN_E_ARG, "bnn_predict: null pointer");
    BNN_REQUIRE(n_systems > 0 && n_units > 0, BNN_E_ARG, "bnn_predict: empty problem (N=%lld, U=%lld)",
                (long long)n_systems, (long long)n_units);
    BNN_REQUIRE(aligned16(d_theta_packed) && aligned16(d_x), BNN_E_ALIGN, "bnn_predict: pointers must be 16-byte aligned");
    // every (mu, std) pair is one 8-byte store
    BNN_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 7u) == 0 && out_unit_stride >= 2 && out_system_stride >= 2 &&
                    out_unit_stride % 2 == 0 && out_system_stride % 2 == 0,
                BNN_E_ALIGN, "bnn_predict: d_out must be 8-byte aligned and the output strides even (>= 2 floats)");
    PredictParams prm;
    prm.X = d_x;
    prm.thp = d_theta_packed;
    prm.eps = d_eps;
    prm.eps_sum = d_eps_sum;
    prm.summary = d_summary_out;
    prm.out = d_out;
    prm.N = n_systems;
    prm.U = n_units;
    prm.unit_offset = unit_offset;
    prm.system_offset = system_offset;
    prm.out_unit_stride = out_unit_stride;
    prm.out_sys_stride = out_system_stride;
    prm.seed = seed;
    prm.F = cfg->n_features;
    LiveCols lc = live_columns(cfg);
    prm.kin = lc.n;
    for (int c = 0; c < MAXF; ++c) prm.cm.inv[c] = -1;
    for (int k = 0; k < lc.n; ++k) prm.cm.inv[(int)lc.col[k]] = (int8_t)k;
    for (int j = 0; j < 16; ++j) prm.wide_col[j] = (32 + j < lc.n) ? lc.col[32 + j] : (int8_t)-1;
    prm.hc = HeadConsts{cfg->lo_mu, cfg->hi_mu, cfg->lo_sd, cfg->hi_sd};
#ifdef BNN_TC_TIMELINE
    {   // diagnostic build only (make TIMELINE=1): device pointer (decimal) to 8*512 int64 of clock stamps
        const char* dbgp = getenv("BNN_TC_TIMELINE_PTR");
        prm.dbg = dbgp ? reinterpret_cast<long long*>(strtoull(dbgp, nullptr, 10)) : nullptr;
    }
#endif
    // Units are processed in chunks whose weights stay resident in L2: every system tile streams the weights of all
    // units of a launch (~52 kB per unit: tensor-core operands + head), so with 60,000 units (BASELINE configs[2]) one
    // launch would pull 3 GB per tile from HBM; per chunk only x is re-read (16.4 kB per system).  Results do not
    // depend on the chunking (Philox is keyed on the global unit index).  bnn_set_predict_unit_chunk overrides (tests).
    int64_t unit_chunk = predict_unit_chunk() > 0 ? predict_unit_chunk() : 1024;
    if (unit_chunk < 1 || n_units <= unit_chunk + unit_chunk / 2) unit_chunk = n_units;
    else if (predict_unit_chunk() <= 0) {
        // equal launches (no short last one): ceil(n / 1024) launches of ceil(n / launches) units
        const int64_t launches = (n_units + unit_chunk - 1) / unit_chunk;
        unit_chunk = (n_units + launches - 1) / launches;
    }
    const PredictParams prm_all = prm;
    const int L2_ = 2 * L;
    int n_sms = 148;
    {
        int dev = 0;
        BNN_CUDA(cudaGetDevice(&dev));
        BNN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    for (int64_t u0 = 0; u0 < n_units; u0 += unit_chunk) {
    const int64_t n_units_launch = n_units - u0 < unit_chunk ? n_units - u0 : unit_chunk;
    prm = prm_all;
    prm.U = n_units_launch;
    prm.unit_offset = unit_offset + u0;
    prm.thp = d_theta_packed + u0 * (int64_t)PackedLayout(lc.n, cfg->n_features).P;
    prm.out = d_out + u0 * prm_all.out_unit_stride;
    if (d_eps) prm.eps = d_eps + u0 * n_systems * L2_;
    if (d_eps_sum) prm.eps_sum = d_eps_sum + u0 * n_systems * L2_;
    if (d_summary_out) prm.summary = d_summary_out + u0 * n_systems * L2_;
    int rc_launch = BNN_OK;
    // split units over CTAs only when the tiles alone cannot fill the GPU twice
    const int64_t tiles = (n_systems + SYS_TILE - 1) / SYS_TILE;
    int64_t chunks = 1;
    if (tiles < 2 * n_sms) chunks = (2 * n_sms + tiles - 1) / tiles;
    if (chunks > n_units_launch) chunks = n_units_launch;
    prm.units_per_cta = (int)((n_units_launch + chunks - 1) / chunks);
    const int force = predict_variant();
    const int T = cfg->n_times;
    cudaStream_t st = (cudaStream_t)stream;
    auto fits = [&](int nc) { return v2_smem_bytes(prm.kin, prm.F, T, nc) <= 227 * 1024; };
    if (tc::tc_fits(prm, T) && (force == PV_AUTO || force == PV_TC))
        rc_launch = prm.kin <= TC_K1 ? tc::launch_tc<2, false>(prm, st) : tc::launch_tc<2, true>(prm, st);
    else if (force == PV_V1) rc_launch = launch_v1<8>(prm, T, st);
    else if (fits(12)) rc_launch = launch_v2<12>(prm, T, st);
    else if (fits(8)) rc_launch = launch_v2<8>(prm, T, st);
    else rc_launch = launch_v1<8>(prm, T, st);
    if (rc_launch != BNN_OK) return rc_launch;
    }  // unit chunks
    return BNN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Small companions of the noisy / split forward.
// ---------------------------------------------------------------------------------------
namespace bnn {

__global__ void add_input_noise_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                       const float* __restrict__ lv_in, int F, uint64_t zero_mask, int64_t total,
                                       float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % F);
    // x - mask keeps NaN/Inf as NaN in a zeroed column (:452-478)
    const float xv = ((zero_mask >> c) & 1ull) ? __fsub_rn(x[i], x[i]) : x[i];
    out[i] = __fadd_rn(xv, __fmul_rn(eps[i], expf(__fdiv_rn(__ldg(lv_in + c), 2.0f))));
}

// one warp per group of 8 systems: regress_nn + soft_clamp from given summary statistics
__global__ void __launch_bounds__(32) head_only_kernel(const float* __restrict__ summary, int64_t B,
                                                       const float* __restrict__ thp, PackedLayout pl, HeadConsts hc,
                                                       float* __restrict__ out) {
    __shared__ float sA[SYS_TILE * 41], sB[SYS_TILE * 41];
    const int lane = threadIdx.x, p = lane >> 2, q = lane & 3;
    const int64_t n0 = (int64_t)blockIdx.x * SYS_TILE;
    const int n_valid = (int)min((int64_t)SYS_TILE, B - n0);
    for (int idx = lane; idx < SYS_TILE * S2; idx += 32) {
        const int s = idx / S2, j = idx % S2;
        sA[s * 41 + j] = (s < n_valid) ? summary[(n0 + s) * S2 + j] : 0.f;
    }
    __syncwarp();
    head_layer<true>(sA, thp + pl.V0p, thp + pl.c0p, S2, p, q, sB);
    __syncwarp();
    head_layer<true>(sB, thp + pl.V1p, thp + pl.c1p, H, p, q, sA);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = q * 10 + i;
        const float r = sA[p * 41 + k];
        o0 = fmaf(r, __ldg(thp + pl.V2 + k), o0);
        o1 = fmaf(r, __ldg(thp + pl.V2 + H + k), o1);
    }
    o0 += __shfl_xor_sync(0xffffffffu, o0, 1);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 1);
    o0 += __shfl_xor_sync(0xffffffffu, o0, 2);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
    if (q == 0 && p < n_valid) {
        o0 += __ldg(thp + pl.c2);
        o1 += __ldg(thp + pl.c2 + 1);
        out[(n0 + p) * 2] = soft_clamp_dev(o0, hc.lo_mu, hc.hi_mu);
        out[(n0 + p) * 2 + 1] = soft_clamp_dev(o1, hc.lo_sd, hc.hi_sd);
    }
}

}  // namespace bnn

extern "C" {

int bnn_add_input_noise(const bnn_model_config* cfg, const float* d_x, const float* d_eps_in, const float* d_lv_in,
                        int64_t n_rows, float* d_x_noisy, void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_x && d_eps_in && d_lv_in && d_x_noisy && n_rows > 0, BNN_E_ARG, "bnn_add_input_noise: null pointer");
    const int64_t total = n_rows * cfg->n_features;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "bnn_add_input_noise: too many rows");
    add_input_noise_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        d_x, d_eps_in, d_lv_in, cfg->n_features, cfg->zero_mask, total, d_x_noisy);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

int bnn_predict_instability(const bnn_model_config* cfg, const float* d_summary, int64_t B,
                            const float* d_theta_packed, float* d_out, void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_summary && d_theta_packed && d_out && B > 0, BNN_E_ARG, "bnn_predict_instability: null pointer");
    PackedLayout pl(live_columns(cfg).n, cfg->n_features);
    HeadConsts hc{cfg->lo_mu, cfg->hi_mu, cfg->lo_sd, cfg->hi_sd};
    const int64_t blocks = (B + SYS_TILE - 1) / SYS_TILE;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "bnn_predict_instability: B too large");
    head_only_kernel<<<(unsigned)blocks, 32, 0, (cudaStream_t)stream>>>(d_summary, B, d_theta_packed, pl, hc, d_out);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

}  // extern "C"
