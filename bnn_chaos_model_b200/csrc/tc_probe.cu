// Hardware probe for the tcgen05 building blocks used by the tensor-core predictive kernel:
// D[128,N] = A[128,K] * B[N,K]^T with kind::tf32, A in TMEM (written with tcgen05.st), B in
// shared memory in the canonical K-major no-swizzle layout, D read back with tcgen05.ld.
// tests/test_gpu_tc.py compares against numpy for several (N, K, descriptor) variants.
#include "common.cuh"
#include "tc.cuh"

namespace bnn {

__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ D, int K, int N, int variant) {
    extern __shared__ __align__(128) float bs[];  // B canonical: chunk-major [K/4][N][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 256);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init_fence();
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        bs[((k >> 2) * N + n) * 4 + (k & 3)] = B[i];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t a_col = 0, d_col = 128;
    // A row `tid` -> TMEM lane tid, columns [0,K)
    for (int c = 0; c < K; c += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[tid * K + c + j]);
        tmem_st8(tmem + lane_base + a_col + c, v);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0 && !(variant & 4)) {
        const uint32_t idesc = idesc_tf32(128, N);
        const uint32_t chunk_bytes = (uint32_t)N * 16u;  // distance between 16-byte K chunks
        const uint32_t lbo = (variant & 1) ? 128u : chunk_bytes;
        const uint32_t sbo = (variant & 1) ? chunk_bytes : 128u;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t bdesc = smem_desc_kmajor(smem_u32(bs) + ks * 2 * chunk_bytes, lbo, sbo);
            mma_tf32_ts(tmem + d_col, tmem + a_col + ks * 8, bdesc, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    if (!(variant & 4)) mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t rd_col = (variant & 2) ? a_col : d_col;  // variant&2: read back the A region instead of D
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + lane_base + rd_col + c, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// SS probe: D[j][k] = sum_r G[r][j] * Hm[r][k] (the weight-gradient GEMM of the training step: the contraction runs over
// the time-step ROWS) with BOTH operands in shared memory in the canonical K-major no-swizzle layout with K = rows:
// [row quad][feature][4 rows] fp32, feature pitch 16 B (one core matrix = 8 features x 4 rows), 8-feature groups 128 B
// apart (SBO), row quads `pitch` features apart (LBO).  M = 128: feature rows beyond MJ read whatever follows in shared
// memory (their D lanes are garbage and ignored), likewise N beyond NK.  bias_round: operands are stored as
// bits + 0x1000, so that the tensor core's truncation of the 13 low mantissa bits rounds to nearest.  two_batches: the
// k-steps are issued in two halves with a commit + wait in between (accumulation across commits).
__global__ void __launch_bounds__(128) tc_probe_ss_kernel(const float* __restrict__ G, const float* __restrict__ Hm,
                                                          float* __restrict__ D, int R, int MJ, int NK, int N,
                                                          int bias_round, int two_batches) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NQ = R / 4, PG = MJ | 1, PH = NK | 1;
    float* ga = sm;
    float* hb = sm + NQ * PG * 4;
    float* tail = hb + NQ * PH * 4;
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 256);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init_fence();
    }
    const uint32_t add = bias_round ? 0x1000u : 0u;
    for (int i = tid; i < NQ * PG * 4; i += 128) ga[i] = 0.f;
    for (int i = tid; i < NQ * PH * 4; i += 128) hb[i] = 0.f;
    for (int i = tid; i < 1024; i += 128) tail[i] = 3.0e30f;   // what the out-of-range feature rows of the last quad read
    __syncthreads();
    for (int i = tid; i < R * MJ; i += 128) {
        const int r = i / MJ, j = i - r * MJ;
        ga[((r >> 2) * PG + j) * 4 + (r & 3)] = __uint_as_float(__float_as_uint(G[i]) + add);
    }
    for (int i = tid; i < R * NK; i += 128) {
        const int r = i / NK, k = i - r * NK;
        hb[((r >> 2) * PH + k) * 4 + (r & 3)] = __uint_as_float(__float_as_uint(Hm[i]) + add);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int KS = R / 8, half = two_batches ? KS / 2 : KS;
    uint32_t par = 0;
    for (int batch = 0; batch < (two_batches ? 2 : 1); ++batch) {
        if (tid == 0) {
            const uint32_t idesc = idesc_tf32(128, N);
            const int k0 = batch ? half : 0, k1 = batch ? KS : half;
            for (int ks = k0; ks < k1; ++ks) {
                const uint64_t ad = smem_desc_kmajor(smem_u32(ga) + (uint32_t)ks * 2u * PG * 16u, (uint32_t)PG * 16u, 128u);
                const uint64_t bd = smem_desc_kmajor(smem_u32(hb) + (uint32_t)ks * 2u * PH * 16u, (uint32_t)PH * 16u, 128u);
                mma_tf32_ss(tmem, ad, bd, idesc, ks > 0);
            }
            mma_commit(&bar);
        }
        mbar_wait(&bar, par);
        par ^= 1;
        __syncthreads();
    }
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + lane_base + c, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) D[tid * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// Timing probe: `reps` x (K/8) dependent tcgen05.mma (same D) issued back to back by one thread;
// out[0] = cycles from first issue to commit completion, out[1] = cycles spent issuing.
__global__ void __launch_bounds__(128) tc_time_kernel(int K, int N, int reps, int from_smem, long long* out) {
    extern __shared__ __align__(128) float bs[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        mbar_init_fence();
    }
    for (int i = tid; i < (N + 128) * K; i += 128) bs[i] = 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // from_smem bit 0: A from shared memory; bit 8: a second warp issues an identical, independent chain (its own D and
    // A columns) at the same time -- tells a per-issuing-thread cost from a shared tensor-pipe service time
    const bool two = (from_smem & 256) != 0;
    if (warp == 0 || (two && warp == 1)) {
        const uint32_t idesc = idesc_tf32(128, N);
        const uint32_t chunk_bytes = (uint32_t)N * 16u;
        const uint32_t a_chunk = 128u * 16u;
        const uint32_t a_base = smem_u32(bs) + (uint32_t)N * K * 4;
        long long t0 = 0, t1 = 0;
        const uint64_t bd0 = smem_desc_kmajor(smem_u32(bs), chunk_bytes, 128u);
        const uint64_t ad0 = smem_desc_kmajor(a_base, a_chunk, 128u);
        const uint64_t bstep = 2ull * (uint64_t)N, astep = 2ull * 128ull;
        const uint32_t dcol = tmem + (warp == 0 ? 256u : 384u), acol = tmem + (warp == 0 ? 0u : 64u);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 6; ++ks) {
                    if (from_smem & 1)
                        mma_tf32_ss(dcol, ad0 + ks * astep, bd0 + ks * bstep, idesc, (r | ks) > 0);
                    else
                        mma_tf32_ts(dcol, acol + ks * 8, bd0 + ks * bstep, idesc, (r | ks) > 0);
                }
            }
            __syncwarp();
        }
        t1 = clock64();
        if (elect_one_sync()) mma_commit(warp == 0 ? &bar : &bar2);
        __syncwarp();
        mbar_wait(warp == 0 ? &bar : &bar2, 0);
        const long long t2 = clock64();
        if (tid == 0) {
            out[0] = t2 - t0;
            out[1] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace bnn

extern "C" int bnn_tc_time(int32_t K, int32_t N, int32_t reps, int32_t from_smem, long long* d_out, void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_out && K % 8 == 0 && K >= 8 && K <= 64 && N % 16 == 0 && N >= 16 && N <= 256, BNN_E_ARG,
                "bnn_tc_time: bad arguments");
    const size_t smem = (size_t)(N + 128) * K * 4;
    cudaFuncSetAttribute(tc_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    tc_time_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(K, N, reps, from_smem, d_out);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

extern "C" int bnn_tc_probe_ss(const float* d_G, const float* d_H, float* d_D, int32_t R, int32_t MJ, int32_t NK, int32_t N,
                               int32_t bias_round, int32_t two_batches, void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_G && d_H && d_D, BNN_E_ARG, "bnn_tc_probe_ss: null pointer");
    BNN_REQUIRE(R % 8 == 0 && R >= 8 && R <= 128 && MJ >= 1 && MJ <= 128 && NK >= 1 && NK <= N && N % 16 == 0 && N >= 16 &&
                    N <= 256, BNN_E_ARG, "bnn_tc_probe_ss: need R%%8==0 (8..128), MJ<=128, NK<=N, N%%16==0 (16..256)");
    const size_t smem = ((size_t)(R / 4) * ((MJ | 1) + (NK | 1)) * 4 + 1024) * sizeof(float);
    BNN_REQUIRE(smem <= 200 * 1024, BNN_E_ARG, "bnn_tc_probe_ss: operands do not fit in shared memory");
    cudaFuncSetAttribute(tc_probe_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    tc_probe_ss_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d_G, d_H, d_D, R, MJ, NK, N, bias_round, two_batches);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

extern "C" int bnn_tc_probe(const float* d_A, const float* d_B, float* d_D, int32_t K, int32_t N, int32_t variant,
                            void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_A && d_B && d_D, BNN_E_ARG, "bnn_tc_probe: null pointer");
    BNN_REQUIRE(K % 8 == 0 && K >= 8 && K <= 128 && N % 16 == 0 && N >= 16 && N <= 128, BNN_E_ARG,
                "bnn_tc_probe: need K%%8==0 (8..128), N%%16==0 (16..128)");
    tc_probe_kernel<<<1, 128, (size_t)N * K * 4, (cudaStream_t)stream>>>(d_A, d_B, d_D, K, N, variant);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

// ---------------------------------------------------------------------------------------
// Warp-level mma.sync.m16n8k8 (tf32, fp32 accumulate) issue-rate probe: every warp of the CTA runs `iters` rounds of
// NACC independent MMAs (operands in registers).  out[0] = cycles of warp 0, out[1] = MMAs per warp.
// ---------------------------------------------------------------------------------------
namespace bnn {
template <int NACC>
__global__ void mma_sync_rate_kernel(int iters, long long* out, float* sink) {
    float c[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.001f * (float)((threadIdx.x + i) & 7));
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + 0.001f * (float)((threadIdx.x + i) & 3));
#pragma unroll
    for (int n = 0; n < NACC; ++n)
        for (int i = 0; i < 4; ++i) c[n][i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[n][0]), "+f"(c[n][1]), "+f"(c[n][2]), "+f"(c[n][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < NACC; ++n) s += c[n][0] + c[n][1] + c[n][2] + c[n][3];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = (long long)iters * NACC;
    }
}
}  // namespace bnn

extern "C" int bnn_mma_sync_rate(int32_t warps_per_cta, int32_t n_acc, int32_t iters, long long* d_out, float* d_sink,
                                 void* stream) {
    using namespace bnn;
    int rc = check_device();
    if (rc != BNN_OK) return rc;
    BNN_REQUIRE(d_out && d_sink && warps_per_cta >= 1 && warps_per_cta <= 32 && iters >= 1, BNN_E_ARG,
                "bnn_mma_sync_rate: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_acc == 1) mma_sync_rate_kernel<1><<<148, 32 * warps_per_cta, 0, st>>>(iters, d_out, d_sink);
    else if (n_acc == 4) mma_sync_rate_kernel<4><<<148, 32 * warps_per_cta, 0, st>>>(iters, d_out, d_sink);
    else if (n_acc == 8) mma_sync_rate_kernel<8><<<148, 32 * warps_per_cta, 0, st>>>(iters, d_out, d_sink);
    else if (n_acc == 15) mma_sync_rate_kernel<15><<<148, 32 * warps_per_cta, 0, st>>>(iters, d_out, d_sink);
    else BNN_REQUIRE(false, BNN_E_ARG, "bnn_mma_sync_rate: n_acc must be 1, 4, 8 or 15");
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}
