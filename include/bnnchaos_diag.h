/*
 * bnnchaos_diag.h -- DIAGNOSTIC entry points of libbnnchaos.so: hardware probes and timers used by tests/test_gpu_tc.py,
 * bench.py's FFMA-peak cross-check and tools/.  They are NOT part of the product ABI (include/bnnchaos.h): nothing on
 * the MultiSWAG predictive / SWAG-training path calls them, and a caller binding the reference-facing interface does
 * not need this header.
 */
#ifndef BNNCHAOS_DIAG_H_
#define BNNCHAOS_DIAG_H_

#include <stddef.h>
#include <stdint.h>

#include "bnnchaos.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Measured-FFMA-peak micro-kernel (roofline denominator check): runs `iters` dependent-free
 * fma.rn.f32x2 (packed=1) or fma.rn.f32 (packed=0) per thread on every SM, returns via
 * d_sink.  flops = 2 * grid*block*iters*16 ; time it with events around the call. */
int bnn_ffma_peak(int32_t packed, int64_t iters, float* d_sink, int64_t* flops_out, void* stream);

/* Diagnostic: one tcgen05 (kind::tf32) GEMM D[128,N] = A[128,K] B[N,K]^T with A staged in TMEM and
 * B in shared memory (canonical K-major, no swizzle); variant 0 is the descriptor convention the
 * tensor-core predictive kernel uses.  Used by tests/test_gpu_tc.py to pin the hardware layouts. */
int bnn_tc_probe(const float* d_A, const float* d_B, float* d_D, int32_t K, int32_t N, int32_t variant,
                 void* stream);

/* Diagnostic: cycles for reps x (K/8) dependent tcgen05.mma (M=128, kind::tf32) issued back to back;
 * d_out[0] = issue-to-completion cycles, d_out[1] = cycles spent issuing.  from_smem: A from shared memory. */
int bnn_tc_time(int32_t K, int32_t N, int32_t reps, int32_t from_smem, long long* d_out, void* stream);

/* Diagnostic: issue rate of the warp-level mma.sync.m16n8k8 (tf32 x tf32 -> fp32, operands in registers): 148 CTAs of
 * warps_per_cta warps each run iters rounds of n_acc (1, 4, 8 or 15) independent MMAs; d_out[0] = cycles of warp 0 of
 * CTA 0, d_out[1] = MMAs per warp; d_sink: 148 * 32 * warps_per_cta floats. */
int bnn_mma_sync_rate(int32_t warps_per_cta, int32_t n_acc, int32_t iters, long long* d_out, float* d_sink, void* stream);

/* Diagnostic: per-phase cycle totals of CTA (0,0) of the last bnn_train_step (v3 kernel), copied to host_out[n].
 * All zeros unless the library was built with `make TRAIN_TIMELINE=1` (the stamps are compiled out by default).
 * Synchronises the device. */
int bnn_train_timeline(unsigned long long* host_out, int32_t n);

/* Diagnostic: D[128, N] (lanes j < MJ, columns k < NK meaningful) = sum_r G[r][j] * Hm[r][k] -- the training step's
 * weight-gradient GEMM, contraction over the R time-step rows -- as tcgen05.mma kind::tf32 with BOTH operands in shared
 * memory, canonical K-major no-swizzle layout with K = rows ([row quad][feature][4 rows]), M = 128 with the surplus
 * feature rows reading whatever follows.  d_G [R, MJ], d_H [R, NK] row-major fp32.  bias_round: store bits + 0x1000 so
 * that the tensor core's truncation rounds to nearest tf32; two_batches: accumulate across two commits. */
int bnn_tc_probe_ss(const float* d_G, const float* d_H, float* d_D, int32_t R, int32_t MJ, int32_t NK, int32_t N,
                    int32_t bias_round, int32_t two_batches, void* stream);

/* Diagnostic: the unfused K1 -- swag_sample_kernel writes theta [U,d], pack_theta_kernel gathers it into the packed layout
 * -- with bnn_swag_sample's arguments (d_theta required).  bnn_swag_sample itself runs the fused kernel, which must give
 * bit-identical results (tests/test_gpu_predict.py). */
int bnn_swag_sample_unfused(const bnn_model_config* cfg, const float* d_w_avg, const float* d_w2_avg,
                            const float* d_pre_D, int32_t n_models, int32_t K, const int32_t* d_unit_model,
                            int64_t n_units, int64_t unit_offset, int32_t samples_per_model, float scale, uint64_t seed,
                            const float* d_z1, const float* d_z2, float* d_theta, float* d_theta_packed, void* stream);

/* Diagnostic: force the kernel behind bnn_predict for this process: 0 = automatic (tensor cores when T = 100 and at most 32
 * live columns, else FP32 FFMA2), 1 = tensor-core, 2 = FFMA2 v2 (warp-specialised), 3 = FFMA2 v1 (synchronous).  Default:
 * read once from BNN_PREDICT_VARIANT (tc | v2 | v1).  bnn_set_predict_unit_chunk: units per launch of bnn_predict's
 * L2-sized unit chunks (0 = default 1024; the results do not depend on it). */
int bnn_set_predict_variant(int32_t variant);
int bnn_set_predict_unit_chunk(int64_t units);

/* Diagnostic: 1 forces the radix-select path of bnn_summarize_instability (default 0: shared-memory sort while it fits). */
int bnn_set_summary_variant(int32_t variant);

/* Diagnostic: force the kernel behind bnn_train_step / bnn_train_noise for this process: 0 = automatic (the tensor-core
 * kernel where its shared-memory plan fits, else the FP32 FFMA kernel), 1 = tensor-core, 2 = FP32 FFMA.  The default is
 * read once from the environment variable BNN_TRAIN_VARIANT (tc | v3). */
int bnn_set_train_variant(int32_t variant);

/* Diagnostic: launches per training step.  A CTA of the training kernels belongs to one seed, so floor(SMs / n_seeds) CTAs
 * per seed can leave SMs idle (30 seeds: 4 x 30 = 120 of 148); bnn_train_step then runs the seeds in groups, one launch
 * per group, chosen by a cost model (0, the default).  1..64 forces that many groups.  A seed's gradient depends on the
 * number of CTAs that summed it (rounding order), not on the grouping as such. */
int bnn_set_train_seed_groups(int32_t groups);
/* The plan bnn_train_step uses for n_seeds seeds of batch B on the current device (148 SMs are assumed when no device is
 * present): CTAs per seed, launches per step, seeds per launch.  Host-only query. */
int bnn_train_seed_plan(const bnn_model_config* cfg, int64_t B, int32_t n_seeds, int32_t* n_cta, int32_t* groups,
                        int32_t* seeds_per_group);

#ifdef __cplusplus
}
#endif
#endif /* BNNCHAOS_DIAG_H_ */
