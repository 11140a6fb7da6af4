// K4: fused SWAG training step -- noisy forward + analytic backward + clip + SGD-momentum for n_seeds
// independent models in one call, and the validation loss (bnn_eval_loss).
//
// Reference: SWAGModel.training_step (/root/reference/spock_reg_model.py:722-732) -> lossfnc (:579-583) ->
// VarModel.forward(noisy_val=True) (:486-528) -> _lossfnc (:547-577); loss.backward(); Lightning's
// gradient_clip_val = clip_grad_norm_ (run_swag.py:61,74-79); torch.optim.SGD(momentum, weight_decay) (:709-711).
//
// total = sum_b nll_b + beta_in * B * 1/2 sum_c (e^{lvin_c} - lvin_c - 1) + beta_out * sum_{b,j} 1/2 (s_bj^2 + e^{lvs_j} - lvs_j - 1)
//
// Backward (per system; T rows, n = T), derived by hand and checked against the reference's autograd gradient
// (tests/golden/train_v50.npz):
//   head:   g_r = (g_mu * (hi-lo)/2 (1 - tanh^2 r0), g_sd * ...);  dV2 += g_r r2^T;  g_a2 = (V2^T g_r) . [r2>0];
//           dV1 += g_a2 r1^T;  g_a1 = (V1^T g_a2) . [r1>0];  dV0 += g_a1 s'^T;  g_s' = V0^T g_a1
//   noise:  s' = s + eps_sum e^{lvs/2}:  g_s = g_s' + beta_out s;  dlvs += g_s' (s'-s)/2  (+ beta_out B (e^{lvs}-1)/2)
//   stats:  s = [mu_s, sqrt(|v_s|+1e-5)], mu_s = eps1 sqrt(v/n) + m, v_s = eps2 sqrt(2 v^2/(n-1)) + v:
//           g_m = g_mus;  g_v = g_mus eps1 / (2 n sqrt(v/n)) + g_vs (1 + eps2 2 v / ((n-1) sqrt(2 v^2/(n-1)))),
//           g_vs = g_sds sign(v_s) / (2 sd_s);   g_f[t] = g_m / n + g_v 2 (f_t - m) / (n-1)      (v = std^2, unbiased)
//   MLP:    dW2 += g_f^T h2; g_a2 = (g_f W2) . [h2>0]; dW1 += g_a2^T h1; g_a1 = (g_a2 W1) . [h1>0]; dW0 += g_a1^T x';
//           g_x = g_a1 W0;  x' = mask(x) + eps_in e^{lvin/2}:  dlvin += g_x . (x' - mask(x)) / 2  (+ beta_in B (e^{lvin}-1)/2)
//
// Decomposition: one CTA owns a seed and walks that seed's systems (grid = n_cta x n_seeds); a system's T x 41
// input, both hidden activations and the latent rows stay in shared memory (feature-major [feature][row]) for the
// whole forward + backward; every thread accumulates its own block of every weight gradient in registers over all
// systems of the CTA, writes one partial gradient vector per CTA, and a second kernel adds the partials in a fixed
// order (bit-reproducible SWAG moments), a third computes the global norm, clips and applies SGD.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "loss_device.cuh"
#include "predict_device.cuh"

namespace bnn {
namespace train {

constexpr int NTHR = 288;          // 9 warps: 25 row quads x 11 column groups of the widest row GEMM (T = 100)
constexpr int DPAD = 8;            // metric slots appended to each gradient partial
constexpr int SLOT_NLL = 0, SLOT_SKL = 1;

struct Params {
    const float* theta;        // [n_seeds, d]
    const float* X;            // [n_data, T, F]
    const float* Y;            // [n_data, 2]
    const int32_t* batch_index;  // [n_seeds, B] rows of X / Y, or null (row = b)
    const float* eps_in;       // [n_seeds, B, T, F] or null
    const float* eps12;        // [n_seeds, B, 2L] or null
    const float* eps_sum;      // [n_seeds, B, 2L] or null
    float* partial;            // [n_seeds, n_cta, d + DPAD]
    float* head_rec;           // [n_seeds, B, REC] per-system head vectors (v3 only)
    float* xprod;              // [n_seeds, n_cta, 2, 2, F, 2T] input-tile scratch of the producer warps (v3 only)
    int saliency;              // v3 only: d mu / d x instead of the training gradient (bnn_saliency)
    float* gx_out;             // [n_seeds, B, T, F] or null
    float* mu_out;             // [n_seeds, B]
    int B, T, F, FP, n_cta;
    uint64_t seed, step;
    uint64_t zero_mask;
    HeadConsts hc;
    float beta_out;
};

// Philox key of one seed-model's noise streams; counters: (block, batch position, step, stream)
__host__ __device__ inline uint64_t seed_key(uint64_t seed, int seed_index) {
    return seed + 0x9E3779B97F4A7C15ull * (uint64_t)(seed_index + 1);
}

struct Smem {
    // offsets in floats
    int xT, nT, h1T, h2T, fT, W0T, b0, W1T, b1, W2T, b2, W2n, W1n, W0n, small, total;
    __host__ __device__ Smem(int T, int F, int FP) {
        int o = 0;
        xT = o; o += F * T;
        nT = o; o += F * T;
        h1T = o; o += H * T;
        h2T = o; o += H * T;
        fT = o; o += L * T;
        W0T = o; o += F * H;
        b0 = o; o += H;
        W1T = o; o += H * H;
        b1 = o; o += H;
        W2T = o; o += H * L;
        b2 = o; o += L;
        W2n = o; o += L * H;
        W1n = o; o += H * H;
        W0n = o; o += H * FP;
        small = o; o += 768;
        total = o;
    }
};
// layout of the `small` region (floats)
enum { SM_M = 0, SM_VAR = 20, SM_SIM = 40, SM_SIV = 60, SM_VS = 80, SM_S = 100, SM_SP = 140, SM_R1 = 180, SM_R2 = 220,
       SM_G2 = 260, SM_G1 = 300, SM_GS = 340, SM_GM = 380, SM_GV = 400, SM_E12 = 420, SM_ESN = 460, SM_ELVH = 500,
       SM_LVS = 540, SM_NSC = 580 /* exp(lv_in/2), up to 64 */, SM_R = 644, SM_GR = 648, SM_Y = 652 };

// C[4 rows of quad q][4 columns of group cg] += sum_k AT[k][4q..4q+3] * W[k][4cg..4cg+3]
// Accumulators are column pairs (fma.rn.f32x2): 8 FFMA2 + 4 operand packs per k instead of 16 FFMA.
__device__ __forceinline__ void rowgemm4x4(const float* __restrict__ AT, int RP, int K, const float* __restrict__ W,
                                           int NP, int q, int cg, float (&acc)[4][4]) {
    const float* ap = AT + 4 * q;
    const float* wp = W + 4 * cg;
    u64 a2[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) { a2[r][0] = pack2(acc[r][0], acc[r][1]); a2[r][1] = pack2(acc[r][2], acc[r][3]); }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(ap + k * RP);
        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(wp + k * NP);
        const u64 av[4] = {pack2(a.x, a.x), pack2(a.y, a.y), pack2(a.z, a.z), pack2(a.w, a.w)};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            a2[r][0] = fma2(av[r], w.x, a2[r][0]);
            a2[r][1] = fma2(av[r], w.y, a2[r][1]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) { unpack2(a2[r][0], acc[r][0], acc[r][1]); unpack2(a2[r][1], acc[r][2], acc[r][3]); }
}

// acc[jj][kk] += sum_r G[j0+jj][r] * Hm[krow[kk]][r]   (both feature-major with row pitch RP = T)
__device__ __forceinline__ void outer_acc(const float* __restrict__ G, const float* __restrict__ Hm, int T, int j0,
                                          const int (&krow)[4], float (&acc)[2][4]) {
    const float* g0p = G + j0 * T;
    const float* g1p = g0p + T;
#pragma unroll 2
    for (int r = 0; r < T; r += 4) {
        const float4 g0 = *reinterpret_cast<const float4*>(g0p + r);
        const float4 g1 = *reinterpret_cast<const float4*>(g1p + r);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 h = *reinterpret_cast<const float4*>(Hm + krow[kk] * T + r);
            acc[0][kk] = fmaf(g0.x, h.x, acc[0][kk]); acc[0][kk] = fmaf(g0.y, h.y, acc[0][kk]);
            acc[0][kk] = fmaf(g0.z, h.z, acc[0][kk]); acc[0][kk] = fmaf(g0.w, h.w, acc[0][kk]);
            acc[1][kk] = fmaf(g1.x, h.x, acc[1][kk]); acc[1][kk] = fmaf(g1.y, h.y, acc[1][kk]);
            acc[1][kk] = fmaf(g1.z, h.z, acc[1][kk]); acc[1][kk] = fmaf(g1.w, h.w, acc[1][kk]);
        }
    }
}

__device__ __forceinline__ float row_sum(const float* __restrict__ G, int T) {
    float s = 0.f;
    for (int r = 0; r < T; r += 4) {
        const float4 g = *reinterpret_cast<const float4*>(G + r);
        s += (g.x + g.y) + (g.z + g.w);
    }
    return s;
}

__global__ void __launch_bounds__(NTHR, 2) train_fwd_bwd_kernel(const Params prm) {
    extern __shared__ __align__(16) float sm[];
    const int T = prm.T, F = prm.F, FP = prm.FP, NQ = T >> 2;
    const Smem L_(T, F, FP);
    const FlatLayout fl(F);
    const int tid = threadIdx.x, lane = tid & 31;
    const int sidx = blockIdx.y;
    const float* th = prm.theta + (int64_t)sidx * fl.d;
    float* xT = sm + L_.xT; float* nT = sm + L_.nT; float* h1T = sm + L_.h1T; float* h2T = sm + L_.h2T; float* fT = sm + L_.fT;
    float* W0T = sm + L_.W0T; float* b0 = sm + L_.b0; float* W1T = sm + L_.W1T; float* b1 = sm + L_.b1;
    float* W2T = sm + L_.W2T; float* b2 = sm + L_.b2; float* W2n = sm + L_.W2n; float* W1n = sm + L_.W1n; float* W0n = sm + L_.W0n;
    float* sv = sm + L_.small;

    // ---- stage this seed's feature weights (natural and transposed) and noise scales ----
    for (int i = tid; i < H * F; i += NTHR) {
        const int j = i / F, c = i - j * F;
        const float w = __ldg(th + fl.W0 + i);
        W0T[c * H + j] = w;
        W0n[j * FP + c] = w;
    }
    for (int i = tid; i < H * (FP - F); i += NTHR) W0n[(i / (FP - F)) * FP + F + i % (FP - F)] = 0.f;
    for (int i = tid; i < H * H; i += NTHR) {
        const int j = i / H, k = i - j * H;
        const float w = __ldg(th + fl.W1 + i);
        W1T[k * H + j] = w;
        W1n[i] = w;
    }
    for (int i = tid; i < L * H; i += NTHR) {
        const int j = i / H, k = i - j * H;
        const float w = __ldg(th + fl.W2 + i);
        W2T[k * L + j] = w;
        W2n[i] = w;
    }
    if (tid < H) { b0[tid] = __ldg(th + fl.b0 + tid); b1[tid] = __ldg(th + fl.b1 + tid); }
    if (tid < L) b2[tid] = __ldg(th + fl.b2 + tid);
    if (tid < S2) {
        const float lv = __ldg(th + fl.lv_sum + tid);
        sv[SM_LVS + tid] = lv;
        sv[SM_ELVH + tid] = expf(__fdiv_rn(lv, 2.0f));
    }
    if (tid < F) sv[SM_NSC + tid] = expf(__fdiv_rn(__ldg(th + fl.lv_in + tid), 2.0f));

    // ---- thread-owned gradient accumulators (summed over this CTA's systems) ----
    float aW0[2][4] = {}, aW1[2][4] = {}, aW2[2][4] = {}, aV0[2][4] = {}, aV1[2][4] = {};
    float ab0 = 0.f, ab1 = 0.f, ab2 = 0.f, aV2 = 0.f, ac0 = 0.f, ac1 = 0.f, ac2 = 0.f, alvs = 0.f;
    float alvin[4] = {0.f, 0.f, 0.f, 0.f};
    float a_nll = 0.f, a_skl = 0.f;
    // (j2 x k4) blocks of the 40x40 / 40xF / 20x40 gradients
    const int jb = tid / 10, kb10 = tid % 10;      // dW1, dV0, dV1, dW2: 10 k-groups
    const int KG0 = FP >> 2;                       // dW0: FP/4 k-groups (F = 41 -> 11); 20 * KG0 <= 248 threads
    const int jb0 = tid / KG0, kb11 = tid % KG0;
    // a thread's 4 k's are strided by the group count (k = kb + KG*kk): for a fixed kk the lanes of a warp then read
    // CONSECUTIVE feature rows, whose 400-byte pitch (T = 100) walks the 16-byte bank groups -- contiguous k's per
    // thread put rows 4 apart on the same banks (5-way conflicts, 79 % of the LSU wavefront peak in ncu)
    int krow0[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) krow0[kk] = min(kb11 + KG0 * kk, F - 1);
    const int krow10[4] = {kb10, kb10 + 10, kb10 + 20, kb10 + 30};
    const int q_rg = tid % NQ, cg_rg = tid / NQ;   // row-GEMM item of this thread
    const float Tf = (float)T, Tm1 = (float)(T - 1);
    const uint64_t key = seed_key(prm.seed, sidx);
    __syncthreads();

    for (int b = blockIdx.x; b < prm.B; b += gridDim.x) {
        const int64_t sb = (int64_t)sidx * prm.B + b;
        const int64_t row = prm.batch_index ? (int64_t)prm.batch_index[sb] : (int64_t)b;
        // ---- S0: x' = mask(x) + eps_in * exp(lv_in/2), feature-major; noise term kept for dlv_in ----
        {
            const float* xs = prm.X + row * (int64_t)T * F;
            const float* es = prm.eps_in ? prm.eps_in + sb * (int64_t)T * F : nullptr;
            if (es) {
                for (int i = tid; i < T * F; i += NTHR) {
                    const int t = i / F, c = i - t * F;
                    float xv = __ldg(xs + i);
                    if ((prm.zero_mask >> c) & 1ull) xv = __fsub_rn(xv, xv);  // x - mask keeps NaN (:452-478)
                    const float nz = __fmul_rn(__ldg(es + i), sv[SM_NSC + c]);
                    nT[c * T + t] = nz;
                    xT[c * T + t] = __fadd_rn(xv, nz);
                }
            } else {
                const int F4 = (F + 3) >> 2;
                for (int i = tid; i < T * F4; i += NTHR) {
                    const int t = i / F4, c4 = i - t * F4;
                    const float4 n4 = philox_normal4_fast(key, STREAM_EPS_IN, (uint32_t)b, (uint32_t)prm.step, (uint32_t)i);
                    const float e[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = 4 * c4 + u;
                        if (c < F) {
                            float xv = __ldg(xs + t * F + c);
                            if ((prm.zero_mask >> c) & 1ull) xv = __fsub_rn(xv, xv);
                            const float nz = __fmul_rn(e[u], sv[SM_NSC + c]);
                            nT[c * T + t] = nz;
                            xT[c * T + t] = __fadd_rn(xv, nz);
                        }
                    }
                }
            }
            if (tid < S2 / 4) {
                float4 a, c;
                if (prm.eps12) {
                    a = __ldg(reinterpret_cast<const float4*>(prm.eps12 + sb * S2) + tid);
                    c = __ldg(reinterpret_cast<const float4*>(prm.eps_sum + sb * S2) + tid);
                } else {
                    a = philox_normal4(key, STREAM_EPS, (uint32_t)b, (uint32_t)prm.step, (uint32_t)tid);
                    c = philox_normal4(key, STREAM_EPS_SUM, (uint32_t)b, (uint32_t)prm.step, (uint32_t)tid);
                }
                *reinterpret_cast<float4*>(sv + SM_E12 + 4 * tid) = a;
                *reinterpret_cast<float4*>(sv + SM_ESN + 4 * tid) = c;
            }
            if (tid == 32) {
                sv[SM_Y] = __ldg(prm.Y + row * 2);
                sv[SM_Y + 1] = __ldg(prm.Y + row * 2 + 1);
            }
        }
        __syncthreads();
        // ---- S1..S3: feature_nn forward ----
        if (cg_rg < 10) {
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float bv = b0[4 * cg_rg + c]; for (int r = 0; r < 4; ++r) acc[r][c] = bv; }
            rowgemm4x4(xT, T, F, W0T, H, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(h1T + (4 * cg_rg + c) * T + 4 * q_rg) =
                    make_float4(relu_nan(acc[0][c]), relu_nan(acc[1][c]), relu_nan(acc[2][c]), relu_nan(acc[3][c]));
        }
        __syncthreads();
        if (cg_rg < 10) {
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float bv = b1[4 * cg_rg + c]; for (int r = 0; r < 4; ++r) acc[r][c] = bv; }
            rowgemm4x4(h1T, T, H, W1T, H, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(h2T + (4 * cg_rg + c) * T + 4 * q_rg) =
                    make_float4(relu_nan(acc[0][c]), relu_nan(acc[1][c]), relu_nan(acc[2][c]), relu_nan(acc[3][c]));
        }
        __syncthreads();
        if (cg_rg < 5) {
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float bv = b2[4 * cg_rg + c]; for (int r = 0; r < 4; ++r) acc[r][c] = bv; }
            rowgemm4x4(h2T, T, H, W2T, L, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(fT + (4 * cg_rg + c) * T + 4 * q_rg) =
                    make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
        }
        __syncthreads();
        // ---- S4: pooling (two-pass mean / unbiased variance per latent column, :418-419) ----
        if (tid < L * 8) {
            const int c = tid >> 3, part = tid & 7;
            const float* fc = fT + c * T;
            float s = 0.f;
            for (int r = part; r < T; r += 8) s += fc[r];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            const float mean = __fdiv_rn(s, Tf);
            float m2 = 0.f;
            for (int r = part; r < T; r += 8) { const float dl = fc[r] - mean; m2 = fmaf(dl, dl, m2); }
            m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
            if (part == 0) {
                const float sd = sqrtf(__fdiv_rn(m2, Tm1));
                const float var = __fmul_rn(sd, sd);
                const float sim = sqrtf(__fdiv_rn(var, Tf));                                   // :422
                const float siv = sqrtf(__fdiv_rn(__fmul_rn(2.0f, __fmul_rn(var, var)), Tm1));   // :423
                const float e1 = sv[SM_E12 + c], e2 = sv[SM_E12 + L + c];
                const float mus = __fadd_rn(__fmul_rn(e1, sim), mean);                          // :426
                const float vs = __fadd_rn(__fmul_rn(e2, siv), var);                            // :427
                const float sds = sqrtf(__fadd_rn(fabsf(vs), 1e-5f));                           // :430
                sv[SM_M + c] = mean; sv[SM_VAR + c] = var; sv[SM_SIM + c] = sim; sv[SM_SIV + c] = siv; sv[SM_VS + c] = vs;
                sv[SM_S + c] = mus; sv[SM_S + L + c] = sds;
                // summary noise (:448-450) and the KL terms of the clean summary (:515-520)
                const float lv0 = sv[SM_LVS + c], lv1 = sv[SM_LVS + L + c];
                sv[SM_SP + c] = __fadd_rn(mus, __fmul_rn(sv[SM_ESN + c], sv[SM_ELVH + c]));
                sv[SM_SP + L + c] = __fadd_rn(sds, __fmul_rn(sv[SM_ESN + L + c], sv[SM_ELVH + L + c]));
                a_skl += 0.5f * (mus * mus + expf(lv0) - lv0 - 1.0f) + 0.5f * (sds * sds + expf(lv1) - lv1 - 1.0f);
            }
        }
        __syncthreads();
        // ---- S6: regress_nn forward (head weights through L2: 13 kB per seed, shared by every CTA of the seed) ----
        if (tid < H * 4) {
            const int j = tid >> 2, part = tid & 3;
            const float* w = th + fl.V0 + j * S2 + 10 * part;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) a = fmaf(sv[SM_SP + 10 * part + k], __ldg(w + k), a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[SM_R1 + j] = relu_nan(a + __ldg(th + fl.c0 + j));
        }
        __syncthreads();
        if (tid < H * 4) {
            const int j = tid >> 2, part = tid & 3;
            const float* w = th + fl.V1 + j * H + 10 * part;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) a = fmaf(sv[SM_R1 + 10 * part + k], __ldg(w + k), a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[SM_R2 + j] = relu_nan(a + __ldg(th + fl.c1 + j));
        }
        __syncthreads();
        if (tid < 32) {
            // o = lane >> 4 (two outputs), 16 lanes each: k = l16, l16+16, l16+32
            const int o = lane >> 4, l16 = lane & 15;
            float a = 0.f;
            for (int k = l16; k < H; k += 16) a = fmaf(sv[SM_R2 + k], __ldg(th + fl.V2 + o * H + k), a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            const float r0 = __shfl_sync(0xffffffffu, a, 0) + __ldg(th + fl.c2);
            const float r1 = __shfl_sync(0xffffffffu, a, 16) + __ldg(th + fl.c2 + 1);
            if (lane == 0) {
                const float t0 = tanhf(r0), t1 = tanhf(r1);
                const float mu = soft_clamp_dev(r0, prm.hc.lo_mu, prm.hc.hi_mu);
                const float sd = soft_clamp_dev(r1, prm.hc.lo_sd, prm.hc.hi_sd);
                float l0, l1, dm0, dm1, ds0, ds1;
                nll_terms(mu, sd, sv[SM_Y], l0, dm0, ds0);
                nll_terms(mu, sd, sv[SM_Y + 1], l1, dm1, ds1);
                a_nll += -(l0 + l1);
                const float gmu = -(dm0 + dm1), gsd = -(ds0 + ds1);
                sv[SM_GR] = gmu * 0.5f * (prm.hc.hi_mu - prm.hc.lo_mu) * (1.0f - t0 * t0);
                sv[SM_GR + 1] = gsd * 0.5f * (prm.hc.hi_sd - prm.hc.lo_sd) * (1.0f - t1 * t1);
            }
        }
        __syncthreads();
        // ---- S7: regress_nn backward ----
        const float gr0 = sv[SM_GR], gr1 = sv[SM_GR + 1];
        if (tid < 2 * H) {
            const int o = tid / H, k = tid - o * H;
            aV2 = fmaf(o ? gr1 : gr0, sv[SM_R2 + k], aV2);
            if (tid < H) {
                const float g = gr0 * __ldg(th + fl.V2 + tid) + gr1 * __ldg(th + fl.V2 + H + tid);
                sv[SM_G2 + tid] = sv[SM_R2 + tid] > 0.f ? g : 0.f;
            }
        }
        if (tid == 2 * H) ac2 += gr0;
        if (tid == 2 * H + 1) ac2 += gr1;
        __syncthreads();
        if (tid < H * 4) {
            const int k = tid >> 2, part = tid & 3;
            float a = 0.f;
#pragma unroll
            for (int jj = 0; jj < 10; ++jj) {
                const int j = 10 * part + jj;
                a = fmaf(sv[SM_G2 + j], __ldg(th + fl.V1 + j * H + k), a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[SM_G1 + k] = sv[SM_R1 + k] > 0.f ? a : 0.f;
        }
        if (tid < 200) {  // dV1[j][k] += g_a2[j] r1[k]
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    aV1[jj][kk] = fmaf(sv[SM_G2 + 2 * jb + jj], sv[SM_R1 + krow10[kk]], aV1[jj][kk]);
        }
        if (tid >= 200 && tid < 200 + H) ac1 += sv[SM_G2 + tid - 200];
        __syncthreads();
        if (tid < S2 * 4) {
            const int k = tid >> 2, part = tid & 3;
            float a = 0.f;
#pragma unroll
            for (int jj = 0; jj < 10; ++jj) {
                const int j = 10 * part + jj;
                a = fmaf(sv[SM_G1 + j], __ldg(th + fl.V0 + j * S2 + k), a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) {
                const float s = sv[SM_S + k], sp = sv[SM_SP + k];
                alvs = fmaf(a, 0.5f * (sv[SM_ESN + k] * sv[SM_ELVH + k]), alvs);  // ds'/dlv = eps e^{lv/2} / 2
                sv[SM_GS + k] = a + prm.beta_out * s;
                (void)sp;
            }
        }
        if (tid < 200) {  // dV0[j][k] += g_a1[j] s'[k]
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    aV0[jj][kk] = fmaf(sv[SM_G1 + 2 * jb + jj], sv[SM_SP + krow10[kk]], aV0[jj][kk]);
        }
        if (tid >= 200 && tid < 200 + H) ac0 += sv[SM_G1 + tid - 200];
        __syncthreads();
        if (tid < L) {
            const int c = tid;
            const float gmus = sv[SM_GS + c], gsds = sv[SM_GS + L + c];
            const float vs = sv[SM_VS + c], sds = sv[SM_S + L + c], var = sv[SM_VAR + c];
            const float sgn = vs > 0.f ? 1.0f : (vs < 0.f ? -1.0f : 0.f);
            const float gvs = gsds * sgn / (2.0f * sds);
            const float e1 = sv[SM_E12 + c], e2 = sv[SM_E12 + L + c];
            const float gv = gmus * e1 / (2.0f * Tf * sv[SM_SIM + c]) +
                             gvs * (1.0f + e2 * (2.0f * var) / (Tm1 * sv[SM_SIV + c]));
            sv[SM_GM + c] = gmus / Tf;              // coefficient of 1
            sv[SM_GV + c] = 2.0f * gv / Tm1;        // coefficient of (f - m)
        }
        __syncthreads();
        // ---- S8: g_f in place over f ----
        for (int i = tid; i < L * NQ; i += NTHR) {
            const int c = i / NQ, q = i - c * NQ;
            float4* p = reinterpret_cast<float4*>(fT + c * T + 4 * q);
            const float m = sv[SM_M + c], A = sv[SM_GM + c], Bc = sv[SM_GV + c];
            float4 f = *p;
            f.x = fmaf(Bc, f.x - m, A); f.y = fmaf(Bc, f.y - m, A); f.z = fmaf(Bc, f.z - m, A); f.w = fmaf(Bc, f.w - m, A);
            *p = f;
        }
        __syncthreads();
        // ---- S9: dW2 += g_f^T h2, db2 ----
        if (tid < 100) outer_acc(fT, h2T, T, 2 * jb, krow10, aW2);
        else if (tid < 100 + L) ab2 += row_sum(fT + (tid - 100) * T, T);
        __syncthreads();
        // ---- S10: g_a2 = (g_f W2) . [h2 > 0], in place over h2 ----
        if (cg_rg < 10) {
            float acc[4][4] = {};
            rowgemm4x4(fT, T, L, W2n, H, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4* p = reinterpret_cast<float4*>(h2T + (4 * cg_rg + c) * T + 4 * q_rg);
                const float4 h = *p;
                *p = make_float4(h.x > 0.f ? acc[0][c] : 0.f, h.y > 0.f ? acc[1][c] : 0.f, h.z > 0.f ? acc[2][c] : 0.f,
                                 h.w > 0.f ? acc[3][c] : 0.f);
            }
        }
        __syncthreads();
        // ---- S11: dW1 += g_a2^T h1, db1 ----
        if (tid < 200) outer_acc(h2T, h1T, T, 2 * jb, krow10, aW1);
        else if (tid < 200 + H) ab1 += row_sum(h2T + (tid - 200) * T, T);
        __syncthreads();
        // ---- S12: g_a1 = (g_a2 W1) . [h1 > 0], in place over h1 ----
        if (cg_rg < 10) {
            float acc[4][4] = {};
            rowgemm4x4(h2T, T, H, W1n, H, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4* p = reinterpret_cast<float4*>(h1T + (4 * cg_rg + c) * T + 4 * q_rg);
                const float4 h = *p;
                *p = make_float4(h.x > 0.f ? acc[0][c] : 0.f, h.y > 0.f ? acc[1][c] : 0.f, h.z > 0.f ? acc[2][c] : 0.f,
                                 h.w > 0.f ? acc[3][c] : 0.f);
            }
        }
        __syncthreads();
        // ---- S13: dW0 += g_a1^T x', db0;  S14: dlv_in += sum_r (g_a1 W0)[r][c] * noise[r][c] / 2 ----
        if (tid < 20 * KG0) outer_acc(h1T, xT, T, 2 * jb0, krow0, aW0);
        else if (tid >= 248) ab0 += row_sum(h1T + (tid - 248) * T, T);
        if (cg_rg < (FP >> 2)) {
            float acc[4][4] = {};
            rowgemm4x4(h1T, T, H, W0n, FP, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = min(4 * cg_rg + c, F - 1);
                const float4 nz = *reinterpret_cast<const float4*>(nT + col * T + 4 * q_rg);
                alvin[c] += (acc[0][c] * nz.x + acc[1][c] * nz.y) + (acc[2][c] * nz.z + acc[3][c] * nz.w);
            }
        }
        __syncthreads();
    }

    // ---- write this CTA's partial gradient (flatten() order) ----
    float* part = prm.partial + ((int64_t)sidx * prm.n_cta + blockIdx.x) * (fl.d + DPAD);
    // dlv_in: fixed-order sum over the row quads through shared memory (xT is free now)
    float* red = xT;  // [FP][NQ]
    if (cg_rg < (FP >> 2)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[(4 * cg_rg + c) * NQ + q_rg] = alvin[c];
    }
    __syncthreads();
    if (tid < F) {
        float s = 0.f;
        for (int q = 0; q < NQ; ++q) s += red[tid * NQ + q];
        part[fl.lv_in + tid] = 0.5f * s;
    }
    // lv_sum: owner = thread 4k (part == 0 of the k-th group of the g_s' GEMV)
    if (tid < S2 * 4 && (tid & 3) == 0) part[fl.lv_sum + (tid >> 2)] = alvs;
    if (tid < 20 * KG0) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int c = kb11 + KG0 * kk;
                if (c < F) part[fl.W0 + (2 * jb0 + jj) * F + c] = aW0[jj][kk];
            }
    }
    if (tid < 200) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                part[fl.W1 + (2 * jb + jj) * H + krow10[kk]] = aW1[jj][kk];
                part[fl.V0 + (2 * jb + jj) * S2 + krow10[kk]] = aV0[jj][kk];
                part[fl.V1 + (2 * jb + jj) * H + krow10[kk]] = aV1[jj][kk];
            }
    }
    if (tid < 100) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) part[fl.W2 + (2 * jb + jj) * H + krow10[kk]] = aW2[jj][kk];
    }
    if (tid >= 248) part[fl.b0 + tid - 248] = ab0;
    if (tid >= 200 && tid < 200 + H) {
        part[fl.b1 + tid - 200] = ab1;
        part[fl.c0 + tid - 200] = ac0;
        part[fl.c1 + tid - 200] = ac1;
    }
    if (tid >= 100 && tid < 100 + L) part[fl.b2 + tid - 100] = ab2;
    if (tid < 2 * H) part[fl.V2 + tid] = aV2;
    if (tid == 2 * H || tid == 2 * H + 1) part[fl.c2 + tid - 2 * H] = ac2;
    // metrics: nll (lane 0 of warp 0) and the summary KL terms (threads 8c of the pooling groups)
    float skl = (tid < L * 8 && (tid & 7) == 0) ? a_skl : 0.f;
    __syncthreads();
    red[tid] = skl;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int c = 0; c < L; ++c) s += red[8 * c];
        part[fl.d + SLOT_NLL] = a_nll;
        part[fl.d + SLOT_SKL] = s;
        for (int i = 2; i < DPAD; ++i) part[fl.d + i] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------
// v2: one CTA per SM, 576 threads, TWO systems per iteration (200 rows).  Differences from v1:
//   * activation gradients get their own buffers (g_a2, g_a1) instead of overwriting h2 / h1, so the three
//     weight-gradient outer products run in ONE phase at the end with every warp busy on its own matrix
//     (dW0: 220 threads, dW1: 200, dW2: 100, bias sums: 56) -- 9 block barriers per system instead of 17;
//   * the stored noise tile is gone: dlv_in uses x' - mask(x) (re-read through L2), which frees the room;
//   * row pitch 2T + 4: consecutive feature rows walk the 16-byte bank groups for the outer products;
//   * a thread owns one feature-matrix block only: 36 accumulator registers instead of 58.
// The two halves of the CTA (threads [0,288) / [288,576)) run the per-system pooling / head code for system 0 / 1.
// ---------------------------------------------------------------------------------------
constexpr int NTHR2 = 576, HALF2 = 288;

struct Smem2 {
    int RP, xT, h1T, h2T, fT, g2T, g1T, W0T, b0, W1T, b1, W2T, b2, W2n, W1n, W0n, small, total;
    __host__ __device__ Smem2(int T, int F, int FP) {
        RP = 2 * T + 4;
        int o = 0;
        xT = o; o += F * RP;
        h1T = o; o += H * RP;
        h2T = o; o += H * RP;
        fT = o; o += L * RP;
        g2T = o; o += H * RP;
        g1T = o; o += H * RP;
        W0T = o; o += F * H;
        b0 = o; o += H;
        W1T = o; o += H * H;
        b1 = o; o += H;
        W2T = o; o += H * L;
        b2 = o; o += L;
        W2n = o; o += L * H;
        W1n = o; o += H * H;
        W0n = o; o += H * FP;
        small = o; o += 2 * 768;
        total = o;
    }
};

template <int RP, int K, int NP>
__device__ __forceinline__ void rowgemm4x4_c(const float* __restrict__ AT, const float* __restrict__ W, int q, int cg,
                                             float (&acc)[4][4]) {
    const float* ap = AT + 4 * q;
    const float* wp = W + 4 * cg;
    u64 a2[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) { a2[r][0] = pack2(acc[r][0], acc[r][1]); a2[r][1] = pack2(acc[r][2], acc[r][3]); }
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(ap + k * RP);
        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(wp + k * NP);
        const u64 av[4] = {pack2(a.x, a.x), pack2(a.y, a.y), pack2(a.z, a.z), pack2(a.w, a.w)};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            a2[r][0] = fma2(av[r], w.x, a2[r][0]);
            a2[r][1] = fma2(av[r], w.y, a2[r][1]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) { unpack2(a2[r][0], acc[r][0], acc[r][1]); unpack2(a2[r][1], acc[r][2], acc[r][3]); }
}

// acc[jj][kk] += sum_r G[j0+jj][r] * Hm[krow[kk]][r], feature-major with row pitch RP, r < n_rows (multiple of 4).
// Packed fp32x2 accumulators over even / odd rows (row pairs come straight out of the LDS.128): 16 FFMA2 per
// 4-row step instead of 32 FFMA; the two halves are added when the step's 2T rows are done.
__device__ __forceinline__ void outer_acc_p(const float* __restrict__ G, const float* __restrict__ Hm, int RP, int n_rows,
                                            int j0, const int (&krow)[4], float (&acc)[2][4]) {
    const float* g0p = G + j0 * RP;
    const float* g1p = g0p + RP;
    u64 a2[2][4];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) a2[jj][kk] = pack2(acc[jj][kk], 0.f);
#pragma unroll 2
    for (int r = 0; r < n_rows; r += 4) {
        const ulonglong2 g0 = *reinterpret_cast<const ulonglong2*>(g0p + r);
        const ulonglong2 g1 = *reinterpret_cast<const ulonglong2*>(g1p + r);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const ulonglong2 h = *reinterpret_cast<const ulonglong2*>(Hm + krow[kk] * RP + r);
            a2[0][kk] = fma2(g0.x, h.x, a2[0][kk]);
            a2[0][kk] = fma2(g0.y, h.y, a2[0][kk]);
            a2[1][kk] = fma2(g1.x, h.x, a2[1][kk]);
            a2[1][kk] = fma2(g1.y, h.y, a2[1][kk]);
        }
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float lo, hi;
            unpack2(a2[jj][kk], lo, hi);
            acc[jj][kk] = lo + hi;
        }
}

// compiled for one (T, F): every pitch and loop bound is a constant, which removes the address arithmetic that made up
// a quarter of the executed instructions of the run-time-shaped version (ncu: IMAD + LEA + IADD3 = 22 %)
template <int T, int F>
__global__ void __launch_bounds__(NTHR2, 1) train_fwd_bwd2_kernel(const Params prm) {
    extern __shared__ __align__(16) float sm[];
    constexpr int FP = (F + 3) & ~3, NQ = T >> 2, NQ2 = 2 * NQ, RT = 2 * T, RP = 2 * T + 4;
    const Smem2 L_(T, F, FP);
    const FlatLayout fl(F);
    const int tid = threadIdx.x, lane = tid & 31;
    const int half = tid >= HALF2 ? 1 : 0, lt = tid - half * HALF2;  // system slot of the pair, thread within the half
    const int sidx = blockIdx.y;
    const float* th = prm.theta + (int64_t)sidx * fl.d;
    float* xT = sm + L_.xT; float* h1T = sm + L_.h1T; float* h2T = sm + L_.h2T; float* fT = sm + L_.fT;
    float* g2T = sm + L_.g2T; float* g1T = sm + L_.g1T;
    float* W0T = sm + L_.W0T; float* b0 = sm + L_.b0; float* W1T = sm + L_.W1T; float* b1 = sm + L_.b1;
    float* W2T = sm + L_.W2T; float* b2 = sm + L_.b2; float* W2n = sm + L_.W2n; float* W1n = sm + L_.W1n; float* W0n = sm + L_.W0n;
    float* sv = sm + L_.small + half * 768;   // this half's per-system scalars / vectors
    float* sv0 = sm + L_.small;               // shared constants live in slot 0 (lv_sum, exp(lv/2) vectors)

    // ---- stage this seed's feature weights (natural and transposed) and noise scales ----
    for (int i = tid; i < H * F; i += NTHR2) {
        const int j = i / F, c = i - j * F;
        const float w = __ldg(th + fl.W0 + i);
        W0T[c * H + j] = w;
        W0n[j * FP + c] = w;
    }
    for (int i = tid; i < H * (FP - F); i += NTHR2) W0n[(i / (FP - F)) * FP + F + i % (FP - F)] = 0.f;
    for (int i = tid; i < H * H; i += NTHR2) {
        const int j = i / H, k = i - j * H;
        const float w = __ldg(th + fl.W1 + i);
        W1T[k * H + j] = w;
        W1n[i] = w;
    }
    for (int i = tid; i < L * H; i += NTHR2) {
        const int j = i / H, k = i - j * H;
        const float w = __ldg(th + fl.W2 + i);
        W2T[k * L + j] = w;
        W2n[i] = w;
    }
    if (tid < H) { b0[tid] = __ldg(th + fl.b0 + tid); b1[tid] = __ldg(th + fl.b1 + tid); }
    if (tid < L) b2[tid] = __ldg(th + fl.b2 + tid);
    if (tid < S2) {
        const float lv = __ldg(th + fl.lv_sum + tid);
        sv0[SM_LVS + tid] = lv;
        sv0[SM_ELVH + tid] = expf(__fdiv_rn(lv, 2.0f));
    }
    if (tid < F) sv0[SM_NSC + tid] = expf(__fdiv_rn(__ldg(th + fl.lv_in + tid), 2.0f));
    // zero the 4 pad rows of every feature row once (the row GEMMs never touch them, the outer products skip them)
    for (int i = tid; i < (F + 4 * H + L) * 4; i += NTHR2) {
        const int f = i >> 2, r = RT + (i & 3);
        (sm + L_.xT)[f * RP + r] = 0.f;  // xT, h1T, h2T, fT, g2T, g1T are contiguous: f indexes all their rows
    }

    // ---- thread-owned gradient accumulators ----
    float aW[2][4] = {}, aV0[2][4] = {}, aV1[2][4] = {};   // aW: this thread's block of dW0 | dW1 | dW2
    float abias[2] = {0.f, 0.f}, aV2 = 0.f, ac0 = 0.f, ac1 = 0.f, ac2 = 0.f, alvs = 0.f;
    float alvin[4] = {0.f, 0.f, 0.f, 0.f};
    float a_nll = 0.f, a_skl = 0.f;
    constexpr int KG0 = FP >> 2, N0 = 20 * KG0;             // dW0 owners: tid < N0 (220 at F = 41)
    // outer-product role of this thread
    int role, jb_o, krow[4];
    if (tid < N0) {
        role = 0; jb_o = tid / KG0;
        for (int kk = 0; kk < 4; ++kk) krow[kk] = min(tid % KG0 + KG0 * kk, F - 1);
    } else if (tid < N0 + 200) {
        role = 1; jb_o = (tid - N0) / 10;
        for (int kk = 0; kk < 4; ++kk) krow[kk] = (tid - N0) % 10 + 10 * kk;
    } else if (tid < N0 + 300) {
        role = 2; jb_o = (tid - N0 - 200) / 10;
        for (int kk = 0; kk < 4; ++kk) krow[kk] = (tid - N0 - 200) % 10 + 10 * kk;
    } else {
        role = 3; jb_o = tid - N0 - 300;  // bias rows: jb_o and jb_o + n_bias_threads over the 100 rows b0 | b1 | b2
        for (int kk = 0; kk < 4; ++kk) krow[kk] = 0;
    }
    constexpr int n_bias_thr = NTHR2 - N0 - 300;
    const int jbh = lt / 10;                                // head blocks (dV0, dV1): lt < 200
    const int krow10[4] = {lt % 10, lt % 10 + 10, lt % 10 + 20, lt % 10 + 30};
    const int q_rg = tid % NQ2, cg_rg = tid / NQ2;          // row-GEMM item of this thread (q over both systems)
    const float Tf = (float)T, Tm1 = (float)(T - 1);
    const uint64_t key = seed_key(prm.seed, sidx);
    __syncthreads();

    for (int b0i = 2 * blockIdx.x; b0i < prm.B; b0i += 2 * gridDim.x) {
        const int b = b0i + half;
        const bool act = b < prm.B;
        const int64_t sb = (int64_t)sidx * prm.B + (act ? b : b0i);
        const int64_t row = prm.batch_index ? (int64_t)prm.batch_index[sb] : (int64_t)(act ? b : b0i);
        // ---- S0: x' = mask(x) + eps_in * exp(lv_in/2), feature-major, both systems (each half loads its own) ----
        {
            const float* xs = prm.X + row * (int64_t)T * F;
            const float* es = prm.eps_in ? prm.eps_in + sb * (int64_t)T * F : nullptr;
            float* xh = xT + half * T;
            if (!act) {
                for (int i = lt; i < T * F; i += HALF2) { const int t = i / F, c = i - t * F; xh[c * RP + t] = 0.f; }
            } else if (es) {
                for (int i = lt; i < T * F; i += HALF2) {
                    const int t = i / F, c = i - t * F;
                    float xv = __ldg(xs + i);
                    if ((prm.zero_mask >> c) & 1ull) xv = __fsub_rn(xv, xv);  // x - mask keeps NaN (:452-478)
                    xh[c * RP + t] = __fadd_rn(xv, __fmul_rn(__ldg(es + i), sv0[SM_NSC + c]));
                }
            } else {
                const int F4 = (F + 3) >> 2;
                for (int i = lt; i < T * F4; i += HALF2) {
                    const int t = i / F4, c4 = i - t * F4;
                    const float4 n4 = philox_normal4_fast(key, STREAM_EPS_IN, (uint32_t)b, (uint32_t)prm.step, (uint32_t)i);
                    const float e[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = 4 * c4 + u;
                        if (c < F) {
                            float xv = __ldg(xs + t * F + c);
                            if ((prm.zero_mask >> c) & 1ull) xv = __fsub_rn(xv, xv);
                            xh[c * RP + t] = __fadd_rn(xv, __fmul_rn(e[u], sv0[SM_NSC + c]));
                        }
                    }
                }
            }
            if (lt < S2 / 4) {
                float4 a, c;
                if (prm.eps12) {
                    a = __ldg(reinterpret_cast<const float4*>(prm.eps12 + sb * S2) + lt);
                    c = __ldg(reinterpret_cast<const float4*>(prm.eps_sum + sb * S2) + lt);
                } else {
                    a = philox_normal4(key, STREAM_EPS, (uint32_t)b, (uint32_t)prm.step, (uint32_t)lt);
                    c = philox_normal4(key, STREAM_EPS_SUM, (uint32_t)b, (uint32_t)prm.step, (uint32_t)lt);
                }
                *reinterpret_cast<float4*>(sv + SM_E12 + 4 * lt) = a;
                *reinterpret_cast<float4*>(sv + SM_ESN + 4 * lt) = c;
            }
            if (lt == 32) {
                sv[SM_Y] = __ldg(prm.Y + row * 2);
                sv[SM_Y + 1] = __ldg(prm.Y + row * 2 + 1);
            }
        }
        __syncthreads();
        // ---- S1..S3: feature_nn forward over the 2T rows ----
        if (cg_rg < 10) {
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float bv = b0[4 * cg_rg + c]; for (int r = 0; r < 4; ++r) acc[r][c] = bv; }
            rowgemm4x4_c<RP, F, H>(xT, W0T, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(h1T + (4 * cg_rg + c) * RP + 4 * q_rg) =
                    make_float4(relu_nan(acc[0][c]), relu_nan(acc[1][c]), relu_nan(acc[2][c]), relu_nan(acc[3][c]));
        }
        __syncthreads();
        if (cg_rg < 10) {
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float bv = b1[4 * cg_rg + c]; for (int r = 0; r < 4; ++r) acc[r][c] = bv; }
            rowgemm4x4_c<RP, H, H>(h1T, W1T, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(h2T + (4 * cg_rg + c) * RP + 4 * q_rg) =
                    make_float4(relu_nan(acc[0][c]), relu_nan(acc[1][c]), relu_nan(acc[2][c]), relu_nan(acc[3][c]));
        }
        __syncthreads();
        if (cg_rg < 5) {
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float bv = b2[4 * cg_rg + c]; for (int r = 0; r < 4; ++r) acc[r][c] = bv; }
            rowgemm4x4_c<RP, H, L>(h2T, W2T, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(fT + (4 * cg_rg + c) * RP + 4 * q_rg) =
                    make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
        }
        __syncthreads();
        // ---- S4: pooling per system (two-pass mean / unbiased variance per latent column, :418-419) ----
        if (lt < L * 8) {
            const int c = lt >> 3, part = lt & 7;
            const float* fc = fT + c * RP + half * T;
            float s = 0.f;
            for (int r = part; r < T; r += 8) s += fc[r];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            const float mean = __fdiv_rn(s, Tf);
            float m2 = 0.f;
            for (int r = part; r < T; r += 8) { const float dl = fc[r] - mean; m2 = fmaf(dl, dl, m2); }
            m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
            if (part == 0) {
                const float sd = sqrtf(__fdiv_rn(m2, Tm1));
                const float var = __fmul_rn(sd, sd);
                const float sim = sqrtf(__fdiv_rn(var, Tf));                                   // :422
                const float siv = sqrtf(__fdiv_rn(__fmul_rn(2.0f, __fmul_rn(var, var)), Tm1));   // :423
                const float e1 = sv[SM_E12 + c], e2 = sv[SM_E12 + L + c];
                const float mus = __fadd_rn(__fmul_rn(e1, sim), mean);                          // :426
                const float vs = __fadd_rn(__fmul_rn(e2, siv), var);                            // :427
                const float sds = sqrtf(__fadd_rn(fabsf(vs), 1e-5f));                           // :430
                sv[SM_M + c] = mean; sv[SM_VAR + c] = var; sv[SM_SIM + c] = sim; sv[SM_SIV + c] = siv; sv[SM_VS + c] = vs;
                sv[SM_S + c] = mus; sv[SM_S + L + c] = sds;
                const float lv0 = sv0[SM_LVS + c], lv1 = sv0[SM_LVS + L + c];
                sv[SM_SP + c] = __fadd_rn(mus, __fmul_rn(sv[SM_ESN + c], sv0[SM_ELVH + c]));
                sv[SM_SP + L + c] = __fadd_rn(sds, __fmul_rn(sv[SM_ESN + L + c], sv0[SM_ELVH + L + c]));
                if (act) a_skl += 0.5f * (mus * mus + expf(lv0) - lv0 - 1.0f) + 0.5f * (sds * sds + expf(lv1) - lv1 - 1.0f);
            }
        }
        __syncthreads();
        // ---- S6: regress_nn forward (head weights through L2) ----
        if (lt < H * 4) {
            const int j = lt >> 2, part = lt & 3;
            const float* w = th + fl.V0 + j * S2 + 10 * part;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) a = fmaf(sv[SM_SP + 10 * part + k], __ldg(w + k), a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[SM_R1 + j] = relu_nan(a + __ldg(th + fl.c0 + j));
        }
        __syncthreads();
        if (lt < H * 4) {
            const int j = lt >> 2, part = lt & 3;
            const float* w = th + fl.V1 + j * H + 10 * part;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) a = fmaf(sv[SM_R1 + 10 * part + k], __ldg(w + k), a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[SM_R2 + j] = relu_nan(a + __ldg(th + fl.c1 + j));
        }
        __syncthreads();
        if (lt < 32) {
            const int o = lane >> 4, l16 = lane & 15;
            float a = 0.f;
            for (int k = l16; k < H; k += 16) a = fmaf(sv[SM_R2 + k], __ldg(th + fl.V2 + o * H + k), a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            const float r0 = __shfl_sync(0xffffffffu, a, 0) + __ldg(th + fl.c2);
            const float r1 = __shfl_sync(0xffffffffu, a, 16) + __ldg(th + fl.c2 + 1);
            if (lane == 0) {
                const float t0 = tanhf(r0), t1 = tanhf(r1);
                const float mu = soft_clamp_dev(r0, prm.hc.lo_mu, prm.hc.hi_mu);
                const float sd = soft_clamp_dev(r1, prm.hc.lo_sd, prm.hc.hi_sd);
                float l0, l1, dm0, dm1, ds0, ds1;
                nll_terms(mu, sd, sv[SM_Y], l0, dm0, ds0);
                nll_terms(mu, sd, sv[SM_Y + 1], l1, dm1, ds1);
                if (act) a_nll += -(l0 + l1);
                const float gmu = -(dm0 + dm1), gsd = -(ds0 + ds1);
                // an inactive slot (odd batch tail) contributes nothing: zero upstream gradient
                sv[SM_GR] = act ? gmu * 0.5f * (prm.hc.hi_mu - prm.hc.lo_mu) * (1.0f - t0 * t0) : 0.f;
                sv[SM_GR + 1] = act ? gsd * 0.5f * (prm.hc.hi_sd - prm.hc.lo_sd) * (1.0f - t1 * t1) : 0.f;
            }
        }
        __syncthreads();
        // ---- S7: regress_nn backward ----
        const float gr0 = sv[SM_GR], gr1 = sv[SM_GR + 1];
        if (lt < 2 * H) {
            const int o = lt / H, k = lt - o * H;
            aV2 = fmaf(o ? gr1 : gr0, sv[SM_R2 + k], aV2);
            if (lt < H) {
                const float g = gr0 * __ldg(th + fl.V2 + lt) + gr1 * __ldg(th + fl.V2 + H + lt);
                sv[SM_G2 + lt] = sv[SM_R2 + lt] > 0.f ? g : 0.f;
            }
        }
        if (lt == 2 * H) ac2 += gr0;
        if (lt == 2 * H + 1) ac2 += gr1;
        __syncthreads();
        if (lt < H * 4) {
            const int k = lt >> 2, part = lt & 3;
            float a = 0.f;
#pragma unroll
            for (int jj = 0; jj < 10; ++jj) {
                const int j = 10 * part + jj;
                a = fmaf(sv[SM_G2 + j], __ldg(th + fl.V1 + j * H + k), a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[SM_G1 + k] = sv[SM_R1 + k] > 0.f ? a : 0.f;
        }
        if (lt < 200) {  // dV1[j][k] += g_a2[j] r1[k]
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    aV1[jj][kk] = fmaf(sv[SM_G2 + 2 * jbh + jj], sv[SM_R1 + krow10[kk]], aV1[jj][kk]);
        }
        if (lt >= 200 && lt < 200 + H) ac1 += sv[SM_G2 + lt - 200];
        __syncthreads();
        if (lt < S2 * 4) {
            const int k = lt >> 2, part = lt & 3;
            float a = 0.f;
#pragma unroll
            for (int jj = 0; jj < 10; ++jj) {
                const int j = 10 * part + jj;
                a = fmaf(sv[SM_G1 + j], __ldg(th + fl.V0 + j * S2 + k), a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) {
                alvs = fmaf(a, 0.5f * (sv[SM_ESN + k] * sv0[SM_ELVH + k]), alvs);  // ds'/dlv = eps e^{lv/2} / 2
                sv[SM_GS + k] = a + (act ? prm.beta_out * sv[SM_S + k] : 0.f);
            }
        }
        if (lt < 200) {  // dV0[j][k] += g_a1[j] s'[k]
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    aV0[jj][kk] = fmaf(sv[SM_G1 + 2 * jbh + jj], sv[SM_SP + krow10[kk]], aV0[jj][kk]);
        }
        if (lt >= 200 && lt < 200 + H) ac0 += sv[SM_G1 + lt - 200];
        __syncthreads();
        if (lt < L) {
            const int c = lt;
            const float gmus = sv[SM_GS + c], gsds = sv[SM_GS + L + c];
            const float vs = sv[SM_VS + c], sds = sv[SM_S + L + c], var = sv[SM_VAR + c];
            const float sgn = vs > 0.f ? 1.0f : (vs < 0.f ? -1.0f : 0.f);
            const float gvs = gsds * sgn / (2.0f * sds);
            const float e1 = sv[SM_E12 + c], e2 = sv[SM_E12 + L + c];
            const float gv = gmus * e1 / (2.0f * Tf * sv[SM_SIM + c]) +
                             gvs * (1.0f + e2 * (2.0f * var) / (Tm1 * sv[SM_SIV + c]));
            sv[SM_GM + c] = act ? gmus / Tf : 0.f;        // coefficient of 1
            sv[SM_GV + c] = act ? 2.0f * gv / Tm1 : 0.f;  // coefficient of (f - m)
        }
        __syncthreads();
        // ---- S8: g_f in place over f (each half its own system) ----
        for (int i = lt; i < L * NQ; i += HALF2) {
            const int c = i / NQ, q = i - c * NQ;
            float4* p = reinterpret_cast<float4*>(fT + c * RP + half * T + 4 * q);
            const float m = sv[SM_M + c], A = sv[SM_GM + c], Bc = sv[SM_GV + c];
            float4 f = *p;
            f.x = fmaf(Bc, f.x - m, A); f.y = fmaf(Bc, f.y - m, A); f.z = fmaf(Bc, f.z - m, A); f.w = fmaf(Bc, f.w - m, A);
            *p = f;
        }
        __syncthreads();
        // ---- S10: g_a2 = (g_f W2) . [h2 > 0] ----
        if (cg_rg < 10) {
            float acc[4][4] = {};
            rowgemm4x4_c<RP, L, H>(fT, W2n, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 h = *reinterpret_cast<const float4*>(h2T + (4 * cg_rg + c) * RP + 4 * q_rg);
                *reinterpret_cast<float4*>(g2T + (4 * cg_rg + c) * RP + 4 * q_rg) =
                    make_float4(h.x > 0.f ? acc[0][c] : 0.f, h.y > 0.f ? acc[1][c] : 0.f, h.z > 0.f ? acc[2][c] : 0.f,
                                h.w > 0.f ? acc[3][c] : 0.f);
            }
        }
        __syncthreads();
        // ---- S12: g_a1 = (g_a2 W1) . [h1 > 0] ----
        if (cg_rg < 10) {
            float acc[4][4] = {};
            rowgemm4x4_c<RP, H, H>(g2T, W1n, q_rg, cg_rg, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 h = *reinterpret_cast<const float4*>(h1T + (4 * cg_rg + c) * RP + 4 * q_rg);
                *reinterpret_cast<float4*>(g1T + (4 * cg_rg + c) * RP + 4 * q_rg) =
                    make_float4(h.x > 0.f ? acc[0][c] : 0.f, h.y > 0.f ? acc[1][c] : 0.f, h.z > 0.f ? acc[2][c] : 0.f,
                                h.w > 0.f ? acc[3][c] : 0.f);
            }
        }
        __syncthreads();
        // ---- S13: all weight-gradient outer products in one phase; S14: dlv_in += sum (g_a1 W0) . (x' - mask(x)) / 2 ----
        if (role == 0) outer_acc_p(g1T, xT, RP, RT, 2 * jb_o, krow, aW);
        else if (role == 1) outer_acc_p(g2T, h1T, RP, RT, 2 * jb_o, krow, aW);
        else if (role == 2) outer_acc_p(fT, h2T, RP, RT, 2 * jb_o, krow, aW);
        else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = jb_o + e * n_bias_thr;  // bias row: b0 [0,40) | b1 [40,80) | b2 [80,100)
                if (r < 2 * H + L) {
                    const float* g = r < H ? g1T + r * RP : (r < 2 * H ? g2T + (r - H) * RP : fT + (r - 2 * H) * RP);
                    abias[e] += row_sum(g, RT);
                }
            }
        }
        if (cg_rg < (FP >> 2)) {
            float acc[4][4] = {};
            rowgemm4x4_c<RP, H, FP>(g1T, W0n, q_rg, cg_rg, acc);
            // rows 4 q_rg .. +3 belong to system slot q_rg / NQ; its data row and noise-free input come back through L2
            const int hs = q_rg / NQ, t0r = 4 * (q_rg - hs * NQ);
            const int bb = b0i + hs;
            if (bb < prm.B) {
                const int64_t sbb = (int64_t)sidx * prm.B + bb;
                const int64_t rowb = prm.batch_index ? (int64_t)prm.batch_index[sbb] : (int64_t)bb;
                const float* xs = prm.X + rowb * (int64_t)T * F;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int col = 4 * cg_rg + c;
                    if (col < F) {
                        const bool zeroed = (prm.zero_mask >> col) & 1ull;
                        float s = 0.f;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            float x0 = __ldg(xs + (t0r + r) * F + col);
                            if (zeroed) x0 = __fsub_rn(x0, x0);
                            s = fmaf(acc[r][c], xT[col * RP + 4 * q_rg + r] - x0, s);
                        }
                        alvin[c] += s;
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- write this CTA's partial gradient (flatten() order) ----
    float* part = prm.partial + ((int64_t)sidx * prm.n_cta + blockIdx.x) * (fl.d + DPAD);
    float* red = xT;  // scratch: the activations are dead now
    // dlv_in: fixed-order sum over the row quads of both systems
    if (cg_rg < (FP >> 2)) {
#pragma unroll
        for (int c = 0; c < 4; ++c) red[(4 * cg_rg + c) * NQ2 + q_rg] = alvin[c];
    }
    __syncthreads();
    if (tid < F) {
        float s = 0.f;
        for (int q = 0; q < NQ2; ++q) s += red[tid * NQ2 + q];
        part[fl.lv_in + tid] = 0.5f * s;
    }
    __syncthreads();
    // head accumulators of the upper half are added to the lower half's through shared memory
    // layout per lt: [0..7] aV0, [8..15] aV1, 16 aV2, 17 ac0, 18 ac1, 19 ac2, 20 alvs, 21 a_nll, 22 a_skl
    float* hx = red + HALF2 * 0;
    if (half == 1) {
        float* o = hx + lt * 24;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { o[jj * 4 + kk] = aV0[jj][kk]; o[8 + jj * 4 + kk] = aV1[jj][kk]; }
        o[16] = aV2; o[17] = ac0; o[18] = ac1; o[19] = ac2; o[20] = alvs; o[21] = a_nll; o[22] = a_skl;
    }
    __syncthreads();
    if (half == 0) {
        const float* o = hx + lt * 24;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { aV0[jj][kk] += o[jj * 4 + kk]; aV1[jj][kk] += o[8 + jj * 4 + kk]; }
        aV2 += o[16]; ac0 += o[17]; ac1 += o[18]; ac2 += o[19]; alvs += o[20]; a_nll += o[21]; a_skl += o[22];
        if (lt < S2 * 4 && (lt & 3) == 0) part[fl.lv_sum + (lt >> 2)] = alvs;
        if (lt < 200) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    part[fl.V0 + (2 * jbh + jj) * S2 + krow10[kk]] = aV0[jj][kk];
                    part[fl.V1 + (2 * jbh + jj) * H + krow10[kk]] = aV1[jj][kk];
                }
        }
        if (lt >= 200 && lt < 200 + H) {
            part[fl.c0 + lt - 200] = ac0;
            part[fl.c1 + lt - 200] = ac1;
        }
        if (lt < 2 * H) part[fl.V2 + lt] = aV2;
        if (lt == 2 * H || lt == 2 * H + 1) part[fl.c2 + lt - 2 * H] = ac2;
    }
    // feature-matrix blocks and biases
    if (role == 0) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int c = tid % KG0 + KG0 * kk;
                if (c < F) part[fl.W0 + (2 * jb_o + jj) * F + c] = aW[jj][kk];
            }
    } else if (role == 1) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) part[fl.W1 + (2 * jb_o + jj) * H + krow[kk]] = aW[jj][kk];
    } else if (role == 2) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) part[fl.W2 + (2 * jb_o + jj) * H + krow[kk]] = aW[jj][kk];
    } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int r = jb_o + e * n_bias_thr;
            if (r < H) part[fl.b0 + r] = abias[e];
            else if (r < 2 * H) part[fl.b1 + r - H] = abias[e];
            else if (r < 2 * H + L) part[fl.b2 + r - 2 * H] = abias[e];
        }
    }
    // metrics: nll (lane 0 of warp 0 of each half, already merged) and the summary KL terms (threads 8c of the pooling groups)
    __syncthreads();
    red[tid] = (half == 0 && lt < L * 8 && (lt & 7) == 0) ? a_skl : 0.f;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int c = 0; c < L; ++c) s += red[8 * c];
        part[fl.d + SLOT_NLL] = a_nll;
        part[fl.d + SLOT_SKL] = s;
        for (int i = 2; i < DPAD; ++i) part[fl.d + i] = 0.f;
    }
}

}  // namespace train
}  // namespace bnn
#include "train_v3.cuh"
#include "train_v4.cuh"
namespace bnn {
namespace train {

// grad[s][i] = sum_c partial[s][c][i] (fixed order) + analytic KL terms; per-block sum of squares.
__global__ void __launch_bounds__(256) train_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ theta,
                                                           int n_cta, int d, int F, float kl_in_scale, float kl_sum_scale,
                                                           float* __restrict__ grad, float* __restrict__ sq) {
    __shared__ float sh[256];
    const int s = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    const int DP = d + DPAD;
    float g = 0.f;
    if (i < DP) {
        const float* p = partial + (int64_t)s * n_cta * DP + i;
        for (int c = 0; c < n_cta; ++c) g += p[(int64_t)c * DP];
        if (i < F) {  // input_kl * beta_in * B (:585-590, :726): d/dlv = (e^lv - 1)/2
            g += kl_in_scale * 0.5f * (expf(theta[(int64_t)s * d + i]) - 1.0f);
        } else if (i < F + S2) {  // summary_kl * beta_out, summed over the batch (:515-520, :727)
            g += kl_sum_scale * 0.5f * (expf(theta[(int64_t)s * d + i]) - 1.0f);
        }
        grad[(int64_t)s * DP + i] = g;
    }
    sh[threadIdx.x] = (i < d) ? g * g : 0.f;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) sq[s * gridDim.x + blockIdx.x] = sh[0];
}

// clip_grad_norm_ (coef = clip / (norm + 1e-6), applied when < 1) + torch.optim.SGD with momentum / weight decay.
__global__ void __launch_bounds__(256) train_update_kernel(const float* __restrict__ grad, const float* __restrict__ sq,
                                                           int n_blocks, int d, int F, int B, bnn_train_hparams hp,
                                                           float* __restrict__ theta, float* __restrict__ mom,
                                                           float* __restrict__ grad_out, float* __restrict__ metrics) {
    const int s = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    const int DP = d + DPAD;
    float ss = 0.f;
    for (int b = 0; b < n_blocks; ++b) ss += sq[s * n_blocks + b];
    const float norm = sqrtf(ss);
    const float coef = hp.clip_norm / (norm + 1e-6f);
    const float* g = grad + (int64_t)s * DP;
    if (i < d) {
        const float gi = g[i];
        if (grad_out) grad_out[(int64_t)s * d + i] = gi;
        if (hp.apply_update) {
            const int64_t o = (int64_t)s * d + i;
            const float th = theta[o];
            float dp = coef < 1.0f ? gi * coef : gi;
            if (hp.weight_decay != 0.f) dp = fmaf(hp.weight_decay, th, dp);
            const float buf = hp.first_step ? dp : fmaf(hp.momentum, mom[o], dp);
            mom[o] = buf;
            theta[o] = th - hp.lr * buf;
        }
    }
    if (metrics && blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) {
        float* m = metrics + s * 8;
        const float nll = g[d + SLOT_NLL], skl = g[d + SLOT_SKL] * hp.beta_out;
        m[0] = nll / (float)B;
        m[3] = skl / (float)B;
        m[4] = norm;
        m[5] = coef < 1.0f ? coef : 1.0f;
        m[6] = isfinite(nll + skl + norm) ? 0.f : 1.f;
        m[7] = 0.f;
        // m[1] (loss with reg) and m[2] (input_kl) are completed by train_metrics_kernel, which reads lv_in before the update
    }
}

// input_kl * beta_in * B / B and the total, from the weights the step was evaluated at (runs BEFORE the update kernel)
__global__ void train_metrics_kernel(const float* __restrict__ theta, const float* __restrict__ grad, int d, int F, int B,
                                     bnn_train_hparams hp, float* __restrict__ metrics) {
    const int s = blockIdx.x;
    if (threadIdx.x != 0) return;
    float kl = 0.f;
    for (int c = 0; c < F; ++c) {
        const float lv = theta[(int64_t)s * d + c];
        kl += expf(lv) - lv - 1.0f;
    }
    const float ikl = 0.5f * kl * hp.beta_in * (float)B;
    const float* g = grad + (int64_t)s * (d + DPAD);
    const float nll = g[d + SLOT_NLL], skl = g[d + SLOT_SKL] * hp.beta_out;
    metrics[s * 8 + 1] = (nll + (ikl + skl)) / (float)B;
    metrics[s * 8 + 2] = ikl / (float)B;
}

// the Philox draws of one training step, written out (parity tests feed them to the oracle)
__global__ void train_noise_kernel(int B, int T, int F, uint64_t seed, uint64_t step, float* __restrict__ eps_in,
                                   float* __restrict__ eps12, float* __restrict__ eps_sum) {
    const int sidx = blockIdx.y, b = blockIdx.x;
    const uint64_t key = seed_key(seed, sidx);
    const int64_t sb = (int64_t)sidx * B + b;
    const int F4 = (F + 3) >> 2;
    for (int i = threadIdx.x; i < T * F4; i += blockDim.x) {
        const int t = i / F4, c4 = i - t * F4;
        const float4 n4 = philox_normal4_fast(key, STREAM_EPS_IN, (uint32_t)b, (uint32_t)step, (uint32_t)i);
        const float e[4] = {n4.x, n4.y, n4.z, n4.w};
        for (int u = 0; u < 4; ++u)
            if (4 * c4 + u < F) eps_in[(sb * T + t) * F + 4 * c4 + u] = e[u];
    }
    if (threadIdx.x < S2 / 4) {
        reinterpret_cast<float4*>(eps12 + sb * S2)[threadIdx.x] =
            philox_normal4(key, STREAM_EPS, (uint32_t)b, (uint32_t)step, threadIdx.x);
        reinterpret_cast<float4*>(eps_sum + sb * S2)[threadIdx.x] =
            philox_normal4(key, STREAM_EPS_SUM, (uint32_t)b, (uint32_t)step, threadIdx.x);
    }
}

// loss_sum[u] = sum_b _lossfnc(out[u][b], y[b]) -- one block per unit, fixed order
__global__ void __launch_bounds__(256) nll_sum_kernel(const float2* __restrict__ out, const float2* __restrict__ y, int64_t B,
                                                      float* __restrict__ loss_sum) {
    __shared__ float sh[256];
    const float2* o = out + (int64_t)blockIdx.x * B;
    float acc = 0.f;
    for (int64_t b = threadIdx.x; b < B; b += 256) {
        float l0, l1, a, c;
        const float2 p = o[b], yy = y[b];
        nll_terms(p.x, p.y, yy.x, l0, a, c);
        nll_terms(p.x, p.y, yy.y, l1, a, c);
        acc += -(l0 + l1);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_sum[blockIdx.x] = sh[0];
}

// sumsq[s][c] = sum over the CTAs' partial sums of (d mu / d x)^2 in column c (the v3 kernel parks them, halved, in the
// dlv_in slot of its partial), fixed order
__global__ void saliency_finish_kernel(const float* __restrict__ partial, int n_cta, int d, int F, float* __restrict__ sumsq) {
    const int s = blockIdx.x, c = threadIdx.x;
    if (c >= F) return;
    const float* p = partial + (int64_t)s * n_cta * (d + DPAD) + c;
    float a = 0.f;
    for (int i = 0; i < n_cta; ++i) a += p[(int64_t)i * (d + DPAD)];
    sumsq[s * F + c] = 2.0f * a;
}

// v4 (two CTAs per SM, one system per iteration): BNN_TRAIN_VARIANT=v4
static bool use_v4(const bnn_model_config* cfg) {
    const char* force = getenv("BNN_TRAIN_VARIANT");
    return force && !strcmp(force, "v4") && cfg->n_times == 100 && cfg->n_features == 41;
}

// v3 (large register tiles) for the reference's shape; BNN_TRAIN_VARIANT=v1|v2|v3|v4 forces one
static bool use_v3(const bnn_model_config* cfg) {
    const char* force = getenv("BNN_TRAIN_VARIANT");
    if (force && (!strcmp(force, "v1") || !strcmp(force, "v2") || !strcmp(force, "v4"))) return false;
    return cfg->n_times == 100 && cfg->n_features == 41;
}

// v2 (two systems per iteration, one CTA per SM) when its tile fits
static bool use_v2(const bnn_model_config* cfg) {
    const char* force = getenv("BNN_TRAIN_VARIANT");
    if (force && !strcmp(force, "v1")) return false;
    const int F = cfg->n_features, T = cfg->n_times, FP = (F + 3) & ~3;
    const bool fits = (size_t)Smem2(T, F, FP).total * sizeof(float) <= 227 * 1024 && (2 * (T / 4)) * (FP / 4) <= NTHR2 &&
                      20 * (FP / 4) + 300 < NTHR2 && NTHR2 - 20 * (FP / 4) - 300 >= 50;
    return fits && T == 100 && F == 41;  // the shape the kernel is compiled for (the reference's only one)
}

static int pick_n_cta(const bnn_model_config* cfg, int64_t B, int n_seeds) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (use_v4(cfg)) {   // two resident CTAs per SM, one system per iteration
        int64_t n4 = 2ll * sms / n_seeds;
        if (n4 > B) n4 = B;
        return (int)(n4 < 1 ? 1 : n4);
    }
    const bool v2 = use_v2(cfg) || use_v3(cfg);
    // v1: two resident CTAs per SM, one system per iteration; v2: one CTA per SM, two systems per iteration
    // one-CTA-per-SM kernels: never more CTAs than SMs (a 149th CTA would run alone in a second wave)
    int64_t n = v2 ? sms / n_seeds : (2ll * sms + n_seeds - 1) / n_seeds;
    const int64_t cap = v2 ? (B + 1) / 2 : B;
    if (n > cap) n = cap;
    if (n < 1) n = 1;
    return (int)n;
}

}  // namespace train
}  // namespace bnn

extern "C" {

size_t bnn_train_workspace_bytes(const bnn_model_config* cfg, int64_t B, int32_t n_seeds) {
    using namespace bnn;
    if (validate_config(cfg) != BNN_OK || B <= 0 || n_seeds <= 0) return 0;
    const int d = FlatLayout(cfg->n_features).d;
    const int n_cta = train::pick_n_cta(cfg, B, n_seeds);
    const size_t DP = d + train::DPAD;
    const size_t nb = (DP + 255) / 256;
    // partial | grad | sq | pad | head records (v3)
    return ((size_t)n_seeds * n_cta * DP + (size_t)n_seeds * DP + (size_t)n_seeds * nb + 64 +
            (size_t)n_seeds * B * train::REC + 4 +
            (size_t)n_seeds * n_cta * (8 * cfg->n_features * cfg->n_times + 2 * train::XSM)) * sizeof(float);
}

int bnn_train_step(const bnn_model_config* cfg, const bnn_train_hparams* hp, int32_t n_seeds, float* d_theta,
                   float* d_momentum, const float* d_x, const float* d_y, const int32_t* d_batch_index, int64_t B,
                   const float* d_eps_in, const float* d_eps12, const float* d_eps_sum, uint64_t seed, uint64_t step,
                   float* d_grad_out, float* d_metrics, void* d_workspace, void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(hp && d_theta && d_x && d_y && d_workspace, BNN_E_ARG, "bnn_train_step: null pointer");
    BNN_REQUIRE(n_seeds >= 1 && n_seeds <= 65535 && B >= 1 && B < (1ll << 30), BNN_E_ARG,
                "bnn_train_step: n_seeds=%d B=%lld out of range", n_seeds, (long long)B);
    BNN_REQUIRE(!hp->apply_update || d_momentum, BNN_E_ARG, "bnn_train_step: momentum buffer required to update");
    const bool any = d_eps_in || d_eps12 || d_eps_sum, all = d_eps_in && d_eps12 && d_eps_sum;
    BNN_REQUIRE(!any || all, BNN_E_ARG, "bnn_train_step: give all of eps_in, eps12, eps_sum or none");
    BNN_REQUIRE(aligned16(d_workspace) && (!d_eps12 || (aligned16(d_eps12) && aligned16(d_eps_sum))), BNN_E_ALIGN,
                "bnn_train_step: workspace / eps12 / eps_sum must be 16-byte aligned");
    const int F = cfg->n_features, T = cfg->n_times, FP = (F + 3) & ~3;
    BNN_REQUIRE((T / 4) * (FP / 4) <= train::NTHR && FP <= 48, BNN_E_CONFIG,
                "bnn_train_step: T/4 * ceil(F/4) = %d exceeds the %d-thread tile, or F > 48", (T / 4) * (FP / 4),
                train::NTHR);
    const FlatLayout fl(F);
    const train::Smem sl(T, F, FP);
    const size_t smem = (size_t)sl.total * sizeof(float);
    BNN_REQUIRE(smem <= 227 * 1024, BNN_E_CONFIG, "bnn_train_step: tile needs %zu bytes of shared memory", smem);
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        BNN_CUDA(cudaFuncSetAttribute(train::train_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n_cta = train::pick_n_cta(cfg, B, n_seeds);
    const int DP = fl.d + train::DPAD, nb = (DP + 255) / 256;
    float* partial = (float*)d_workspace;
    float* grad = partial + (size_t)n_seeds * n_cta * DP;
    float* sq = grad + (size_t)n_seeds * DP;
    float* head_rec = sq + (size_t)n_seeds * nb;
    head_rec += (4 - ((uintptr_t)head_rec / sizeof(float)) % 4) % 4;  // 16-byte aligned records

    train::Params prm;
    prm.theta = d_theta; prm.X = d_x; prm.Y = d_y; prm.batch_index = d_batch_index;
    prm.eps_in = d_eps_in; prm.eps12 = d_eps12; prm.eps_sum = d_eps_sum;
    prm.partial = partial;
    prm.head_rec = head_rec;
    prm.xprod = head_rec + (size_t)n_seeds * B * train::REC;   // REC is a multiple of 4: stays 16-byte aligned
    prm.B = (int)B; prm.T = T; prm.F = F; prm.FP = FP; prm.n_cta = n_cta;
    prm.seed = seed; prm.step = step; prm.zero_mask = cfg->zero_mask;
    prm.hc = HeadConsts{cfg->lo_mu, cfg->hi_mu, cfg->lo_sd, cfg->hi_sd};
    prm.beta_out = hp->beta_out;
    prm.saliency = 0; prm.gx_out = nullptr; prm.mu_out = nullptr;
    if (train::use_v4(cfg)) {
        const size_t smem4 = (size_t)train::Smem4(T, F).total * sizeof(float);
        static PerDeviceOnce attr4_done;
        if (attr4_done.need()) {
            BNN_CUDA(cudaFuncSetAttribute(train::train_fwd_bwd4_kernel<100, 41>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem4));
            BNN_CUDA(cudaFuncSetAttribute(train::train_fwd_bwd4_kernel<100, 41>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        }
        train::train_fwd_bwd4_kernel<100, 41><<<dim3(n_cta, n_seeds), train::NTHR4, smem4, st>>>(prm);
    } else if (train::use_v3(cfg)) {
        const size_t smem3 = (size_t)train::Smem3(T, F).total * sizeof(float);
        static PerDeviceOnce attr3_done;
        if (attr3_done.need()) {
            BNN_CUDA(cudaFuncSetAttribute(train::train_fwd_bwd3_kernel<100, 41>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
        }
        train::train_fwd_bwd3_kernel<100, 41><<<dim3(n_cta, n_seeds), train::NTHR3, smem3, st>>>(prm);
    } else if (train::use_v2(cfg)) {
        const size_t smem2 = (size_t)train::Smem2(T, F, FP).total * sizeof(float);
        static PerDeviceOnce attr2_done;
        if (attr2_done.need()) {
            BNN_CUDA(cudaFuncSetAttribute(train::train_fwd_bwd2_kernel<100, 41>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          227 * 1024));
        }
        train::train_fwd_bwd2_kernel<100, 41><<<dim3(n_cta, n_seeds), train::NTHR2, smem2, st>>>(prm);
    } else {
        train::train_fwd_bwd_kernel<<<dim3(n_cta, n_seeds), train::NTHR, smem, st>>>(prm);
    }
    BNN_CUDA(cudaGetLastError());
    train::train_reduce_kernel<<<dim3(nb, n_seeds), 256, 0, st>>>(partial, d_theta, n_cta, fl.d, F,
                                                                 hp->beta_in * (float)B, hp->beta_out * (float)B, grad, sq);
    BNN_CUDA(cudaGetLastError());
    if (d_metrics) {
        train::train_metrics_kernel<<<n_seeds, 32, 0, st>>>(d_theta, grad, fl.d, F, (int)B, *hp, d_metrics);
        BNN_CUDA(cudaGetLastError());
    }
    train::train_update_kernel<<<dim3(nb, n_seeds), 256, 0, st>>>(grad, sq, nb, fl.d, F, (int)B, *hp, d_theta, d_momentum,
                                                                 d_grad_out, d_metrics);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

int bnn_saliency(const bnn_model_config* cfg, int32_t n_models, const float* d_theta, const float* d_x, int64_t B,
                 const float* d_eps12, uint64_t seed, float* d_grad_x, float* d_sumsq, float* d_mu, void* d_workspace,
                 void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_theta && d_x && d_sumsq && d_mu && d_workspace, BNN_E_ARG, "bnn_saliency: null pointer");
    BNN_REQUIRE(n_models >= 1 && n_models <= 65535 && B >= 1 && B < (1ll << 30), BNN_E_ARG,
                "bnn_saliency: n_models=%d B=%lld out of range", n_models, (long long)B);
    BNN_REQUIRE(cfg->n_times == 100 && cfg->n_features == 41, BNN_E_CONFIG,
                "bnn_saliency: compiled for T=100, F=41 (got T=%d, F=%d)", cfg->n_times, cfg->n_features);
    BNN_REQUIRE(aligned16(d_workspace) && (!d_eps12 || aligned16(d_eps12)), BNN_E_ALIGN, "bnn_saliency: alignment");
    const int F = cfg->n_features, T = cfg->n_times;
    const FlatLayout fl(F);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t n_cta = sms / n_models;
    if (n_cta > (B + 1) / 2) n_cta = (B + 1) / 2;
    if (n_cta < 1) n_cta = 1;
    const int DP = fl.d + train::DPAD;
    float* partial = (float*)d_workspace;
    float* xprod = partial + (size_t)n_models * n_cta * DP;
    xprod += (4 - ((uintptr_t)xprod / sizeof(float)) % 4) % 4;
    train::Params prm;
    prm.theta = d_theta; prm.X = d_x; prm.Y = nullptr; prm.batch_index = nullptr;
    prm.eps_in = nullptr; prm.eps12 = d_eps12; prm.eps_sum = nullptr;
    prm.partial = partial; prm.head_rec = nullptr; prm.xprod = xprod;
    prm.B = (int)B; prm.T = T; prm.F = F; prm.FP = (F + 3) & ~3; prm.n_cta = (int)n_cta;
    prm.seed = seed; prm.step = 0; prm.zero_mask = cfg->zero_mask;
    prm.hc = HeadConsts{cfg->lo_mu, cfg->hi_mu, cfg->lo_sd, cfg->hi_sd};
    prm.beta_out = 0.f;
    prm.saliency = 1; prm.gx_out = d_grad_x; prm.mu_out = d_mu;
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        BNN_CUDA(cudaFuncSetAttribute(train::train_fwd_bwd3_kernel<100, 41>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem3 = (size_t)train::Smem3(T, F).total * sizeof(float);
    train::train_fwd_bwd3_kernel<100, 41><<<dim3((unsigned)n_cta, n_models), train::NTHR3, smem3, st>>>(prm);
    BNN_CUDA(cudaGetLastError());
    train::saliency_finish_kernel<<<n_models, 64, 0, st>>>(partial, (int)n_cta, fl.d, F, d_sumsq);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

int bnn_train_timeline(unsigned long long* host_out, int32_t n) {
    using namespace bnn;
    BNN_REQUIRE(host_out && n >= 0, BNN_E_ARG, "bnn_train_timeline: null pointer");
    for (int i = 0; i < n; ++i) host_out[i] = 0;
#ifdef BNN_TRAIN_TIMELINE
    BNN_CUDA(cudaDeviceSynchronize());
    unsigned long long tmp[train::TL_N];
    BNN_CUDA(cudaMemcpyFromSymbol(tmp, train::g_train_tl, sizeof(tmp)));
    for (int i = 0; i < n && i < train::TL_N; ++i) host_out[i] = tmp[i];
#endif
    return BNN_OK;
}

int bnn_train_noise(const bnn_model_config* cfg, int32_t n_seeds, int64_t B, uint64_t seed, uint64_t step,
                    float* d_eps_in, float* d_eps12, float* d_eps_sum, void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_eps_in && d_eps12 && d_eps_sum && n_seeds >= 1 && n_seeds <= 65535 && B >= 1, BNN_E_ARG,
                "bnn_train_noise: null pointer or empty problem");
    BNN_REQUIRE(aligned16(d_eps12) && aligned16(d_eps_sum), BNN_E_ALIGN, "bnn_train_noise: eps12 / eps_sum alignment");
    train::train_noise_kernel<<<dim3((unsigned)B, n_seeds), 128, 0, (cudaStream_t)stream>>>(
        (int)B, cfg->n_times, cfg->n_features, seed, step, d_eps_in, d_eps12, d_eps_sum);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

int bnn_eval_loss(const bnn_model_config* cfg, const float* d_x, const float* d_y, int64_t B,
                  const float* d_theta_packed, int64_t n_units, const float* d_eps, uint64_t seed,
                  float* d_out_mu_sd, float* d_loss_sum, void* d_workspace, void* stream) {
    using namespace bnn;
    BNN_REQUIRE(d_y && d_loss_sum, BNN_E_ARG, "bnn_eval_loss: null pointer");
    BNN_REQUIRE(d_out_mu_sd || d_workspace, BNN_E_ARG,
                "bnn_eval_loss: give d_out_mu_sd or a workspace of n_units*B*2 floats");
    BNN_REQUIRE(n_units >= 1 && n_units < (1ll << 31), BNN_E_ARG, "bnn_eval_loss: n_units out of range");
    float* out = d_out_mu_sd ? d_out_mu_sd : (float*)d_workspace;
    int rc = bnn_predict(cfg, d_x, B, d_theta_packed, n_units, d_eps, nullptr, seed, 0, 0, 0, out, nullptr, nullptr, stream);
    if (rc != BNN_OK) return rc;
    train::nll_sum_kernel<<<(unsigned)n_units, 256, 0, (cudaStream_t)stream>>>((const float2*)out, (const float2*)d_y, B,
                                                                              d_loss_sum);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

}  // extern "C"
