// K4 v4: two independent CTAs per SM (included by train.cu after train_v3.cuh, whose tile / stash helpers it uses).
//
// v3 is latency-bound at three pipeline warps per scheduler (ncu: issue active 44 %): its phases are separated by block
// barriers and most of them cannot use all warps.  v4 trades the barrier-synchronous pair of systems for TWO CTAs per SM
// that each walk ONE system per iteration (T = 100 rows, 256 threads, 128 registers): whenever one CTA sits in a
// latency-bound phase (the pooling / head chain, a barrier, the tail of a GEMM) the other one issues.  To fit two CTAs
// the activation gradients overwrite the activations they belong to, which forces the weight-gradient outer products
// into three phases in backward order:
//     forward x -> h1 -> h2 -> f;  head;  g_f over f;
//     dW2 += g_f^T h2;  g_a2 over h2;  dW1 += g_a2^T h1;  g_a1 over h1;  dW0 += g_a1^T x', g_x, dlv_in
// A thread therefore owns one 4 x 8 block in EACH of the three matrices; the 96 accumulators live in the thread's TMEM
// lane (256 columns per CTA, 512 per SM: all of tensor memory) and are in registers only inside their phase.
// Shared memory per CTA: 141 feature rows x 100 (56.4 kB) + feature weights natural and transposed (33.2 kB) + head
// weights (13.4 kB) + scratch = 105.9 kB.  The input noise is drawn in the load phase by all threads (five Philox blocks
// interleaved per thread): the other CTA hides it, which is what v3 needs its four producer warps for.
// Same math, Philox streams, records and partial-gradient format as v3 (reference: spock_reg_model.py:486-528,547-593,
// 722-732; derivation in the header of train.cu).
#pragma once

namespace bnn {
namespace train {

constexpr int NTHR4 = 256;

struct Smem4 {
    int RP, xT, h1T, h2T, fT, W0T, b0, W1T, b1, W2T, b2, W2n, W1n, W0n, V0, V1, V2, cb, consts, small, tslot, total;
    __host__ __device__ Smem4(int T, int F) {
        RP = T;   // T = 100: consecutive feature rows are 4 banks apart -> conflict-free 16-byte accesses over 8 rows
        int o = 0;
        xT = o; o += F * RP;
        h1T = o; o += H * RP;
        h2T = o; o += H * RP;
        fT = o; o += L * RP;
        W0T = o; o += F * H;
        b0 = o; o += H;
        W1T = o; o += H * H;
        b1 = o; o += H;
        W2T = o; o += H * L;
        b2 = o; o += L;
        W2n = o; o += L * H;
        W1n = o; o += H * H;
        W0n = o; o += H * W0NP;
        V0 = o; o += H * S2;
        V1 = o; o += H * H;
        V2 = o; o += 2 * H;
        cb = o; o += 2 * H + 4;
        consts = o; o += C3_TOTAL;
        small = o; o += SMALL3;
        tslot = o; o += 4;
        total = o;
    }
};

template <int T, int F>
__global__ void __launch_bounds__(NTHR4, 2) train_fwd_bwd4_kernel(const Params prm) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NQ = T >> 2, RP = T, F4 = (F + 3) >> 2;
    static_assert(F == 41 && T == 100, "thread maps below are laid out for the reference's shape");
    const Smem4 L_(T, F);
    const FlatLayout fl(F);
    const int tid = threadIdx.x, lane = tid & 31;
    const int sidx = blockIdx.y;
    const float* th = prm.theta + (int64_t)sidx * fl.d;
    float* xT = sm + L_.xT; float* h1T = sm + L_.h1T; float* h2T = sm + L_.h2T; float* fT = sm + L_.fT;
    float* W0T = sm + L_.W0T; float* b0 = sm + L_.b0; float* W1T = sm + L_.W1T; float* b1 = sm + L_.b1;
    float* W2T = sm + L_.W2T; float* b2 = sm + L_.b2; float* W2n = sm + L_.W2n; float* W1n = sm + L_.W1n; float* W0n = sm + L_.W0n;
    float* V0s = sm + L_.V0; float* V1s = sm + L_.V1; float* V2s = sm + L_.V2; float* cbs = sm + L_.cb;
    float* cst = sm + L_.consts;
    float* sv = sm + L_.small;

    // ---- stage this seed's weights ----
    for (int i = tid; i < H * F; i += NTHR4) {
        const int j = i / F, c = i - j * F;
        const float w = __ldg(th + fl.W0 + i);
        W0T[c * H + j] = w;
        W0n[j * W0NP + c] = w;
    }
    for (int i = tid; i < H * (W0NP - F); i += NTHR4) W0n[(i / (W0NP - F)) * W0NP + F + i % (W0NP - F)] = 0.f;
    for (int i = tid; i < H * H; i += NTHR4) {
        const int j = i / H, k = i - j * H;
        const float w = __ldg(th + fl.W1 + i);
        W1T[k * H + j] = w;
        W1n[i] = w;
        V1s[i] = __ldg(th + fl.V1 + i);
        V0s[i] = __ldg(th + fl.V0 + i);   // H * S2 == H * H
    }
    for (int i = tid; i < L * H; i += NTHR4) {
        const int j = i / H, k = i - j * H;
        const float w = __ldg(th + fl.W2 + i);
        W2T[k * L + j] = w;
        W2n[i] = w;
    }
    if (tid < H) {
        b0[tid] = __ldg(th + fl.b0 + tid); b1[tid] = __ldg(th + fl.b1 + tid);
        cbs[tid] = __ldg(th + fl.c0 + tid); cbs[H + tid] = __ldg(th + fl.c1 + tid);
    }
    if (tid < 2 * H) V2s[tid] = __ldg(th + fl.V2 + tid);
    if (tid < 2) cbs[2 * H + tid] = __ldg(th + fl.c2 + tid);
    if (tid < L) b2[tid] = __ldg(th + fl.b2 + tid);
    if (tid < S2) {
        const float lv = __ldg(th + fl.lv_sum + tid);
        cst[C3_LVS + tid] = lv;
        cst[C3_ELVH + tid] = expf(__fdiv_rn(lv, 2.0f));
    }
    if (tid < F) cst[C3_NSC + tid] = expf(__fdiv_rn(__ldg(th + fl.lv_in + tid), 2.0f));

    // ---- thread roles ----
    const int cg_rg = tid / NQ, q_rg = tid - NQ * cg_rg;   // row GEMMs: 5 column groups (tid < 125), g_x: 6 (tid < 150)
    // outer products: dW0 / dW1: block tid % 50 over row group tid / 50 (5 quads), tid < 250;
    //                 dW2:       block tid % 25 over row group tid / 25, tid < 125
    const int rgA = tid / 50, blkA = tid - 50 * rgA, jbA = blkA / 5, kbA = blkA - 5 * jbA;
    const int rgC = tid / 25, blkC = tid - 25 * rgC, jbC = blkC / 5, kbC = blkC - 5 * jbC;
    const bool hasA = tid < 250, hasC = tid < 125;

    // TMEM stash: columns [0,32) dW0 block, [32,64) dW1 block, [64,96) dW2 block of this thread
    constexpr int TCOLS = 256;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + L_.tslot);
    if (tid < 32) { tmem_alloc(tslot, TCOLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t taddr = tbase + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)((tid >> 7) * 96);
    {
        float z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0.f;
        stash_store<32>(taddr, z);
        stash_store<32>(taddr + 32, z);
        stash_store<32>(taddr + 64, z);
    }
    // tid < 150: dlv_in partial sums of this thread's 8 g_x columns; 160 <= tid < 240: one of the 80 row sums
    // (0..39: b0 = sum g_a1, 40..79: column 40 of dW0); 128 <= tid < 168: b1 = sum g_a2 (own register)
    float aux[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float ab1 = 0.f, ab2 = 0.f, a_nll = 0.f, a_skl = 0.f;
    const float Tf = (float)T, Tm1 = (float)(T - 1);
    const uint64_t key = seed_key(prm.seed, sidx);
    __syncthreads();

    for (int b = blockIdx.x; b < prm.B; b += gridDim.x) {
        const int64_t sb = (int64_t)sidx * prm.B + b;
        const int64_t row = prm.batch_index ? (int64_t)prm.batch_index[sb] : (int64_t)b;
        const float* xs = prm.X + row * (int64_t)T * F;
        // ---- S0: x' = mask(x) + eps_in * exp(lv_in/2), feature-major; items (4-column group, time step) ----
        {
            constexpr int NR = 5;   // 1100 items over 256 threads
            float xv[NR][4], ev[NR][4];
            int tt[NR], c4s[NR];
            uint4 ctr[NR];
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                const int id0 = tid + NTHR4 * k;
                const int id = id0 < T * F4 ? id0 : T * F4 - 1;
                const int c4 = id / T, t = id - c4 * T;
                c4s[k] = c4;
                tt[k] = id0 < T * F4 ? t : -1;
                ctr[k] = make_uint4((uint32_t)(t * F4 + c4), (uint32_t)b, (uint32_t)prm.step, STREAM_EPS_IN);
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[k][u] = __ldg(xs + t * F + min(4 * c4 + u, F - 1));
            }
            if (prm.eps_in) {
                const float* es = prm.eps_in + sb * (int64_t)T * F;
#pragma unroll
                for (int k = 0; k < NR; ++k) {
                    const int t = (int)ctr[k].x / F4;
#pragma unroll
                    for (int u = 0; u < 4; ++u) ev[k][u] = __ldg(es + t * F + min(4 * c4s[k] + u, F - 1));
                }
            } else {
                constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
                uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
                for (int r = 0; r < 10; ++r) {
#pragma unroll
                    for (int k = 0; k < NR; ++k) {
                        const uint4 c = ctr[k];
                        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
                        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
                        ctr[k] = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
                    }
                    k0 += W0;
                    k1 += W1;
                }
#pragma unroll
                for (int k = 0; k < NR; ++k) {
                    const float4 n4 = box_muller_fast(ctr[k]);
                    ev[k][0] = n4.x; ev[k][1] = n4.y; ev[k][2] = n4.z; ev[k][3] = n4.w;
                }
            }
#pragma unroll
            for (int k = 0; k < NR; ++k)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = 4 * c4s[k] + u;
                    if (c < F && tt[k] >= 0) {
                        float x = xv[k][u];
                        if ((prm.zero_mask >> c) & 1ull) x = __fsub_rn(x, x);  // x - mask keeps NaN (:452-478)
                        xT[c * RP + tt[k]] = __fadd_rn(x, __fmul_rn(ev[k][u], cst[C3_NSC + c]));
                    }
                }
            if (tid >= 224 && tid < 224 + S2 / 4) {   // a warp with a short item list
                const int l = tid - 224;
                float4 a, c;
                if (prm.eps12) {
                    a = __ldg(reinterpret_cast<const float4*>(prm.eps12 + sb * S2) + l);
                    c = __ldg(reinterpret_cast<const float4*>(prm.eps_sum + sb * S2) + l);
                } else {
                    a = philox_normal4(key, STREAM_EPS, (uint32_t)b, (uint32_t)prm.step, (uint32_t)l);
                    c = philox_normal4(key, STREAM_EPS_SUM, (uint32_t)b, (uint32_t)prm.step, (uint32_t)l);
                }
                *reinterpret_cast<float4*>(sv + V3_E12 + 4 * l) = a;
                *reinterpret_cast<float4*>(sv + V3_ESN + 4 * l) = c;
            }
            if (tid == 255) {
                sv[V3_Y] = __ldg(prm.Y + row * 2);
                sv[V3_Y + 1] = __ldg(prm.Y + row * 2 + 1);
            }
        }
        __syncthreads();
        // ---- S1..S3: feature_nn forward ----
        if (tid < 5 * NQ) {
            u64 a2[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u64 bv = *reinterpret_cast<const u64*>(b0 + 8 * cg_rg + 2 * i);
#pragma unroll
                for (int r = 0; r < 4; ++r) a2[r][i] = bv;
            }
            rowgemm4<RP, F, H, 8>(xT, W0T, q_rg, 8 * cg_rg, a2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) unpack2(a2[r][i], v[r][0], v[r][1]);
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    *reinterpret_cast<float4*>(h1T + (8 * cg_rg + 2 * i + e) * RP + 4 * q_rg) =
                        make_float4(relu_nan(v[0][e]), relu_nan(v[1][e]), relu_nan(v[2][e]), relu_nan(v[3][e]));
            }
        }
        __syncthreads();
        if (tid < 5 * NQ) {
            u64 a2[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u64 bv = *reinterpret_cast<const u64*>(b1 + 8 * cg_rg + 2 * i);
#pragma unroll
                for (int r = 0; r < 4; ++r) a2[r][i] = bv;
            }
            rowgemm4<RP, H, H, 8>(h1T, W1T, q_rg, 8 * cg_rg, a2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) unpack2(a2[r][i], v[r][0], v[r][1]);
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    *reinterpret_cast<float4*>(h2T + (8 * cg_rg + 2 * i + e) * RP + 4 * q_rg) =
                        make_float4(relu_nan(v[0][e]), relu_nan(v[1][e]), relu_nan(v[2][e]), relu_nan(v[3][e]));
            }
        }
        __syncthreads();
        if (tid < 5 * NQ) {
            u64 a2[4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const u64 bv = *reinterpret_cast<const u64*>(b2 + 4 * cg_rg + 2 * i);
#pragma unroll
                for (int r = 0; r < 4; ++r) a2[r][i] = bv;
            }
            rowgemm4<RP, H, L, 4>(h2T, W2T, q_rg, 4 * cg_rg, a2);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float v[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) unpack2(a2[r][i], v[r][0], v[r][1]);
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    *reinterpret_cast<float4*>(fT + (4 * cg_rg + 2 * i + e) * RP + 4 * q_rg) =
                        make_float4(v[0][e], v[1][e], v[2][e], v[3][e]);
            }
        }
        __syncthreads();
        // ---- S4: pooling (two-pass mean / unbiased variance per latent column, :418-419) ----
        if (tid < L * 8) {
            const int c = tid >> 3, part = tid & 7;
            const float* fc = fT + c * RP;
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                v[i] = (part + 8 * i < NQ) ? *reinterpret_cast<const float4*>(fc + 4 * (part + 8 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            const float mean = __fdiv_rn(s, Tf);
            float m2 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (part + 8 * i < NQ) {
                    const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
                    m2 = fmaf(d0, d0, m2); m2 = fmaf(d1, d1, m2); m2 = fmaf(d2, d2, m2); m2 = fmaf(d3, d3, m2);
                }
            m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
            m2 += __shfl_xor_sync(0xffffffffu, m2, 4);
            if (part == 0) {
                const float sd = sqrtf(__fdiv_rn(m2, Tm1));
                const float var = __fmul_rn(sd, sd);
                const float sim = sqrtf(__fdiv_rn(var, Tf));                                   // :422
                const float siv = sqrtf(__fdiv_rn(__fmul_rn(2.0f, __fmul_rn(var, var)), Tm1));   // :423
                const float e1 = sv[V3_E12 + c], e2 = sv[V3_E12 + L + c];
                const float mus = __fadd_rn(__fmul_rn(e1, sim), mean);                          // :426
                const float vs = __fadd_rn(__fmul_rn(e2, siv), var);                            // :427
                const float sds = sqrtf(__fadd_rn(fabsf(vs), 1e-5f));                           // :430
                sv[V3_M + c] = mean; sv[V3_VAR + c] = var; sv[V3_SIM + c] = sim; sv[V3_SIV + c] = siv; sv[V3_VS + c] = vs;
                sv[V3_S + c] = mus; sv[V3_S + L + c] = sds;
                const float lv0 = cst[C3_LVS + c], lv1 = cst[C3_LVS + L + c];
                sv[V3_SP + c] = __fadd_rn(mus, __fmul_rn(sv[V3_ESN + c], cst[C3_ELVH + c]));
                sv[V3_SP + L + c] = __fadd_rn(sds, __fmul_rn(sv[V3_ESN + L + c], cst[C3_ELVH + L + c]));
                a_skl += 0.5f * (mus * mus + expf(lv0) - lv0 - 1.0f) + 0.5f * (sds * sds + expf(lv1) - lv1 - 1.0f);
            }
        }
        __syncthreads();
        // ---- S6: regress_nn forward ----
        if (tid < H * 4) {
            const int j = tid >> 2, part = tid & 3;
            const float* w = V0s + j * S2 + 10 * part;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) a = fmaf(sv[V3_SP + 10 * part + k], w[k], a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[V3_R1 + j] = relu_nan(a + cbs[j]);
        }
        __syncthreads();
        if (tid < H * 4) {
            const int j = tid >> 2, part = tid & 3;
            const float* w = V1s + j * H + 10 * part;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 10; ++k) a = fmaf(sv[V3_R1 + 10 * part + k], w[k], a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[V3_R2 + j] = relu_nan(a + cbs[H + j]);
        }
        __syncthreads();
        float gr0 = 0.f, gr1 = 0.f;   // valid in warp 0
        if (tid < 32) {
            const int o = lane >> 4, l16 = lane & 15;
            float a = 0.f;
            for (int k = l16; k < H; k += 16) a = fmaf(sv[V3_R2 + k], V2s[o * H + k], a);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 8);
            const float r0 = __shfl_sync(0xffffffffu, a, 0) + cbs[2 * H];
            const float r1 = __shfl_sync(0xffffffffu, a, 16) + cbs[2 * H + 1];
            const float t0 = tanhf(r0), t1 = tanhf(r1);
            const float mu = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, __fadd_rn(t0, 1.0f)), __fsub_rn(prm.hc.hi_mu, prm.hc.lo_mu)), prm.hc.lo_mu);
            const float sd = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, __fadd_rn(t1, 1.0f)), __fsub_rn(prm.hc.hi_sd, prm.hc.lo_sd)), prm.hc.lo_sd);
            float l = 0.f, dm = 0.f, ds = 0.f;
            if (lane < 2) nll_terms(mu, sd, sv[V3_Y + lane], l, dm, ds);   // one label per lane
            const float l1 = __shfl_sync(0xffffffffu, l, 1), dm1 = __shfl_sync(0xffffffffu, dm, 1), ds1 = __shfl_sync(0xffffffffu, ds, 1);
            if (lane == 0) {
                a_nll += -(l + l1);
                const float gmu = -(dm + dm1), gsd = -(ds + ds1);
                gr0 = gmu * 0.5f * (prm.hc.hi_mu - prm.hc.lo_mu) * (1.0f - t0 * t0);
                gr1 = gsd * 0.5f * (prm.hc.hi_sd - prm.hc.lo_sd) * (1.0f - t1 * t1);
            }
            // ---- S7: regress_nn backward: g_a2 = (V2^T g_r) . [r2 > 0] ----
            gr0 = __shfl_sync(0xffffffffu, gr0, 0);
            gr1 = __shfl_sync(0xffffffffu, gr1, 0);
            for (int k = lane; k < H; k += 32) {
                const float g = gr0 * V2s[k] + gr1 * V2s[H + k];
                sv[V3_G2 + k] = sv[V3_R2 + k] > 0.f ? g : 0.f;
            }
        }
        __syncthreads();
        if (tid < H * 4) {
            const int k = tid >> 2, part = tid & 3;
            float a = 0.f;
#pragma unroll
            for (int jj = 0; jj < 10; ++jj) {
                const int j = 10 * part + jj;
                a = fmaf(sv[V3_G2 + j], V1s[j * H + k], a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) sv[V3_G1 + k] = sv[V3_R1 + k] > 0.f ? a : 0.f;
        }
        __syncthreads();
        float* rec = prm.head_rec + sb * REC;
        if (tid < S2 * 4) {
            const int k = tid >> 2, part = tid & 3;
            float a = 0.f;
#pragma unroll
            for (int jj = 0; jj < 10; ++jj) {
                const int j = 10 * part + jj;
                a = fmaf(sv[V3_G1 + j], V0s[j * S2 + k], a);
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (part == 0) {
                rec[R_DLVS + k] = a * (0.5f * (sv[V3_ESN + k] * cst[C3_ELVH + k]));  // ds'/dlv = eps e^{lv/2} / 2
                sv[V3_GS + k] = a + prm.beta_out * sv[V3_S + k];
            }
        } else if (tid < S2 * 4 + 64) {
            // two other warps write the record (s', r1, r2, g_a1, g_a2 are complete)
            for (int i = tid - S2 * 4; i < R_DLVS; i += 64) rec[i] = sv[V3_REC0 + i];
        }
        if (tid == 0) { rec[R_GR] = gr0; rec[R_GR + 1] = gr1; }
        __syncthreads();
        if (tid < L) {
            const int c = tid;
            const float gmus = sv[V3_GS + c], gsds = sv[V3_GS + L + c];
            const float vs = sv[V3_VS + c], sds = sv[V3_S + L + c], var = sv[V3_VAR + c];
            const float sgn = vs > 0.f ? 1.0f : (vs < 0.f ? -1.0f : 0.f);
            const float gvs = gsds * sgn / (2.0f * sds);
            const float e1 = sv[V3_E12 + c], e2 = sv[V3_E12 + L + c];
            const float gv = gmus * e1 / (2.0f * Tf * sv[V3_SIM + c]) +
                             gvs * (1.0f + e2 * (2.0f * var) / (Tm1 * sv[V3_SIV + c]));
            sv[V3_GM + c] = gmus / Tf;        // coefficient of 1
            sv[V3_GV + c] = 2.0f * gv / Tm1;  // coefficient of (f - m)
        }
        __syncthreads();
        // ---- S8: g_f in place over f ----
        for (int i = tid; i < L * NQ; i += NTHR4) {
            const int c = i / NQ, q = i - c * NQ;
            float4* p = reinterpret_cast<float4*>(fT + c * RP + 4 * q);
            const float m = sv[V3_M + c], A = sv[V3_GM + c], Bc = sv[V3_GV + c];
            float4 f = *p;
            f.x = fmaf(Bc, f.x - m, A); f.y = fmaf(Bc, f.y - m, A); f.z = fmaf(Bc, f.z - m, A); f.w = fmaf(Bc, f.w - m, A);
            *p = f;
        }
        __syncthreads();
        // ---- S9: dW2 += g_f^T h2 (tid < 125); b2 = column sums of g_f (tid 128..147) ----
        if (tid < 128) {   // whole warps: tcgen05.ld / .st are warp-collective; lanes without a block run zero quads
            float aW[4][8];
            stash_load<32>(taddr + 64, &aW[0][0]);
            outer4x8(fT + (hasC ? jbC : 0) * RP, 5 * RP, h2T + kbC * RP, 5 * RP, hasC ? 5 * rgC : 0, hasC ? 5 * rgC + 5 : 0, aW);
            stash_store<32>(taddr + 64, &aW[0][0]);
        } else if (tid < 128 + L) {
            const float* g = fT + (tid - 128) * RP;
            float s = 0.f;
            for (int r = 0; r < T; r += 4) {
                const float4 v = *reinterpret_cast<const float4*>(g + r);
                s += (v.x + v.y) + (v.z + v.w);
            }
            ab2 += s;
        }
        __syncthreads();
        // ---- S10: g_a2 = (g_f W2) . [h2 > 0], in place over h2 ----
        if (tid < 5 * NQ) {
            u64 a2[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) a2[r][i] = 0ull;
            rowgemm4<RP, L, H, 8>(fT, W2n, q_rg, 8 * cg_rg, a2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) unpack2(a2[r][i], v[r][0], v[r][1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float4* hp = reinterpret_cast<float4*>(h2T + (8 * cg_rg + 2 * i + e) * RP + 4 * q_rg);
                    const float4 h = *hp;
                    *hp = make_float4(h.x > 0.f ? v[0][e] : 0.f, h.y > 0.f ? v[1][e] : 0.f, h.z > 0.f ? v[2][e] : 0.f,
                                      h.w > 0.f ? v[3][e] : 0.f);
                }
            }
        }
        __syncthreads();
        // ---- S11: dW1 += g_a2^T h1 (tid < 250) ----
        {
            float aW[4][8];
            stash_load<32>(taddr + 32, &aW[0][0]);
            outer4x8(h2T + (hasA ? jbA : 0) * RP, 10 * RP, h1T + kbA * RP, 5 * RP, hasA ? 5 * rgA : 0, hasA ? 5 * rgA + 5 : 0, aW);
            stash_store<32>(taddr + 32, &aW[0][0]);
        }
        __syncthreads();
        // ---- S12: g_a1 = (g_a2 W1) . [h1 > 0], in place over h1; b1 = row sums of g_a2 (tid 128..167) ----
        if (tid < 5 * NQ) {
            u64 a2[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) a2[r][i] = 0ull;
            rowgemm4<RP, H, H, 8>(h2T, W1n, q_rg, 8 * cg_rg, a2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) unpack2(a2[r][i], v[r][0], v[r][1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float4* hp = reinterpret_cast<float4*>(h1T + (8 * cg_rg + 2 * i + e) * RP + 4 * q_rg);
                    const float4 h = *hp;
                    *hp = make_float4(h.x > 0.f ? v[0][e] : 0.f, h.y > 0.f ? v[1][e] : 0.f, h.z > 0.f ? v[2][e] : 0.f,
                                      h.w > 0.f ? v[3][e] : 0.f);
                }
            }
        } else if (tid >= 128 && tid < 128 + H) {
            const float* g = h2T + (tid - 128) * RP;
            float s0 = 0.f, s1 = 0.f;
            for (int t = 0; t < T; t += 4) {
                const float4 v = *reinterpret_cast<const float4*>(g + t);
                s0 += v.x + v.z; s1 += v.y + v.w;
            }
            ab1 += s0 + s1;
        }
        __syncthreads();
        // ---- S13: dW0 += g_a1^T x' (tid < 250); S14: g_x = g_a1 W0 and dlv_in (tid < 150); row sums (160..239) ----
        {
            float aW[4][8];
            stash_load<32>(taddr, &aW[0][0]);
            outer4x8(h1T + (hasA ? jbA : 0) * RP, 10 * RP, xT + kbA * RP, 5 * RP, hasA ? 5 * rgA : 0, hasA ? 5 * rgA + 5 : 0, aW);
            stash_store<32>(taddr, &aW[0][0]);
        }
        if (tid < 6 * NQ) {
            u64 a2[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) a2[r][i] = 0ull;
            rowgemm4<RP, H, W0NP, 8>(h1T, W0n, q_rg, 8 * cg_rg, a2);
            // dlv_in needs x' - mask(x): the noise-free input comes back through L2
            const float* xr = xs + 4 * q_rg * F;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v[4][2];
#pragma unroll
                for (int r = 0; r < 4; ++r) unpack2(a2[r][i], v[r][0], v[r][1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * cg_rg + 2 * i + e;
                    if (col < F) {
                        const bool zeroed = (prm.zero_mask >> col) & 1ull;
                        const float4 xp = *reinterpret_cast<const float4*>(xT + col * RP + 4 * q_rg);
                        const float xpr[4] = {xp.x, xp.y, xp.z, xp.w};
                        float s = 0.f;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            float x0 = __ldg(xr + r * F + col);
                            if (zeroed) x0 = __fsub_rn(x0, x0);
                            s = fmaf(v[r][e], xpr[r] - x0, s);
                        }
                        aux[2 * i + e] += s;
                    }
                }
            }
        } else if (tid >= 160 && tid < 160 + 2 * H) {
            const int r = tid - 160;
            const float* g = h1T + (r < H ? r : r - H) * RP;
            float s0 = 0.f, s1 = 0.f;
            if (r >= H) {
                const float* xc = xT + (F - 1) * RP;
                for (int t = 0; t < T; t += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(g + t);
                    const float4 x = *reinterpret_cast<const float4*>(xc + t);
                    s0 = fmaf(v.x, x.x, s0); s1 = fmaf(v.y, x.y, s1); s0 = fmaf(v.z, x.z, s0); s1 = fmaf(v.w, x.w, s1);
                }
            } else {
                for (int t = 0; t < T; t += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(g + t);
                    s0 += v.x + v.z; s1 += v.y + v.w;
                }
            }
            aux[0] += s0 + s1;
        }
        __syncthreads();
    }

    // =====================================================================================================
    // Epilogue: this CTA's partial gradient in flatten() order.  The activation buffers are dead: scratch.
    // =====================================================================================================
    float* part = prm.partial + ((int64_t)sidx * prm.n_cta + blockIdx.x) * (fl.d + DPAD);
    float* red = sm;  // 14,100 floats of activation area
    // (1) feature matrices: add the five row groups in a fixed order, one matrix at a time (8,000 floats of scratch)
#pragma unroll 1
    for (int mtx = 0; mtx < 3; ++mtx) {
        const bool has = mtx == 2 ? hasC : hasA;
        const int nblk = mtx == 2 ? 25 : 50;
        float aW[4][8];
        stash_load<32>(taddr + 32 * mtx, &aW[0][0]);
        if (has) {
            float* o = red + tid * 32;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) o[jj * 8 + kk] = aW[jj][kk];
        }
        __syncthreads();
        if (has && tid < nblk) {
            const int jb = mtx == 2 ? jbC : jbA, kb = mtx == 2 ? kbC : kbA;
            const int jstr = mtx == 2 ? 5 : 10;
            const int off = mtx == 0 ? fl.W0 : (mtx == 1 ? fl.W1 : fl.W2);
            const int pitch = mtx == 0 ? F : H;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    float a = aW[jj][kk];
                    for (int g = 1; g < 5; ++g) a += red[(tid + g * nblk) * 32 + jj * 8 + kk];
                    part[off + (jb + jstr * jj) * pitch + kb + 5 * kk] = a;
                }
        }
        __syncthreads();
    }
    // (2) dlv_in (fixed-order sum over the row quads), b0 / column 40 of dW0, b1, b2
    if (tid < 6 * NQ) {
#pragma unroll
        for (int c = 0; c < 8; ++c) red[(8 * cg_rg + c) * NQ + q_rg] = aux[c];
    } else if (tid >= 160 && tid < 160 + 2 * H) {
        const int r = tid - 160;
        if (r < H) part[fl.b0 + r] = aux[0];
        else part[fl.W0 + (r - H) * F + (F - 1)] = aux[0];
    }
    if (tid >= 128 && tid < 128 + H) part[fl.b1 + tid - 128] = ab1;
    if (tid >= 128 && tid < 128 + L) part[fl.b2 + tid - 128] = ab2;
    __syncthreads();
    if (tid < F) {
        float s = 0.f;
        for (int q = 0; q < NQ; ++q) s += red[tid * NQ + q];
        part[fl.lv_in + tid] = 0.5f * s;
    }
    __syncthreads();
    // (3) metrics
    red[tid] = (tid < L * 8 && (tid & 7) == 0) ? a_skl : 0.f;
    if (tid == 0) red[NTHR4] = a_nll;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int c = 0; c < L; ++c) s += red[8 * c];
        part[fl.d + SLOT_NLL] = red[NTHR4];
        part[fl.d + SLOT_SKL] = s;
        for (int i = 2; i < DPAD; ++i) part[fl.d + i] = 0.f;
    }
    __syncthreads();
    // (4) head gradients from the records of this CTA's systems, in system order
    {
        constexpr int CH = 48;  // records staged per chunk (48 kB)
        // roles: tid < 200: 2 x 4 blocks of dV0 and dV1; 200..239: c0, c1; 240..255 + second pass: dV2, c2, dlv_sum
        const int jbh = tid / 10, kb = tid % 10;
        float aV0[2][4] = {}, aV1[2][4] = {};
        float s0 = 0.f, s1 = 0.f;
        float t2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // tid >= 240: 16 threads x 8 of the 122 (dV2 | c2 | dlv_sum) sums
        const int n_sys = (prm.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        for (int c0i = 0; c0i < n_sys; c0i += CH) {
            const int nc = min(CH, n_sys - c0i);
            for (int i = tid; i < nc * (REC / 4); i += NTHR4) {
                const int s = c0i + i / (REC / 4), w = i % (REC / 4);
                const int bsys = blockIdx.x + gridDim.x * s;
                reinterpret_cast<float4*>(red)[i] =
                    __ldcg(reinterpret_cast<const float4*>(prm.head_rec + ((int64_t)sidx * prm.B + bsys) * REC) + w);
            }
            __syncthreads();
            if (tid < 200) {
                for (int s = 0; s < nc; ++s) {
                    const float* r = red + s * REC;
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            aV0[jj][kk] = fmaf(r[R_G1 + 2 * jbh + jj], r[R_SP + kb + 10 * kk], aV0[jj][kk]);
                            aV1[jj][kk] = fmaf(r[R_G2 + 2 * jbh + jj], r[R_R1 + kb + 10 * kk], aV1[jj][kk]);
                        }
                }
            } else if (tid < 240) {
                for (int s = 0; s < nc; ++s) { s0 += red[s * REC + R_G1 + tid - 200]; s1 += red[s * REC + R_G2 + tid - 200]; }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int id = (tid - 240) * 8 + e;   // 0..79: dV2[o][k]; 80, 81: c2; 82..121: dlv_sum
                    if (id < 2 * H) {
                        const int o = id / H, k = id - o * H;
                        for (int s = 0; s < nc; ++s) t2[e] = fmaf(red[s * REC + R_GR + o], red[s * REC + R_R2 + k], t2[e]);
                    } else if (id < 2 * H + 2) {
                        for (int s = 0; s < nc; ++s) t2[e] += red[s * REC + R_GR + id - 2 * H];
                    } else if (id < 2 * H + 2 + S2) {
                        for (int s = 0; s < nc; ++s) t2[e] += red[s * REC + R_DLVS + id - 2 * H - 2];
                    }
                }
            }
            __syncthreads();
        }
        if (tid < 200) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    part[fl.V0 + (2 * jbh + jj) * S2 + kb + 10 * kk] = aV0[jj][kk];
                    part[fl.V1 + (2 * jbh + jj) * H + kb + 10 * kk] = aV1[jj][kk];
                }
        } else if (tid < 240) {
            part[fl.c0 + tid - 200] = s0;
            part[fl.c1 + tid - 200] = s1;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int id = (tid - 240) * 8 + e;
                if (id < 2 * H) part[fl.V2 + id] = t2[e];
                else if (id < 2 * H + 2) part[fl.c2 + id - 2 * H] = t2[e];
                else if (id < 2 * H + 2 + S2) part[fl.lv_sum + id - 2 * H - 2] = t2[e];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tbase, TCOLS);
}

}  // namespace train
}  // namespace bnn
