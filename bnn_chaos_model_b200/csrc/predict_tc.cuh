// K2-TC: the fused MultiSWAG predictive kernel on the 5th-generation tensor cores.
//
// Same reference path as predict_device.cuh (/root/reference/spock_reg_model.py:416-442,
// :452-478, :878-908); feature_nn's three layers run as tcgen05.mma kind::tf32 with the
// 3xTF32 split (a = a_hi + a_lo, w = w_hi + w_lo; a w ~= a_hi w_hi + a_lo w_hi + a_hi w_lo,
// fp32 accumulation in TMEM), which keeps the per-system outputs within the 1e-5 relative
// tolerance of the fp32 reference (measured: tests/test_gpu_predict.py).
//
// Why: the FFMA2 kernel is bound by the shared-memory -> register path (ncu: 93 % of the LSU
// wavefront peak at 52 % FMA utilisation, profiles/), because every lane needs r + c operand words
// per r*c FMAs.  tcgen05 reads B from shared memory and A from tensor memory itself; the CUDA cores
// only run the per-row epilogue (ReLU + hi/lo split), 3 instructions per activation.
//
// One CTA per SM, 5 systems (500 time-step rows -> four 128-row M tiles) resident in shared memory
// as fp32; per (unit, M tile) "job" a 3-deep ring of TMEM slots [A_hi 48 | A_lo 40 | D 48 columns]:
//   epilogue warps (4 per slot, thread = row = TMEM lane):
//       stage x: smem -> hi/lo -> tcgen05.st A          -> arrive a_ready
//       after each layer: tcgen05.ld D -> ReLU -> hi/lo -> tcgen05.st A   -> arrive a_ready
//       last layer: D -> per-32-row-block pooled (mean, M2) records
//   MMA warp (one thread): waits a_ready, issues 12/17/17 tcgen05.mma per layer, tcgen05.commit -> d_ready
//   producer warp: cp.async.bulk of each unit's weights (hi/lo B operands + fp32 head) into a 2-slot ring
//   2 tail warps: Chan-merge the records per system, sampled summary statistics, regress_nn, store.
// Biases ride in the GEMMs: x carries a ones column (index 31), the hidden activations a constant
// ones block in TMEM columns 40..47 of A_hi.
#pragma once
#include "predict_device.cuh"
#include "tc.cuh"

namespace bnn {
namespace tc {

constexpr int SYS = 5;            // systems per CTA tile
constexpr int MT = 4;             // 128-row M tiles per tile (5 * 100 rows -> 512)
constexpr int T_FIXED = 100;      // time steps (the tiling is specific to T = 100)
constexpr int ROWS = SYS * T_FIXED;
constexpr int TM_AHI = 0, TM_ALO = 48, TM_D = 88, TM_SLOT = 136;
constexpr int N_BLOCKS = MT * 4;  // 32-row blocks per tile
constexpr int REC_FLOATS = N_BLOCKS * 2 * L * 2;  // [block][segment][col][mean, M2]
constexpr int FB_FLOATS = 32 * L;                 // per epilogue warp
constexpr int TAIL_SCRATCH = 1024;

struct Bars {
    uint64_t w_full[2], w_empty[2], unit_done[2];
    uint64_t a_ready[3], d_ready[3];
    uint32_t tmem_base;
    uint32_t pad;
};

// block b covers tile rows [32b, 32b+32): rows of system sysA up to `split`, then system sysA+1
__device__ __forceinline__ void block_geom(int b, int& sysA, int& split, int& nvalid) {
    const int R0 = 32 * b;
    sysA = R0 / T_FIXED;
    split = min(32, T_FIXED * (sysA + 1) - R0);
    nvalid = min(32, ROWS - R0);
}

// ---------------------------------------------------------------------------------------
// x tile: X[N,T,F] -> xs[row][32] fp32, 16-byte chunks XOR-swizzled by (row & 7) so that a warp's
// 32 rows read conflict-free; live columns packed first, column 31 = 1.0 (bias), rest 0.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int xs_index(int row, int c) { return row * 32 + ((((c >> 2) ^ (row & 7))) << 2) + (c & 3); }

__device__ __forceinline__ void load_x_tile_tc(const float* __restrict__ X, int64_t n0, int n_valid, int F, int kin,
                                               const ColMap& cm, float* __restrict__ xs, int* __restrict__ poison) {
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) xs[i] = 0.f;
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) poison[r] = 0;
    __syncthreads();
    const float* src = X + n0 * (int64_t)T_FIXED * F;
    const int total = n_valid * T_FIXED * F;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int row = idx / F, c = idx - row * F;
        const float v = __ldg(src + idx);
        const int k = cm.inv[c];
        if (k >= 0)
            xs[xs_index(row, k)] = v;
        else if (!isfinite(v))
            poison[row] = 1;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
        xs[xs_index(r, TC_K1 - 1)] = 1.0f;
        if (poison[r]) xs[xs_index(r, 0)] = __int_as_float(0x7fc00000);  // x - mask keeps NaN/Inf as NaN (:452-478)
    }
    __syncthreads();
}

// hi/lo split of 8 values and store to the A_hi / A_lo columns of this thread's TMEM lane.
// hi = v rounded to nearest tf32 (|lo| <= 2^-12 |v|), lo = v - hi exactly.
__device__ __forceinline__ void split_store8(const float (&v)[8], uint32_t t_hi, uint32_t t_lo) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float hi = tf32_rna(v[j]);
        h[j] = __float_as_uint(hi);
        l[j] = __float_as_uint(v[j] - hi);
    }
    tmem_st8(t_hi, h);
    tmem_st8(t_lo, l);
}

// ---------------------------------------------------------------------------------------
// MMA issue (one thread).  wb: shared-memory byte address of this unit's ring slot; offsets in floats
// relative to the slot start.  ts: TMEM address of the slot (lane 0).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t bdesc(uint32_t base_addr, int N, int ks) {
    const uint32_t chunk = (uint32_t)N * 16u;
    return smem_desc_kmajor(base_addr + (uint32_t)ks * 2u * chunk, chunk, 128u);
}

template <int N, int KS_HI, int KS_LO>
__device__ __forceinline__ void issue_layer(uint32_t ts, uint32_t bh_addr, uint32_t bl_addr) {
    constexpr uint32_t idesc = idesc_tf32(128, N);
    const uint32_t d = ts + TM_D, ahi = ts + TM_AHI, alo = ts + TM_ALO;
    // descriptors advance by two 16-byte K chunks (2*N*16 bytes -> 2*N in the 16-byte address field) per
    // K = 8 step; the loops stay rolled so that the 46 descriptors of a job are not all kept in registers
    const uint64_t dh = bdesc(bh_addr, N, 0), dl = bdesc(bl_addr, N, 0);
    constexpr uint64_t step = 2ull * N;
    // called by the whole (converged) MMA warp; one elected lane issues
    if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < KS_HI; ++ks) mma_tf32_ts(d, ahi + 8 * ks, dh + ks * step, idesc, ks > 0);
#pragma unroll
        for (int ks = 0; ks < KS_LO; ++ks) mma_tf32_ts(d, alo + 8 * ks, dh + ks * step, idesc, true);
#pragma unroll
        for (int ks = 0; ks < KS_HI; ++ks) mma_tf32_ts(d, ahi + 8 * ks, dl + ks * step, idesc, true);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// Tail of one unit from the per-block records (one warp; lane = p*4+q, p = system slot, q = 5 columns).
// ring: this unit's ring slot (starts at PackedLayout::V0p); scratch: TAIL_SCRATCH floats.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tail_unit_tc(const float* __restrict__ rec, const float* __restrict__ ring,
                                             const PackedLayout& pl, const float* __restrict__ eps_u,
                                             const float* __restrict__ eps_sum_u, float* __restrict__ summary_u,
                                             uint64_t seed, uint32_t gunit, int64_t gsys0, int64_t n0, int n_valid,
                                             const HeadConsts& hc, float* __restrict__ scratch,
                                             float* __restrict__ out_unit, int64_t out_sys_stride) {
    const int lane = threadIdx.x & 31;
    const int p = lane >> 2, q = lane & 3;
    const float* thp = ring - pl.V0p;  // so that thp + pl.<head field> addresses the ring slot
    float* sA = scratch;
    float* sB = scratch + SYS_TILE * 41;
    float* eS = scratch + 2 * SYS_TILE * 41;

    if (eps_u) {
        for (int idx = lane; idx < SYS_TILE * S2; idx += 32) {
            const int s = idx / S2, j = idx % S2;
            eS[idx] = (s < n_valid) ? __ldg(eps_u + (n0 + s) * S2 + j) : 0.f;
        }
    } else {
        for (int b = lane; b < SYS_TILE * (S2 / 4); b += 32) {
            const int s = b / (S2 / 4), blk = b % (S2 / 4);
            const float4 n4 = philox_normal4(seed, STREAM_EPS, gunit, (uint32_t)(gsys0 + s), (uint32_t)blk);
            *reinterpret_cast<float4*>(eS + s * S2 + blk * 4) = n4;
        }
    }
    __syncwarp();

    const float Tf = (float)T_FIXED, Tm1 = (float)(T_FIXED - 1);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int col = q * 5 + i;
        float n = 0.f, mean = 0.f, m2 = 0.f;
        if (p < SYS) {
            const int bfirst = (T_FIXED * p) / 32, blast = (T_FIXED * p + T_FIXED - 1) / 32;
            for (int b = bfirst; b <= blast; ++b) {
                int sysA, split, nvalid;
                block_geom(b, sysA, split, nvalid);
                const int seg = (sysA == p) ? 0 : 1;
                const int cnt = (seg == 0) ? min(split, nvalid) : (nvalid - split);
                if (cnt <= 0) continue;
                const float2 r = *reinterpret_cast<const float2*>(rec + ((b * 2 + seg) * L + col) * 2);
                const float nb = (float)cnt;
                if (n == 0.f) {
                    n = nb; mean = r.x; m2 = r.y;
                } else {
                    const float nn = n + nb, delta = r.x - mean;
                    mean = mean + delta * (nb / nn);
                    m2 = (m2 + r.y) + delta * delta * (n * nb / nn);
                    n = nn;
                }
            }
        }
        const float sd = sqrtf(__fdiv_rn(m2, Tm1));  // torch.std(x, dim=1)**2 (:419)
        const float var = __fmul_rn(sd, sd);
        const float std_in_mu = sqrtf(__fdiv_rn(var, Tf));
        const float std_in_var = sqrtf(__fdiv_rn(__fmul_rn(2.0f, __fmul_rn(var, var)), Tm1));
        const float mu_s = __fadd_rn(__fmul_rn(eS[p * S2 + col], std_in_mu), mean);
        const float var_s = __fadd_rn(__fmul_rn(eS[p * S2 + L + col], std_in_var), var);
        float s_mu = mu_s;
        float s_sd = sqrtf(__fadd_rn(fabsf(var_s), 1e-5f));
        if (p < n_valid) {
            if (summary_u) {
                summary_u[(n0 + p) * S2 + col] = s_mu;
                summary_u[(n0 + p) * S2 + L + col] = s_sd;
            }
            if (eps_sum_u) {
                const float e0 = __ldg(eps_sum_u + (n0 + p) * S2 + col);
                const float e1 = __ldg(eps_sum_u + (n0 + p) * S2 + L + col);
                s_mu = __fadd_rn(s_mu, __fmul_rn(e0, expf(__fdiv_rn(thp[pl.lv_sum + col], 2.0f))));
                s_sd = __fadd_rn(s_sd, __fmul_rn(e1, expf(__fdiv_rn(thp[pl.lv_sum + L + col], 2.0f))));
            }
        }
        sA[p * 41 + col] = s_mu;
        sA[p * 41 + L + col] = s_sd;
    }
    __syncwarp();
    head_layer<false>(sA, thp + pl.V0p, thp + pl.c0p, S2, p, q, sB);
    __syncwarp();
    head_layer<false>(sB, thp + pl.V1p, thp + pl.c1p, H, p, q, sA);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = q * 10 + i;
        const float r = sA[p * 41 + k];
        o0 = fmaf(r, thp[pl.V2 + k], o0);
        o1 = fmaf(r, thp[pl.V2 + H + k], o1);
    }
    o0 += __shfl_xor_sync(0xffffffffu, o0, 1);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 1);
    o0 += __shfl_xor_sync(0xffffffffu, o0, 2);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
    if (q == 0 && p < n_valid) {
        o0 += thp[pl.c2];
        o1 += thp[pl.c2 + 1];
        float2 o = make_float2(soft_clamp_dev(o0, hc.lo_mu, hc.hi_mu), soft_clamp_dev(o1, hc.lo_sd, hc.hi_sd));
        *reinterpret_cast<float2*>(out_unit + (n0 + p) * out_sys_stride) = o;
    }
    __syncwarp();
}

}  // namespace tc
}  // namespace bnn
