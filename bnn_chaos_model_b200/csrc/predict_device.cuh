// Device building blocks of the fused MultiSWAG predictive kernel (K2).
//
// Reference path per (weight sample, system) -- /root/reference/spock_reg_model.py:
//   zero_* masks (:452-478) -> feature_nn per time step (:301-321,:359,:417)
//   -> mean / unbiased variance over time + sampled summary statistics (:418-435)
//   -> regress_nn + soft_clamp (:437-442, :295-296).
//
// Work decomposition (one CTA = SYS_TILE systems resident in shared memory, transposed
// xT[k][row] with only the live input columns):
//   * a TASK is 32 time-step rows x all 40 hidden units for ONE unit (weight sample):
//     lane = p*4 + q;  p in 0..7 owns a quad of 4 consecutive rows, q in 0..3 owns 10 of the
//     40 output columns (5 of the 20 latent columns in the last layer).
//     With T = 100 a system is 25 quads = 3 full tasks (8 quads) + 1 left-over quad; the
//     left-over quads of the 8 systems of the tile form one more task -> 25 equal tasks.
//   * layer k-loop per lane: 1 LDS.128 of x (4 rows), 3 LDS of weights (10 columns),
//     20 fma.rn.f32x2 (FFMA2) -> 40 FMA-pipe cycles against 28 issue slots.
//   * activations between layers go through a warp-private transposed buffer hT[col][32]
//     (only __syncwarp, never a CTA barrier).
//   * pooling: each quad's (mean, M2) is merged with Chan's pairwise update, first across
//     the 8 quad-lanes of a task by shuffles, then across a system's records in the tail.
#pragma once
#include "common.cuh"

namespace bnn {

constexpr int TASK_ROWS = 32;
constexpr int HT_FLOATS = H * TASK_ROWS;  // warp-private activation buffer (5 KB)

struct TileGeom {
    int T;        // time steps per system
    int RP;       // row pitch of xT (floats)
    int QF;       // full tasks per system  = (T/4) / 8
    int RM;       // remainder quads per system = (T/4) % 8  (-> RM extra tasks per tile)
    int n_tasks;  // SYS_TILE*QF + RM
    int n_rec;    // records per system = QF + RM
    __host__ __device__ explicit TileGeom(int T_) : T(T_) {
        RP = SYS_TILE * T + 4;
        QF = (T / 4) / 8;
        RM = (T / 4) % 8;
        n_tasks = SYS_TILE * QF + RM;
        n_rec = QF + RM;
    }
};

// rec[sys][r][col][2] floats per unit slot
__host__ __device__ inline int rec_floats(const TileGeom& g) { return SYS_TILE * g.n_rec * L * 2; }

// ---------------------------------------------------------------------------------------
// One dense 40-wide layer for a 32-row task.  src: [nk][src_pitch] floats (k-major), lane's
// 4 rows start at src + roff.  w: [nk][HP] shared, bias [HP].  Result: acc[r][jp] holds
// columns (q*10 + 2jp, +1) of row r, bias added, no activation.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void dense40(const float* __restrict__ src, int src_pitch, int roff, int nk,
                                        const float* __restrict__ w, const float* __restrict__ bias, int q,
                                        u64 (&acc)[4][5]) {
    const float* bq = bias + q * GC;
#pragma unroll
    for (int jp = 0; jp < 5; ++jp) {
        const u64 b = *reinterpret_cast<const u64*>(bq + 2 * jp);
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][jp] = b;
    }
    const float* xp = src + roff;
    const float* wp = w + q * GC;
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + k * src_pitch);
        const ulonglong2 w01 = *reinterpret_cast<const ulonglong2*>(wp + k * HP);
        const ulonglong2 w23 = *reinterpret_cast<const ulonglong2*>(wp + k * HP + 4);
        const u64 w4 = *reinterpret_cast<const u64*>(wp + k * HP + 8);
        const u64 wv[5] = {w01.x, w01.y, w23.x, w23.y, w4};
        const u64 xd[4] = {pack2(xv.x, xv.x), pack2(xv.y, xv.y), pack2(xv.z, xv.z), pack2(xv.w, xv.w)};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int jp = 0; jp < 5; ++jp) acc[r][jp] = fma2(xd[r], wv[jp], acc[r][jp]);
    }
}

// torch.relu propagates NaN; CUDA's fmaxf returns the non-NaN operand, so use max.NaN.f32.
__device__ __forceinline__ float relu_nan(float x) {
    float r;
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
    return r;
}

// ReLU + transposed store of a lane's 4x10 block into the warp buffer hT[col][32].
__device__ __forceinline__ void relu_store(const u64 (&acc)[4][5], float* __restrict__ hT, int p, int q) {
#pragma unroll
    for (int jp = 0; jp < 5; ++jp) {
        float lo[4], hi[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) unpack2(acc[r][jp], lo[r], hi[r]);
        float4 v0 = make_float4(relu_nan(lo[0]), relu_nan(lo[1]), relu_nan(lo[2]), relu_nan(lo[3]));
        float4 v1 = make_float4(relu_nan(hi[0]), relu_nan(hi[1]), relu_nan(hi[2]), relu_nan(hi[3]));
        const int col = q * 10 + 2 * jp;
        *reinterpret_cast<float4*>(hT + col * TASK_ROWS + p * 4) = v0;
        *reinterpret_cast<float4*>(hT + (col + 1) * TASK_ROWS + p * 4) = v1;
    }
}

// Last feature layer (40 -> 20): lane owns 5 latent columns x 4 rows; accumulators are row
// pairs so x pairs come straight out of the LDS.128; W2p / b2p are stored duplicated.
__device__ __forceinline__ void dense20(const float* __restrict__ hT, int p, const float* __restrict__ w,
                                        const float* __restrict__ bias, int q, u64 (&a3)[5][2]) {
    const float* bq = bias + q * GC;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const u64 b = *reinterpret_cast<const u64*>(bq + 2 * i);
        a3[i][0] = b;
        a3[i][1] = b;
    }
    const float* xp = hT + p * 4;
    const float* wp = w + q * GC;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
        const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(xp + k * TASK_ROWS);
        const ulonglong2 w01 = *reinterpret_cast<const ulonglong2*>(wp + k * HP);
        const ulonglong2 w23 = *reinterpret_cast<const ulonglong2*>(wp + k * HP + 4);
        const u64 w4 = *reinterpret_cast<const u64*>(wp + k * HP + 8);
        const u64 wv[5] = {w01.x, w01.y, w23.x, w23.y, w4};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            a3[i][0] = fma2(xv.x, wv[i], a3[i][0]);
            a3[i][1] = fma2(xv.y, wv[i], a3[i][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// One task: feature_nn on 32 rows, then per-quad (mean, M2) for the lane's 5 latent columns.
//   xT/geom: CTA tile; wsm: this unit's feature weights (PackedLayout order) in shared.
//   hT: warp-private buffer.  task index t -> rows as described in the header comment.
// Writes records into rec (this unit's slot): rec[sys][r][col][2].
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mlp_task(const float* __restrict__ xT, const TileGeom& g, int kin,
                                         const float* __restrict__ wsm, const PackedLayout& pl,
                                         float* __restrict__ hT, int t, float* __restrict__ rec) {
    const int lane = threadIdx.x & 31;
    const int p = lane >> 2, q = lane & 3;
    int sys, r_idx, quad;
    const bool full = t < SYS_TILE * g.QF;
    if (full) {
        sys = t / g.QF;
        r_idx = t % g.QF;
        quad = r_idx * 8 + p;
    } else {
        sys = p;
        r_idx = g.QF + (t - SYS_TILE * g.QF);
        quad = g.QF * 8 + (t - SYS_TILE * g.QF);
    }
    const int roff = sys * g.T + quad * 4;

    u64 acc[4][5];
    dense40(xT, g.RP, roff, kin, wsm + pl.W0p, wsm + pl.b0p, q, acc);
    relu_store(acc, hT, p, q);
    __syncwarp();
    dense40(hT, TASK_ROWS, p * 4, H, wsm + pl.W1p, wsm + pl.b1p, q, acc);
    __syncwarp();  // every lane is done reading h1 before it is overwritten
    relu_store(acc, hT, p, q);
    __syncwarp();
    u64 a3[5][2];
    dense20(hT, p, wsm + pl.W2p, wsm + pl.b2p, q, a3);
    __syncwarp();  // hT may be rewritten by this warp's next task

#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float f0, f1, f2, f3;
        unpack2(a3[i][0], f0, f1);
        unpack2(a3[i][1], f2, f3);
        float mean = ((f0 + f1) + (f2 + f3)) * 0.25f;
        float d0 = f0 - mean, d1 = f1 - mean, d2 = f2 - mean, d3 = f3 - mean;
        float m2 = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        if (full) {
            // Chan pairwise merge over the 8 quad-lanes (equal counts n = 4, 8, 16)
#pragma unroll
            for (int lvl = 0; lvl < 3; ++lvl) {
                const float om = __shfl_xor_sync(0xffffffffu, mean, 4 << lvl);
                const float o2 = __shfl_xor_sync(0xffffffffu, m2, 4 << lvl);
                const float delta = om - mean;
                m2 = (m2 + o2) + delta * delta * (float)(2 << lvl);  // n_a*n_b/(n_a+n_b) = n/2
                mean = 0.5f * (mean + om);
            }
        }
        if (!full || p == 0) {
            float* rp = rec + ((sys * g.n_rec + r_idx) * L + (q * 5 + i)) * 2;
            rp[0] = mean;
            rp[1] = m2;
        }
    }
}

// ---------------------------------------------------------------------------------------
// v2 task: same arithmetic as mlp_task, but the activations of a layer are handed to the next
// one a QUARTER at a time (the 10 columns owned by lane group j), through two ping-pong
// buffers hq[2][10][32] (2.5 KB per warp instead of 5 KB), so that 12+ consumer warps fit
// next to the resident x tile.  The previous layer's accumulators stay in registers while the
// next layer accumulates.
// ---------------------------------------------------------------------------------------
constexpr int HQ_FLOATS = 2 * 10 * TASK_ROWS;  // per consumer warp

__device__ __forceinline__ void store_quarter(const u64 (&acc)[4][5], float* __restrict__ hq, int p) {
#pragma unroll
    for (int jp = 0; jp < 5; ++jp) {
        float lo[4], hi[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) unpack2(acc[r][jp], lo[r], hi[r]);
        *reinterpret_cast<float4*>(hq + (2 * jp) * TASK_ROWS + p * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(hq + (2 * jp + 1) * TASK_ROWS + p * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    }
}

__device__ __forceinline__ void relu_inplace(u64 (&acc)[4][5]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int jp = 0; jp < 5; ++jp) {
            float lo, hi;
            unpack2(acc[r][jp], lo, hi);
            acc[r][jp] = pack2(relu_nan(lo), relu_nan(hi));
        }
}

__device__ __forceinline__ void mlp_task_v2(const float* __restrict__ xT, const TileGeom& g, int kin,
                                            const float* __restrict__ wsm, const PackedLayout& pl,
                                            float* __restrict__ hq, int t, float* __restrict__ rec) {
    const int lane = threadIdx.x & 31;
    const int p = lane >> 2, q = lane & 3;
    int sys, r_idx, quad;
    const bool full = t < SYS_TILE * g.QF;
    if (full) {
        sys = t / g.QF;
        r_idx = t % g.QF;
        quad = r_idx * 8 + p;
    } else {
        sys = p;
        r_idx = g.QF + (t - SYS_TILE * g.QF);
        quad = g.QF * 8 + (t - SYS_TILE * g.QF);
    }
    const int roff = sys * g.T + quad * 4;

    u64 a1[4][5], a2[4][5];
    dense40(xT, g.RP, roff, kin, wsm + pl.W0p, wsm + pl.b0p, q, a1);
    relu_inplace(a1);

    // layer 2: a2 = b1 + h1 W1^T, h1 streamed through hq one lane group at a time
    {
        const float* bq = wsm + pl.b1p + q * GC;
#pragma unroll
        for (int jp = 0; jp < 5; ++jp) {
            const u64 b = *reinterpret_cast<const u64*>(bq + 2 * jp);
#pragma unroll
            for (int r = 0; r < 4; ++r) a2[r][jp] = b;
        }
        const float* wq = wsm + pl.W1p + q * GC;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float* hb = hq + (j & 1) * (10 * TASK_ROWS);
            if (q == j) store_quarter(a1, hb, p);
            __syncwarp();
            const float* xp = hb + p * 4;
            const float* wp = wq + (j * 10) * HP;
#pragma unroll
            for (int kk = 0; kk < 10; ++kk) {
                const float4 xv = *reinterpret_cast<const float4*>(xp + kk * TASK_ROWS);
                const ulonglong2 w01 = *reinterpret_cast<const ulonglong2*>(wp + kk * HP);
                const ulonglong2 w23 = *reinterpret_cast<const ulonglong2*>(wp + kk * HP + 4);
                const u64 w4 = *reinterpret_cast<const u64*>(wp + kk * HP + 8);
                const u64 wv[5] = {w01.x, w01.y, w23.x, w23.y, w4};
                const u64 xd[4] = {pack2(xv.x, xv.x), pack2(xv.y, xv.y), pack2(xv.z, xv.z), pack2(xv.w, xv.w)};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int jp = 0; jp < 5; ++jp) a2[r][jp] = fma2(xd[r], wv[jp], a2[r][jp]);
            }
        }
    }
    relu_inplace(a2);

    // layer 3: 40 -> 20, row-pair accumulators, duplicated W2p
    u64 a3[5][2];
    {
        const float* bq = wsm + pl.b2p + q * GC;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const u64 b = *reinterpret_cast<const u64*>(bq + 2 * i);
            a3[i][0] = b;
            a3[i][1] = b;
        }
        const float* wq = wsm + pl.W2p + q * GC;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float* hb = hq + (j & 1) * (10 * TASK_ROWS);
            // ping-pong: buffer (j&1) was last read in phase j-2, and every lane passed the
            // __syncwarp of phase j-1 since; for j<2 the layer-2 reads are covered the same way
            if (q == j) store_quarter(a2, hb, p);
            __syncwarp();
            const float* xp = hb + p * 4;
            const float* wp = wq + (j * 10) * HP;
#pragma unroll
            for (int kk = 0; kk < 10; ++kk) {
                const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(xp + kk * TASK_ROWS);
                const ulonglong2 w01 = *reinterpret_cast<const ulonglong2*>(wp + kk * HP);
                const ulonglong2 w23 = *reinterpret_cast<const ulonglong2*>(wp + kk * HP + 4);
                const u64 w4 = *reinterpret_cast<const u64*>(wp + kk * HP + 8);
                const u64 wv[5] = {w01.x, w01.y, w23.x, w23.y, w4};
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    a3[i][0] = fma2(xv.x, wv[i], a3[i][0]);
                    a3[i][1] = fma2(xv.y, wv[i], a3[i][1]);
                }
            }
        }
    }
    __syncwarp();  // all reads of hq done before this warp's next task writes it

#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float f0, f1, f2, f3;
        unpack2(a3[i][0], f0, f1);
        unpack2(a3[i][1], f2, f3);
        float mean = ((f0 + f1) + (f2 + f3)) * 0.25f;
        float d0 = f0 - mean, d1 = f1 - mean, d2 = f2 - mean, d3 = f3 - mean;
        float m2 = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        if (full) {
#pragma unroll
            for (int lvl = 0; lvl < 3; ++lvl) {
                const float om = __shfl_xor_sync(0xffffffffu, mean, 4 << lvl);
                const float o2 = __shfl_xor_sync(0xffffffffu, m2, 4 << lvl);
                const float delta = om - mean;
                m2 = (m2 + o2) + delta * delta * (float)(2 << lvl);
                mean = 0.5f * (mean + om);
            }
        }
        if (!full || p == 0) {
            float* rp = rec + ((sys * g.n_rec + r_idx) * L + (q * 5 + i)) * 2;
            rp[0] = mean;
            rp[1] = m2;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Tail of one unit for the tile's 8 systems, executed by ONE warp: merge records, sampled
// summary statistics (:422-432), regress_nn (:437-438), soft_clamp (:440-441).
//   lane = p*4+q: p = system, q = group of 5 latent / 10 hidden columns.
//   thp: this unit's packed weights in GLOBAL memory (head part is read through L2).
//   eps_u: explicit draws for this unit [N][2L] (global) or nullptr -> Philox.
//   eps_sum_u: summary-noise draws [N][2L] (noisy forward, :448-450) or nullptr (none).
//   summary_u: optional output [N][2L] of the summary statistics (compute_summary_stats).
//   scratch: >= 8*41*2 + 8*40 floats of warp-private shared memory.
// ---------------------------------------------------------------------------------------
struct HeadConsts {
    float lo_mu, hi_mu, lo_sd, hi_sd;
};

__device__ __forceinline__ float soft_clamp_dev(float x, float lo, float hi) {
    // 0.5*(tanh(x)+1)*(high-lo) + lo, evaluated left to right like the reference (:295-296)
    return __fadd_rn(__fmul_rn(__fmul_rn(0.5f, __fadd_rn(tanhf(x), 1.0f)), __fsub_rn(hi, lo)), lo);
}

template <bool GLOBAL_W>
__device__ __forceinline__ float4 ldw4(const float* p) {
    if (GLOBAL_W) return __ldg(reinterpret_cast<const float4*>(p));
    return *reinterpret_cast<const float4*>(p);
}
template <bool GLOBAL_W>
__device__ __forceinline__ float2 ldw2(const float* p) {
    if (GLOBAL_W) return __ldg(reinterpret_cast<const float2*>(p));
    return *reinterpret_cast<const float2*>(p);
}
template <bool GLOBAL_W>
__device__ __forceinline__ float ldw1(const float* p) {
    if (GLOBAL_W) return __ldg(p);
    return *p;
}

template <bool GLOBAL_W>
__device__ __forceinline__ void head_layer(const float* __restrict__ sin_, const float* __restrict__ wg,
                                           const float* __restrict__ bg, int nk, int p, int q,
                                           float* __restrict__ sout) {
    float acc[10];
    const float* bq = bg + q * GC;
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = ldw1<GLOBAL_W>(bq + i);
    const float* wq = wg + q * GC;
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
        const float sv = sin_[p * 41 + k];
        const float4 a = ldw4<GLOBAL_W>(wq + k * HP);
        const float4 b = ldw4<GLOBAL_W>(wq + k * HP + 4);
        const float2 c = ldw2<GLOBAL_W>(wq + k * HP + 8);
        acc[0] = fmaf(sv, a.x, acc[0]); acc[1] = fmaf(sv, a.y, acc[1]);
        acc[2] = fmaf(sv, a.z, acc[2]); acc[3] = fmaf(sv, a.w, acc[3]);
        acc[4] = fmaf(sv, b.x, acc[4]); acc[5] = fmaf(sv, b.y, acc[5]);
        acc[6] = fmaf(sv, b.z, acc[6]); acc[7] = fmaf(sv, b.w, acc[7]);
        acc[8] = fmaf(sv, c.x, acc[8]); acc[9] = fmaf(sv, c.y, acc[9]);
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) sout[p * 41 + q * 10 + i] = relu_nan(acc[i]);
}

template <bool GLOBAL_W>
__device__ __forceinline__ void tail_unit(const float* __restrict__ rec, const TileGeom& g,
                                          const float* __restrict__ thp, const PackedLayout& pl,
                                          const float* __restrict__ eps_u, const float* __restrict__ eps_sum_u,
                                          float* __restrict__ summary_u, uint64_t seed, uint32_t gunit,
                                          int64_t gsys0, int64_t n0, int n_valid, const HeadConsts& hc,
                                          float* __restrict__ scratch, float* __restrict__ out_unit,
                                          int64_t out_sys_stride) {
    const int lane = threadIdx.x & 31;
    const int p = lane >> 2, q = lane & 3;
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

    const float Tf = (float)g.T, Tm1 = (float)(g.T - 1);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int col = q * 5 + i;
        const float* rp = rec + ((p * g.n_rec) * L + col) * 2;
        float n = (g.QF > 0) ? 32.f : 4.f;
        float mean = rp[0], m2 = rp[1];
        for (int r = 1; r < g.n_rec; ++r) {
            const float nb = (r < g.QF) ? 32.f : 4.f;
            const float mb = rp[r * L * 2], m2b = rp[r * L * 2 + 1];
            const float nn = n + nb, delta = mb - mean;
            mean = mean + delta * (nb / nn);
            m2 = (m2 + m2b) + delta * delta * (n * nb / nn);
            n = nn;
        }
        // sample_var = torch.std(x, dim=1)**2 (:419): unbiased, then sqrt and square
        const float sd = sqrtf(__fdiv_rn(m2, Tm1));
        const float var = __fmul_rn(sd, sd);
        const float std_in_mu = sqrtf(__fdiv_rn(var, Tf));                                   // :422
        const float std_in_var = sqrtf(__fdiv_rn(__fmul_rn(2.0f, __fmul_rn(var, var)), Tm1));  // :423
        const float mu_s = __fadd_rn(__fmul_rn(eS[p * S2 + col], std_in_mu), mean);           // :426
        const float var_s = __fadd_rn(__fmul_rn(eS[p * S2 + L + col], std_in_var), var);      // :427
        float s_mu = mu_s;
        float s_sd = sqrtf(__fadd_rn(fabsf(var_s), 1e-5f));                                   // :430
        if (p < n_valid) {
            if (summary_u) {  // summary statistics before the summary noise (what _summary_kl sees, :515)
                summary_u[(n0 + p) * S2 + col] = s_mu;
                summary_u[(n0 + p) * S2 + L + col] = s_sd;
            }
            if (eps_sum_u) {  // add_summary_noise (:448-450): s + eps * exp(logvar/2)
                const float e0 = __ldg(eps_sum_u + (n0 + p) * S2 + col);
                const float e1 = __ldg(eps_sum_u + (n0 + p) * S2 + L + col);
                s_mu = __fadd_rn(s_mu, __fmul_rn(e0, expf(__fdiv_rn(ldw1<GLOBAL_W>(thp + pl.lv_sum + col), 2.0f))));
                s_sd = __fadd_rn(s_sd, __fmul_rn(e1, expf(__fdiv_rn(ldw1<GLOBAL_W>(thp + pl.lv_sum + L + col), 2.0f))));
            }
        }
        sA[p * 41 + col] = s_mu;
        sA[p * 41 + L + col] = s_sd;
    }
    __syncwarp();
    head_layer<GLOBAL_W>(sA, thp + pl.V0p, thp + pl.c0p, S2, p, q, sB);
    __syncwarp();
    head_layer<GLOBAL_W>(sB, thp + pl.V1p, thp + pl.c1p, H, p, q, sA);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = q * 10 + i;
        const float r = sA[p * 41 + k];
        o0 = fmaf(r, ldw1<GLOBAL_W>(thp + pl.V2 + k), o0);
        o1 = fmaf(r, ldw1<GLOBAL_W>(thp + pl.V2 + H + k), o1);
    }
    o0 += __shfl_xor_sync(0xffffffffu, o0, 1);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 1);
    o0 += __shfl_xor_sync(0xffffffffu, o0, 2);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
    if (q == 0 && p < n_valid) {
        o0 += ldw1<GLOBAL_W>(thp + pl.c2);
        o1 += ldw1<GLOBAL_W>(thp + pl.c2 + 1);
        float2 o = make_float2(soft_clamp_dev(o0, hc.lo_mu, hc.hi_mu), soft_clamp_dev(o1, hc.lo_sd, hc.hi_sd));
        *reinterpret_cast<float2*>(out_unit + (n0 + p) * out_sys_stride) = o;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// Load SYS_TILE systems of X[N,T,F] into xT[k][row] (live columns only, transposed).
// A non-finite value in a zeroed column poisons the row (the reference's x - mask keeps
// NaN/Inf as NaN, :452-478).  `poison` is RP ints of shared memory.
// ---------------------------------------------------------------------------------------
struct ColMap {
    int8_t inv[MAXF];  // column -> live index or -1
};

__device__ __forceinline__ void load_x_tile(const float* __restrict__ X, int64_t n0, int n_valid, int F,
                                            const TileGeom& g, int kin, const ColMap& cm,
                                            float* __restrict__ xT, int* __restrict__ poison) {
    const int rows = SYS_TILE * g.T;
    for (int r = threadIdx.x; r < g.RP; r += blockDim.x) poison[r] = 0;
    const int valid_rows = n_valid * g.T;
    // zero-fill rows of missing systems and the pad
    for (int idx = threadIdx.x; idx < kin * (g.RP - valid_rows); idx += blockDim.x) {
        const int k = idx / (g.RP - valid_rows), r = valid_rows + idx % (g.RP - valid_rows);
        xT[k * g.RP + r] = 0.f;
    }
    __syncthreads();
    const float* src = X + n0 * (int64_t)g.T * F;
    const int total = valid_rows * F;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int row = idx / F, c = idx - row * F;
        const float v = __ldg(src + idx);
        const int k = cm.inv[c];
        if (k >= 0)
            xT[k * g.RP + row] = v;
        else if (!isfinite(v))
            poison[row] = 1;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += blockDim.x)
        if (poison[r]) xT[r] = __int_as_float(0x7fc00000);
    __syncthreads();
}

}  // namespace bnn
