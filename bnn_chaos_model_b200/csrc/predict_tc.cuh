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
// as fp32; per (unit, M tile) "job" one of NSLOT TMEM slots [A_hi 40 | A_lo 40 | D 48 columns]:
//   epilogue warps (4 per slot, thread = row = TMEM lane; quadrant-0 warp also issues the slot's MMAs):
//       stage x: smem -> hi/lo -> tcgen05.st A
//       after each layer: tcgen05.ld D -> + bias -> ReLU -> hi/lo -> tcgen05.st A
//       last layer: D + bias -> per-32-row-block pooled (mean, M2) records
//       hand-offs inside a slot are named hardware barriers (bar.sync: waiting warps cost no issue slots); only
//       the issuing warp polls the tcgen05.commit mbarrier
//   MMA issue (one elected lane): 12/15/15 tcgen05.mma per layer (M=128, N=48/48/32, K=8)
//   producer warp: cp.async.bulk of each unit's hi/lo B operands + biases (38 KB) into a 2-slot ring
//   NT tail warps (unit i -> warp i % NT): merge the records per system, sampled summary statistics,
//       regress_nn with the fp32 head weights read through L2, store (mu, std); records live in an NT-deep ring
//       (unit_done / rec_free barriers give the epilogue back-pressure when the tails fall behind).
#pragma once
#include <cuda_fp16.h>

#include "predict_device.cuh"
#include "tc.cuh"

namespace bnn {
namespace tc {

constexpr int SYS = 5;            // systems per CTA tile
constexpr int MT = 4;             // 128-row M tiles per tile (5 * 100 rows -> 512)
constexpr int T_FIXED = 100;      // time steps (the tiling is specific to T = 100)
constexpr int ROWS = SYS * T_FIXED;
constexpr int XS_ROWS = MT * 128;   // rows of the shared-memory x tile: the 12 rows behind the tile stay zero (no row test in the stage)
constexpr int TM_AHI = 0, TM_ALO = 40, TM_D = 80, TM_SLOT = 128;  // TMEM columns of one slot
constexpr int N_BLOCKS = MT * 4;  // 32-row blocks per tile
constexpr int REC_FLOATS = N_BLOCKS * 2 * L * 2;  // [block][segment][col][mean, M2]
constexpr int FB_PITCH = 36;                     // latent staging for the pooling, per epilogue warp: 10 columns x (32 rows + 4 pad) -- column-major, so that
constexpr int FB_FLOATS = 10 * FB_PITCH;         // a 4-row granule of a column is one 16-byte load (pitch 36: the 8 lanes of a quarter warp hit 8 bank groups)
constexpr int TAIL_SCRATCH = 4 * 208;            // per tail pair: eS[5*40] | sum[5*41] | sB[5*41] | sC[5*41] (208-float areas)
constexpr int HEAD_FLOATS = S2 * HP + HP + H * HP + HP + 2 * H + 4 + S2;  // PackedLayout V0p .. lv_sum: one contiguous block
constexpr int NREC = 4;                          // depth of the block-record ring (units the epilogue may run ahead of the tails; 2 in the wide variant)
constexpr int N_SLOT = 4;                        // TMEM slots = jobs in flight (2 per team)

struct Bars {
    uint64_t w_full[2];                                // B operands of unit i landed in ring slot i & 1
    uint64_t unit_done[NREC], rec_free[NREC];          // record ring slot i % NREC: written by the epilogue / read by the tail
    uint64_t d_ready[N_SLOT];                          // tcgen05.commit of the slot's current layer
    uint64_t a_ready[N_SLOT];                          // the team's 8 epilogue warps: A of the slot's next layer is in place
    uint64_t h_full[2];                                // head warp j: the head block of its next unit landed in its buffer
    uint64_t sum_full[2], sum_free[2];                 // tail pair j: summary statistics written by the stats warp / consumed
    uint32_t tmem_base;
    int next_item;                                     // unused since the static partition (kept: the layout of Bars is measured)
    unsigned int x_maxbits[SYS];                       // fp16 x tile: bits of the largest finite |x| of each system
    float x_scale[SYS];                                // ... and the power of two that undoes the system's down-scaling
};

// The 4 records of system p (T = 100, 5 systems per tile, 32-row blocks): record index (block*2 + segment) and
// row count.  System p starts at row 100p: a leading partial block (segment 1 of block (100p)/32 when 100p % 32
// != 0), full blocks, and a trailing partial block (segment 0).
struct SysRec { int idx[4]; float cnt[4]; };
__device__ __forceinline__ SysRec sys_records(int p) {
    SysRec r;
    const int first = T_FIXED * p, last = T_FIXED * p + T_FIXED - 1;
    const int b0 = first >> 5, b1 = last >> 5;  // 4 blocks for every p in 0..4
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = b0 + k;
        const int lo = max(first, 32 * b), hi = min(last, 32 * b + 31);
        const int seg = ((32 * b) / T_FIXED == p) ? 0 : 1;
        r.idx[k] = b * 2 + seg;
        r.cnt[k] = (b <= b1) ? (float)(hi - lo + 1) : 0.f;
    }
    return r;
}

// Pooling geometry of the 16 32-row blocks, fixed by T = 100 and 5 systems per tile: rows [0, e0) of block b belong to
// one system, rows [e0, nvalid) (if any) to the next; n0 / n1 the two row counts as floats with their correctly rounded
// reciprocals (mean = S / n as S * r corrected by one FMA pair).  In constant memory: the epilogue warps recomputed it
// per (unit, block) with integer divisions -- 40 of the ~190 instructions of a pooling call.
struct PoolGeom { int e0, nvalid; float n0, n1, rcp0, rcp1; int pad0, pad1; };
constexpr PoolGeom pool_geom(int b) {
    const int R0 = 32 * b, sysA = R0 / T_FIXED;
    const int split = (T_FIXED * (sysA + 1) - R0) < 32 ? (T_FIXED * (sysA + 1) - R0) : 32;
    const int nvalid = (ROWS - R0) < 32 ? (ROWS - R0) : 32;
    const int e0 = split < nvalid ? split : nvalid;
    const int n1 = nvalid > split ? nvalid - split : 1;
    return PoolGeom{e0, nvalid, (float)e0, (float)n1, 1.0f / (float)e0, 1.0f / (float)n1, 0, 0};
}
__constant__ PoolGeom c_pool_geom[N_BLOCKS] = {
    pool_geom(0), pool_geom(1), pool_geom(2), pool_geom(3), pool_geom(4), pool_geom(5), pool_geom(6), pool_geom(7),
    pool_geom(8), pool_geom(9), pool_geom(10), pool_geom(11), pool_geom(12), pool_geom(13), pool_geom(14), pool_geom(15)};
static_assert(N_BLOCKS == 16, "c_pool_geom lists 16 blocks");

// block b covers tile rows [32b, 32b+32): rows of system sysA up to `split`, then system sysA+1
__device__ __forceinline__ void block_geom(int b, int& sysA, int& split, int& nvalid) {
    const int R0 = 32 * b;
    sysA = R0 / T_FIXED;
    split = min(32, T_FIXED * (sysA + 1) - R0);
    nvalid = min(32, ROWS - R0);
}

// ---------------------------------------------------------------------------------------
// x tile: X[N,T,F] -> xs[row][32] fp32, 16-byte chunks XOR-swizzled by (row & 7) so that a warp's
// 32 rows read conflict-free; live columns packed first, the rest 0.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int xs_index(int row, int c) { return row * 32 + ((((c >> 2) ^ (row & 7))) << 2) + (c & 3); }

__device__ __forceinline__ void load_x_tile_tc(const float* __restrict__ X, int64_t n0, int n_valid, int F, int kin,
                                               const ColMap& cm, float* __restrict__ xs, int* __restrict__ poison) {
    for (int i = threadIdx.x; i < XS_ROWS * 32; i += blockDim.x) xs[i] = 0.f;
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) poison[r] = 0;
    __syncthreads();
    // one warp per row, lane = column (columns lane and lane + 32): coalesced reads, the column map looked up once per lane
    // (see load_x_tile_f16); live columns 32.. (wide variant) are read from L2 by the stage
    const float* src = X + n0 * (int64_t)T_FIXED * F;
    const int rows = n_valid * T_FIXED;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int c1 = lane + 32;
    const int k0 = lane < F ? (int)cm.inv[lane] : -2;
    const int k1 = c1 < F ? (int)cm.inv[c1] : -2;
#pragma unroll 4
    for (int row = warp; row < rows; row += n_warps) {
        const float* p = src + (int64_t)row * F;
        const float v0 = k0 != -2 ? __ldg(p + lane) : 0.f;
        const float v1 = k1 != -2 ? __ldg(p + c1) : 0.f;
        if (k0 >= 0 && k0 < 32) xs[xs_index(row, k0)] = v0;
        if (k1 >= 0 && k1 < 32) xs[xs_index(row, k1)] = v1;
        if ((k0 == -1 && !isfinite(v0)) || (k1 == -1 && !isfinite(v1))) poison[row] = 1;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
        if (poison[r]) xs[xs_index(r, 0)] = __int_as_float(0x7fc00000);  // x - mask keeps NaN/Inf as NaN (:452-478)
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// x tile for the fp16 layer 1 (at most 32 live inputs): the tile is split ONCE into fp16 hi = RN(x), lo = RN(x - hi)
// -- 11 + 11 significant bits, the same as the tf32 split -- and written as the A operand of tcgen05.mma kind::f16 in
// the canonical K-major no-swizzle layout [M tile][k / 8][128 rows][8 halves] (hi at byte 0, lo at byte 32768 of the
// x area), which the tensor core reads straight from shared memory: no per-unit staging through registers and tensor
// memory (16 values split + two tcgen05.st per thread, unit and slot before), and K = 16 per instruction.
// fp16 range: lo below 2^-14 loses bits to gradual underflow (absolute error <= 2^-25 per input, ~1e-8 of a
// pre-activation); a SYSTEM whose largest finite |x| reaches 2^15 is scaled down by a power of two and the layer-1
// epilogue scales its rows of the accumulator back (x_scale[system]; 1 for in-distribution inputs), so large inputs stay
// finite like in the fp32 reference and cost precision only in their own system.  NaN / Inf propagate (Inf: hi = Inf,
// lo = NaN).
// ---------------------------------------------------------------------------------------
constexpr int XH_BYTES = XS_ROWS * 32 * 2;   // bytes of the hi (and of the lo) array
__device__ __forceinline__ int x16_offset(int row, int k) {   // byte offset of (row, k) inside the hi or lo array
    return ((((row >> 7) * 4 + (k >> 3)) * 128 + (row & 127)) << 4) + ((k & 7) << 1);
}
// One warp per time-step row, lane = input column (columns lane and lane + 32): coalesced row reads, the column -> live
// index map is looked up once per lane.  (Element-indexed loops with a per-element lookup of the map in constant memory --
// 32 different addresses per warp load, serialised -- made the tile load 8 % of the kernel's time: ncu source view.)
__device__ __forceinline__ void load_x_tile_f16(const float* __restrict__ X, int64_t n0, int n_valid, int F, int kin,
                                                const ColMap& cm, unsigned char* __restrict__ xs16,
                                                int* __restrict__ poison, unsigned int* __restrict__ maxbits,
                                                float* __restrict__ scale_out) {
    for (int i = threadIdx.x; i < 2 * XH_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(xs16)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) poison[r] = 0;
    if (threadIdx.x < SYS) maxbits[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int c1 = lane + 32;
    // role of this lane's two columns: live index k >= 0 (staged when k < 32), -1 = zeroed by the model's flags, -2 = none
    const int k0 = lane < F ? (int)cm.inv[lane] : -2;
    const int k1 = c1 < F ? (int)cm.inv[c1] : -2;
    const bool live0 = k0 >= 0 && k0 < 32, live1 = k1 >= 0 && k1 < 32;
    const float* src = X + n0 * (int64_t)T_FIXED * F;
    const int rows = n_valid * T_FIXED;
    // pass 1 (L2 hits: the tile was bulk-prefetched while the previous item ran): largest finite |x| of the live columns,
    // non-finite values in zeroed columns
#pragma unroll 4
    for (int row = warp; row < rows; row += n_warps) {
        const float* p = src + (int64_t)row * F;
        const float v0 = k0 != -2 ? __ldg(p + lane) : 0.f;
        const float v1 = k1 != -2 ? __ldg(p + c1) : 0.f;
        unsigned int mx = 0u;
        if (live0 && isfinite(v0)) mx = __float_as_uint(fabsf(v0));
        if (live1 && isfinite(v1)) mx = max(mx, __float_as_uint(fabsf(v1)));
        if ((k0 == -1 && !isfinite(v0)) || (k1 == -1 && !isfinite(v1))) poison[row] = 1;
        if (__any_sync(0xffffffffu, mx >= 0x47000000u)) {   // only |x| >= 2^15 matters (rare)
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0) atomicMax(maxbits + row / T_FIXED, mx);
        }
    }
    __syncthreads();
    // per system: |x| 2^-kk < 2^15
    auto shift_of = [&](int sys) {
        const int e = (int)(maxbits[sys] >> 23) - 127;
        return e > 14 ? e - 14 : 0;
    };
    if (threadIdx.x < SYS) scale_out[threadIdx.x] = __uint_as_float((uint32_t)(127 + shift_of(threadIdx.x)) << 23);
    unsigned char* xh = xs16;
    unsigned char* xl = xs16 + XH_BYTES;
    auto put = [&](int row, int k, float v) {
        const __half h = __float2half_rn(v);
        const int off = x16_offset(row, k);
        *reinterpret_cast<__half*>(xh + off) = h;
        *reinterpret_cast<__half*>(xl + off) = __float2half_rn(v - __half2float(h));
    };
#pragma unroll 4
    for (int row = warp; row < rows; row += n_warps) {
        const float* p = src + (int64_t)row * F;
        const int kk = shift_of(row / T_FIXED);
        const float down = __uint_as_float((uint32_t)(127 - kk) << 23);
        const float v0 = live0 ? __ldg(p + lane) * down : 0.f;
        const float v1 = live1 ? __ldg(p + c1) * down : 0.f;
        if (live0) put(row, k0, v0);
        if (live1) put(row, k1, v1);
        // a spare K column (at most 31 live inputs) carries 1.0 and W0's row there the layer-1 bias (hi + lo): the bias comes
        // out of the MMA and the layer-1 epilogue has nothing to add -- for rows whose block holds no scaled-down system
        // (decided per 32-row block = per epilogue warp, whose tcgen05.st must not diverge: no system touching the block is scaled)
        if (lane == 0 && kin < TC_K1 && shift_of((row & ~31) / T_FIXED) == 0 && shift_of(min((row | 31) / T_FIXED, SYS - 1)) == 0)
            *reinterpret_cast<unsigned short*>(xh + x16_offset(row, TC_K1 - 1)) = 0x3c00u;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
        if (poison[r]) *reinterpret_cast<unsigned short*>(xh + x16_offset(r, 0)) = 0x7e00u;  // x - mask keeps NaN/Inf as NaN (:452-478)
    }
    fence_proxy_async_smem();   // the tensor core reads the tile through the async proxy
    __syncthreads();
}

// (d + bias) -> ReLU -> hi/lo of 4 consecutive columns; bias: 16-byte aligned shared memory, same for every lane.
// hi = v rounded to the 11 significant bits of tf32, lo = v - hi exactly, on the FMA pipe with packed fp32x2 instructions:
//   c = fma(v, 8192, v) = RN(8193 v);  hi = fma(v, -8192, c) = c - 8192 v (exact: Veltkamp's splitting with the exact
//   product 8192 v in place of the rounded c - v, so every step is one FMA and no contraction can change a value);
//   lo = v - hi (exact).  Three instructions per PAIR of values.  The integer form ((bits + 0x1000) & ~0x1fff: IADD3 +
// LOP3 per value) kept the half-rate ALU pipe at 62 % of its peak -- the busiest pipe of the kernel, `math pipe throttle`
// the top stall of the epilogue.  Inf becomes NaN (Inf - Inf), as it did in lo = v - hi before.
// EPI: 0 = plain split, 1 = (d * scale + bias) -> ReLU -> split, 2 = ReLU -> split (the bias is already in the accumulator)
template <int EPI>
__device__ __forceinline__ void split_group4(const uint32_t* __restrict__ d, const float* __restrict__ bias,
                                             uint32_t* __restrict__ h, uint32_t* __restrict__ l, u64 scale2 = 0x3f8000003f800000ull) {
    u64 v01 = pack2(__uint_as_float(d[0]), __uint_as_float(d[1]));
    u64 v23 = pack2(__uint_as_float(d[2]), __uint_as_float(d[3]));
    if (EPI) {
        if (EPI == 1) {
            // d * scale + bias: scale = 1 (one rounding, the same value as d + bias) except in layer 1 of a down-scaled system
            const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(bias);
            v01 = fma2(v01, scale2, b.x);
            v23 = fma2(v23, scale2, b.y);
        }
        float v[4];
        unpack2(v01, v[0], v[1]);
        unpack2(v23, v[2], v[3]);
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = relu_nan(v[u]);
        v01 = pack2(v[0], v[1]);
        v23 = pack2(v[2], v[3]);
    }
    const u64 K = pack2(8192.0f, 8192.0f), Kn = pack2(-8192.0f, -8192.0f);
    const u64 h01 = fma2(v01, Kn, fma2(v01, K, v01)), h23 = fma2(v23, Kn, fma2(v23, K, v23));
    const u64 l01 = sub2(v01, h01), l23 = sub2(v23, h23);
    float f[4];
    unpack2(h01, f[0], f[1]);
    unpack2(h23, f[2], f[3]);
#pragma unroll
    for (int u = 0; u < 4; ++u) h[u] = __float_as_uint(f[u]);
    unpack2(l01, f[0], f[1]);
    unpack2(l23, f[2], f[3]);
#pragma unroll
    for (int u = 0; u < 4; ++u) l[u] = __float_as_uint(f[u]);
}
template <int EPI>
__device__ __forceinline__ void split_store16(const uint32_t (&d)[16], const float* __restrict__ bias, uint32_t t_hi,
                                              uint32_t t_lo, u64 scale2 = 0x3f8000003f800000ull) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) split_group4<EPI>(&d[4 * g4], bias + 4 * g4, &h[4 * g4], &l[4 * g4], scale2);
    tmem_st16(t_hi, h);
    tmem_st16(t_lo, l);
}
template <int EPI>
__device__ __forceinline__ void split_store8(const uint32_t (&d)[8], const float* __restrict__ bias, uint32_t t_hi,
                                             uint32_t t_lo) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int g4 = 0; g4 < 2; ++g4) split_group4<EPI>(&d[4 * g4], bias + 4 * g4, &h[4 * g4], &l[4 * g4]);
    tmem_st8(t_hi, h);
    tmem_st8(t_lo, l);
}
template <int EPI>
__device__ __forceinline__ void split_store4(const uint32_t (&d)[4], const float* __restrict__ bias, uint32_t t_hi,
                                             uint32_t t_lo, u64 scale2 = 0x3f8000003f800000ull) {
    uint32_t h[4], l[4];
    split_group4<EPI>(d, bias, h, l, scale2);
    tmem_st4(t_hi, h);
    tmem_st4(t_lo, l);
}

// ---------------------------------------------------------------------------------------
// MMA issue (one thread).  wb: shared-memory byte address of this unit's ring slot; offsets in floats
// relative to the slot start.  ts: TMEM address of the slot (lane 0).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t bdesc(uint32_t base_addr, int N, int ks) {
    const uint32_t chunk = (uint32_t)N * 16u;
    return smem_desc_kmajor(base_addr + (uint32_t)ks * 2u * chunk, chunk, 128u);
}

// K steps [KS0, KS1) of one layer; FIRST: the very first MMA overwrites D (no accumulate)
template <int N, int KS0, int KS1, bool FIRST>
__device__ __forceinline__ void issue_layer_part(uint32_t ts, uint32_t bh_addr, uint32_t bl_addr) {
    constexpr uint32_t idesc = idesc_tf32(128, N);
    const uint32_t d = ts + TM_D, ahi = ts + TM_AHI, alo = ts + TM_ALO;
    // descriptors advance by two 16-byte K chunks (2*N*16 bytes -> 2*N in the 16-byte address field) per K = 8 step;
    // A always starts at column 0 of the slot's A regions (a second pass restages its columns there)
    const uint64_t dh = bdesc(bh_addr, N, 0), dl = bdesc(bl_addr, N, 0);
    constexpr uint64_t step = 2ull * N;
    // called by a whole (converged) warp; one elected lane issues a_lo w_hi + a_hi w_lo + a_hi w_hi, the two
    // small (2^-11) correction terms FIRST: the tensor core rounds its fp32 accumulator toward zero after every
    // MMA, so only the last KS steps round at the full magnitude of the sum (measured: 3x smaller error than
    // accumulating the corrections after the main term, tools/accuracy_study.py)
    if (elect_one_sync()) {
#pragma unroll
        for (int ks = KS0; ks < KS1; ++ks) mma_tf32_ts(d, alo + 8 * (ks - KS0), dh + ks * step, idesc, !FIRST || ks > KS0);
#pragma unroll
        for (int ks = KS0; ks < KS1; ++ks) mma_tf32_ts(d, ahi + 8 * (ks - KS0), dl + ks * step, idesc, true);
#pragma unroll
        for (int ks = KS0; ks < KS1; ++ks) mma_tf32_ts(d, ahi + 8 * (ks - KS0), dh + ks * step, idesc, true);
    }
    __syncwarp();
}
// Layer 1 as kind::f16 with BOTH operands in shared memory: A = the tile's fp16 hi / lo rows of M tile m (ah_addr /
// al_addr: byte address of the M tile's first K chunk; chunks 2048 bytes apart), B = the unit's fp16 hi / lo W0.
// K = 32 = two K = 16 steps per term; corrections first, like issue_layer_part.
template <int N>
__device__ __forceinline__ void issue_layer1_f16(uint32_t ts, uint32_t ah_addr, uint32_t al_addr, uint32_t bh_addr,
                                                 uint32_t bl_addr) {
    constexpr uint32_t idesc = idesc_f16(128, N);
    const uint32_t d = ts + TM_D;
    const uint64_t ah = smem_desc_kmajor(ah_addr, 2048u, 128u), al = smem_desc_kmajor(al_addr, 2048u, 128u);
    const uint64_t bh = smem_desc_kmajor(bh_addr, (uint32_t)N * 16u, 128u), bl = smem_desc_kmajor(bl_addr, (uint32_t)N * 16u, 128u);
    constexpr uint64_t astep = 2ull * 2048ull / 16ull, bstep = 2ull * N;   // two 16-byte K chunks per K = 16 step
    if (elect_one_sync()) {
#pragma unroll
        for (int ks = 0; ks < TC_K1 / 16; ++ks) mma_f16_ss(d, al + ks * astep, bh + ks * bstep, idesc, ks > 0);
#pragma unroll
        for (int ks = 0; ks < TC_K1 / 16; ++ks) mma_f16_ss(d, ah + ks * astep, bl + ks * bstep, idesc, true);
#pragma unroll
        for (int ks = 0; ks < TC_K1 / 16; ++ks) mma_f16_ss(d, ah + ks * astep, bh + ks * bstep, idesc, true);
    }
    __syncwarp();
}

template <int N, int KS>
__device__ __forceinline__ void issue_layer(uint32_t ts, uint32_t bh_addr, uint32_t bl_addr) {
    issue_layer_part<N, 0, KS, true>(ts, bh_addr, bl_addr);
}

// ---------------------------------------------------------------------------------------
// Tail of one unit from the per-block records (one warp; lane = p*4+q, p = system slot, q = 10 hidden / 5 latent
// columns).  hw: this unit's head block (PackedLayout V0p .. lv_sum, HEAD_FLOATS floats) in SHARED memory -- the tail
// warp prefetches it with one bulk copy while it waits for the unit's records; read through L2 (round 1) the 40 + 40
// dependent weight rows made the tail the longest chain of the kernel.  scratch: TAIL_SCRATCH floats, warp-private.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void head_layer_s(const float* __restrict__ sin_, const float* __restrict__ ws,
                                             const float* __restrict__ bs, int p, int q, float* __restrict__ sout) {
    // packed fp32x2 FMAs: 5 instead of 10 per k; (w[2j], w[2j+1]) pairs are adjacent in the padded row
    u64 acc[5];
    const float* bq = bs + q * GC;
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[j] = pack2(bq[2 * j], bq[2 * j + 1]);
    const float* wq = ws + q * GC;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(wq + k * HP);
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(wq + k * HP + 4);
        const u64 c = *reinterpret_cast<const u64*>(wq + k * HP + 8);
        const float sv = sin_[p * 41 + k];
        const u64 s2 = pack2(sv, sv);
        acc[0] = fma2(s2, a.x, acc[0]);
        acc[1] = fma2(s2, a.y, acc[1]);
        acc[2] = fma2(s2, b.x, acc[2]);
        acc[3] = fma2(s2, b.y, acc[3]);
        acc[4] = fma2(s2, c, acc[4]);
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        float lo, hi;
        unpack2(acc[j], lo, hi);
        sout[p * 41 + q * 10 + 2 * j] = relu_nan(lo);
        sout[p * 41 + q * 10 + 2 * j + 1] = relu_nan(hi);
    }
}

// ---- tail of one unit, in two stages run by two different warps (a "stats" warp and a "head" warp; lane = p*4+q,
// p = system slot, q = 10 hidden / 5 latent columns) ----
// Stage 1a: the unit's 5 x 40 standard normals for the sampled summary statistics (explicit eps, or Philox keyed on the
// global unit / system index) -> eS.  Independent of the records: runs BEFORE the stats warp waits for them.
__device__ __forceinline__ void tail_draw_eps(const float* __restrict__ eps_u, uint64_t seed, uint32_t gunit, int64_t gsys0,
                                              int64_t n0, int n_valid, float* __restrict__ eS) {
    const int lane = threadIdx.x & 31;
    if (eps_u) {
        for (int idx = lane; idx < SYS * S2; idx += 32) {
            const int s = idx / S2, j = idx % S2;
            eS[idx] = (s < n_valid) ? __ldg(eps_u + (n0 + s) * S2 + j) : 0.f;
        }
    } else {
#pragma unroll 1
        for (int b = lane; b < SYS * (S2 / 4); b += 32) {
            const int s = b / (S2 / 4), blk = b % (S2 / 4);
            const float4 n4 = philox_normal4(seed, STREAM_EPS, gunit, (uint32_t)(gsys0 + s), (uint32_t)blk);
            *reinterpret_cast<float4*>(eS + s * S2 + blk * 4) = n4;
        }
    }
    __syncwarp();
}

// Stage 1b: merge the 4 block records of every system (exact two-level mean / M2), sampled summary statistics
// (:416-430) -> sum[SYS][41] in shared memory.  lv_sum_g: the unit's summary_noise_logvar in GLOBAL memory (noisy forward).
__device__ __forceinline__ void tail_stats(const float* __restrict__ rec, const float* __restrict__ eS,
                                           const float* __restrict__ lv_sum_g, const float* __restrict__ eps_sum_u,
                                           float* __restrict__ summary_u, int64_t n0, int n_valid, float* __restrict__ sum) {
    const int lane = threadIdx.x & 31;
    const int p = min(lane >> 2, SYS - 1), q = lane & 3;  // lanes of p >= SYS shadow system SYS-1 (stores are masked)
    const bool live = (lane >> 2) < SYS;
    const SysRec sr = sys_records(p);
    const float Tf = (float)T_FIXED, Tm1 = (float)(T_FIXED - 1);
    // unrolled: the five columns' sqrt / divide chains interleave
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int col = q * 5 + i;
        float2 r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = *reinterpret_cast<const float2*>(rec + (sr.idx[k] * L + col) * 2);
        // exact two-level mean / M2: mean = sum n_k m_k / T, M2 = sum M2_k + sum n_k (m_k - mean)^2
        const float mean = ((sr.cnt[0] * r[0].x + sr.cnt[1] * r[1].x) + (sr.cnt[2] * r[2].x + sr.cnt[3] * r[3].x)) / Tf;
        float m2 = (r[0].y + r[1].y) + (r[2].y + r[3].y);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dlt = r[k].x - mean;
            m2 = fmaf(sr.cnt[k] * dlt, dlt, m2);
        }
        const float sd = sqrtf(__fdiv_rn(m2, Tm1));  // torch.std(x, dim=1)**2 (:419)
        const float var = __fmul_rn(sd, sd);
        const float std_in_mu = sqrtf(__fdiv_rn(var, Tf));                                    // :422
        const float std_in_var = sqrtf(__fdiv_rn(__fmul_rn(2.0f, __fmul_rn(var, var)), Tm1));   // :423
        const float mu_s = __fadd_rn(__fmul_rn(eS[p * S2 + col], std_in_mu), mean);            // :426
        const float var_s = __fadd_rn(__fmul_rn(eS[p * S2 + L + col], std_in_var), var);       // :427
        float s_mu = mu_s;
        float s_sd = sqrtf(__fadd_rn(fabsf(var_s), 1e-5f));                                    // :430
        if (live && p < n_valid) {
            if (summary_u) {
                summary_u[(n0 + p) * S2 + col] = s_mu;
                summary_u[(n0 + p) * S2 + L + col] = s_sd;
            }
            if (eps_sum_u) {
                const float e0 = __ldg(eps_sum_u + (n0 + p) * S2 + col);
                const float e1 = __ldg(eps_sum_u + (n0 + p) * S2 + L + col);
                s_mu = __fadd_rn(s_mu, __fmul_rn(e0, expf(__fdiv_rn(__ldg(lv_sum_g + col), 2.0f))));
                s_sd = __fadd_rn(s_sd, __fmul_rn(e1, expf(__fdiv_rn(__ldg(lv_sum_g + L + col), 2.0f))));
            }
        }
        if (live) {
            sum[p * 41 + col] = s_mu;
            sum[p * 41 + L + col] = s_sd;
        }
    }
    __syncwarp();
}

// Stage 2: regress_nn + soft_clamp (:301-321, :432-442) from sum[SYS][41]; hw: the unit's head block in shared memory;
// sB, sC: [SYS][41] scratch of the head warp.  `after_layer1()` runs once `sum` has been consumed.
template <class F>
__device__ __forceinline__ void tail_head(const float* __restrict__ sum, const float* __restrict__ hw, const PackedLayout& pl,
                                          int64_t n0, int n_valid, const HeadConsts& hc, float* __restrict__ sB,
                                          float* __restrict__ sC, float* __restrict__ out_unit, int64_t out_sys_stride,
                                          F after_layer1) {
    const int lane = threadIdx.x & 31;
    const int p = min(lane >> 2, SYS - 1), q = lane & 3;
    const bool live = (lane >> 2) < SYS;
    if (live) head_layer_s(sum, hw, hw + pl.c0p - pl.V0p, p, q, sB);
    __syncwarp();
    after_layer1();
    if (live) head_layer_s(sB, hw + pl.V1p - pl.V0p, hw + pl.c1p - pl.V0p, p, q, sC);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = q * 10 + i;
        const float r = sC[p * 41 + k];
        o0 = fmaf(r, hw[pl.V2 - pl.V0p + k], o0);
        o1 = fmaf(r, hw[pl.V2 - pl.V0p + H + k], o1);
    }
    o0 += __shfl_xor_sync(0xffffffffu, o0, 1);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 1);
    o0 += __shfl_xor_sync(0xffffffffu, o0, 2);
    o1 += __shfl_xor_sync(0xffffffffu, o1, 2);
    if (q == 0 && live && p < n_valid) {
        o0 += hw[pl.c2 - pl.V0p];
        o1 += hw[pl.c2 - pl.V0p + 1];
        float2 o = make_float2(soft_clamp_dev(o0, hc.lo_mu, hc.hi_mu), soft_clamp_dev(o1, hc.lo_sd, hc.hi_sd));
        *reinterpret_cast<float2*>(out_unit + (n0 + p) * out_sys_stride) = o;
    }
    __syncwarp();
}

}  // namespace tc
}  // namespace bnn
