// Shared device/host helpers for libbnnchaos (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/bnnchaos.h"
#include "../../include/bnnchaos_diag.h"

namespace bnn {

// ---------------------------------------------------------------------------------------
// Error plumbing (C ABI: no exceptions).
// ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_device();  // BNN_OK or BNN_E_ARCH

#define BNN_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            bnn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

#define BNN_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            bnn::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

// One-time, per-device setup (cudaFuncSetAttribute applies to the current device): need() is true the first time it
// is called with a given device current.  Two racing first calls may both return true; the setup is idempotent.
struct PerDeviceOnce {
    unsigned long long mask = 0;
    bool need() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long b = 1ull << (d & 63);
        const unsigned long long old = __atomic_fetch_or(&mask, b, __ATOMIC_RELAXED);
        return !(old & b);
    }
};

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------
// Compiled model shape.  The kernels are specialised for the reference's only shipped
// architecture (hidden=40, latent=20, in=out=1; find_minima.py:36-41); n_features, the
// zero mask and T stay runtime parameters.
// ---------------------------------------------------------------------------------------
constexpr int H = 40;        // hidden
constexpr int L = 20;        // latent
constexpr int S2 = 2 * L;    // summary width
constexpr int GC = 12;       // padded columns per lane group (10 real + 2 pad)
constexpr int HP = 4 * GC;   // padded hidden row: 4 lane groups x 12
constexpr int MAXF = 64;     // zero_mask is 64 bits wide
constexpr int SYS_TILE = 8;  // systems per CTA tile
// tensor-core path: padded GEMM shapes (N is a multiple of 16; the biases are added in the epilogue)
constexpr int TC_K1 = 32;   // layer-1 K: up to 32 live inputs (the v50 flag set has 31)
constexpr int TC_K1W = 48;  // layer-1 K of the "wide" variant: 33..48 live inputs (all 41 columns: noisy forward, other flag sets)
constexpr int TC_K2 = 40;   // layer-2/3 K: 40 hidden
constexpr int TC_N = 48;    // layer-1/2 N: 40 hidden + 8 zero rows
constexpr int TC_N3 = 32;   // layer-3 N: 20 latent + 12 zero rows
constexpr int TC_BIAS = 128;  // bias block riding behind the B operands: b0[48] | b1[48] | b2[32] (zero padded)

// Flat (SWAGModel.flatten, spock_reg_model.py:734-746) offsets for in=out=1.
struct FlatLayout {
    int F, d;
    int lv_in, lv_sum, W0, b0, W1, b1, W2, b2, V0, c0, V1, c1, V2, c2;
    __host__ __device__ explicit FlatLayout(int F_) : F(F_) {
        int o = 0;
        lv_in = o;  o += F;
        lv_sum = o; o += S2;
        W0 = o; o += H * F;
        b0 = o; o += H;
        W1 = o; o += H * H;
        b1 = o; o += H;
        W2 = o; o += L * H;
        b2 = o; o += L;
        V0 = o; o += H * S2;
        c0 = o; o += H;
        V1 = o; o += H * H;
        c1 = o; o += H;
        V2 = o; o += 2 * H;
        c2 = o; o += 2;
        d = o;
    }
};

// Kernel-side ("packed") layout of one unit's weights.  All matrices are stored k-major
// (input index first) so that a lane group's output columns are contiguous:
//   W0p[kin][48]  kin = live input columns only;  [k][q*12+i] = W0[q*10+i][col(k)], i<10
//   W1p[40][48], V0p[40][48], V1p[40][48] likewise
//   W2p[40][48]   [k][q*12+2i+{0,1}] = W2[q*5+i][k], i<5   (duplicated: f32x2 row pairs)
//   biases b0p/b1p/c0p/c1p[48] ([q*12+i]), b2p[48] (duplicated like W2p), V2[2][40], c2[4],
//   then the two logvar vectors verbatim: lv_sum[40], lv_in[F rounded up to 4]
// The feature part (first feat_floats) is what the MLP warps stage in shared memory.
struct PackedLayout {
    int kin;  // live input columns
    int W0p, b0p, W1p, b1p, W2p, b2p, feat_floats;
    int V0p, c0p, V1p, c1p, V2, c2, lv_sum, lv_in;
    // tensor-core section (only when kin <= 48): tf32 hi/lo splits of the feature weights as tcgen05 B
    // operands, canonical K-major chunk layout [k/4][n][4], then the fp32 bias block.
    int tc_ok, tc_k1, b1_f16, B1h, B1l, B2h, B2l, B3h, B3l, Bb, P;
    __host__ __device__ explicit PackedLayout(int kin_, int F_ = MAXF) : kin(kin_) {
        int o = 0;
        W0p = o; o += kin * HP;
        b0p = o; o += HP;
        W1p = o; o += H * HP;
        b1p = o; o += HP;
        W2p = o; o += H * HP;
        b2p = o; o += HP;
        feat_floats = o;
        V0p = o; o += S2 * HP;
        c0p = o; o += HP;
        V1p = o; o += H * HP;
        c1p = o; o += HP;
        V2 = o; o += 2 * H;
        c2 = o; o += 4;
        lv_sum = o; o += S2;                 // summary_noise_logvar (noisy forward only)
        lv_in = o; o += (F_ + 3) & ~3;       // input_noise_logvar, all F columns
        tc_ok = (kin <= TC_K1W) ? 1 : 0;
        tc_k1 = kin <= TC_K1 ? TC_K1 : TC_K1W;
        // layer 1 with at most 32 live inputs runs as kind::f16 (x is staged once per tile as fp16 hi / lo, see
        // predict_tc.cuh): its B operands are fp16 hi / lo, [K1/8][48][8 halves], two halves per 32-bit word; the wide
        // variant keeps tf32 operands [K1/4][48][4]
        b1_f16 = (kin <= TC_K1) ? 1 : 0;
        const int b1 = tc_ok ? (b1_f16 ? tc_k1 * TC_N / 2 : tc_k1 * TC_N) : 0;
        B1h = o; o += b1;
        B1l = o; o += b1;
        B2h = o; o += tc_ok ? TC_K2 * TC_N : 0;          // [40/4][48][4]
        B2l = o; o += tc_ok ? TC_K2 * TC_N : 0;
        B3h = o; o += tc_ok ? TC_K2 * TC_N3 : 0;         // [40/4][32][4]
        B3l = o; o += tc_ok ? TC_K2 * TC_N3 : 0;
        Bb = o; o += tc_ok ? TC_BIAS : 0;
        P = o;
    }
};

struct LiveCols {
    int n;
    int8_t col[MAXF];
};

static inline LiveCols live_columns(const bnn_model_config* cfg) {
    LiveCols lc;
    lc.n = 0;
    for (int c = 0; c < cfg->n_features; ++c)
        if (!((cfg->zero_mask >> c) & 1ull)) lc.col[lc.n++] = (int8_t)c;
    for (int i = lc.n; i < MAXF; ++i) lc.col[i] = -1;
    return lc;
}

int validate_config(const bnn_model_config* cfg);

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) + Box-Muller; mirrors oracle/restatement.py.
// counter = (blk, a, b, stream), key = (seed lo, seed hi).
// ---------------------------------------------------------------------------------------
enum : uint32_t { STREAM_Z1 = 1, STREAM_Z2 = 2, STREAM_EPS = 3, STREAM_EPS_IN = 4, STREAM_EPS_SUM = 5 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

__device__ __forceinline__ float u01(uint32_t u) {
    return (__uint2float_rn(u >> 8) + 1.0f) * 5.9604644775390625e-08f;  // 2^-24, in (0,1]
}

__device__ __forceinline__ float4 box_muller(uint4 u) {
    float a0 = u01(u.x), b0 = u01(u.y), a1 = u01(u.z), b1 = u01(u.w);
    float r0 = sqrtf(-2.0f * logf(a0)), r1 = sqrtf(-2.0f * logf(a1));
    float s0, c0, s1, c1;
    sincosf(6.283185307179586f * b0, &s0, &c0);
    sincosf(6.283185307179586f * b1, &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// Box-Muller on the special-function unit (lg2 / sin / cos approximations, ~1e-6 absolute): for the bulk training
// noise (STREAM_EPS_IN), where 4 normals otherwise cost ~370 instructions.  The angle is reduced to (-pi, pi], the
// range the approximations are specified on.  Every consumer of a stream must use the same variant.
__device__ __forceinline__ float4 box_muller_fast(uint4 u) {
    float a0 = u01(u.x), b0 = u01(u.y), a1 = u01(u.z), b1 = u01(u.w);
    float r0 = sqrtf(-2.0f * __logf(a0)), r1 = sqrtf(-2.0f * __logf(a1));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * (b0 - 0.5f), &s0, &c0);   // sin(2 pi b) = -sin(2 pi (b - 1/2))
    __sincosf(6.283185307179586f * (b1 - 0.5f), &s1, &c1);
    return make_float4(-r0 * c0, -r0 * s0, -r1 * c1, -r1 * s1);
}

__device__ __forceinline__ float4 philox_normal4_fast(uint64_t seed, uint32_t stream, uint32_t a, uint32_t b,
                                                      uint32_t blk) {
    uint4 r = philox4x32_10(make_uint4(blk, a, b, stream), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return box_muller_fast(r);
}

__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t stream, uint32_t a, uint32_t b,
                                                 uint32_t blk) {
    uint4 r = philox4x32_10(make_uint4(blk, a, b, stream), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return box_muller(r);
}

// fp32 -> tf32 (10 explicit mantissa bits), round to nearest; result is an fp32 bit pattern
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ---------------------------------------------------------------------------------------
// Packed fp32x2 FMA (Blackwell FFMA2): one issue slot, two IEEE fp32 FMAs.
// ---------------------------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

}  // namespace bnn
