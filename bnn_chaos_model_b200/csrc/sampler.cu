// K1: SWAG weight sampling for all (model, weight-sample) units at once, and the
// flatten()-order -> packed-layout gather.
//
// Reference: SWAGModel.sample_weights, /root/reference/spock_reg_model.py:815-838
//   D = pre_D - w_avg[:,None]; z1 ~ N(0,I_d); z2 ~ N(0,I_K)
//   w  = w_avg + (scale/sqrt2) * z1 * sqrt|w2_avg - w_avg^2|     (dense diag matmul in the
//        reference, :832-834; every off-diagonal term is an exact +0)
//   w += scale * (D z2) / sqrt(2 (K-1))                           (:835)
// and SWAGModel.load (:748-761), which installs the flat vector into the module.
//
// HBM-bound: writes d floats per unit; pre_D rows are re-read from L2.
#include <cuda_fp16.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace bnn {

constexpr int SAMPLER_THREADS = 128;
constexpr int MAXK = 64;

// One block = 512 consecutive flat indices (4 per thread = one Philox block) of one unit.
__global__ void __launch_bounds__(SAMPLER_THREADS)
swag_sample_kernel(const float* __restrict__ w_avg, const float* __restrict__ w2_avg,
                   const float* __restrict__ pre_D, int d, int K, const int32_t* __restrict__ unit_model,
                   int64_t unit_offset, int samples_per_model, int n_models, float c1, float scale, float c2div,
                   uint64_t seed, const float* __restrict__ z1, const float* __restrict__ z2,
                   float* __restrict__ theta, int blocks_per_unit) {
    __shared__ float z2s[MAXK];
    const int64_t u = blockIdx.x / blocks_per_unit;
    const int jb = (blockIdx.x % blocks_per_unit) * (SAMPLER_THREADS * 4);
    const int64_t gu = unit_offset + u;
    int m = unit_model ? unit_model[u] : (int)(gu / samples_per_model);
    m = min(max(m, 0), n_models - 1);

    if (threadIdx.x < (K + 3) / 4) {
        float4 n4;
        if (z2) {
            const float* p = z2 + u * K + threadIdx.x * 4;
            int rem = K - threadIdx.x * 4;
            n4.x = p[0];
            n4.y = rem > 1 ? p[1] : 0.f;
            n4.z = rem > 2 ? p[2] : 0.f;
            n4.w = rem > 3 ? p[3] : 0.f;
        } else {
            n4 = philox_normal4(seed, STREAM_Z2, (uint32_t)gu, 0u, threadIdx.x);
        }
        z2s[threadIdx.x * 4 + 0] = n4.x;
        z2s[threadIdx.x * 4 + 1] = n4.y;
        z2s[threadIdx.x * 4 + 2] = n4.z;
        z2s[threadIdx.x * 4 + 3] = n4.w;
    }
    __syncthreads();

    const int j0 = jb + threadIdx.x * 4;
    if (j0 >= d) return;
    float zz[4];
    if (z1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) zz[i] = (j0 + i < d) ? z1[u * (int64_t)d + j0 + i] : 0.f;
    } else {
        float4 n4 = philox_normal4(seed, STREAM_Z1, (uint32_t)gu, 0u, (uint32_t)(j0 >> 2));
        zz[0] = n4.x; zz[1] = n4.y; zz[2] = n4.z; zz[3] = n4.w;
    }
    const float* wa = w_avg + (int64_t)m * d;
    const float* w2a = w2_avg + (int64_t)m * d;
    const float* pd = pre_D + (int64_t)m * d * K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j0 + i;
        if (j >= d) break;
        const float w = wa[j];
        // explicit _rn intrinsics: keep the reference's separate mul/sub/add roundings (no FMA contraction)
        const float sig = fabsf(__fsub_rn(w2a[j], __fmul_rn(w, w)));
        float th = __fadd_rn(w, __fmul_rn(__fmul_rn(c1, zz[i]), sqrtf(sig)));
        const float* row = pd + (int64_t)j * K;
        float dot = 0.f;
        for (int k = 0; k < K; ++k) dot = fmaf(__fsub_rn(__ldg(row + k), w), z2s[k], dot);
        th = __fadd_rn(th, __fdiv_rn(__fmul_rn(scale, dot), c2div));
        theta[u * (int64_t)d + j] = th;
    }
}

// Where packed float i of a unit comes from: index into the flat (flatten()-order) vector and how it is transformed.
enum PackKind : int { PK_ZERO = 0, PK_COPY = 1, PK_TF32_HI = 2, PK_TF32_LO = 3, PK_F16_HI = 4, PK_F16_LO = 5 };
struct alignas(16) PackSrc {
    int src;
    int kind;
    int src2;   // fp16 kinds: the word's second half (k + 1); -1 = zero
    int pad_ = 0;
};

__device__ __forceinline__ PackSrc pack_source(int i, const FlatLayout& fl, const PackedLayout& pl, const LiveCols& lc) {
    int src = -1;
    auto grp = [](int c, int& q, int& ii) { q = c / GC; ii = c % GC; };
    int q, ii;
    if (i < pl.b0p) {  // W0p[kin][48]
        int k = i / HP;
        grp(i % HP, q, ii);
        if (ii < 10) src = fl.W0 + (q * 10 + ii) * fl.F + lc.col[k];
    } else if (i < pl.W1p) {
        grp(i - pl.b0p, q, ii);
        if (ii < 10) src = fl.b0 + q * 10 + ii;
    } else if (i < pl.b1p) {
        int r = i - pl.W1p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.W1 + (q * 10 + ii) * H + k;
    } else if (i < pl.W2p) {
        grp(i - pl.b1p, q, ii);
        if (ii < 10) src = fl.b1 + q * 10 + ii;
    } else if (i < pl.b2p) {  // W2p duplicated pairs
        int r = i - pl.W2p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.W2 + (q * 5 + ii / 2) * H + k;
    } else if (i < pl.V0p) {
        grp(i - pl.b2p, q, ii);
        if (ii < 10) src = fl.b2 + q * 5 + ii / 2;
    } else if (i < pl.c0p) {
        int r = i - pl.V0p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.V0 + (q * 10 + ii) * S2 + k;
    } else if (i < pl.V1p) {
        grp(i - pl.c0p, q, ii);
        if (ii < 10) src = fl.c0 + q * 10 + ii;
    } else if (i < pl.c1p) {
        int r = i - pl.V1p, k = r / HP;
        grp(r % HP, q, ii);
        if (ii < 10) src = fl.V1 + (q * 10 + ii) * H + k;
    } else if (i < pl.V2) {
        grp(i - pl.c1p, q, ii);
        if (ii < 10) src = fl.c1 + q * 10 + ii;
    } else if (i < pl.c2) {
        src = fl.V2 + (i - pl.V2);
    } else if (i < pl.lv_sum) {
        int r = i - pl.c2;
        if (r < 2) src = fl.c2 + r;
    } else if (i < pl.lv_in) {
        src = fl.lv_sum + (i - pl.lv_sum);
    } else if (i < pl.B1h) {
        int r = i - pl.lv_in;
        if (r < fl.F) src = fl.lv_in + r;
    } else if (i >= pl.Bb) {  // fp32 bias block b0[48] | b1[48] | b2[32]
        const int r = i - pl.Bb;
        if (r < TC_N) { if (r < H) src = fl.b0 + r; }
        else if (r < 2 * TC_N) { if (r - TC_N < H) src = fl.b1 + r - TC_N; }
        else if (r - 2 * TC_N < L) src = fl.b2 + r - 2 * TC_N;
    } else {
        // tensor-core B operands: element (n, k) of a [N][K] matrix sits at ((k/4)*N + n)*4 + k%4
        int r, N, nreal, wsrc, kreal, ld;
        bool lo;
        if (i < pl.B2h && pl.b1_f16) {
            // fp16 hi / lo of W0: word r of [K1/8][48][4 words] holds the halves k = 8 chunk + 2 (r & 3) + {0, 1} of row n
            r = i - pl.B1h;
            const int b1 = pl.tc_k1 * TC_N / 2;
            lo = r >= b1; r -= lo ? b1 : 0;
            const int chunk = r / (TC_N * 4), n = (r / 4) % TC_N, k = chunk * 8 + 2 * (r & 3);
            PackSrc ps{-1, lo ? PK_F16_LO : PK_F16_HI, -1};
            if (n < H) {
                if (k < pl.kin) ps.src = fl.W0 + n * fl.F + (int)lc.col[k];
                if (k + 1 < pl.kin) ps.src2 = fl.W0 + n * fl.F + (int)lc.col[k + 1];
                // spare K column (at most 31 live inputs): the layer-1 bias, multiplied by the ones column of the x tile
                if (k + 1 == TC_K1 - 1 && pl.kin < TC_K1) ps.src2 = fl.b0 + n;
            }
            return ps;
        } else if (i < pl.B2h) {
            r = i - pl.B1h; N = TC_N; lo = r >= pl.tc_k1 * TC_N; r -= lo ? pl.tc_k1 * TC_N : 0;
            nreal = H; kreal = pl.kin; wsrc = fl.W0; ld = fl.F;
        } else if (i < pl.B3h) {
            r = i - pl.B2h; N = TC_N; lo = r >= TC_K2 * TC_N; r -= lo ? TC_K2 * TC_N : 0;
            nreal = H; kreal = H; wsrc = fl.W1; ld = H;
        } else {
            r = i - pl.B3h; N = TC_N3; lo = r >= TC_K2 * TC_N3; r -= lo ? TC_K2 * TC_N3 : 0;
            nreal = L; kreal = H; wsrc = fl.W2; ld = H;
        }
        const int chunk = r / (N * 4), n = (r / 4) % N, k = chunk * 4 + (r & 3);
        if (n < nreal && k < kreal)
            return PackSrc{wsrc + n * ld + ((wsrc == fl.W0) ? (int)lc.col[k] : k), lo ? PK_TF32_LO : PK_TF32_HI, -1};
        return PackSrc{-1, PK_ZERO, -1};
    }
    return PackSrc{src, src >= 0 ? PK_COPY : PK_ZERO, -1};
}

// fp16 hi (round to nearest: 11 significant bits, like tf32) or lo = fp16(w - hi) of one weight, as 16 bits
__device__ __forceinline__ uint32_t f16_part_bits(float w, bool lo) {
    const __half h = __float2half_rn(w);
    return (uint32_t)__half_as_ushort(lo ? __float2half_rn(w - __half2float(h)) : h);
}

__device__ __forceinline__ float pack_value(const float* th, PackSrc ps) {
    if (ps.kind == PK_ZERO) return 0.f;
    if (ps.kind >= PK_F16_HI) {
        const bool lo = ps.kind == PK_F16_LO;
        const uint32_t a = ps.src >= 0 ? f16_part_bits(th[ps.src], lo) : 0u;
        const uint32_t b = ps.src2 >= 0 ? f16_part_bits(th[ps.src2], lo) : 0u;
        return __uint_as_float(a | (b << 16));
    }
    const float w = th[ps.src];
    if (ps.kind == PK_COPY) return w;
    const float hi = tf32_rna(w);
    return ps.kind == PK_TF32_HI ? hi : tf32_rna(w - hi);  // lo pre-rounded: the tensor core would truncate it
}

// theta [U,d] (flatten order) -> theta_packed [U,P].  One thread per packed float.
__global__ void pack_theta_kernel(const float* __restrict__ theta, int64_t n_units, FlatLayout fl, PackedLayout pl,
                                  LiveCols lc, float* __restrict__ packed) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n_units * pl.P) return;
    const int64_t u = idx / pl.P;
    const int i = (int)(idx % pl.P);
    packed[idx] = pack_value(theta + u * fl.d, pack_source(i, fl, pl, lc));
}

// ---------------------------------------------------------------------------------------
// K1 fused: sample G units per CTA and write their packed layout directly (and the flat vector when asked).
//   phase A: the CTA walks theta in chunks of 512 elements; the chunk's 512 x K block of pre_D is brought into shared
//            memory once with coalesced 16-byte loads (the per-thread rows of the unfused kernel were 32 distinct
//            lines per warp load: the L1 tag stage, not the arithmetic, set its 0.57 ms per 1000 units) and shared by
//            the CTA's G units; thread (g, quad) draws one Philox block = 4 elements of unit g; theta lands in shared
//            memory (30 kB per unit);
//   phase B: packed index i -> flat source (pack_source, computed once per i) -> one coalesced store per unit.
// Arithmetic identical to swag_sample_kernel + pack_theta_kernel (bit-equal results; tests/test_gpu_predict.py).
// ---------------------------------------------------------------------------------------
constexpr int SP_QUADS = 128;                     // element quads per chunk (one Philox block each)
constexpr int SP_CHUNK = SP_QUADS * 4;            // 512 flat elements per chunk

// floats per quad of pre_D rows in the shared-memory tile: 4 K + V with V = the widest vector K allows (4 | 2 | 1), so
// that the stride between two lanes' quads is an odd number of V-float words: conflict-free V-wide reads
static inline int sample_pack_vec(int K) { return (K & 3) == 0 ? 4 : ((K & 1) == 0 ? 2 : 1); }
static inline size_t sample_pack_smem_bytes(int G, int d, int K) {
    return ((size_t)G * ((d + 3) & ~3) + (size_t)SP_QUADS * (4 * K + sample_pack_vec(K)) + (size_t)G * MAXK) * sizeof(float);
}

template <int V> struct VecT;
template <> struct VecT<4> { typedef float4 type; };
template <> struct VecT<2> { typedef float2 type; };
template <> struct VecT<1> { typedef float type; };

template <int G, int V>
__global__ void __launch_bounds__(SP_QUADS * G)
swag_sample_pack_kernel(const float* __restrict__ w_avg, const float* __restrict__ w2_avg,
                        const float* __restrict__ pre_D, int d, int K, const int32_t* __restrict__ unit_model,
                        int64_t unit_offset, int samples_per_model, int n_models, float c1, float scale, float c2div,
                        uint64_t seed, const float* __restrict__ z1, const float* __restrict__ z2, int64_t n_units,
                        float* __restrict__ theta, float* __restrict__ packed, FlatLayout fl, PackedLayout pl,
                        LiveCols lc, const PackSrc* __restrict__ pack_table) {
    typedef typename VecT<V>::type vec_t;
    extern __shared__ __align__(16) float sp_smem[];
    const int dpad = (d + 3) & ~3;
    const int pitch = 4 * K + V;                  // floats per quad of pre_D rows
    float* th_s = sp_smem;                        // [G][dpad]
    float* tile = th_s + G * dpad;                // [SP_QUADS][pitch]
    float* z2s = tile + SP_QUADS * pitch;         // [G][MAXK]
    const int tid = threadIdx.x, g = tid / SP_QUADS, q = tid % SP_QUADS;
    const int64_t u0 = (int64_t)blockIdx.x * G;
    const int64_t u = u0 + g;
    const bool live = u < n_units;
    const int64_t gu = unit_offset + u;
    auto model_of = [&](int64_t uu) {
        int m = unit_model ? unit_model[uu] : (int)((unit_offset + uu) / samples_per_model);
        return min(max(m, 0), n_models - 1);
    };
    const int m = live ? model_of(u) : 0;
    // the units of one CTA normally share a model (consecutive units, samples_per_model of them per model)
    bool same = true;
#pragma unroll
    for (int gg = 1; gg < G; ++gg)
        if (u0 + gg < n_units && model_of(u0 + gg) != model_of(u0)) same = false;

    if (live && q < (K + 3) / 4) {
        float4 n4;
        if (z2) {
            const float* p = z2 + u * K + q * 4;
            const int rem = K - q * 4;
            n4.x = p[0];
            n4.y = rem > 1 ? p[1] : 0.f;
            n4.z = rem > 2 ? p[2] : 0.f;
            n4.w = rem > 3 ? p[3] : 0.f;
        } else {
            n4 = philox_normal4(seed, STREAM_Z2, (uint32_t)gu, 0u, q);
        }
        float* zs = z2s + g * MAXK + q * 4;
        zs[0] = n4.x; zs[1] = n4.y; zs[2] = n4.z; zs[3] = n4.w;
    }

    const float* wa = w_avg + (int64_t)m * d;
    const float* w2a = w2_avg + (int64_t)m * d;
    const int n_pass = same ? 1 : G;
    const int qv = 4 * K / V;                     // V-float words per quad of rows
    for (int jb = 0; jb < d; jb += SP_CHUNK) {
        const int rows = min(SP_CHUNK, d - jb);
        for (int pass = 0; pass < n_pass; ++pass) {
            const int mt = same ? model_of(u0) : (u0 + pass < n_units ? model_of(u0 + pass) : 0);
            // (mt d + jb) K floats from a 16-byte aligned base: V-float aligned because K is a multiple of V (jb K too)
            const vec_t* src = reinterpret_cast<const vec_t*>(pre_D + ((int64_t)mt * d + jb) * K);
            __syncthreads();                      // the previous chunk's readers are done with the tile
            // a warp per quad of rows (qv consecutive V-float words -> one padded tile row), as asynchronous copies: all of a
            // thread's words are in flight at once and land while it draws its normals (as register loads in rounds of
            // four, the L2 round trips of this loop were 31 % of the kernel's warp samples)
            {
                const int wp = tid >> 5, ln = tid & 31, nw = SP_QUADS * G / 32, nq = (rows + 3) >> 2;
                for (int qq = wp; qq < nq; qq += nw) {
                    const int na = min(qv, (rows - 4 * qq) * K / V);
                    const vec_t* sa = src + (int64_t)qq * qv;
                    const uint32_t da = (uint32_t)__cvta_generic_to_shared(tile + qq * pitch);
                    for (int r = ln; r < na; r += 32)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(da + (uint32_t)(r * V * 4)), "l"(sa + r),
                                     "n"(V * 4)
                                     : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            const int j0 = jb + 4 * q;
            const bool mine = live && j0 < d && (same || g == pass);
            // the thread's four elements advance together through k (every element still sums k = 0 .. K-1 in order,
            // like swag_sample_kernel): four independent FMA chains instead of one, z2 read once per k
            float w[4], th[4], dot[4];
            if (mine) {
                float zz[4];
                if (z1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) zz[i] = (j0 + i < d) ? z1[u * (int64_t)d + j0 + i] : 0.f;
                } else {
                    float4 n4 = philox_normal4(seed, STREAM_Z1, (uint32_t)gu, 0u, (uint32_t)(j0 >> 2));
                    zz[0] = n4.x; zz[1] = n4.y; zz[2] = n4.z; zz[3] = n4.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int j = min(j0 + i, d - 1);
                    w[i] = wa[j];
                    // explicit _rn intrinsics: keep the reference's separate mul/sub/add roundings (no FMA contraction)
                    const float sig = fabsf(__fsub_rn(w2a[j], __fmul_rn(w[i], w[i])));
                    th[i] = __fadd_rn(w[i], __fmul_rn(__fmul_rn(c1, zz[i]), sqrtf(sig)));
                    dot[i] = 0.f;
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (mine) {
                const float* zs = z2s + g * MAXK;
                const float* rows4 = tile + q * pitch;
                // software-pipelined: the shared-memory loads of step k + V are in flight while step k is summed
                vec_t zc = *reinterpret_cast<const vec_t*>(zs), rc[4], zn = zc, rn[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) rn[i] = rc[i] = *reinterpret_cast<const vec_t*>(rows4 + i * K);
#pragma unroll 2
                for (int k = 0; k < K; k += V) {
                    if (k + V < K) {
                        zn = *reinterpret_cast<const vec_t*>(zs + k + V);
#pragma unroll
                        for (int i = 0; i < 4; ++i) rn[i] = *reinterpret_cast<const vec_t*>(rows4 + i * K + k + V);
                    }
                    const float* z = reinterpret_cast<const float*>(&zc);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float* r = reinterpret_cast<const float*>(&rc[i]);
#pragma unroll
                        for (int e = 0; e < V; ++e) dot[i] = fmaf(__fsub_rn(r[e], w[i]), z[e], dot[i]);
                    }
                    zc = zn;
#pragma unroll
                    for (int i = 0; i < 4; ++i) rc[i] = rn[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (j0 + i < d) th_s[g * dpad + j0 + i] = __fadd_rn(th[i], __fdiv_rn(__fmul_rn(scale, dot[i]), c2div));
            }
        }
    }
    __syncthreads();
    const int n_live = (int)min((int64_t)G, n_units - u0);
    if (theta) {
        for (int gg = 0; gg < n_live; ++gg)
            for (int j = tid; j < d; j += SP_QUADS * G) theta[(u0 + gg) * (int64_t)d + j] = th_s[gg * dpad + j];
    }
    if (packed) {
        // the next index's table entry is in flight while this one's G values are gathered and stored
        int4 e_next = make_int4(-1, PK_ZERO, -1, 0);
        if (pack_table && tid < pl.P) e_next = __ldg(reinterpret_cast<const int4*>(pack_table) + tid);
        for (int i = tid; i < pl.P; i += SP_QUADS * G) {
            PackSrc ps;
            if (pack_table) {
                const int4 e = e_next;
                if (i + SP_QUADS * G < pl.P) e_next = __ldg(reinterpret_cast<const int4*>(pack_table) + i + SP_QUADS * G);
                ps.src = e.x; ps.kind = e.y; ps.src2 = e.z;
            } else {
                ps = pack_source(i, fl, pl, lc);
            }
#pragma unroll
            for (int gg = 0; gg < G; ++gg)
                if (gg < n_live) packed[(u0 + gg) * (int64_t)pl.P + i] = pack_value(th_s + gg * dpad, ps);
        }
    }
}

int launch_pack_theta(const bnn_model_config* cfg, const float* d_theta, int64_t n_units, float* d_packed,
                      cudaStream_t st) {
    FlatLayout fl(cfg->n_features);
    LiveCols lc = live_columns(cfg);
    PackedLayout pl(lc.n, cfg->n_features);
    const int64_t total = n_units * pl.P;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "too many units for one pack launch");
    pack_theta_kernel<<<(unsigned)blocks, threads, 0, st>>>(d_theta, n_units, fl, pl, lc, d_packed);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

struct SamplePackArgs {
    const float *w_avg, *w2_avg, *pre_D;
    int d, K;
    const int32_t* unit_model;
    int64_t unit_offset;
    int samples_per_model, n_models;
    float c1, scale, c2div;
    uint64_t seed;
    const float *z1, *z2;
    int64_t n_units;
    float *theta, *packed;
    const PackSrc* pack_table;
};

// packed index -> flat source for one (n_features, zero_mask): the same for every unit, CTA and call, so it is built once
// per device and configuration (a 16-byte entry per packed float, ~280 kB) and read by the fused sampler instead of being
// recomputed by every CTA (pack_source is ~100 instructions of branches and divisions: a quarter of the kernel's
// instructions in the ncu source view).  Entries are immutable once built; nullptr (allocation failed) = compute inline.
__global__ void pack_table_kernel(FlatLayout fl, PackedLayout pl, LiveCols lc, PackSrc* __restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pl.P) table[i] = pack_source(i, fl, pl, lc);
}

static const PackSrc* pack_table_for(const bnn_model_config* cfg, const FlatLayout& fl, const PackedLayout& pl,
                                     const LiveCols& lc, cudaStream_t st) {
    struct Entry { int dev, F; unsigned long long mask; const PackSrc* table; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
        if (e.dev == dev && e.F == cfg->n_features && e.mask == (unsigned long long)cfg->zero_mask) return e.table;
    PackSrc* t = nullptr;
    if (cache.size() >= 64 || cudaMalloc(&t, (size_t)pl.P * sizeof(PackSrc)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    pack_table_kernel<<<(pl.P + 255) / 256, 256, 0, st>>>(fl, pl, lc, t);
    // one-time: later launches on other streams read the finished table
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(t);
        return nullptr;
    }
    cache.push_back(Entry{dev, cfg->n_features, (unsigned long long)cfg->zero_mask, t});
    return t;
}

template <int G, int V>
static int launch_sample_pack(const SamplePackArgs& a, const FlatLayout& fl, const PackedLayout& pl, const LiveCols& lc,
                              cudaStream_t st) {
    const size_t smem = sample_pack_smem_bytes(G, a.d, a.K);
    static PerDeviceOnce attr_done;
    if (attr_done.need())
        BNN_CUDA(cudaFuncSetAttribute(swag_sample_pack_kernel<G, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int64_t blocks = (a.n_units + G - 1) / G;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "bnn_swag_sample: too many units for one launch");
    swag_sample_pack_kernel<G, V><<<(unsigned)blocks, SP_QUADS * G, smem, st>>>(
        a.w_avg, a.w2_avg, a.pre_D, a.d, a.K, a.unit_model, a.unit_offset, a.samples_per_model, a.n_models, a.c1, a.scale,
        a.c2div, a.seed, a.z1, a.z2, a.n_units, a.theta, a.packed, fl, pl, lc, a.pack_table);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

template <int G>
static int launch_sample_pack_g(const SamplePackArgs& a, const FlatLayout& fl, const PackedLayout& pl, const LiveCols& lc,
                                cudaStream_t st) {
    switch (sample_pack_vec(a.K)) {
        case 4: return launch_sample_pack<G, 4>(a, fl, pl, lc, st);
        case 2: return launch_sample_pack<G, 2>(a, fl, pl, lc, st);
        default: return launch_sample_pack<G, 1>(a, fl, pl, lc, st);
    }
}

}  // namespace bnn

extern "C" {

int bnn_pack_theta(const bnn_model_config* cfg, const float* d_theta, int64_t n_units, float* d_theta_packed,
                   void* stream) {
    int rc = bnn::validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = bnn::check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_theta && d_theta_packed && n_units > 0, BNN_E_ARG, "bnn_pack_theta: null pointer or n_units<=0");
    return bnn::launch_pack_theta(cfg, d_theta, n_units, d_theta_packed, (cudaStream_t)stream);
}

int bnn_swag_sample(const bnn_model_config* cfg, const float* d_w_avg, const float* d_w2_avg, const float* d_pre_D,
                    int32_t n_models, int32_t K, const int32_t* d_unit_model, int64_t n_units, int64_t unit_offset,
                    int32_t samples_per_model, float scale, uint64_t seed, const float* d_z1, const float* d_z2,
                    float* d_theta, float* d_theta_packed, void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_w_avg && d_w2_avg && d_pre_D, BNN_E_ARG, "bnn_swag_sample: SWAG statistics pointer is NULL");
    BNN_REQUIRE(aligned16(d_pre_D), BNN_E_ALIGN, "bnn_swag_sample: d_pre_D must be 16-byte aligned");
    BNN_REQUIRE(d_theta || d_theta_packed, BNN_E_ARG, "bnn_swag_sample: give d_theta, d_theta_packed or both");
    BNN_REQUIRE(n_models >= 1 && n_units >= 1, BNN_E_ARG, "bnn_swag_sample: n_models/n_units must be >= 1");
    // the reference fails in D @ z2 when pre_D has fewer than K columns (:835); K>=2 for sqrt(2(K-1))
    BNN_REQUIRE(K >= 2 && K <= MAXK, BNN_E_ARG, "bnn_swag_sample: K=%d out of [2,%d]", K, MAXK);
    BNN_REQUIRE((d_z1 == nullptr) == (d_z2 == nullptr), BNN_E_ARG, "bnn_swag_sample: give both z1 and z2 or neither");
    BNN_REQUIRE(d_unit_model || samples_per_model >= 1, BNN_E_ARG, "bnn_swag_sample: samples_per_model must be >= 1");
    const FlatLayout fl(cfg->n_features);
    const LiveCols lc = live_columns(cfg);
    const PackedLayout pl(lc.n, cfg->n_features);
    SamplePackArgs a;
    a.w_avg = d_w_avg; a.w2_avg = d_w2_avg; a.pre_D = d_pre_D;
    a.d = fl.d; a.K = K;
    a.unit_model = d_unit_model; a.unit_offset = unit_offset;
    a.samples_per_model = samples_per_model; a.n_models = n_models;
    // scale * (1/np.sqrt(2.0)) is a double that torch casts to fp32 when it meets the fp32 tensor (:834)
    a.c1 = (float)((double)scale * (1.0 / sqrt(2.0)));
    a.scale = scale;
    a.c2div = (float)sqrt(2.0 * (K - 1));
    a.seed = seed; a.z1 = d_z1; a.z2 = d_z2; a.n_units = n_units; a.theta = d_theta; a.packed = d_theta_packed;
    cudaStream_t st = (cudaStream_t)stream;
    a.pack_table = d_theta_packed ? pack_table_for(cfg, fl, pl, lc, st) : nullptr;
    // G units per CTA share the pre_D tiles and the packed-index arithmetic: 4 when there are enough units to fill the
    // GPU with such CTAs (and the tile fits), else one unit per CTA
    int n_sms = 148;
    {
        int dev = 0;
        BNN_CUDA(cudaGetDevice(&dev));
        BNN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    BNN_REQUIRE(sample_pack_smem_bytes(1, a.d, K) <= 227 * 1024, BNN_E_CONFIG,
                "bnn_swag_sample: d=%d, K=%d exceed the shared-memory plan", a.d, K);
    if (n_units >= 4ll * n_sms && sample_pack_smem_bytes(4, a.d, K) <= 227 * 1024)
        return launch_sample_pack_g<4>(a, fl, pl, lc, st);
    if (n_units >= 2ll * n_sms && sample_pack_smem_bytes(2, a.d, K) <= 227 * 1024)
        return launch_sample_pack_g<2>(a, fl, pl, lc, st);
    return launch_sample_pack_g<1>(a, fl, pl, lc, st);
}

/* The unfused K1 (one launch for theta, one for the packed layout): kept as the cross-check of the fused kernel. */
int bnn_swag_sample_unfused(const bnn_model_config* cfg, const float* d_w_avg, const float* d_w2_avg,
                            const float* d_pre_D, int32_t n_models, int32_t K, const int32_t* d_unit_model,
                            int64_t n_units, int64_t unit_offset, int32_t samples_per_model, float scale, uint64_t seed,
                            const float* d_z1, const float* d_z2, float* d_theta, float* d_theta_packed, void* stream) {
    using namespace bnn;
    int rc = validate_config(cfg);
    if (rc != BNN_OK) return rc;
    if ((rc = check_device()) != BNN_OK) return rc;
    BNN_REQUIRE(d_w_avg && d_w2_avg && d_pre_D && d_theta, BNN_E_ARG, "bnn_swag_sample_unfused: null pointer");
    BNN_REQUIRE(n_models >= 1 && n_units >= 1 && K >= 2 && K <= MAXK, BNN_E_ARG, "bnn_swag_sample_unfused: bad sizes");
    BNN_REQUIRE((d_z1 == nullptr) == (d_z2 == nullptr), BNN_E_ARG, "bnn_swag_sample_unfused: give both z1 and z2 or neither");
    const int d = FlatLayout(cfg->n_features).d;
    const int bpu = (d + SAMPLER_THREADS * 4 - 1) / (SAMPLER_THREADS * 4);
    const int64_t blocks = n_units * bpu;
    BNN_REQUIRE(blocks < (1ll << 31), BNN_E_ARG, "bnn_swag_sample_unfused: too many units for one launch");
    const float c1 = (float)((double)scale * (1.0 / sqrt(2.0)));
    const float c2div = (float)sqrt(2.0 * (K - 1));
    cudaStream_t st = (cudaStream_t)stream;
    swag_sample_kernel<<<(unsigned)blocks, SAMPLER_THREADS, 0, st>>>(
        d_w_avg, d_w2_avg, d_pre_D, d, K, d_unit_model, unit_offset, samples_per_model, n_models, c1, scale, c2div,
        seed, d_z1, d_z2, d_theta, bpu);
    BNN_CUDA(cudaGetLastError());
    if (d_theta_packed) return launch_pack_theta(cfg, d_theta, n_units, d_theta_packed, st);
    return BNN_OK;
}

}  // extern "C"
