// Kernel + launcher of the tensor-core predictive path (included by predict.cu after PredictParams).
#pragma once
#include "predict_tc.cuh"

namespace bnn {
namespace tc {

// timeline probe (tools/tc_timeline.py): role-major [8][512] int64 of clock64() stamps, CTA 0 lane 0 only.  Compiled in
// only with -DBNN_TC_TIMELINE (make TIMELINE=1): the stamps cost registers in the hot loops.
#ifdef BNN_TC_TIMELINE
#define TC_STAMP(role, code)                                                                          \
    do {                                                                                             \
        if (prm.dbg && blockIdx.x == 0 && lane == 0 && dbg_n < 511) {                                 \
            prm.dbg[(role) * 512 + 1 + dbg_n] = ((long long)(code) << 48) | (clock64() & 0xFFFFFFFFFFFFll); \
            prm.dbg[(role) * 512] = ++dbg_n;                                                          \
        }                                                                                            \
    } while (0)
#else
#define TC_STAMP(role, code) do { } while (0)
#endif

// One unit's ring slot: hi + lo B operands of the three layers + the bias block.  K1 = 32 (one layer-1 pass) or, in the
// WIDE variant (33..48 live inputs), 48: a K = 40 pass over A's 40 columns, then a second pass of 8 restaged columns.
template <bool WIDE>
struct RingPlan {
    static constexpr int K1 = WIDE ? TC_K1W : TC_K1;
    static constexpr int B1_BYTES = WIDE ? K1 * TC_N * 4 : K1 * TC_N * 2;   // tf32 (wide) or fp16 (kind::f16 layer 1) operands
    static constexpr int B_FLOATS = 2 * (B1_BYTES / 4 + TC_K2 * TC_N + TC_K2 * TC_N3) + TC_BIAS;
    static constexpr int B_BYTES = B_FLOATS * 4;
    static constexpr int O_B1H = 0, O_B1L = O_B1H + B1_BYTES, O_B2H = O_B1L + B1_BYTES,
                         O_B2L = O_B2H + TC_K2 * TC_N * 4, O_B3H = O_B2L + TC_K2 * TC_N * 4,
                         O_B3L = O_B3H + TC_K2 * TC_N3 * 4, O_BIAS = O_B3L + TC_K2 * TC_N3 * 4;  // byte offsets in a slot
    static constexpr int NREC = WIDE ? 2 : tc::NREC;   // record-ring depth: the wide ring takes the shared memory of two slots
};

constexpr int N_TEAM = 2;              // epilogue teams; team t owns TMEM slots 2t and 2t+1 and the M tiles t and t+2
constexpr int EW = 8 * N_TEAM;         // epilogue warps
constexpr int NISS = N_SLOT;           // MMA-issuing warps (one per TMEM slot)
constexpr int HC = 20;                 // hidden columns per epilogue warp (the two warps of a lane quadrant split 40)

template <int NT, bool WIDE>
struct SmemPlan {
    using RG = RingPlan<WIDE>;
    // byte offsets inside dynamic shared memory
    static constexpr int xs = 0;                                   // XS_ROWS*32 floats
    static constexpr int ring = xs + XS_ROWS * 32 * 4;             // 2 slots of B operands + biases
    static constexpr int fb = ring + 2 * RG::B_BYTES;                  // per epilogue warp: 32 rows x 12 latent columns
    static constexpr int rec = fb + EW * FB_FLOATS * 4;            // NREC slots of block records
    static constexpr int scratch = rec + RG::NREC * REC_FLOATS * 4;    // per tail warp
    static constexpr int head = (scratch + NT * TAIL_SCRATCH * 4 + 127) & ~127;  // per tail warp: head block of its unit
    static constexpr int head_bytes = (HEAD_FLOATS * 4 + 127) & ~127;
    static constexpr int bars = head + NT * head_bytes;
    static constexpr int total = bars + (int)sizeof(Bars);
};

// Warp roles (768 threads at NT = 2 -> 80 registers per thread):
//   [0, 16)         epilogue warps: team = w / 8, half = (w / 4) % 2, TMEM lane quadrant = w % 4 (thread = tile row).  A team
//                   works on TWO jobs at a time -- (unit i, M tile team) in slot 2*team and (unit i, M tile team + 2) in slot
//                   2*team + 1 -- and alternates between them phase by phase, so that the tensor pipe runs one slot's layer
//                   while the team's CUDA cores run the other slot's epilogue; the two warps of a quadrant split the
//                   columns of a phase (hidden 20 + 20, x 16 + 16, latent 10 + 10).  With one job per 4 warps (round 1)
//                   a warp idled through its slot's MMAs: a third of its time, 2.6 of 4 warps per scheduler active.
//   [16, 20)        issuer warps: issuer s issues every tcgen05.mma of TMEM slot s (it sleeps on the slot's "A ready"
//                   mbarrier); issuer 0 also refills the B-operand ring.  One issuer per slot, not per team: a thread
//                   issues one N = 48 MMA per 44 cycles while the pipe takes two threads' MMAs at one per 33
//   [20, 20 + 2 NT) tail warps: NT pairs (stats warp, head warp), unit i -> pair i % NT.  The stats warp draws eps, merges the
//                   unit's block records and samples the summary statistics (then the record slot is free again); the head
//                   warp runs regress_nn from its bulk-copied weight block.  The tails are 12 % of the instructions, but as
//                   two warps they were the kernel's critical path (ncu: 100 % busy at 0.14 IPC behind 16 epilogue warps)
template <int NT, bool WIDE>
__global__ void __launch_bounds__((EW + NISS + 2 * NT) * 32, 1)
predict_tc_kernel(const PredictParams prm, const int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_tc[];
    static_assert(NT >= 1 && NT <= 2 && NT <= RingPlan<WIDE>::NREC && MT == 2 * N_TEAM && N_SLOT == 2 * N_TEAM && N_SLOT * TM_SLOT <= 512, "role layout");
    const PackedLayout pl(prm.kin, prm.F);
    using Plan = SmemPlan<NT, WIDE>;
    using RG = RingPlan<WIDE>;
    constexpr int NREC = RG::NREC, B_FLOATS = RG::B_FLOATS, B_BYTES = RG::B_BYTES, O_B1H = RG::O_B1H, O_B1L = RG::O_B1L,
                  O_B2H = RG::O_B2H, O_B2L = RG::O_B2L, O_B3H = RG::O_B3H, O_B3L = RG::O_B3L, O_BIAS = RG::O_BIAS;
    float* xs = reinterpret_cast<float*>(smem_tc + Plan::xs);
    float* ring = reinterpret_cast<float*>(smem_tc + Plan::ring);
    float* fb = reinterpret_cast<float*>(smem_tc + Plan::fb);
    float* rec = reinterpret_cast<float*>(smem_tc + Plan::rec);
    float* scratch = reinterpret_cast<float*>(smem_tc + Plan::scratch);
    static_assert(HEAD_FLOATS % 4 == 0, "bulk copies move multiples of 16 bytes");
    Bars* bars = reinterpret_cast<Bars*>(smem_tc + Plan::bars);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W_ISS = EW, W_TAIL = EW + NISS;
#ifdef BNN_TC_TIMELINE
    int dbg_n = 0;
#endif

    // ---- one-time setup: TMEM allocation ----
    if (warp == 0) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    // Static, exactly balanced partition: the launch's (system tile, unit) pairs in tile-major order are cut into gridDim.x
    // equal contiguous ranges (every pair costs the same, so equal counts end together: no partial last wave whatever the
    // shape of the launch, and a CTA loads ceil(range / U) + 1 x tiles at most).  An item = the part of one tile's units
    // that lies in the CTA's range; only the first item of a CTA starts behind unit 0 of its tile.
    int tile, u_first, rem;
    {
        const int64_t W = (int64_t)n_tiles * prm.U;
        const int64_t w0 = W * blockIdx.x / gridDim.x, w1 = W * (blockIdx.x + 1) / gridDim.x;
        tile = (int)(w0 / prm.U);
        u_first = (int)(w0 - (int64_t)tile * prm.U);
        rem = (int)(w1 - w0);
    }
    for (int first_item = 1; rem > 0; first_item = 0) {
        const int64_t n0 = (int64_t)tile * SYS;
        const int n_valid = (int)min((int64_t)SYS, prm.N - n0);
        const int64_t u_begin = u_first;
        const int n_units = (int)min((int64_t)rem, prm.U - u_begin);

        tc_fence_before();
        __syncthreads();  // previous item drained by every role
        if (threadIdx.x == 0) {
            if (rem > n_units) {
                // the next item's x tile (82 kB of a 164 MB array, HBM-cold) starts moving to L2 now, a whole item ahead of its load
                const int64_t n1 = n0 + SYS;
                const int64_t nv = min((int64_t)SYS, prm.N - n1);
                const int64_t bytes = nv * T_FIXED * prm.F * 4;
                const float* p1 = prm.X + n1 * (int64_t)T_FIXED * prm.F;
                if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(p1) & 15) == 0) bulk_prefetch_l2(p1, (uint32_t)bytes);
            }
            if (!first_item) {
                for (int s = 0; s < 2; ++s) mbar_inval(&bars->w_full[s]);
                for (int s = 0; s < NREC; ++s) { mbar_inval(&bars->unit_done[s]); mbar_inval(&bars->rec_free[s]); }
                for (int s = 0; s < N_SLOT; ++s) { mbar_inval(&bars->d_ready[s]); mbar_inval(&bars->a_ready[s]); }
                for (int s = 0; s < NT; ++s) { mbar_inval(&bars->h_full[s]); mbar_inval(&bars->sum_full[s]); mbar_inval(&bars->sum_free[s]); }
            }
            for (int s = 0; s < NT; ++s) { mbar_init(&bars->h_full[s], 1); mbar_init(&bars->sum_full[s], 1); mbar_init(&bars->sum_free[s], 1); }
            for (int s = 0; s < 2; ++s) mbar_init(&bars->w_full[s], 1);
            for (int s = 0; s < NREC; ++s) {
                mbar_init(&bars->unit_done[s], EW * 2);  // every epilogue warp pools two jobs per unit
                mbar_init(&bars->rec_free[s], 1);        // the tail warp of the unit
            }
            for (int s = 0; s < N_SLOT; ++s) {
                mbar_init(&bars->d_ready[s], 1);  // tcgen05.commit
                mbar_init(&bars->a_ready[s], 8);  // one lane of each of the team's warps
            }
            mbar_init_fence();
        }
        if (WIDE)
            load_x_tile_tc(prm.X, n0, n_valid, prm.F, prm.kin, prm.cm, xs, reinterpret_cast<int*>(fb));
        else
            load_x_tile_f16(prm.X, n0, n_valid, prm.F, prm.kin, prm.cm, smem_tc + Plan::xs, reinterpret_cast<int*>(fb), bars->x_maxbits,
                            bars->x_scale);
        __syncthreads();
        tc_fence_after();

        if (warp >= W_TAIL) {
            // ---------------- tail warps: pair j = (stats warp, head warp) takes the units i = j (mod NT) ----------------
            const int tw = warp - W_TAIL, j = tw % NT;
            float* pair_scratch = scratch + j * TAIL_SCRATCH;
            float* eS = pair_scratch, * sum = pair_scratch + 208, * sB = pair_scratch + 416, * sC = pair_scratch + 624;
            if (tw < NT) {
                // stats warp: eps draws (before the records are there), record merge + sampled summary statistics
                int k = 0;
                for (int i = j; i < n_units; i += NT, ++k) {
                    const int rs = i % NREC;
                    const int64_t u = u_begin + i;
                    tail_draw_eps(prm.eps ? prm.eps + u * prm.N * S2 : nullptr, prm.seed, (uint32_t)(prm.unit_offset + u),
                                  prm.system_offset + n0, n0, n_valid, eS);
                    mbar_wait_backoff(&bars->unit_done[rs], (uint32_t)((i / NREC) & 1), 200);  // all 16 block records of unit i
                    if (k > 0) mbar_wait(&bars->sum_free[j], (uint32_t)((k - 1) & 1));          // head warp done with sum
                    tail_stats(rec + rs * REC_FLOATS, eS, prm.thp + u * pl.P + pl.lv_sum,
                               prm.eps_sum ? prm.eps_sum + u * prm.N * S2 : nullptr,
                               prm.summary ? prm.summary + u * prm.N * S2 : nullptr, n0, n_valid, sum);
                    if (lane == 0) {
                        mbar_arrive(&bars->rec_free[rs]);   // the record slot may take unit i + NREC
                        mbar_arrive(&bars->sum_full[j]);
                    }
                }
            } else {
                // head warp: regress_nn from the pair's summary buffer; the unit's head block (16 kB) comes by one bulk copy,
                // issued as soon as the buffer is free, so that it lands while the stats warp works
                float* my_head = reinterpret_cast<float*>(smem_tc + Plan::head + j * Plan::head_bytes);
                auto fetch_head = [&](int i) {
                    if (lane == 0 && i < n_units) {
                        fence_proxy_async_smem();  // the generic-proxy reads of the previous unit are ordered before the copy
                        mbar_arrive_expect_tx(&bars->h_full[j], (uint32_t)(HEAD_FLOATS * 4));
                        bulk_g2s(my_head, prm.thp + (u_begin + i) * pl.P + pl.V0p, (uint32_t)(HEAD_FLOATS * 4), &bars->h_full[j]);
                    }
                };
                fetch_head(j);
                uint32_t ph = 0;
                for (int i = j; i < n_units; i += NT) {
                    const int64_t u = u_begin + i;
                    mbar_wait(&bars->h_full[j], ph);
                    mbar_wait(&bars->sum_full[j], ph);
                    ph ^= 1;
                    tail_head(sum, my_head, pl, n0, n_valid, prm.hc, sB, sC, prm.out + u * prm.out_unit_stride,
                              prm.out_sys_stride, [&]() { if (lane == 0) mbar_arrive(&bars->sum_free[j]); });
                    fetch_head(i + NT);
                }
            }
        } else if (warp >= W_ISS) {
            // ---------------- issuer warps: every tcgen05.mma of the team's two slots; ring refill ----------------
            const int slot = warp - W_ISS;
            const uint32_t ring_addr = smem_u32(ring);
            uint32_t pa = 0;  // parity of a_ready[slot]
            // B operands + biases of unit i -> ring slot i & 1, free once every block record of unit i-2 is written
            // (all its MMAs and bias reads are done).  Issued by one lane of issuer 0 whenever it passes.
            int next_load = 0;
            auto refill = [&]() {
                if (slot == 0 && lane == 0) {
                    while (next_load < n_units &&
                           (next_load < 2 ||
                            mbar_test(&bars->unit_done[(next_load - 2) % NREC], (uint32_t)(((next_load - 2) / NREC) & 1)))) {
                        mbar_arrive_expect_tx(&bars->w_full[next_load & 1], (uint32_t)B_BYTES);
                        bulk_g2s(ring + (size_t)(next_load & 1) * B_FLOATS, prm.thp + (u_begin + next_load) * pl.P + pl.B1h,
                                 (uint32_t)B_BYTES, &bars->w_full[next_load & 1]);
                        ++next_load;
                    }
                }
                __syncwarp();
            };
            refill();
#pragma unroll 1
            for (int i = 0; i < n_units; ++i) {
                const int ws = i & 1;
                // every issuer visits every unit in order, so its view of w_full[ws] never skips a phase
                while (!mbar_test(&bars->w_full[ws], (uint32_t)((i >> 1) & 1))) {
                    refill();
                    __nanosleep(40);
                }
                __syncwarp();
#pragma unroll 1
                for (int layer = 0; layer < 3; ++layer) {
                    mbar_wait(&bars->a_ready[slot], pa);
                    pa ^= 1u;
                    tc_fence_after();
                    TC_STAMP(2 + slot, 10 * layer + 1);
                    uint32_t wb = ring_addr + (uint32_t)ws * (uint32_t)B_BYTES;
                    uint32_t tsv = tmem + slot * TM_SLOT;
                    asm volatile("" : "+r"(wb), "+r"(tsv));  // keep the descriptors out of loop-invariant hoisting
                    if (layer == 0) {
                        if (WIDE) {
                            // K = 40 over A's 40 columns, then (after the epilogue warps restaged 8 columns) K steps 5
                            issue_layer_part<TC_N, 0, 5, true>(tsv, wb + O_B1H, wb + O_B1L);
                            if (elect_one_sync()) mma_commit(&bars->d_ready[slot]);
                            __syncwarp();
                            mbar_wait(&bars->a_ready[slot], pa);
                            pa ^= 1u;
                            tc_fence_after();
                            issue_layer_part<TC_N, 5, 6, false>(tsv, wb + O_B1H, wb + O_B1L);
                        } else {
                            // A straight from the tile's fp16 hi / lo arrays: M tile of this slot = team + 2 (slot in team)
                            uint32_t xa = smem_u32(smem_tc + Plan::xs) + (uint32_t)(((slot >> 1) + 2 * (slot & 1)) * 4 * 2048);
                            asm volatile("" : "+r"(xa));
                            issue_layer1_f16<TC_N>(tsv, xa, xa + (uint32_t)XH_BYTES, wb + O_B1H, wb + O_B1L);
                        }
                    } else if (layer == 1)
                        issue_layer<TC_N, TC_K2 / 8>(tsv, wb + O_B2H, wb + O_B2L);
                    else
                        issue_layer<TC_N3, TC_K2 / 8>(tsv, wb + O_B3H, wb + O_B3L);
                    if (elect_one_sync()) mma_commit(&bars->d_ready[slot]);
                    __syncwarp();
                    TC_STAMP(2 + slot, 10 * layer + 2);
                    refill();
                }
            }
        } else {
            // ---------------- epilogue warps ----------------
            const int team = warp >> 3, half = (warp >> 2) & 1, quad = warp & 3;
            const uint32_t tq = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(2 * team * TM_SLOT);  // slot s: + s * TM_SLOT
            float* my_fb = fb + warp * FB_FLOATS;
            uint32_t pd = 0;  // wide variant (an extra commit in layer 1): bit s = parity of d_ready[2 * team + s]

            // the slot's commit number c = 3 unit + layer completes phase c of d_ready: its parity follows from (unit, layer) --
            // (i + layer) & 1 for the hidden-layer epilogues, (i + 1) & 1 for the read-out of unit i - 1 -- without per-slot state
            auto wait_d = [&](int s, uint32_t par) {
                if (WIDE) {
                    par = (pd >> s) & 1u;
                    pd ^= 1u << s;
                }
                mbar_wait(&bars->d_ready[2 * team + s], par);
                tc_fence_after();
            };
            // A of the slot's next layer is written: TMEM stores complete, then one arrival per warp on the issuer's barrier
            auto publish = [&](int s) {
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a_ready[2 * team + s]);
            };
            // layer 3 of the slot's job: D -> this warp's 10 latent columns [10 half, 10 half + 10) of its 32 rows, column-major
            // (the bias is added to the pooled mean: the deviations do not see it)
            auto latent_to_fb = [&](int s) {
                const uint32_t tl = tq + s * TM_SLOT;
                uint32_t d0[8], d1[2];
                tmem_ld8(tl + TM_D + 10 * half, d0);
                tmem_ld2(tl + TM_D + 10 * half + 8, d1);
                tc_wait_ld();
                uint32_t* dst = reinterpret_cast<uint32_t*>(my_fb) + lane;
#pragma unroll
                for (int c = 0; c < 8; ++c) dst[c * FB_PITCH] = d0[c];
                dst[8 * FB_PITCH] = d1[0];
                dst[9 * FB_PITCH] = d1[1];
                __syncwarp();
            };
            // pooled (mean, M2) records of this warp's 10 latent columns of one 32-row block, two-pass per segment like
            // torch.mean / torch.std.  Every boundary inside a block (the next system's first row at 4, 8, 12 or 16, the tile's
            // last row at 20) is a multiple of 4 rows, so a 4-row granule belongs to one segment: lane = (column c = lane % 10,
            // granules lane / 10 + {0, 3, 6}), one predicate per granule instead of per row, and the three lanes of a column
            // meet through shuffles.  Runs in the shadow of the slot's layer-1 MMAs.  Lane constants (column pointer, first
            // granule row) live in two registers, the block's geometry comes from constant memory.
            const int pj3 = lane / 10, pc = lane - 10 * pj3;
            const float* pool_col = my_fb + pc * FB_PITCH + 4 * min(pj3, 2);   // granule i of this lane: 16 bytes at pool_col + 12 i
            const int pool_r0 = lane < 30 ? 4 * pj3 : 64;     // rows of granule i: pool_r0 + 12 i ...; lanes 30, 31 idle
            auto pool_block = [&](int pi, int m) {
                static_assert(T_FIXED % 4 == 0 && ROWS % 4 == 0, "4-row granules must not straddle systems");
                const int rs = pi % NREC;
                if (pi >= NREC) mbar_wait(&bars->rec_free[rs], (uint32_t)((pi / NREC - 1) & 1));  // tail of unit pi-NREC done (suspended wait:
                                                                                              // a polling loop here took 39 % of all issued instructions)
                const int b = m * 4 + quad;
                const PoolGeom pg = c_pool_geom[b];
                const bool two = pg.nvalid > pg.e0;            // the block holds rows of two systems (warp-uniform)
                u64 vlo[3], vhi[3];                            // rows (0, 1) and (2, 3) of granule i
                bool in0[3], in1[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int r0 = pool_r0 + 12 * i;           // r0 >= 32 is beyond nvalid: both predicates false (the load
                    in0[i] = r0 < pg.e0;                       // then reads the column's pad; its values are dropped)
                    in1[i] = !in0[i] && r0 < pg.nvalid;
                    const ulonglong2 g = *reinterpret_cast<const ulonglong2*>(pool_col + 12 * i);
                    vlo[i] = g.x;
                    vhi[i] = g.y;
                }
                auto gather3 = [&](float x) {
                    return __shfl_sync(0xffffffffu, x, pc) + __shfl_sync(0xffffffffu, x, pc + 10) + __shfl_sync(0xffffffffu, x, pc + 20);
                };
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    float pa, pb;
                    unpack2(add2(vlo[i], vhi[i]), pa, pb);     // packed fp32x2: (row 0 + row 2, row 1 + row 3)
                    const float pr = pa + pb;
                    if (in0[i]) s0 += pr;                      // (predicated adds: a granule belongs to at most one segment)
                    if (in1[i]) s1 += pr;
                }
                // S / n as S r corrected by the residual (r = the rounded reciprocal from the table): the correctly rounded
                // quotient without the ~60-cycle divide sequence on the warp's critical path (exact for power-of-two counts)
                auto div_n = [](float S, float n, float r) {
                    const float q = S * r;
                    return fmaf(fmaf(-n, q, S), r, q);
                };
                const float mean0 = div_n(gather3(s0), pg.n0, pg.rcp0);
                const float mean1 = two ? div_n(gather3(s1), pg.n1, pg.rcp1) : 0.f;
                float q0 = 0.f, q1 = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float mu = in1[i] ? mean1 : mean0;
                    const u64 mu2 = pack2(mu, mu);
                    const u64 dlo = sub2(vlo[i], mu2), dhi = sub2(vhi[i], mu2);
                    float qa, qb;
                    unpack2(fma2(dhi, dhi, mul2(dlo, dlo)), qa, qb);
                    const float qp = qa + qb;
                    if (in0[i]) q0 += qp;
                    if (in1[i]) q1 += qp;
                }
                q0 = gather3(q0);
                if (two) q1 = gather3(q1);
                const float bias_c = ring[(size_t)(pi & 1) * B_FLOATS + O_BIAS / 4 + 2 * TC_N + 10 * half + pc];  // b2 of this column
                float* rb = rec + rs * REC_FLOATS + ((b * 2) * L + 10 * half + pc) * 2;
                if (lane < 10) *reinterpret_cast<float2*>(rb) = make_float2(mean0 + bias_c, q0);
                if (two && lane >= 10 && lane < 20) *reinterpret_cast<float2*>(rb + L * 2) = make_float2(mean1 + bias_c, q1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->unit_done[rs]);
            };

            // down-scaling exponents of this thread's rows in the team's two slots (0 unless the row's system holds |x| >= 2^15):
            // the layer-1 epilogue multiplies the accumulator by 2^kk; one register for the whole item
            int x_kk = 0;
            if (!WIDE) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int sys = min(((team + 2 * s) * 128 + quad * 32 + lane) / T_FIXED, SYS - 1);
                    const int kk = (int)((__float_as_uint(bars->x_scale[sys]) >> 23) - 127u);
                    x_kk |= kk << (8 * s);
                    // bit 16 + s: no row of this warp's 32-row block is scaled -> its layer-1 bias comes out of the MMA (warp-uniform:
                    // the epilogue's tcgen05.ld / st are .sync.aligned)
                    if (__all_sync(0xffffffffu, kk == 0)) x_kk |= 1 << (16 + s);
                }
            }
#pragma unroll 1
            for (int i = 0; i <= n_units; ++i) {
                // ---- phase 0 of both slots: latent rows of the previous unit -> fb, x -> A (layer 1 starts), pooling ----
#pragma unroll 1
                for (int s = 0; s < 2; ++s) {
                    const int m = team + 2 * s;
                    const uint32_t tl = tq + s * TM_SLOT;
                    if ((warp & 7) == 0) TC_STAMP(team, 100 * s + 1);
                    if (i > 0) {
                        wait_d(s, (uint32_t)(i + 1) & 1u);
                        if ((warp & 7) == 0) TC_STAMP(team, 100 * s + 2);
                        latent_to_fb(s);
                    }
                    if (i < n_units && !WIDE) {
                        // layer 1 reads x from shared memory: all it needs from this warp is that the slot's accumulator is
                        // free again (the previous unit's latent rows have been read out)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->a_ready[2 * team + s]);
                    }
                    if (i < n_units && WIDE) {
                        const int R = m * 128 + quad * 32 + lane;  // tile row of this thread (rows 500..511 of xs are zero)
                        uint32_t v[16];
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {
                            const uint4 a = *reinterpret_cast<const uint4*>(xs + R * 32 + (((4 * half + g4) ^ (lane & 7)) << 2));
                            v[4 * g4] = a.x; v[4 * g4 + 1] = a.y; v[4 * g4 + 2] = a.z; v[4 * g4 + 3] = a.w;
                        }
                        split_store16<0>(v, nullptr, tl + TM_AHI + 16 * half, tl + TM_ALO + 16 * half);
                        {
                            // live inputs 32..47 of this row from L2 (the x tile was just read by this CTA; they are the same
                            // for every unit, but 16 more columns of shared memory do not exist): half 0 puts 32..39 behind the
                            // 32 staged columns (first pass, K = 40); half 1 restages 40..47 into columns 0..7 once the first
                            // pass has been read by the tensor core
                            uint32_t w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int c = prm.wide_col[8 * half + j];
                                w[j] = (c >= 0 && R < n_valid * T_FIXED) ? __float_as_uint(__ldg(prm.X + (n0 * T_FIXED + R) * (int64_t)prm.F + c)) : 0u;
                            }
                            if (half == 0) {
                                split_store8<0>(w, nullptr, tl + TM_AHI + 32, tl + TM_ALO + 32);
                                publish(s);
                                wait_d(s, 0u);
                            } else {
                                publish(s);
                                wait_d(s, 0u);
                                split_store8<0>(w, nullptr, tl + TM_AHI, tl + TM_ALO);
                            }
                        }
                        publish(s);
                        if ((warp & 7) == 0) TC_STAMP(team, 100 * s + 3);
                    }
                    if (i > 0) pool_block(i - 1, m);
                    if ((warp & 7) == 0) TC_STAMP(team, 100 * s + 4);
                }
                if (i == n_units) break;
                // ---- layers 1 and 2 of both slots: D + bias -> ReLU -> hi/lo -> A ----
                const float* bias = ring + (size_t)(i & 1) * B_FLOATS + O_BIAS / 4 + HC * half;
#pragma unroll 1
                for (int layer = 0; layer < 2; ++layer) {
#pragma unroll 1
                    for (int s = 0; s < 2; ++s) {
                        const uint32_t tl = tq + s * TM_SLOT + HC * half;
                        const float* bl = bias + layer * TC_N;
                        if ((warp & 7) == 0) TC_STAMP(team, 100 * s + 10 * (layer + 1) + 1);
                        wait_d(s, (uint32_t)(i + layer) & 1u);
                        if ((warp & 7) == 0) TC_STAMP(team, 100 * s + 10 * (layer + 1) + 2);
                        uint32_t d0[16], d1[4];
                        tmem_ld16(tl + TM_D, d0);
                        tmem_ld4(tl + TM_D + 16, d1);
                        tc_wait_ld();
                        if (!WIDE && layer == 0 && prm.kin < TC_K1 && ((x_kk >> (16 + s)) & 1)) {
                            // the layer-1 bias came out of the MMA (ones column of x, see load_x_tile_f16)
                            split_store16<2>(d0, nullptr, tl + TM_AHI, tl + TM_ALO);
                            split_store4<2>(d1, nullptr, tl + TM_AHI + 16, tl + TM_ALO + 16);
                        } else {
                            // undo the power-of-two down-scaling of this row's system (1 unless the system holds |x| >= 2^15)
                            const float sc = (!WIDE && layer == 0) ? __uint_as_float((127u + ((uint32_t)(x_kk >> (8 * s)) & 0xffu)) << 23) : 1.0f;
                            const u64 sc2 = pack2(sc, sc);
                            split_store16<1>(d0, bl, tl + TM_AHI, tl + TM_ALO, sc2);
                            split_store4<1>(d1, bl + 16, tl + TM_AHI + 16, tl + TM_ALO + 16, sc2);
                        }
                        publish(s);
                        if ((warp & 7) == 0) TC_STAMP(team, 100 * s + 10 * (layer + 1) + 3);
                    }
                }
            }
        }
        rem -= n_units;
        ++tile;
        u_first = 0;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int NT, bool WIDE>
static int launch_tc(const PredictParams& prm, cudaStream_t st) {
    constexpr size_t smem = (size_t)SmemPlan<NT, WIDE>::total;
    static_assert(smem <= 227 * 1024, "tensor-core tile does not fit in shared memory");
    constexpr int threads = (EW + NISS + 2 * NT) * 32;
    static PerDeviceOnce attr_done;
    if (attr_done.need()) {
        BNN_CUDA(cudaFuncSetAttribute((predict_tc_kernel<NT, WIDE>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int n_sms = 0, dev = 0;
    BNN_CUDA(cudaGetDevice(&dev));
    BNN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = (prm.N + SYS - 1) / SYS;
    BNN_REQUIRE(tiles < (1ll << 24), BNN_E_ARG, "too many system tiles for one launch (%lld)", (long long)tiles);
    const int64_t pairs = tiles * prm.U;   // (tile, unit) pairs, cut into `grid` equal ranges by the kernel
    const int grid = (int)(pairs < n_sms ? pairs : n_sms);
    predict_tc_kernel<NT, WIDE><<<grid, threads, smem, st>>>(prm, (int)tiles);
    BNN_CUDA(cudaGetLastError());
    return BNN_OK;
}

static bool tc_fits(const PredictParams& prm, int T) {
    if (T != T_FIXED) return false;
    const PackedLayout pl(prm.kin, prm.F);
    return pl.tc_ok != 0;
}

}  // namespace tc
}  // namespace bnn
