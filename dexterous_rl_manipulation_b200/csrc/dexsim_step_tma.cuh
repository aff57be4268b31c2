// dexsim_step_tma.cuh -- the step kernel as a TMA / mbarrier pipeline (sm_100a).
//
// Same arithmetic as step_kernel (dexsim_core.cuh: env_step), different data movement.
// The register-resident version is latency-bound: ~100 scalar LDG/STG per thread, 128 registers,
// 16 warps per SM, DRAM at ~50 % (profiles/r01_step_api.md).  Here a persistent CTA owns a ring of
// shared-memory stages; one producer lane moves whole 128-env tiles with bulk-async copies
//   loads : 2-D tensor-map boxes for the obs rows (jp+jv: 30 x 128 floats, ov: 3 x 128), the float64
//           object position (3 x 128) and SoA actions; 1-D bulk copies for the per-env vectors and
//           for AoS actions (one contiguous 7,680-byte run per tile)
//   stores: the in-place updated jp+jv box, step_count, reward and the three flag vectors
// while four compute warps work out of shared memory (conflict-free: lane == column, and the
// AoS action stride of 15 words is odd).  Latency is hidden by the stage ring, not by occupancy.
// Rows that rarely change (object position, velocity, contact flags) are still written straight
// from registers, and only when they changed.
#pragma once

#include <cuda.h>

#include "dexsim_core.cuh"

// included by dexsim_kernels.cu after finish_and_reset() is defined
namespace dexsim {

#ifndef DEXSIM_TMA_TILE
#define DEXSIM_TMA_TILE 128
#endif
constexpr int TILE = DEXSIM_TMA_TILE;              // envs per tile: 128, or 96 (24.7 KB stages, four CTAs per SM)
static_assert(TILE % 32 == 0 && TILE <= 256, "a tile is a whole number of warps and one TMA box wide");
constexpr int TMA_COMPUTE_THREADS = TILE;          // one compute group = one thread per env of a tile
__host__ __device__ constexpr int tma_threads(int groups) { return groups * TMA_COMPUTE_THREADS + 32; }
constexpr int TMA_GROUPS_MAX = 16;     // per-CTA counter staging (keeps 3 CTAs per SM with 2 stages)

// byte offsets inside one stage (all multiples of 128: TMA box destinations need 128-byte alignment)
constexpr int OFF_JPJV = 0;                         // [30][128] f32, in/out
constexpr int OFF_OV = OFF_JPJV + 30 * TILE * 4;    // [3][128] f32, in
constexpr int OFF_OP64 = OFF_OV + 3 * TILE * 4;     // [3][128] f64, in
constexpr int OFF_THR = OFF_OP64 + 3 * TILE * 8;    // [128] f64, in
constexpr int OFF_ACT = OFF_THR + TILE * 8;         // [128][15] (AoS) or [15][128] (SoA) f32, in
constexpr int OFF_DAMP = OFF_ACT + NJ * TILE * 4;   // [128] f32, in
constexpr int OFF_SC = OFF_DAMP + TILE * 4;         // [128] i32, in/out
constexpr int OFF_REWARD = OFF_SC + TILE * 4;       // [128] f32, out
constexpr int OFF_CMASK = OFF_REWARD + TILE * 4;    // [128] u8, in
constexpr int OFF_TERM = OFF_CMASK + TILE;          // [128] u8, out
constexpr int OFF_TRUNC = OFF_TERM + TILE;          // [128] u8, out
constexpr int OFF_NC = OFF_TRUNC + TILE;            // [128] u8, out
constexpr int OFF_EPRET = OFF_NC + TILE;            // [128] f64, in/out (tracking)
constexpr int OFF_EPST0 = OFF_EPRET + TILE * 8;     // [128] u32, in/out (tracking)
constexpr int OFF_EPST1 = OFF_EPST0 + TILE * 4;     // [128] u32, in/out (tracking)
constexpr int OFF_FIN = OFF_EPST1 + TILE * 4;       // [128] u8, out (auto-reset)
constexpr int STAGE_BYTES = (OFF_FIN + TILE + 127) / 128 * 128;      // every stage starts 128-byte aligned
static_assert(OFF_OV % 128 == 0 && OFF_OP64 % 128 == 0 && OFF_ACT % 128 == 0, "tensor-map box destinations");
static_assert(OFF_THR % 16 == 0 && OFF_DAMP % 16 == 0 && OFF_SC % 16 == 0 && OFF_REWARD % 16 == 0 && OFF_CMASK % 16 == 0 &&
              OFF_TERM % 16 == 0 && OFF_TRUNC % 16 == 0 && OFF_NC % 16 == 0 && OFF_EPRET % 16 == 0 && OFF_EPST0 % 16 == 0 &&
              OFF_EPST1 % 16 == 0 && OFF_FIN % 16 == 0, "1-D bulk copy destinations");

struct StepMaps {             // tensor maps live in kernel parameter space (__grid_constant__)
    CUtensorMap obs_jpjv;     // obs [45, ld] f32, box {128, 30}
    CUtensorMap obs_ov;       // obs [45, ld] f32, box {128, 3}
    CUtensorMap op64;         // op64 [3, ld] f64, box {128, 3}
    CUtensorMap act_soa;      // action [15, ld] f32, box {128, 15} (only valid for layout 0)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Columns a 1-D bulk copy of a tile may touch: the tile's envs rounded up to 32 (copy sizes must be multiples
// of 16 bytes, also for the 1-byte arrays).  This never leaves the allocation: every array has ld = round_up(n, 32)
// columns, and sub-states made by dexsim_step_host start at multiples of 1,024 envs of the parent, so
// round_up(offset + n, 32) <= parent ld.  (Bounding by `ld - base` instead is wrong for sub-states, whose `ld`
// is the PARENT's row pitch.)
__device__ __forceinline__ uint32_t tile_cols(int64_t n, int64_t base) {
    const int64_t left = n - base;
    return left >= TILE ? (uint32_t)TILE : (uint32_t)((left + 31) & ~(int64_t)31);
}

// "Stage filled" barriers.  A parity wait can only tell the last two phases of a barrier apart, so every barrier
// must be watched phase by phase by ONE waiter.  With two compute groups a stage alternates between them
// (3 stages, tile k -> stage k % 3, group k % 2), hence one barrier per (stage, group): tile k uses barrier
// k % (S * G) and it is that barrier's (k / (S * G))-th fill.  (S and G are coprime or G == 1.)
template <int STAGES, int GROUPS>
__device__ __forceinline__ int full_index(int k) { return GROUPS == 1 ? k % STAGES : k % (STAGES * GROUPS); }
template <int STAGES, int GROUPS>
__device__ __forceinline__ uint32_t full_parity(int k) {
    return (uint32_t)((GROUPS == 1 ? k / STAGES : k / (STAGES * GROUPS)) & 1);
}

// DENSE: reward type.  AOS: action layout [n,15].  TRACK: 0 = plain step; 1 = episode tracking (return + history
// summary per env -> failure labels), auto-reset, counters; 2 = auto-reset and counters only (what a curriculum
// needs: episodes, successes, lengths) without the per-env return / history arrays.
// GROUPS: compute groups per CTA.  With 2 groups (8 compute warps, 3 stages, 2 CTAs per SM) group g works on the
// CTA's tiles k = g, g + 2, ... so that two tiles are in their compute phase while a third one loads.
template <bool DENSE, bool AOS, int TRACK, int STAGES, int GROUPS>
__global__ void __launch_bounds__(tma_threads(GROUPS), (GROUPS == 2) ? 2 : ((STAGES == 2) ? (TILE <= 96 ? 4 : 3) : 2))
step_tma_kernel(const DexsimState st, const DexsimParams p, const DexsimGroup* __restrict__ groups,
                const uint16_t* __restrict__ group_of_env, const DexsimStepIO io,
                const __grid_constant__ StepMaps maps, const int num_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stage_base = smem;
    constexpr int NFULL = STAGES * GROUPS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);      // full[S * G], out_ready[S]
    unsigned long long* sh_cnt = reinterpret_cast<unsigned long long*>(bars + NFULL + STAGES);
    double* sh_rs = reinterpret_cast<double*>(sh_cnt + (TRACK ? TMA_GROUPS_MAX * DEXSIM_NCOUNTERS : 0));

    constexpr int TMA_THREADS = tma_threads(GROUPS);
    constexpr int PRODUCER_TID = GROUPS * TMA_COMPUTE_THREADS;
    const int tid = threadIdx.x;
    const int64_t n = st.n, ld = st.ld;
    const bool count_episodes = TRACK && p.auto_reset && io.counters != nullptr;
    const bool staged_cnt = count_episodes && p.num_groups <= TMA_GROUPS_MAX;

    if (tid == 0) {
        for (int b = 0; b < NFULL; ++b) mbar_init(smem_u32(&bars[b]), 1);
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[NFULL + s]), TMA_COMPUTE_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (TRACK && staged_cnt) {
        for (int w = tid; w < p.num_groups * DEXSIM_NCOUNTERS; w += TMA_THREADS) sh_cnt[w] = 0ull;
        for (int w = tid; w < p.num_groups * 2; w += TMA_THREADS) sh_rs[w] = 0.0;
    }
    __syncthreads();

    const int my_tiles = (num_tiles > (int)blockIdx.x) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid >= PRODUCER_TID) {
        // ===== producer warp: one lane issues every bulk copy of this CTA =====
        if (tid == PRODUCER_TID) {
            auto issue_stores = [&](int k) {
                const int s = k % STAGES;
                const int64_t base = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * TILE;
                const uint32_t sb = smem_u32(stage_base + (size_t)s * STAGE_BYTES);
                const uint32_t cols = tile_cols(n, base);
                tma_store_2d(&maps.obs_jpjv, (int)base, DEXSIM_ROW_JP, sb + OFF_JPJV);
                bulk_store(st.step_count + base, sb + OFF_SC, cols * 4);
                bulk_store(io.reward + base, sb + OFF_REWARD, cols * 4);
                bulk_store(io.terminated + base, sb + OFF_TERM, cols);
                bulk_store(io.truncated + base, sb + OFF_TRUNC, cols);
                bulk_store(io.num_contacts + base, sb + OFF_NC, cols);
                if (TRACK == 1) {
                    bulk_store(st.ep_return + base, sb + OFF_EPRET, cols * 8);
                    bulk_store(st.ep_stats + base, sb + OFF_EPST0, cols * 4);
                    bulk_store(st.ep_stats + ld + base, sb + OFF_EPST1, cols * 4);
                }
                if (TRACK && io.finished) bulk_store(io.finished + base, sb + OFF_FIN, cols);
                bulk_commit();
            };
            for (int k = 0; k < my_tiles; ++k) {
                const int s = k % STAGES, use = k / STAGES;
                const uint32_t sb = smem_u32(stage_base + (size_t)s * STAGE_BYTES);
                const uint32_t full = smem_u32(&bars[full_index<STAGES, GROUPS>(k)]);
                if (use > 0) {
                    // the tile that used this stage last: wait for its outputs, send them, and let
                    // the copy engine finish READING the stage before new data lands in it
                    mbar_wait(smem_u32(&bars[NFULL + s]), (uint32_t)((use - 1) & 1));
                    issue_stores(k - STAGES);
                    bulk_wait_read0();
                }
                const int64_t base = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * TILE;
                const uint32_t cols = tile_cols(n, base);
                const bool full_tile = (n - base) >= TILE;
                uint32_t tx = (30 + 3) * TILE * 4 + 3 * TILE * 8 + cols * (8 + 4 + 4 + 1);
                if (AOS) tx += full_tile ? NJ * TILE * 4 : 0;
                else tx += NJ * TILE * 4;
                if (TRACK == 1) tx += cols * (8 + 4 + 4);
                mbar_expect_tx(full, tx);
                tma_load_2d(sb + OFF_JPJV, &maps.obs_jpjv, (int)base, DEXSIM_ROW_JP, full);
                tma_load_2d(sb + OFF_OV, &maps.obs_ov, (int)base, DEXSIM_ROW_OV, full);
                tma_load_2d(sb + OFF_OP64, &maps.op64, (int)base, 0, full);
                if (AOS) { if (full_tile) bulk_load(sb + OFF_ACT, io.action + base * NJ, NJ * TILE * 4, full); }
                else tma_load_2d(sb + OFF_ACT, &maps.act_soa, (int)base, 0, full);
                bulk_load(sb + OFF_THR, st.thr + base, cols * 8, full);
                bulk_load(sb + OFF_DAMP, st.damp + base, cols * 4, full);
                bulk_load(sb + OFF_SC, st.step_count + base, cols * 4, full);
                bulk_load(sb + OFF_CMASK, st.cmask + base, cols, full);
                if (TRACK == 1) {
                    bulk_load(sb + OFF_EPRET, st.ep_return + base, cols * 8, full);
                    bulk_load(sb + OFF_EPST0, st.ep_stats + base, cols * 4, full);
                    bulk_load(sb + OFF_EPST1, st.ep_stats + ld + base, cols * 4, full);
                }
            }
            // drain: outputs of the last min(STAGES, my_tiles) tiles
            const int first = my_tiles > STAGES ? my_tiles - STAGES : 0;
            for (int k = first; k < my_tiles; ++k) {
                mbar_wait(smem_u32(&bars[NFULL + k % STAGES]), (uint32_t)((k / STAGES) & 1));
                issue_stores(k);
            }
            bulk_wait0();       // shared memory must outlive every outstanding bulk store
        }
    } else {
        // ===== compute warps: lane == column of the tile =====
        const int col = tid % TMA_COMPUTE_THREADS;
        for (int k = tid / TMA_COMPUTE_THREADS; k < my_tiles; k += GROUPS) {
            const int s = k % STAGES;
            unsigned char* sp = stage_base + (size_t)s * STAGE_BYTES;
            const int64_t base = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * TILE;
            const int64_t i = base + col;
            mbar_wait(smem_u32(&bars[full_index<STAGES, GROUPS>(k)]), full_parity<STAGES, GROUPS>(k));
            if (i < n) {
                float* s_jpjv = reinterpret_cast<float*>(sp + OFF_JPJV);
                const float* s_ov = reinterpret_cast<const float*>(sp + OFF_OV);
                const double* s_op = reinterpret_cast<const double*>(sp + OFF_OP64);
                EnvRegs e;
#pragma unroll
                for (int j = 0; j < NJ; ++j) { e.jp[j] = s_jpjv[j * TILE + col]; e.jv[j] = s_jpjv[(NJ + j) * TILE + col]; }
#pragma unroll
                for (int c = 0; c < 3; ++c) { e.ov[c] = s_ov[c * TILE + col]; e.op[c] = s_op[c * TILE + col]; }
                e.thr = reinterpret_cast<const double*>(sp + OFF_THR)[col];
                e.damp = reinterpret_cast<const float*>(sp + OFF_DAMP)[col];
                e.sc = reinterpret_cast<const int*>(sp + OFF_SC)[col];
                e.cmask = reinterpret_cast<const uint8_t*>(sp + OFF_CMASK)[col];
                float a[NJ];
                const float* s_act = reinterpret_cast<const float*>(sp + OFF_ACT);
                if (AOS) {
                    if ((n - base) >= TILE) {
#pragma unroll
                        for (int j = 0; j < NJ; ++j) a[j] = s_act[col * NJ + j];
                    } else {        // ragged last tile: its byte count need not be a multiple of 16
#pragma unroll
                        for (int j = 0; j < NJ; ++j) a[j] = __ldg(io.action + i * NJ + j);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) a[j] = s_act[j * TILE + col];
                }
                double op_old[3]; float ov_old[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { op_old[c] = e.op[c]; ov_old[c] = e.ov[c]; }
                const unsigned cmask_old = e.cmask;

                StepResult r;
                env_step<DENSE>(e, a, p, r);

                reinterpret_cast<float*>(sp + OFF_REWARD)[col] = (float)r.total;
                (sp + OFF_TERM)[col] = r.terminated ? 1 : 0;
                (sp + OFF_TRUNC)[col] = r.truncated ? 1 : 0;
                (sp + OFF_NC)[col] = (unsigned char)r.n_c;

                bool did_reset = false;
                if (TRACK) {
                    double ep_return = 0.0;
                    EpStats es{0u, 0u};
                    if (TRACK == 1) {
                        ep_return = __dadd_rn(reinterpret_cast<double*>(sp + OFF_EPRET)[col], r.total);
                        es.w0 = reinterpret_cast<uint32_t*>(sp + OFF_EPST0)[col];
                        es.w1 = reinterpret_cast<uint32_t*>(sp + OFF_EPST1)[col];
                        epstats_push(es, e.sc - 1, r.n_c);
                    }
                    const bool done = r.terminated || r.truncated || (p.loop_max_steps > 0 && e.sc >= p.loop_max_steps);
                    if (p.auto_reset && done) {
                        did_reset = true;
                        const int64_t gid = p.env_gid0 + i;
                        const int g = group_of_env ? (int)group_of_env[i] : (int)((uint32_t)gid % (uint32_t)p.num_groups);
                        uint32_t episode = st.episode[i];
                        double size, mass, friction;
                        unsigned long long* cnt = nullptr;
                        double* rs = nullptr;
                        if (count_episodes) {
                            cnt = (staged_cnt ? sh_cnt : reinterpret_cast<unsigned long long*>(io.counters)) + (int64_t)g * DEXSIM_NCOUNTERS;
                            if (io.ret_sums) rs = (staged_cnt ? sh_rs : io.ret_sums) + 2 * g;
                        }
                        finish_and_reset(e, p, groups[g], (uint32_t)gid, episode, ep_return, es, r.terminated, r.n_c,
                                         cnt, rs, size, mass, friction, nullptr, /*classify=*/TRACK == 1);
                        st.episode[i] = episode;
                        st.size[i] = size; st.mass[i] = mass; st.friction[i] = friction;
                        st.thr[i] = e.thr; st.damp[i] = e.damp;
                    }
                    if (TRACK == 1) {
                        reinterpret_cast<double*>(sp + OFF_EPRET)[col] = ep_return;
                        reinterpret_cast<uint32_t*>(sp + OFF_EPST0)[col] = es.w0;
                        reinterpret_cast<uint32_t*>(sp + OFF_EPST1)[col] = es.w1;
                    }
                    (sp + OFF_FIN)[col] = did_reset ? 1 : 0;
                }
                // always-changing state goes back through the stage (one bulk store per tile)
#pragma unroll
                for (int j = 0; j < NJ; ++j) { s_jpjv[j * TILE + col] = e.jp[j]; s_jpjv[(NJ + j) * TILE + col] = e.jv[j]; }
                reinterpret_cast<int*>(sp + OFF_SC)[col] = e.sc;
                // rarely-changing rows: straight from registers, only when they changed
                float* __restrict__ obs = st.obs;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (__double_as_longlong(e.op[c]) != __double_as_longlong(op_old[c])) {
                        st.op64[c * ld + i] = e.op[c];
                        obs[(DEXSIM_ROW_OP + c) * ld + i] = (float)e.op[c];
                    }
                    if (__float_as_uint(e.ov[c]) != __float_as_uint(ov_old[c])) obs[(DEXSIM_ROW_OV + c) * ld + i] = e.ov[c];
                }
                const unsigned flip = e.cmask ^ cmask_old;
                if (flip) {
#pragma unroll
                    for (int f = 0; f < NF; ++f)
                        if ((flip >> f) & 1u) obs[(DEXSIM_ROW_CONTACT + f) * ld + i] = ((e.cmask >> f) & 1u) ? 1.0f : 0.0f;
                    st.cmask[i] = (uint8_t)e.cmask;
                }
            }
            fence_async_smem();                       // generic-proxy writes -> visible to the copy engine
            mbar_arrive(smem_u32(&bars[NFULL + s]));
        }
    }
    if (TRACK && staged_cnt) {
        __syncthreads();
        unsigned long long* gc = reinterpret_cast<unsigned long long*>(io.counters);
        for (int w = tid; w < p.num_groups * DEXSIM_NCOUNTERS; w += TMA_THREADS)
            if (sh_cnt[w]) atomicAdd(&gc[w], sh_cnt[w]);
        if (io.ret_sums)
            for (int w = tid; w < p.num_groups * 2; w += TMA_THREADS)
                if (sh_rs[w] != 0.0) atomicAdd(&io.ret_sums[w], sh_rs[w]);
    }
}

// ---- host: tensor maps ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D row-major [rows, n] view with row pitch ld elements; box = {TILE columns, box_rows rows}.
static bool make_map_2d(CUtensorMap* m, CUtensorMapDataType dt, size_t elem, void* base, int64_t n, int64_t ld,
                        int rows, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * elem};
    const cuuint32_t box[2] = {(cuuint32_t)TILE, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dexsim
