/* c_abi_demo.c -- the C ABI of include/dexsim.h used from plain C (no Python, no torch): what a binding in any
 * host language does.  All buffers are the caller's (cudaMalloc); the library allocates nothing.
 *
 *   gcc -O2 -std=c99 -Iinclude -I/usr/local/cuda/include examples/c_abi_demo.c -o examples/c_abi_demo \
 *       -Ldexterous_rl_manipulation_b200 -ldexsim_b200 -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,'$ORIGIN/../dexterous_rl_manipulation_b200' -Wl,-rpath,/usr/local/cuda/lib64
 *   ./examples/c_abi_demo [num_envs] [steps] [seed]
 *
 * Runs config_default-style stepping (dense reward, hard preset, Philox resets, random policy through the exposed
 * policy stream, auto-reset with counters) and prints counters + a checksum of the observations, which
 * tests/test_gpu_parity.py::test_c_abi_demo_matches_python_face compares with the Python face on the same inputs. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dexsim.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA: %s (%s:%d)\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
#define DX(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "dexsim: %s (%s:%d)\n", dexsim_error_string(rc_), __FILE__, __LINE__); return 3; } } while (0)

static void* dalloc(size_t bytes) {
    void* p = NULL;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return NULL;
    cudaMemset(p, 0, bytes);
    return p;
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 5000;
    const int steps = argc > 2 ? atoi(argv[2]) : 120;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 7;
    const int64_t ld = (n + 31) / 32 * 32;
    if (dexsim_version() != DEXSIM_ABI_VERSION || dexsim_sizeof_state() != (int)sizeof(DexsimState) ||
        dexsim_sizeof_step_io() != (int)sizeof(DexsimStepIO) || dexsim_sizeof_group() != (int)sizeof(DexsimGroup)) {
        fprintf(stderr, "header / library mismatch\n");
        return 1;
    }

    DexsimState st;
    memset(&st, 0, sizeof st);
    st.n = n; st.ld = ld;
    st.obs = (float*)dalloc((size_t)DEXSIM_OBS * ld * 4);
    st.op64 = (double*)dalloc((size_t)3 * ld * 8);
    st.thr = (double*)dalloc((size_t)ld * 8);
    st.damp = (float*)dalloc((size_t)ld * 4);
    st.step_count = (int32_t*)dalloc((size_t)ld * 4);
    st.cmask = (uint8_t*)dalloc((size_t)ld);
    st.size = (double*)dalloc((size_t)ld * 8);
    st.mass = (double*)dalloc((size_t)ld * 8);
    st.friction = (double*)dalloc((size_t)ld * 8);
    st.episode = (uint32_t*)dalloc((size_t)ld * 4);
    /* ep_return / ep_stats stay NULL: counts-only tracking */
    {   /* the constant quaternion rows (1, 0, 0, 0) are the caller's to initialise once */
        float* ones = (float*)malloc((size_t)ld * 4);
        for (int64_t i = 0; i < ld; ++i) ones[i] = 1.0f;
        CK(cudaMemcpy(st.obs + (size_t)DEXSIM_ROW_QUAT * ld, ones, (size_t)ld * 4, cudaMemcpyHostToDevice));
        free(ones);
    }

    DexsimParams p;
    memset(&p, 0, sizeof p);
    p.w_distance = 1.0; p.w_contact = 0.5; p.w_closure = 0.3; p.w_stability = 0.2;
    p.reward_type = 1; p.max_episode_steps = 50; p.success_threshold = 3;
    p.auto_reset = 1; p.respawn = 1; p.success_is_terminated = 1; p.loop_max_steps = 50;
    p.num_groups = 1; p.seed = seed; p.env_gid0 = 0;

    DexsimGroup g;                       /* CurriculumConfig.hard(): experiments/config.py:203-217 */
    memset(&g, 0, sizeof g);
    g.size = 0.03; g.mass = 0.2; g.friction = 0.3;
    g.spawn_lo[0] = -0.1; g.spawn_hi[0] = 0.1; g.spawn_lo[1] = -0.1; g.spawn_hi[1] = 0.1; g.spawn_lo[2] = 0.05; g.spawn_hi[2] = 0.2;
    DexsimGroup* d_groups = (DexsimGroup*)dalloc(sizeof g);
    CK(cudaMemcpy(d_groups, &g, sizeof g, cudaMemcpyHostToDevice));

    DexsimStepIO io;
    memset(&io, 0, sizeof io);
    float* d_action = (float*)dalloc((size_t)DEXSIM_NJ * ld * 4);
    io.action = d_action; io.action_layout = 0;                   /* SoA [15, ld] */
    io.reward = (float*)dalloc((size_t)ld * 4);
    io.terminated = (uint8_t*)dalloc((size_t)ld);
    io.truncated = (uint8_t*)dalloc((size_t)ld);
    io.num_contacts = (uint8_t*)dalloc((size_t)ld);
    io.finished = (uint8_t*)dalloc((size_t)ld);
    io.counters = (int64_t*)dalloc((size_t)DEXSIM_NCOUNTERS * 8);
    io.ret_sums = (double*)dalloc(2 * 8);

    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    DX(dexsim_reset_philox(&st, &p, d_groups, NULL, NULL, 1, s));
    for (int t = 0; t < steps; ++t) {
        DX(dexsim_fill_policy_actions(&st, &p, DEXSIM_POLICY_RANDOM, d_action, s));   /* a ~ U(-1, 1), Philox */
        DX(dexsim_step(&st, &p, d_groups, NULL, &io, s));
    }
    CK(cudaStreamSynchronize(s));

    int64_t cnt[DEXSIM_NCOUNTERS];
    CK(cudaMemcpy(cnt, io.counters, sizeof cnt, cudaMemcpyDeviceToHost));
    float* h_obs = (float*)malloc((size_t)DEXSIM_OBS * ld * 4);
    CK(cudaMemcpy(h_obs, st.obs, (size_t)DEXSIM_OBS * ld * 4, cudaMemcpyDeviceToHost));
    uint64_t h = 1469598103934665603ull;                          /* FNV-1a over the observation bits, env-major */
    for (int64_t i = 0; i < n; ++i)
        for (int r = 0; r < DEXSIM_OBS; ++r) {
            uint32_t w;
            memcpy(&w, &h_obs[(size_t)r * ld + i], 4);
            if (w == 0x80000000u) w = 0;                          /* -0.0 == +0.0 */
            for (int b = 0; b < 4; ++b) { h ^= (w >> (8 * b)) & 0xFFu; h *= 1099511628211ull; }
        }
    printf("{\"envs\": %lld, \"steps\": %d, \"episodes\": %lld, \"successes\": %lld, \"sum_steps\": %lld, \"obs_fnv1a\": \"%016llx\"}\n",
           (long long)n, steps, (long long)cnt[DEXSIM_CNT_EPISODES], (long long)cnt[DEXSIM_CNT_SUCCESSES],
           (long long)cnt[DEXSIM_CNT_SUM_STEPS], (unsigned long long)h);
    free(h_obs);
    return 0;
}
