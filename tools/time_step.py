"""Kernel timing helper for experiments (CUDA events, not a bench number source)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
CC = dx.CurriculumConfig
for variant in ("api", "api_counts", "api_track"):
    kw = dict(max_episode_steps=200, reward_type="dense", curriculum_config=CC.hard(), seed=42)
    if variant != "api":
        kw.update(auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=variant == "api_track")
    env = dx.BatchedManipulationEnv(n, "cuda", **kw)
    env.reset(seed=42)
    g = torch.Generator(device="cuda").manual_seed(0)
    pool = [torch.rand(n, 15, device="cuda", generator=g) * 2 - 1 for _ in range(4)]
    for t in range(20):
        env.step(pool[t % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 100
    for t in range(reps):
        env.step(pool[t % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{os.environ.get('DEXSIM_LIB_PATH', 'default'):40s} {variant:10s} n={n} {ms * 1e3:8.1f} us/step  "
          f"{n / ms / 1e6:8.2f} G env-steps/s  algo {410 * n / ms / 1e6:7.0f} GB/s")
    del env, pool
