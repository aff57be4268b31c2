"""Per-step CUDA-event times of API-mode stepping around a reset wave: every env of a batch that started together is
truncated in the same step (step max_episode_steps), which makes that one step a full-batch episode end + reset
(experiments only).

    python tools/time_wave.py [num_envs] [variants,comma,separated]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["counts", "track"]
CC = dx.CurriculumConfig
g = torch.Generator(device="cuda").manual_seed(0)
pool = [torch.rand(n, 15, device="cuda", generator=g) * 2 - 1 for _ in range(4)]
T = 40
for variant in variants:
    env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=T, reward_type="dense", curriculum_config=CC.hard(), seed=42,
                                    auto_reset=True, respawn=True, loop_max_steps=T, track_episodes=variant == "track")
    env.reset(seed=42)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * T + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for t in range(3 * T):
        env.step(pool[t % 4])
        ev[t + 1].record()
    torch.cuda.synchronize()
    us = [ev[t].elapsed_time(ev[t + 1]) * 1e3 for t in range(3 * T)]
    quiet = sorted(us)[len(us) // 2]
    waves = [(t + 1, round(u, 1)) for t, u in enumerate(us) if u > 1.5 * quiet]
    eps = int(env.counters[:, 0].sum())
    print(f"{variant:7s} n={n}: median step {quiet:7.2f} us, mean {sum(us) / len(us):7.2f} us, episodes {eps}; steps above 1.5x median: {waves}", flush=True)
    del env
