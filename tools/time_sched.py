"""Static round-robin vs dynamic (counter-fed) tile assignment of the pipelined step kernel, and the cost of reset
waves: CUDA-event timings for experiments (not a bench number source).

    python tools/time_sched.py [num_envs] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
CC = dx.CurriculumConfig
g = torch.Generator(device="cuda").manual_seed(0)
pool = [torch.rand(n, 15, device="cuda", generator=g) * 2 - 1 for _ in range(4)]
for cfg_name in ("hard", "easy"):
    for variant in ("api", "api_counts", "api_track"):
        for sched in ("static", "dynamic", "dyn+alt"):
            kw = dict(max_episode_steps=200, reward_type="dense", curriculum_config=getattr(CC, cfg_name)(), seed=42)
            if variant != "api":
                kw.update(auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=variant == "api_track")
            elif cfg_name == "easy":
                continue
            env = dx.BatchedManipulationEnv(n, "cuda", **kw)
            if sched == "static":
                env._io.sched = None
            if sched != "dyn+alt":          # "dyn+alt": dynamic tiles + the walk direction alternating from step to step (the default)
                env._alternate_tiles = 0
            env.reset(seed=42)
            for t in range(20):
                env.step(pool[t % 4])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(reps):
                env.step(pool[t % 4])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            eps = int(env.counters[:, 0].sum()) if variant != "api" else 0
            print(f"{cfg_name:5s} {variant:10s} {sched:8s} n={n} {ms * 1e3:8.2f} us/step  algo {410 * n / ms / 1e6:7.0f} GB/s  "
                  f"frac {410 * n / ms / 1e6 / 6552:5.3f}  resets/env-step {eps / (n * (reps + 20)):7.4f}", flush=True)
            del env
