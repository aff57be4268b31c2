"""Small-batch timing helper (experiments only)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402
from dexterous_rl_manipulation_b200 import _lib  # noqa: E402

CC = dx.CurriculumConfig
for n in (4096, 16384, 65536, 131072):
    for impl in ("register", "tma"):
        for cfg_name, cfg in (("hard", CC.hard()), ("easy", CC.easy())):
            _lib.set_step_impl(impl)
            env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=200, reward_type="dense", curriculum_config=cfg,
                                            auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=True, seed=42)
            env.reset(seed=42)
            g = torch.Generator(device="cuda").manual_seed(0)
            pool = [torch.rand(n, 15, device="cuda", generator=g) * 2 - 1 for _ in range(4)]
            for t in range(50):
                env.step(pool[t % 4])
            torch.cuda.synchronize()
            reps = 500
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for t in range(reps):
                env.step(pool[t % 4])
            e1.record()
            t_issue = time.perf_counter() - t0
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"n={n:7d} {impl:8s} {cfg_name:4s} gpu {ms * 1e3:7.2f} us/step  host issue {t_issue / reps * 1e6:6.2f} us/step  "
                  f"{n / ms / 1e6:7.2f} G env-steps/s")
            del env, pool
_lib.set_step_impl("auto")
