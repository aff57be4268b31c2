"""CUDA-graph replay timing for small batches (experiments only)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

CC = dx.CurriculumConfig
from dexterous_rl_manipulation_b200 import _lib  # noqa: E402
for n in (4096, 16384, 65536):
  for impl in ("register", "tma"):
    _lib.set_step_impl(impl)
    for steps in (1, 8):
        env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=200, reward_type="dense", curriculum_config=CC.hard(),
                                        auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=True, seed=42)
        env.reset(seed=42)
        buf = torch.rand(steps, n, 15, device="cuda") * 2 - 1
        replay = env.capture_step(buf, steps=steps)
        for _ in range(20):
            replay()
        torch.cuda.synchronize()
        reps = 300
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            replay()
        e1.record()
        host = time.perf_counter() - t0
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * steps)
        print(f"{impl:8s} n={n:6d} graph of {steps} step(s): gpu {ms * 1e3:6.2f} us/step, host {host / (reps * steps) * 1e6:6.2f} us/step, "
              f"{n / ms / 1e6:6.2f} G env-steps/s")
        # correctness: graph replay == eager stepping
        a = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=200, reward_type="dense", curriculum_config=CC.hard(),
                                      auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=True, seed=7)
        b = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=200, reward_type="dense", curriculum_config=CC.hard(),
                                      auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=True, seed=7)
        a.reset(seed=7); b.reset(seed=7)
        rb = b.capture_step(buf, steps=steps)
        for _ in range(5):
            for k in range(steps):
                a.step(buf[k])
            rb()
        assert torch.equal(a._obs, b._obs) and torch.equal(a._step_count, b._step_count) and torch.equal(a.counters, b.counters)
print("graph replay == eager: ok")
