"""Smallest run that drives every kernel family once (for compute-sanitizer: memcheck / racecheck / synccheck)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402
from dexterous_rl_manipulation_b200 import _lib  # noqa: E402

CC = dx.CurriculumConfig
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for impl in ("tma", "register"):
    _lib.set_step_impl(impl)
    for track in (None, False, True):
        kw = dict(max_episode_steps=6, reward_type="dense", seed=3, groups=[CC.easy(), CC(object_size_range=(0.03, 0.08))])
        if track is not None:
            kw.update(auto_reset=True, respawn=True, loop_max_steps=6, track_episodes=track)
        env = dx.BatchedManipulationEnv(n, "cuda", **kw)
        env.reset(seed=3)
        for t in range(14):
            env.step(torch.rand(n, 15, device="cuda") * 2 - 1)
        torch.cuda.synchronize()
    env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=6, reward_type="dense", seed=3, observation_noise_std=0.05,
                                    dynamics_noise_std=0.1, auto_reset=True, respawn=True, loop_max_steps=6, track_episodes=True)
    env.reset(seed=3)
    for t in range(8):
        env.step(torch.rand(n, 15, device="cuda") * 2 - 1)
    torch.cuda.synchronize()
_lib.set_step_impl("auto")
env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=6, reward_type="dense", seed=3, track_episodes=True)
env.reset(seed=3)
env.enable_episode_log(4 * n)
env.enable_history(12)
env.rollout(12, policy="heuristic")
env.step_host(torch.rand(n, 15).pin_memory(), chunks=2)
torch.cuda.synchronize()
print("sanitize case done", n)
