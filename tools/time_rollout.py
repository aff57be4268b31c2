"""Fused rollout timing (experiments): env-steps/s of rollout(50) launches at several batch sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

for n in (int(a) for a in (sys.argv[1:] or ["1048576", "65536"])):
    for policy in ("random", "heuristic"):
        env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=200, reward_type="dense", track_episodes=True,
                                        curriculum_config=dx.CurriculumConfig.hard(), seed=42)
        env.reset(seed=42)
        env.rollout(50, policy=policy)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            env.rollout(50, policy=policy)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{os.environ.get('DEXSIM_LIB_PATH', 'default'):32s} n={n:8d} {policy:9s} {n * 400 / ms / 1e6:7.2f} G env-steps/s "
              f"({ms / 400 * 1e3:6.2f} us per fused step)", flush=True)
        del env
