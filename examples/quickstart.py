"""The snippets of INTEGRATION.md as one runnable script (needs a B200).

    python examples/quickstart.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

CC = dx.CurriculumConfig

# 1. drop-in single env (what unmodified reference callers see)
env = dx.BatchedManipulationEnv(1, "cuda", reward_type="dense", curriculum_config=CC.easy())
obs, info = env.reset(seed=0)
total, steps = 0.0, 0
for _ in range(env.max_episode_steps):
    obs, reward, terminated, truncated, info = env.step(env.action_space.sample() * 0.1 - 0.5)
    total += reward; steps += 1
    if terminated or truncated:
        break
print(f"single env: {steps} steps, return {total:.3f}, contacts {info['num_contacts']}, obs {obs.shape} {obs.dtype}")

# 2. batched stepping with a device policy, auto-reset and the curriculum scheduler
N = 1 << 16
env = dx.BatchedManipulationEnv(N, "cuda", reward_type="dense", max_episode_steps=200, curriculum_config=CC.easy(),
                                auto_reset=True, respawn=True, loop_max_steps=200, seed=42)
sched = dx.CurriculumScheduler(CC.easy(), CC.hard(), success_rate_threshold=0.3, min_episodes_before_progression=20,
                               window_size=15, progression_steps=5)
driver = dx.BatchedCurriculumDriver(env, sched)
obs, info = env.reset(seed=42)
policy = lambda o: (-0.5 + 0.1 * torch.tanh(o[:, :15])).contiguous()          # any CUDA policy: [N,45] -> [N,15]
for t in range(300):
    obs, reward, terminated, truncated, info = env.step(policy(obs))
    if t % 10 == 9:
        driver.poll()
print(f"batched: difficulty {sched.current_difficulty_level:.1f} after {int(env.counters[:, 0].sum())} episodes, "
      f"mean reward of last step {float(reward.mean()):.3f}")

# 3. evaluation rollouts that never leave the device: 20 objects x noise cells, fused heuristic policy
objs = [CC(object_size=0.03 + 0.004 * k, object_mass=0.2, friction_coefficient=0.2) for k in range(20)]
env = dx.BatchedManipulationEnv(20 * 4096, "cuda", reward_type="dense", track_episodes=True, groups=objs, seed=7)
env.reset(seed=7)
counters, ret_sums = env.rollout(200, policy="heuristic")
dx.distributed.allreduce_counters(counters, ret_sums)              # no-op on one GPU
metrics = dx.distributed.summarize_counters(counters, ret_sums)
print("held-out style eval: success rate per object size:",
      " ".join(f"{m['grasp_success_rate']:.2f}" for m in metrics))

# 4. per-seed learner training runs in parallel (the reference's component-ablation loop, 256 runs at once)
out = dx.training.train_learners_batched(256, 10, curriculum_config=CC.medium(), reward_type="dense", max_episode_steps=100, seed=3)
print(f"learner training: episode_rewards {out['episode_rewards'].shape}, mean return {out['episode_rewards'].mean():.2f}")
