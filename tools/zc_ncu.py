import os, sys, torch
sys.path.insert(0, "/root/repo")
import dexterous_rl_manipulation_b200 as dx
n = 1 << 20
env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=200, reward_type="dense", seed=1, auto_reset=True, respawn=True,
                                loop_max_steps=200, track_episodes=False, curriculum_config=dx.CurriculumConfig.easy())
env.reset(seed=1)
env.host_zero_copy = True
a = torch.rand(n, 15).mul_(2).sub_(1).pin_memory()
for t in range(30):
    env.step(a.cuda())
env.step_host(a, chunks=1, packed_contacts=True)
for t in range(3):
    env.step_host(a, chunks=1, packed_contacts=True)
for t in range(2):
    env.step_host(a, chunks=8, packed_contacts=True)
