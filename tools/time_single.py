"""num_envs == 1 drop-in path: microseconds per env.step (experiments only)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx
env = dx.BatchedManipulationEnv(1, "cuda", reward_type="dense", max_episode_steps=200)
obs, info = env.reset(seed=0)
a = np.full(15, -0.5, np.float32)
for _ in range(50): env.step(a)
t0 = time.perf_counter(); n = 2000
for _ in range(n): obs, r, te, tr, info = env.step(a)
print(f"num_envs=1 drop-in: {(time.perf_counter() - t0) / n * 1e6:.1f} us per step; last reward {r:.6f}, info keys {sorted(info)}")
