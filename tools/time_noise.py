"""API-mode stepping of a noisy env: Philox noise drawn inside the step kernel vs separate fill kernels (experiments)."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
for so, sd in ((0.0, 0.1), (0.05, 0.0), (0.05, 0.1)):
    for fused in (True, False):
        env = dx.BatchedManipulationEnv(n, "cuda", max_episode_steps=200, reward_type="dense", seed=1,
                                        curriculum_config=dx.CurriculumConfig.hard(), observation_noise_std=so,
                                        dynamics_noise_std=sd)
        env.fused_noise = fused
        env.reset(seed=1)
        a = torch.rand(n, 15, device="cuda") * 2 - 1
        for _ in range(10):
            env.step(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            env.step(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"sigma_obs {so:4.2f} sigma_dyn {sd:4.2f} {'in-kernel' if fused else 'separate '} : {ms * 1e3:7.1f} us/step "
              f"{n / ms / 1e6:6.2f} G env-steps/s", flush=True)
        del env
