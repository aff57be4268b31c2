"""Small / mid-size batches: register-resident kernel vs TMA pipeline per variant (experiments only; picks the
crossover used by the auto choice in dexsim_kernels.cu::launch_step)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402
from dexterous_rl_manipulation_b200 import _lib  # noqa: E402

CC = dx.CurriculumConfig
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 16384, 32768, 65536, 98304, 131072, 196608, 262144, 524288]
for n in sizes:
    g = torch.Generator(device="cuda").manual_seed(0)
    pool = [torch.rand(n, 15, device="cuda", generator=g) * 2 - 1 for _ in range(4)]
    for variant in ("plain", "counts", "track"):
        res = {}
        for impl in ("register", "tma"):
            _lib.set_step_impl(impl)
            kw = dict(max_episode_steps=200, reward_type="dense", curriculum_config=CC.hard(), seed=42)
            if variant != "plain":
                kw.update(auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=variant == "track")
            env = dx.BatchedManipulationEnv(n, "cuda", **kw)
            env.reset(seed=42)
            for t in range(50):
                env.step(pool[t % 4])
            torch.cuda.synchronize()
            reps = 400
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for t in range(reps):
                env.step(pool[t % 4])
            e1.record()
            t_issue = time.perf_counter() - t0
            torch.cuda.synchronize()
            res[impl] = (e0.elapsed_time(e1) / reps * 1e3, t_issue / reps * 1e6)
            del env
        print(f"n={n:7d} {variant:7s} register {res['register'][0]:7.2f} us  tma {res['tma'][0]:7.2f} us  "
              f"(host issue {res['register'][1]:5.2f} / {res['tma'][1]:5.2f} us)  frac(best) "
              f"{410 * n / min(res['register'][0], res['tma'][0]) / 1e3 / 6552:5.3f}", flush=True)
_lib.set_step_impl("auto")
