"""Run BASELINE.json's configs 2-5 at their full sizes on the B200 path and print one JSON object per config.

    python examples/run_baseline_configs.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        examples/run_baseline_configs.py                          # envs sharded over the ranks, counters all-reduced

Not the bench (bench.py measures the headline metric); this shows every named configuration running end to end
with its own throughput and result summary.  Times are CUDA events on the launching stream, max over ranks.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import dexterous_rl_manipulation_b200 as dx  # noqa: E402

CC = dx.CurriculumConfig
RANK = int(os.environ.get("RANK", 0))
WORLD = int(os.environ.get("WORLD_SIZE", 1))
LOCAL = int(os.environ.get("LOCAL_RANK", 0))


def timed(fn):
    """fn() between two events; returns (result, seconds as the max over ranks)."""
    torch.cuda.synchronize()
    if WORLD > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device="cuda", dtype=torch.float64)
    if WORLD > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return out, float(t.item())


def emit(name, **kw):
    if RANK == 0:
        print(json.dumps({"config": name, "n_gpus": WORLD, **kw}), flush=True)


def reduce_counters(env):
    c, r = env.counters.clone(), env.ret_sums.clone()
    if WORLD > 1:
        dx.distributed.allreduce_counters(c, r)
    return c, r


def config2_default_curriculum(n_total=4096, steps=2000, warmup=200):
    """config_default.json: dense reward + CurriculumScheduler(easy -> hard), random policy, auto-reset (respawn)."""
    lo, hi = dx.distributed.shard_range(n_total, RANK, WORLD)
    easy, hard = CC.easy(), CC.hard()
    env = dx.BatchedManipulationEnv(hi - lo, "cuda", reward_type="dense", max_episode_steps=200, curriculum_config=easy,
                                    auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=True,
                                    info_success=True, seed=42, env_gid0=lo)
    sched = dx.CurriculumScheduler(easy, hard, success_rate_threshold=0.3, window_size=15,
                                   min_episodes_before_progression=20, progression_steps=5)
    driver = dx.BatchedCurriculumDriver(env, sched)
    env.reset(seed=42)

    def run(k):
        done = 0
        while done < k:
            chunk = min(100, k - done)
            env.rollout(chunk, policy="random", zero_counters=False)
            done += chunk
            driver.poll()
    run(warmup)
    _, sec = timed(lambda: run(steps))
    c, r = reduce_counters(env)
    m = dx.distributed.summarize_counters(c, r)[0]
    emit("2: config_default dense + curriculum, random policy", envs=n_total, steps=steps,
         env_steps_per_sec=n_total * steps / sec, difficulty=sched.get_difficulty_level(),
         episodes=m["total_episodes"], success_rate=m["grasp_success_rate"], mean_episode_length=m["mean_episode_length"])


def config3_heldout_noise_sweep(envs_per_cell=65536, noise=(0.0, 0.01, 0.05, 0.10)):
    """20 held-out objects x (sigma_obs, sigma_dyn) grid, fused heuristic policy, one 200-step episode per env.
    One launch per sigma_obs level (observation noise only changes what the policy sees); the dynamics-noise levels
    of a launch are groups of the same batch: env -> (object, sigma_dyn) = id mod 80."""
    rng = np.random.default_rng(123)
    objs = [CC(object_size=float(rng.uniform(0.08, 0.12)), object_mass=float(rng.uniform(0.16, 0.26)),
               friction_coefficient=float(rng.uniform(0.0, 0.29))) for _ in range(20)]
    cfgs = [o for o in objs for _ in noise]
    sd = [s for _ in objs for s in noise]
    n_total = envs_per_cell * len(noise)        # envs of one launch: every (object, sigma_dyn) cell gets envs_per_cell / 20
    lo, hi = dx.distributed.shard_range(n_total, RANK, WORLD)
    table, total_sec = {}, 0.0
    for so in noise:
        env = dx.BatchedManipulationEnv(hi - lo, "cuda", reward_type="dense", max_episode_steps=200, track_episodes=True,
                                        groups=cfgs, group_sigma_dyn=sd, group_sigma_obs=[so] * len(cfgs), seed=42, env_gid0=lo)
        env.reset(seed=42)
        _, sec = timed(lambda: env.rollout(200, policy="heuristic", one_episode=True))
        total_sec += sec
        c, _ = reduce_counters(env)
        c = c.cpu().numpy().reshape(20, len(noise), -1)
        for j, s_dyn in enumerate(noise):
            table[f"obs{so:g}_dyn{s_dyn:g}"] = round(float(c[:, j, 1].sum() / max(1, c[:, j, 0].sum())), 4)
        del env
    emit("3: held-out eval, 20 objects x noise sweep, fused heuristic policy", envs_per_launch=n_total,
         launches=len(noise), env_steps_per_sec=n_total * 200 * len(noise) / total_sec, success_rate_by_cell=table)


def config4_variable(n_total=1 << 20, steps=400, warmup=100):
    """config_variable.json: per-reset size / mass / friction ranges, fused random policy, counters all-reduced."""
    cfg = CC(object_size_range=(0.03, 0.07), object_mass_range=(0.05, 0.15), friction_range=(0.3, 0.7))
    lo, hi = dx.distributed.shard_range(n_total, RANK, WORLD)
    env = dx.BatchedManipulationEnv(hi - lo, "cuda", reward_type="dense", max_episode_steps=200, curriculum_config=cfg,
                                    auto_reset=True, respawn=True, loop_max_steps=200, track_episodes=True, seed=42, env_gid0=lo)
    env.reset(seed=42)
    env.rollout(warmup, policy="random")
    _, sec = timed(lambda: [env.rollout(50, policy="random", zero_counters=False) for _ in range(steps // 50)])
    c, r = reduce_counters(env)
    m = dx.distributed.summarize_counters(c, r)[0]
    emit("4: config_variable, randomized size/mass/friction, fused random policy", envs=n_total, steps=steps,
         env_steps_per_sec=n_total * steps / sec, episodes=m["total_episodes"], success_rate=m["grasp_success_rate"],
         failure_modes={k: v["count"] for k, v in m["failure_mode_frequency"].items()})


def config5_learner_seeds(runs_per_cell=4096, episodes=20,
                          seeds=(42, 123, 456, 789, 1000, 2024, 3000, 4096, 5150, 6001)):
    """sparse vs dense x 10 seeds, one SimpleLearner per env (fused), `episodes` training episodes per run.
    Seeds beyond config_default.json's seven are 4096, 5150, 6001."""
    cells = [(rt, s) for rt in ("dense", "sparse") for s in seeds]
    mine = cells[RANK::WORLD]                    # whole cells per rank: learners never communicate
    out, t0 = {}, time.perf_counter()
    torch.cuda.synchronize()
    for rt, s in mine:
        res = dx.training.train_learners_batched(runs_per_cell, episodes, curriculum_config=CC.medium(), reward_type=rt,
                                                 seed=s, success_is_terminated=True)
        out[f"{rt}/{s}"] = [float(res["episode_rewards"][:, -1].mean()), float(res["successes"][:, -1].mean()),
                            float(res["episode_steps"].sum())]
        del res
    torch.cuda.synchronize()
    sec = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    gathered = [out]
    if WORLD > 1:
        torch.distributed.all_reduce(sec, op=torch.distributed.ReduceOp.MAX)
        gathered = [None] * WORLD
        torch.distributed.all_gather_object(gathered, out)
    merged = {k: v for g in gathered for k, v in g.items()}
    merged = {k: merged[k] for k in sorted(merged)}          # cell order independent of the rank count (float sums below)
    total_steps = sum(v[2] for v in merged.values())
    summary = {rt: {"final_mean_return": float(np.mean([v[0] for k, v in merged.items() if k.startswith(rt)])),
                    "final_success_rate": float(np.mean([v[1] for k, v in merged.items() if k.startswith(rt)])),
                    "return_std_over_seeds": float(np.std([v[0] for k, v in merged.items() if k.startswith(rt)]))}
               for rt in ("dense", "sparse")}
    emit("5: sparse vs dense x 10 seeds, fused SimpleLearner", runs_per_cell=runs_per_cell, episodes=episodes,
         env_steps_per_sec=total_steps / float(sec.item()), timing="wall clock incl. per-episode log read-back", **summary)


class _HeldOut:
    """Duck type of evaluation/heldout_objects.py::HeldOutObjectSet (the reference's class works the same way)."""

    class _Obj:
        def __init__(self, size, mass, friction):
            self.size, self.mass, self.friction = size, mass, friction

    def __init__(self, n=20, seed=123):
        rng = np.random.default_rng(seed)
        self.heldout_objects = [self._Obj(float(rng.uniform(0.08, 0.12)), float(rng.uniform(0.16, 0.26)),
                                          float(rng.uniform(0.0, 0.29))) for _ in range(n)]

    def get_eval_config(self, k):
        o = self.heldout_objects[k]
        return CC(object_size=o.size, object_mass=o.mass, friction_coefficient=o.friction)


def frontend_heldout(episodes_per_object=500):
    """Evaluator.evaluate_heldout_set's result dictionary from the batched front-end; with several ranks every rank
    evaluates a slice of the (object, episode) batch and the per-episode records are all-gathered."""
    t0 = time.perf_counter()
    res = dx.evaluation.evaluate_heldout_set_batched(_HeldOut(), policy="heuristic", seed=42,
                                                     num_episodes_per_object=episodes_per_object, contact_history=False)
    sec = time.perf_counter() - t0
    m = res["metrics"]
    emit("front-end: evaluate_heldout_set_batched, 20 objects x %d episodes" % episodes_per_object,
         episodes=res["overall_stats"]["total_episodes"], success_rate=m["grasp_success_rate"],
         mean_episode_length=m["mean_episode_length"], mean_reward=res["overall_stats"]["mean_reward"],
         failure_types={k: v["count"] for k, v in m["failure_type_frequency"].items()}, wall_seconds=round(sec, 3))


def main():
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    which = sys.argv[1:] or ["2", "3", "4", "5"]
    if "2" in which:
        config2_default_curriculum()
    if "3" in which:
        config3_heldout_noise_sweep()
    if "4" in which:
        config4_variable()
    if "5" in which:
        config5_learner_seeds()
    if "frontend" in which:
        frontend_heldout()
    if WORLD > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
