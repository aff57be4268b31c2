"""Batched counterpart of the reference's learner training loops:
``train_with_config`` (evaluation/component_ablation.py:80-195), ``train_with_curriculum`` /
``train_without_curriculum`` (evaluation/curriculum_ablation.py:25-134) -- each of them is
``for episode: run_episode(env, SimpleLearner)`` on one reused env object.

Here every env of the batch is one such independent training run (its own SimpleLearner, its own
reused env object): ``num_runs`` runs x ``num_episodes`` episodes execute as fused rollouts with the
learner in-kernel, and the per-episode records come back from the device log.
"""
from typing import Dict, Optional, Tuple

import numpy as np

from .env import BatchedManipulationEnv


def run_episodes_batched(env: BatchedManipulationEnv, policy: str = "random", max_steps: Optional[int] = None,
                         actions=None, reset: bool = True, seed=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``run_episode`` (training/episode_utils.py:13-55) for every env of the batch at once: one episode per env in ONE
    fused launch, returning the reference's triple as arrays ``(success [n] bool, steps [n] int32, total_reward [n]
    float64)`` indexed by env.

    ``policy``: "random" | "heuristic" | "learner" (generated in-kernel, Philox) or "external" with ``actions`` of shape
    [max_steps, n, 15].  Semantics are the reference loop's: ``env.reset()`` first (``reset=False`` continues from the
    current state), at most ``max_steps or env.max_episode_steps`` steps, stop at ``terminated or truncated``, the env is
    left un-reset; ``success`` is ``info.get("success", False)`` -- always False unless the env was built with
    ``info_success=True`` (the reference env never sets the key).  Needs ``track_episodes=True``."""
    n = env.num_envs
    k = int(max_steps or env.max_episode_steps)
    if reset:
        env.reset(seed=seed)
    env.enable_episode_log(capacity=n)
    env._rollout_steps = 0
    kw = {"actions": actions} if policy == "external" else {}
    env.rollout(k, policy=policy, loop_max_steps=k, success_is_terminated=bool(env.info_success), one_episode=True, **kw)
    log = env.read_episode_log(sort=False)
    success, steps, total = np.zeros(n, bool), np.full(n, k, np.int32), np.zeros(n)
    i = log["env_gid"].astype(np.int64) - env.env_gid0
    success[i], steps[i], total[i] = log["success"].astype(bool), log["steps"], log["episode_reward"]
    if len(log) != n:
        raise RuntimeError(f"{n - len(log)} envs did not finish an episode within {k} steps")
    return success, steps, total


def train_learners_batched(num_runs: int, num_episodes: int, curriculum_config=None, reward_type: str = "dense",
                           max_episode_steps: int = 200, learning_rate: float = 0.01, exploration_noise: float = 0.3,
                           action_clip_range: float = 0.5, seed: int = 42, device="cuda", env_gid0: int = 0,
                           success_is_terminated: bool = False, scheduler=None) -> Dict:
    """Returns per-run arrays shaped [num_runs, num_episodes]: ``episode_rewards``, ``episode_steps``,
    ``successes`` (always False with the reference's run_episode semantics unless
    ``success_is_terminated=True``) and the final ``mean_action`` [num_runs, 15].

    ``scheduler``: an (unchanged) CurriculumScheduler driven from the finished-episode stream in
    (step, env) order; its current config is applied to later resets of every run."""
    from .curriculum import BatchedCurriculumDriver
    n = max(int(num_runs), 2)
    env = BatchedManipulationEnv(n, device, max_episode_steps=max_episode_steps, reward_type=reward_type,
                                 curriculum_config=curriculum_config, track_episodes=True, seed=seed, env_gid0=env_gid0)
    env.enable_learner(learning_rate, exploration_noise, action_clip_range)
    env.enable_episode_log(capacity=n)
    driver = BatchedCurriculumDriver(env, scheduler) if scheduler is not None else None
    rewards = np.zeros((n, num_episodes))
    steps = np.zeros((n, num_episodes), np.int32)
    succ = np.zeros((n, num_episodes), bool)
    for ep in range(int(num_episodes)):
        # `for episode: run_episode(env, policy)`: one launch = one episode of every run.  The env object is
        # reused, so only the first reset samples the spawn; later ones keep the object where it was left.
        env.reset(seed=seed if ep == 0 else None)
        env._ep_log_count.zero_()
        env.rollout(max_episode_steps, policy="learner", respawn=False, loop_max_steps=max_episode_steps,
                    success_is_terminated=success_is_terminated, one_episode=True)
        log = env.read_episode_log()
        if len(log) != n:
            raise RuntimeError(f"episode {ep}: {n - len(log)} runs did not finish within {max_episode_steps} steps")
        i = log["env_gid"].astype(np.int64) - env.env_gid0
        rewards[i, ep], steps[i, ep], succ[i, ep] = log["episode_reward"], log["steps"], log["success"].astype(bool)
        if driver is not None:
            driver.poll()
    k = int(num_runs)
    return {"episode_rewards": rewards[:k], "episode_steps": steps[:k], "successes": succ[:k],
            "mean_action": env.learner_mean[:k].cpu().numpy(), "env": env}


def train_shared_learner_batched(num_envs: int, num_episodes: int, curriculum_config=None, reward_type: str = "dense",
                                 max_episode_steps: int = 200, learning_rate: float = 0.01, exploration_noise: float = 0.3,
                                 action_clip_range: float = 0.5, seed: int = 42, device="cuda", num_envs_global: Optional[int] = None,
                                 group=None) -> Dict:
    """ONE SimpleLearner trained on a whole batch of rollouts (and on every GPU of the job): each episode all envs start
    from the shared mean action, explore and hill-climb independently inside the fused rollout exactly like
    policies/simple_learner.py:60-95, and after the episode the mean of the env with the highest episode return --
    over all envs of all ranks -- becomes the shared mean (``distributed.share_best_candidate``: one 17-double
    all-gather per episode; the only collective besides the counter all-reduce).  With ``torch.distributed`` initialised
    the ``num_envs_global`` envs (default ``num_envs`` x world size) are sharded by global id; every rank returns the same
    history.  Unlike ``train_learners_batched`` (independent runs, the reference's semantics run many times) this is a
    population search, so its trajectories have no reference counterpart -- only its per-env arithmetic does.

    Returns {"best_return" [num_episodes], "winner_gid" [num_episodes], "mean_action" [15], "env"}."""
    import torch
    import torch.distributed as dist
    from .distributed import shard_range, share_best_candidate
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    total = int(num_envs_global) if num_envs_global is not None else int(num_envs) * world
    lo, hi = shard_range(total, rank, world)
    n = max(hi - lo, 2)
    env = BatchedManipulationEnv(n, device, max_episode_steps=max_episode_steps, reward_type=reward_type,
                                 curriculum_config=curriculum_config, track_episodes=True, seed=seed, env_gid0=lo)
    env.enable_learner(learning_rate, exploration_noise, action_clip_range)
    env.enable_episode_log(capacity=n)
    best_hist, gid_hist = np.zeros(num_episodes), np.zeros(num_episodes, np.int64)
    for ep in range(int(num_episodes)):
        env.reset(seed=seed if ep == 0 else None)
        env._ep_log_count.zero_()
        env.rollout(max_episode_steps, policy="learner", respawn=False, loop_max_steps=max_episode_steps, one_episode=True)
        log = env.read_episode_log(sort=False)
        log = log[log["env_gid"] < hi] if hi - lo < n else log          # padding env of a 1-env shard never wins
        k = int(np.lexsort((log["env_gid"], -log["episode_reward"]))[0])
        gid = int(log["env_gid"][k])
        mean = env._learner_mean[:, gid - lo].clone()
        best, wgid, wmean = share_best_candidate(float(log["episode_reward"][k]), gid, mean, group=group)
        env._learner_mean[:, :] = wmean.to(env._learner_mean.dtype).reshape(15, 1)      # every env continues from the winner
        best_hist[ep], gid_hist[ep] = best, wgid
    return {"best_return": best_hist, "winner_gid": gid_hist, "mean_action": env.learner_mean[0].cpu().numpy(), "env": env}
