"""Driving the reference's UNCHANGED CurriculumScheduler from a batched env.

The scheduler (experiments/curriculum_scheduler.py:13-273) is sequential: one
``update(success, episode_steps)`` per finished episode, at most ``progression_steps``
progressions ever.  A batched env finishes thousands of episodes per step, so the driver
(SURVEY.md section 7 "hard parts", 8a-15)

  * reads the device counters' delta every ``poll()`` (one tiny D2H copy),
  * replays that many ``update`` calls one by one ONLY while the scheduler can still
    progress (difficulty < 1.0), spreading successes evenly through the batch of updates,
  * pushes ``scheduler.get_current_config()`` into ``env.curriculum_config`` when a
    progression happened (component_ablation.py:163-166) -- effective at the next resets,
  * folds the rest of a poll's episodes into the scheduler's totals in bulk (bounded tail of the
    per-episode lists) and then lets the scheduler's own progression test run on the new totals, which
    also drives ``StepBasedScheduler`` (step-count milestones).

The scheduler is duck-typed: ``update``, ``get_current_config`` and the public attributes
``current_difficulty_level``, ``episode_successes``, ``episode_steps``, ``total_steps``,
``total_episodes``, ``window_size`` are all it needs.
"""
from typing import Optional

import torch

from ._lib import CNT_EPISODES, CNT_SUCCESSES, CNT_SUM_STEPS


def spread_flag(k: int, episodes: int, successes: int) -> bool:
    """Element k of the length-``episodes`` boolean sequence that spreads ``successes`` Trues evenly
    (Bresenham): the deterministic stand-in for the unknowable order of parallel episodes.  O(1),
    so a poll never materialises millions of flags."""
    return ((k + 1) * successes) // episodes > (k * successes) // episodes


def spread_successes(episodes: int, successes: int):
    return [spread_flag(k, episodes, successes) for k in range(episodes)]


class BatchedCurriculumDriver:
    def __init__(self, env, scheduler, max_sequential_updates: int = 4096):
        self.env = env
        self.scheduler = scheduler
        self.max_sequential_updates = int(max_sequential_updates)
        self._last = None
        self.progressions = 0
        self.exact_episodes = 0          # every episode fed so far / how many of them succeeded (see feed())
        self.exact_successes = 0
        env.curriculum_config = scheduler.get_current_config()

    def feed(self, episodes: int, successes: int, steps: int) -> int:
        """Feed an aggregate of finished episodes; returns the number of progressions."""
        if episodes <= 0:
            return 0
        sch = self.scheduler
        episodes, successes = int(episodes), int(successes)
        self.exact_episodes += episodes
        self.exact_successes += successes
        base, rem = divmod(int(steps), episodes)
        progressed = 0
        k = 0
        if getattr(sch, "current_difficulty_level", 1.0) < 1.0:
            limit = min(episodes, self.max_sequential_updates)
            while k < limit and sch.current_difficulty_level < 1.0:
                if sch.update(spread_flag(k, episodes, successes), base + (1 if k < rem else 0)):
                    progressed += 1
                k += 1
        if k < episodes:
            # bulk tail: no progression can happen any more (or the sequential budget is spent);
            # the scheduler's totals (total_episodes, total_steps) stay exact, its per-episode lists get a bounded
            # tail -- so statistics the reference derives from the LISTS (get_statistics()["overall_success_rate"] =
            # mean(episode_successes)) describe the retained tail only; exact totals are kept in
            # `self.exact_episodes` / `self.exact_successes` for callers that report overall rates
            rest = episodes - k
            rest_steps = int(steps) - (base * k + min(k, rem))
            keep = min(rest, int(getattr(sch, "window_size", 20)))
            sch.episode_successes.extend(spread_flag(j, episodes, successes) for j in range(episodes - keep, episodes))
            sch.episode_steps.extend([base] * keep)
            sch.total_steps += rest_steps
            sch.total_episodes += rest
            # let the scheduler act on the new totals / window exactly as further update() calls would
            # (CurriculumScheduler: one level per call while the window rate holds; StepBasedScheduler,
            # experiments/curriculum_scheduler.py:276-335: one milestone per call)
            # The reference progresses at most once per update() call, i.e. at most `rest` times over this tail.
            should, prog = getattr(sch, "_should_progress", None), getattr(sch, "_progress", None)
            guard = 0
            while should is not None and prog is not None and guard < min(rest, 1024) and should():
                guard += 1
                if not prog():
                    break
                progressed += 1
        if progressed:
            self.progressions += progressed
            self.env.curriculum_config = sch.get_current_config()
        return progressed

    def poll(self, counters: Optional[torch.Tensor] = None) -> int:
        """Read the counters (summed over groups), feed the delta since the last poll."""
        c = self.env.counters if counters is None else counters
        tot = c[:, [CNT_EPISODES, CNT_SUCCESSES, CNT_SUM_STEPS]].sum(0).cpu().tolist()
        if self._last is None:
            self._last = [0, 0, 0]
        d = [int(a - b) for a, b in zip(tot, self._last)]
        self._last = tot
        return self.feed(d[0], d[1], d[2])


class CurriculumScheduler:
    """Behaviour-compatible stand-in for the reference scheduler for users (and bench.py) that do
    not have the reference on their path; the reference's own class works with the driver
    unchanged.  Semantics restated from experiments/curriculum_scheduler.py:
      * progression needs total_episodes >= min_episodes_before_progression, level < 1.0, a full
        window, and mean(last window_size successes) >= success_rate_threshold (:163-186);
      * each progression adds 1/progression_steps to the level (capped at 1.0) and re-interpolates
        size / mass / friction / spawn_distance linearly between the initial and target configs;
        only ``object_size_range`` is interpolated among the ranges, mass / friction / spawn ranges
        keep the INITIAL config's values (:77-139);  the window is not cleared on progression."""

    def __init__(self, initial_config, target_config, success_rate_threshold=0.7,
                 min_episodes_before_progression=50, window_size=20, progression_steps=5):
        self.initial_config, self.target_config = initial_config, target_config
        self.success_rate_threshold = success_rate_threshold
        self.min_episodes_before_progression = min_episodes_before_progression
        self.window_size, self.progression_steps = window_size, progression_steps
        self.reset()

    def reset(self):
        self.current_config = self._interpolate(0.0)
        self.current_difficulty_level = 0.0
        self.episode_successes, self.episode_steps = [], []
        self.total_steps = self.total_episodes = 0
        self.progression_history = []

    def _interpolate(self, d):
        from .config import CurriculumConfig
        a, b = self.initial_config, self.target_config
        d = min(max(float(d), 0.0), 1.0)
        mix = lambda x, y: x * (1 - d) + y * d
        size_range = None
        if a.object_size_range is not None and b.object_size_range is not None:
            size_range = (mix(a.object_size_range[0], b.object_size_range[0]),
                          mix(a.object_size_range[1], b.object_size_range[1]))
        return CurriculumConfig(
            object_size=mix(a.object_size, b.object_size), object_size_range=size_range,
            object_mass=mix(a.object_mass, b.object_mass), object_mass_range=a.object_mass_range,
            friction_coefficient=mix(a.friction_coefficient, b.friction_coefficient), friction_range=a.friction_range,
            spawn_distance=mix(a.spawn_distance, b.spawn_distance), spawn_distance_range=a.spawn_distance_range,
            spawn_x_range=a.spawn_x_range, spawn_y_range=a.spawn_y_range, spawn_z_range=a.spawn_z_range)

    def _window_rate(self):
        w = self.episode_successes[-self.window_size:]
        return sum(1 for s in w if s) / len(w)

    def update(self, success, episode_steps):
        self.episode_successes.append(bool(success))
        self.episode_steps.append(int(episode_steps))
        self.total_steps += int(episode_steps)
        self.total_episodes += 1
        if (self.total_episodes < self.min_episodes_before_progression or self.current_difficulty_level >= 1.0
                or len(self.episode_successes) < self.window_size):
            return False
        rate = self._window_rate()
        if rate < self.success_rate_threshold:
            return False
        new = min(self.current_difficulty_level + 1.0 / self.progression_steps, 1.0)
        if new <= self.current_difficulty_level:
            return False
        self.current_difficulty_level = new
        self.current_config = self._interpolate(new)
        self.progression_history.append({
            "episode": self.total_episodes, "total_steps": self.total_steps, "difficulty_level": float(new),
            "success_rate": float(rate), "object_size": float(self.current_config.object_size),
            "object_mass": float(self.current_config.object_mass),
            "friction_coefficient": float(self.current_config.friction_coefficient)})
        return True

    def get_current_config(self):
        return self.current_config

    def get_difficulty_level(self):
        return self.current_difficulty_level

    def get_statistics(self):
        w = self.episode_successes[-self.window_size:] if len(self.episode_successes) >= self.window_size \
            else self.episode_successes
        mean = lambda xs: (sum(1 for s in xs if s) / len(xs)) if xs else 0.0
        return {"total_episodes": self.total_episodes, "total_steps": self.total_steps,
                "current_difficulty_level": float(self.current_difficulty_level),
                "recent_success_rate": mean(w), "overall_success_rate": mean(self.episode_successes),
                "num_progressions": len(self.progression_history),
                "progression_history": list(self.progression_history)}
