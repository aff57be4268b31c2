"""Env sharding across the GPUs of one box.

Envs never interact, so the data path has NO collective: rank r owns the contiguous global
env ids ``shard_range(N, r, world)`` and keys its Philox streams by global id, which makes
every env's trajectory independent of the number of GPUs.  The only exchange is an all-reduce
(SUM) of the small per-group counters (episodes, successes, length sums, failure-label counts)
after a rollout -- a few KB, latency-bound, issued once per evaluation, never per step.
Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU tensors (used by the CPU tests).
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of global env ids owned by ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(int(num_envs_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin this process to the CPU cores NVML reports as local to GPU ``local_rank`` (same NUMA node / PCIe
    root).  Call it before allocating pinned host buffers: the host side of ``step_host`` streams ~240 MB per
    step per GPU, and with several ranks on a two-socket host the copies otherwise cross the socket link.
    Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, bits in enumerate(mask) for b in range(64) if (bits >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def allreduce_counters(counters: torch.Tensor, ret_sums: torch.Tensor = None, group=None):
    """In-place SUM over ranks of the int64 counter table (and the float64 return sums)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counters, ret_sums
    work = [dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group, async_op=True)]
    if ret_sums is not None:
        work.append(dist.all_reduce(ret_sums, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in work:
        w.wait()
    return counters, ret_sums


def summarize_counters(counters: torch.Tensor, ret_sums: torch.Tensor = None, max_steps: int = 200):
    """Per-group aggregate metrics in the schema of EvaluationMetrics.compute_aggregate_metrics
    (evaluation/metrics.py:126-203) from the device counters."""
    from ._lib import (CNT_EPISODES, CNT_LABEL_METRICS, CNT_LABEL_TAXONOMY, CNT_SUCCESSES, CNT_SUM_FINAL_CONTACTS,
                       CNT_SUM_STEPS, CNT_SUM_STEPS_SQ, CNT_VAR_TIES, LABELS_METRICS, LABELS_TAXONOMY)
    c = counters.detach().cpu().numpy()
    r = None if ret_sums is None else ret_sums.detach().cpu().numpy()
    out = []
    for g in range(c.shape[0]):
        n = int(c[g, CNT_EPISODES])
        if n == 0:
            out.append({})
            continue
        mean_len = c[g, CNT_SUM_STEPS] / n
        var_len = max(c[g, CNT_SUM_STEPS_SQ] / n - mean_len * mean_len, 0.0)
        m = {
            "grasp_success_rate": c[g, CNT_SUCCESSES] / n,
            "mean_episode_length": float(mean_len),
            "std_episode_length": float(var_len ** 0.5),
            "failure_type_frequency": {name: {"count": int(c[g, CNT_LABEL_METRICS + k]),
                                              "frequency": float(c[g, CNT_LABEL_METRICS + k] / n)}
                                       for k, name in enumerate(LABELS_METRICS)},
            "failure_mode_frequency": {name: {"count": int(c[g, CNT_LABEL_TAXONOMY + k]),
                                              "frequency": float(c[g, CNT_LABEL_TAXONOMY + k] / n)}
                                       for k, name in enumerate(LABELS_TAXONOMY)},
            "total_episodes": n,
            "successful_episodes": int(c[g, CNT_SUCCESSES]),
            "failed_episodes": n - int(c[g, CNT_SUCCESSES]),
            "mean_contacts": float(c[g, CNT_SUM_FINAL_CONTACTS] / n),
            "variance_ties": int(c[g, CNT_VAR_TIES]),
        }
        if r is not None:
            mean_r = r[g, 0] / n
            m["mean_reward"] = float(mean_r)
            m["std_reward"] = float(max(r[g, 1] / n - mean_r * mean_r, 0.0) ** 0.5)
        out.append(m)
    return out


def share_best_candidate(best_value: float, best_gid: int, payload: torch.Tensor, group=None):
    """Cross-rank arg-max with payload: every rank proposes ``(best_value, best_gid, payload)``; all ranks return the
    proposal with the highest value (ties: lowest global id), identical on every rank.

    This is the one exchange of a SHARED SimpleLearner (BASELINE.json north_star: "NCCL ... and, if training is batched,
    the simple_learner update"; policies/simple_learner.py:73-95 keeps whatever adjustment improved the best reward):
    the winning rank's mean action [15] replaces everybody's.  One all-gather of 17 float64 per rank -- latency-bound,
    issued once per shared update, never per env-step.  Works with ``nccl`` (CUDA payloads) and ``gloo`` (CPU)."""
    payload = payload.detach().reshape(-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(best_value), int(best_gid), payload.clone()
    mine = torch.cat([torch.tensor([float(best_value), float(best_gid)], dtype=torch.float64, device=payload.device),
                      payload.to(torch.float64)])
    world = dist.get_world_size(group)
    table = torch.empty(world, mine.numel(), dtype=torch.float64, device=payload.device)
    dist.all_gather_into_tensor(table, mine, group=group) if payload.is_cuda else \
        dist.all_gather(list(table.unbind(0)), mine, group=group)
    rows = table.cpu()
    order = sorted(range(world), key=lambda r: (-float(rows[r, 0]), float(rows[r, 1])))
    w = order[0]
    return float(rows[w, 0]), int(rows[w, 1]), table[w, 2:].to(payload.dtype)
