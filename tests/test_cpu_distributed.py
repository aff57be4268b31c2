"""world_size-2 gloo tests of the N>1 host logic (CPU): shard ranges, the counter all-reduce, and
that per-shard counter tables add up to the single-process table (the oracle produces the counters
here; on the GPU the same is asserted with the CUDA kernels in test_sharding_is_gpu_count_independent)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dexterous_rl_manipulation_b200 as dx


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_global, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    lo, hi = dx.distributed.shard_range(n_global, rank, world)
    cfgs = [dx.CurriculumConfig.easy(), dx.CurriculumConfig.hard()]
    groups = np.concatenate([oracle.make_group(c) for c in cfgs])
    n = hi - lo
    ob = oracle.OracleBatch(n, dense=True, max_episode_steps=30)
    draws = [oracle.reset_draws(9, lo + i, 0, groups[(lo + i) % 2:(lo + i) % 2 + 1]) for i in range(n)]
    ob.reset_predrawn(np.stack([d[0] for d in draws]), np.array([d[1] for d in draws]), np.array([d[2] for d in draws]),
                      np.array([d[3] for d in draws]), np.stack([d[4] for d in draws]))
    cnt, rs = ob.rollout(groups, 70, 9, policy_kind=2, env_gid0=lo, loop_max_steps=30)
    counters, ret_sums = torch.from_numpy(cnt.copy()), torch.from_numpy(rs.copy())
    dx.distributed.allreduce_counters(counters, ret_sums)
    if rank == 0:
        np.save(os.path.join(out_dir, "counters.npy"), counters.numpy())
        np.save(os.path.join(out_dir, "ret_sums.npy"), ret_sums.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_counter_allreduce_matches_single_process(tmp_path):
    from oracle import oracle
    n_global, world = 301, 2
    mp.spawn(_worker, args=(world, _free_port(), n_global, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "counters.npy")
    got_rs = np.load(tmp_path / "ret_sums.npy")
    cfgs = [dx.CurriculumConfig.easy(), dx.CurriculumConfig.hard()]
    groups = np.concatenate([oracle.make_group(c) for c in cfgs])
    ob = oracle.OracleBatch(n_global, dense=True, max_episode_steps=30)
    draws = [oracle.reset_draws(9, i, 0, groups[i % 2:i % 2 + 1]) for i in range(n_global)]
    ob.reset_predrawn(np.stack([d[0] for d in draws]), np.array([d[1] for d in draws]), np.array([d[2] for d in draws]),
                      np.array([d[3] for d in draws]), np.stack([d[4] for d in draws]))
    cnt, rs = ob.rollout(groups, 70, 9, policy_kind=2, env_gid0=0, loop_max_steps=30)
    assert cnt[:, 0].sum() > n_global
    assert np.array_equal(got, cnt)
    np.testing.assert_allclose(got_rs, rs, rtol=1e-12)


def test_allreduce_is_identity_without_process_group():
    c = torch.arange(36, dtype=torch.int64).reshape(2, 18)
    out, _ = dx.distributed.allreduce_counters(c.clone())
    assert torch.equal(out, c)


class _Obj:
    def __init__(self, size, mass, friction):
        self.size, self.mass, self.friction = size, mass, friction


class _HeldOut:
    """Duck type of evaluation/heldout_objects.py::HeldOutObjectSet."""

    def __init__(self, n):
        self.heldout_objects = [_Obj(0.03 + 0.01 * k, 0.1 + 0.01 * k, 0.3 + 0.02 * k) for k in range(n)]

    def get_eval_config(self, k):
        o = self.heldout_objects[k]
        return dx.CurriculumConfig(object_size=o.size, object_mass=o.mass, friction_coefficient=o.friction)


def _fake_shard(heldout_set, policy, n_eps, seeds, reward_type, max_episode_steps, device, policy_seed,
                contact_history, actions, lo, hi):
    """Stands in for the CUDA launch of one rank's slice: records are a pure function of the GLOBAL env index,
    exactly the property the real shard has (Philox keys, reset seeds and groups depend on the global id only)."""
    from dexterous_rl_manipulation_b200 import _lib
    out = {}
    for g in range(lo, hi):
        rng = np.random.default_rng(1000 + g)
        steps = int(rng.integers(1, max_episode_steps + 1))
        counts = rng.integers(0, 6, steps).astype(np.uint8)
        rec = np.zeros((), _lib.EPISODE_RECORD_DTYPE)
        rec["env_gid"], rec["episode"], rec["steps"] = g, 0, steps
        rec["success"] = int(counts[-1] >= 3)
        rec["final_contacts"] = int(counts[-1])
        rec["label_metrics"] = _lib.LABEL_NONE if rec["success"] else int(rng.integers(0, 6))
        rec["label_taxonomy"] = _lib.LABEL_NONE if rec["success"] else int(rng.integers(0, 6))
        rec["episode_reward"] = float(rng.normal())
        out[g] = (rec, counts)
    return out


def _frontend_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dexterous_rl_manipulation_b200 import evaluation
    evaluation._heldout_shard = _fake_shard
    res = evaluation.evaluate_heldout_set_batched(_HeldOut(7), policy="heuristic", num_episodes_per_object=3, seed=11,
                                                  max_episode_steps=40)
    import pickle
    with open(os.path.join(out_dir, f"res{rank}.pkl"), "wb") as fh:
        pickle.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_heldout_front_end_gathers_the_single_process_result(tmp_path):
    """evaluate_heldout_set_batched under a 2-rank gloo group: each rank evaluates its slice of the (object, episode)
    batch, the records are all-gathered, and EVERY rank returns exactly the single-process result."""
    import pickle
    from dexterous_rl_manipulation_b200 import evaluation
    world = 2
    mp.spawn(_frontend_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    orig = evaluation._heldout_shard
    evaluation._heldout_shard = _fake_shard
    try:
        single = evaluation.evaluate_heldout_set_batched(_HeldOut(7), policy="heuristic", num_episodes_per_object=3, seed=11,
                                                         max_episode_steps=40)
    finally:
        evaluation._heldout_shard = orig
    assert single["overall_stats"]["total_episodes"] == 21
    for rank in range(world):
        with open(tmp_path / f"res{rank}.pkl", "rb") as fh:
            got = pickle.load(fh)
        assert got == single, rank


def _share_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 1 holds the better candidate in round 0; round 1 is a tie on the value, decided by the lower global id
    rounds = [((1.5, 7, torch.full((15,), 0.25)), (2.5, 900, torch.arange(15, dtype=torch.float32))),
              ((3.0, 40, torch.full((15,), -1.0)), (3.0, 12, torch.full((15,), 2.0)))]
    got = []
    for cand in rounds:
        v, g, payload = cand[rank]
        got.append(dx.distributed.share_best_candidate(v, g, payload))
    torch.save(got, os.path.join(out_dir, f"share_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shared_learner_exchange_two_ranks(tmp_path):
    """distributed.share_best_candidate (the shared SimpleLearner's one collective): both ranks end up with the same
    winner -- highest value, ties to the lowest global env id -- and its payload."""
    mp.spawn(_share_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "share_0.pt"), torch.load(tmp_path / "share_1.pt")
    for (va, ga, pa), (vb, gb, pb) in zip(a, b):
        assert (va, ga) == (vb, gb) and torch.equal(pa, pb)
    assert (a[0][0], a[0][1]) == (2.5, 900) and torch.equal(a[0][2], torch.arange(15, dtype=torch.float32))
    assert (a[1][0], a[1][1]) == (3.0, 12) and torch.equal(a[1][2], torch.full((15,), 2.0))
    # without a process group the call is the identity
    v, g, p = dx.distributed.share_best_candidate(1.0, 3, torch.ones(15))
    assert (v, g) == (1.0, 3) and torch.equal(p, torch.ones(15))
