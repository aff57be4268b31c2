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
