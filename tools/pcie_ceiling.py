"""Host <-> device copy ceiling of the end-to-end step (no kernels): per bench step every rank uploads its actions
(E x 15 float32) and downloads what step_host() returns (41 observation rows, reward, three flag bytes per env) from /
to pinned host memory, both directions at once on separate streams -- the traffic of `e2e` in bench.py and nothing else.

    python tools/pcie_ceiling.py [--num-envs E] [--steps K]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py

Prints one JSON line (rank 0): env-steps/s the copies alone would allow (max over ranks), GB/s per direction per GPU.
bench.py imports `measure()` to report e2e.frac_of_copy_ceiling.
"""
import argparse
import json
import os
import time


def measure(num_envs, steps=20, warmup=3, device=None, packed_contacts=False, barrier=None, rows=None, extra_bytes_per_env=None):
    """Seconds per step of the copies alone on this rank (caller takes the max over ranks).
    ``rows`` observation rows of 4 bytes + ``extra_bytes_per_env`` (reward, flags, masks) come down per env; the defaults are
    41 + 7 (every row but the constant quaternion) or 36 + 8 with ``packed_contacts``."""
    import torch
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    E = int(num_envs)
    ld = (E + 31) // 32 * 32
    rows = rows if rows is not None else (36 if packed_contacts else 41)
    h_act = torch.empty(E, 15).pin_memory()
    d_act = torch.empty(E, 15, device=dev)
    d_out = torch.empty(rows, ld, device=dev)
    h_out = torch.empty(rows, ld).pin_memory()
    small = extra_bytes_per_env if extra_bytes_per_env is not None else (4 + 3 + (1 if packed_contacts else 0))
    d_small = torch.empty(E * small, dtype=torch.uint8, device=dev)
    h_small = torch.empty_like(d_small, device="cpu").pin_memory()
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def one():
        with torch.cuda.stream(up):
            d_act.copy_(h_act, non_blocking=True)
        with torch.cuda.stream(down):
            h_out.copy_(d_out, non_blocking=True)
            h_small.copy_(d_small, non_blocking=True)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize(dev)
    if barrier is not None:
        barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    h2d = E * 15 * 4
    d2h = rows * ld * 4 + d_small.numel()
    return {"seconds_per_step": dt, "h2d_bytes": h2d, "d2h_bytes": d2h, "h2d_gbs": h2d / dt / 1e9, "d2h_gbs": d2h / dt / 1e9}


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bar = (lambda: (dist.barrier(), torch.cuda.synchronize(dev))) if world > 1 else None
    out = {}
    for name, packed in (("full", False), ("packed_contacts", True)):
        m = measure(args.num_envs, args.steps, device=dev, packed_contacts=packed, barrier=bar)
        t = torch.tensor([m["seconds_per_step"]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        m["seconds_per_step_max_over_ranks"] = float(t.item())
        m["env_steps_per_sec_ceiling"] = args.num_envs * world / float(t.item())
        out[name] = m
    if rank == 0:
        print(json.dumps({"n_gpus": world, "envs_per_gpu": args.num_envs, **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
