"""Multi-GPU sharding of a rollout: environments are independent (python/mujoco/rollout.cc:85-177 has
no cross-env data), so rank r of G owns the contiguous env range [r*n/G, (r+1)*n/G) and steps it with
its own Batch; there is NO per-step collective.  The only exchange is at the rollout boundary: one
all-gather of the per-env results (episode returns / final states) — SURVEY.md section 8(e).

Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(nenv_total, rank, world):
    """contiguous env range [lo, hi) owned by `rank`; remainders go to the lowest ranks"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(nenv_total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_env_rows(local_rows, nenv_total, group=None, device=None):
    """all-gather per-env rows [n_local, k] (numpy or torch) from every rank into [nenv_total, k], in env
    order.  Shards may differ by one env: rows are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.as_tensor(np.ascontiguousarray(local_rows) if isinstance(local_rows, np.ndarray) else local_rows)
    if device is not None:
        t = t.to(device)
    if t.dim() == 1:
        t = t[:, None]
    nmax = -(-int(nenv_total) // world)
    pad = torch.zeros((nmax, t.shape[1]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = torch.empty((world * nmax, t.shape[1]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_range(nenv_total, r, world)
        rows.append(out[r * nmax: r * nmax + (hi - lo)])
    return torch.cat(rows, 0)


def max_over_ranks(value, group=None, device=None):
    """max of a python float over ranks (timing: a multi-GPU step is as slow as its slowest rank)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
