"""Sharding of independent likelihood evaluations over the GPUs of one box.

The hot path has no inter-GPU data traffic: the tree, counts and root prior
are replicated, the leading batch axis (parameter vectors / model-selection candidates /
mapping replicates) is split contiguously over ranks, and the only collective is an all-gather
of per-vector log-likelihood scalars (or a sum of integer histograms for the simulator).
One process per GPU; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), \
        int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(total, rank, world):
    """Contiguous, balanced split of range(total): the first total % world ranks get one extra."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_ragged(local, total, world=None):
    """Concatenate every rank's 1-D shard (sizes given by shard_bounds) on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size() if world is None else world
    sizes = [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]
    width = max(sizes)
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    out = torch.empty(world * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width:r * width + sizes[r]] for r in range(world)])


def sharded_log_likelihood(evaluate, params_rows):
    """Evaluate this rank's slice of `params_rows` (B, 9) with `evaluate(rows) -> 1-D tensor`
    and return the full (B,) vector on every rank."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(params_rows.shape[0], rank, world)
    local = evaluate(params_rows[lo:hi]) if hi > lo else torch.zeros(0, dtype=torch.float64)
    return all_gather_ragged(local, params_rows.shape[0], world)


def sum_across_ranks(tensor):
    """In-place sum (histograms of the simulator); no-op for a single process."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def max_across_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shutdown():
    """Destroy the default process group if this process created one."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
