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
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device(f"cuda:{local}"))
        else:
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
    if hi > lo:
        local = evaluate(params_rows[lo:hi])
    else:
        # an empty shard (fewer vectors than ranks) still joins the collective, on the device
        # the backend works on
        on_gpu = dist.is_initialized() and dist.get_backend() == "nccl"
        local = torch.zeros(0, dtype=torch.float64,
                            device=torch.device("cuda", torch.cuda.current_device()) if on_gpu else "cpu")
    return all_gather_ragged(local, params_rows.shape[0], world)


class GatherPipeline:
    """All-gather of per-rank log-likelihood shards that overlaps the next step: `submit` starts
    the collective (async on NCCL's stream, ordered after the producing kernels) and returns at
    once; `collect` hands out the gathered vector of the PREVIOUS submit.  One collective is in
    flight at a time, so a step costs its own kernels, not kernels + collective latency."""

    def __init__(self, total, world=None):
        self.total = int(total)
        self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else world
        self.sizes = [shard_bounds(self.total, r, self.world)[1] - shard_bounds(self.total, r, self.world)[0]
                      for r in range(self.world)]
        self.width = max(self.sizes) if self.sizes else 0
        self._pending = None

    def submit(self, local):
        if self.world == 1:
            self._pending = (None, local)
            return
        padded = torch.zeros(self.width, dtype=local.dtype, device=local.device)
        padded[:local.numel()] = local
        out = torch.empty(self.world * self.width, dtype=local.dtype, device=local.device)
        work = dist.all_gather_into_tensor(out, padded, async_op=True)
        self._pending = (work, out)

    def collect(self):
        if self._pending is None:
            return None
        work, out = self._pending
        self._pending = None
        if work is None:
            return out
        work.wait()               # the current stream waits for the collective; the host does not
        return torch.cat([out[r * self.width:r * self.width + self.sizes[r]] for r in range(self.world)])


def chromevol_model_configurations():
    """Candidate set for model selection beyond the reference's two (SURVEY.md section 8(a) notes):
    rate subsets mirroring ChromEvol's named models.  Each entry is an ordinary
    `model_configurations` item of calculate_model_selection_criteria (AR:737-812), so every
    candidate is checkable with the same oracle."""
    zero = dict(gain=0.1, loss=0.1, dupl=0, demidupl_gain=0, demidupl_loss=0, base_number_gain=0,
                base_number_loss=0, linear_rate=0)
    def cfg(name, free, **init):
        return {'name': name, 'initial_params': dict(zero, **init), 'params_to_optimize': free}
    return [
        cfg('GainLossOnly', ['gain', 'loss']),
        cfg('ConstRate', ['gain', 'loss', 'dupl'], dupl=0.05),
        cfg('ConstRateDemi', ['gain', 'loss', 'dupl', 'demidupl_gain'], dupl=0.05, demidupl_gain=0.01),
        cfg('ConstRateDemiEst', ['gain', 'loss', 'dupl', 'demidupl_gain', 'demidupl_loss'],
            dupl=0.05, demidupl_gain=0.01, demidupl_loss=0.01),
        cfg('ConstRateNoDupl', ['gain', 'loss', 'demidupl_gain'], demidupl_gain=0.01),
        cfg('LinearRate', ['gain', 'loss', 'dupl', 'linear_rate'], dupl=0.05, linear_rate=0.01),
        cfg('LinearRateDemi', ['gain', 'loss', 'dupl', 'demidupl_gain', 'linear_rate'], dupl=0.05,
            demidupl_gain=0.01, linear_rate=0.01),
        {'name': 'FullModelDefaultRates', 'initial_params': None,
         'params_to_optimize': ['gain', 'loss', 'dupl', 'demidupl_gain', 'demidupl_loss',
                                'base_number_gain', 'base_number_loss', 'linear_rate']},
    ]


def sharded_model_selection(select, model_configurations):
    """Model-selection candidates are independent optimiser runs (AR:756 loop bodies share only
    the read-only tree): candidate k runs on rank k % world (round robin, SURVEY.md section 8(d) config
    4) through `select(configs) -> dict name -> result` (calculate_model_selection_criteria
    with the tree and counts bound), and every rank receives the merged dict in candidate order."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = [c for k, c in enumerate(model_configurations) if k % world == rank]
    local = select(mine) if mine else {}
    if world == 1:
        return local
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    merged = {}
    for c in model_configurations:
        for part in gathered:
            if c['name'] in part:
                merged[c['name']] = part[c['name']]
    return merged


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
