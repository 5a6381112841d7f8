"""Multi-GPU plumbing: one process per GPU, scenes sharded by rank, no collective on the data path.

The reference's only strategy is data parallelism — DistributedSampler + DDP (tools/train.py:166-168,
btcdet/datasets/__init__.py:54-59), per-GPU batch = global // total_gpus (tools/train.py:82-83).  Scenes are
independent, so the forward hot path needs no exchange; NCCL (NVLink 5 / NVSwitch) only carries the gradient
all-reduce of a training step and the timing reduction of the benchmark.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_* (torchrun); no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"))
    return rank, world


def shard(n_items, rank, world):
    """Indices of the items rank `rank` owns: contiguous, balanced, disjoint, covering range(n_items)
    (same split rule as torch's DistributedSampler without padding)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def max_over_ranks(value, device="cpu"):
    """MAX all-reduce of a python float (benchmark timing: the job is as slow as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, seconds_this_rank, device="cpu"):
    """Whole-job units/s = all ranks' units / slowest rank's time."""
    return sum_over_ranks(units_this_rank, device) / max_over_ranks(seconds_this_rank, device)
