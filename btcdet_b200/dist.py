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


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_gpu(device_index, local_rank=0, local_world=1):
    """Pin this process to cores of the NUMA node its GPU hangs off (a disjoint slice per local rank), BEFORE pinned host
    buffers are allocated, so that first-touch places them on that node: with one process per GPU the host side of the
    end-to-end path (37 MB of pinned H2D + D2H per step and GPU) otherwise crosses the socket interconnect for half of
    the GPUs (round-1 SCALE: e2e efficiency 0.59 at 8 GPUs with every rank on cores 0-31 / node 0).
    Returns a small dict for the bench line; never raises (containers may forbid any of this)."""
    import os
    info = {"numa_node": None, "cores": None, "bound": False}
    try:
        import torch

        def node_of(idx):
            p = torch.cuda.get_device_properties(idx)
            bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
            return int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())

        node = node_of(device_index)
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = os.sched_getaffinity(0)
        mine = sorted(cpus & allowed)
        if not mine:
            info["note"] = "the GPU's node has no core in this process' allowed set (%d allowed cores)" % len(allowed)
            return info
        # the ranks whose GPUs hang off the same node share its cores evenly (at least two cores each: main thread + copy /
        # NCCL threads); with one process per GPU the local rank is the device index
        peers = [g for g in range(min(torch.cuda.device_count(), max(local_world, 1))) if node_of(g) == node] or [device_index]
        pos = peers.index(device_index) if device_index in peers else local_rank % len(peers)
        per = max(2, len(mine) // len(peers))
        lo = (pos * per) % len(mine)
        sl = (mine + mine)[lo:lo + per]
        os.sched_setaffinity(0, set(sl))
        info.update({"cores": len(sl), "bound": True, "ranks_on_node": len(peers)})
    except Exception as e:      # noqa: BLE001
        info["note"] = repr(e)[:120]
    return info
