"""N>1 host logic on CPU: world_size-2 gloo process group (scene sharding, timing reduction, aggregate throughput)."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    from btcdet_b200 import dist as bd
    r, w = bd.init("gloo")
    mine = list(bd.shard(11, r, w))
    slowest = bd.max_over_ranks(0.5 + r)                       # rank 1 is slower
    thr = bd.aggregate_throughput(len(mine), 0.5 + r)
    # a data-parallel "gradient all-reduce": every rank ends with the mean
    g = torch.full((4,), float(r + 1))
    torch.distributed.all_reduce(g)
    g /= w
    out[rank] = (mine, slowest, thr, g.tolist())
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_reductions():
    world, port = 2, 29571
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (m0, s0, t0, g0), (m1, s1, t1, g1) = out[0], out[1]
    assert sorted(m0 + m1) == list(range(11)) and not set(m0) & set(m1) and abs(len(m0) - len(m1)) <= 1
    assert s0 == s1 == 1.5                                   # max over ranks
    assert t0 == t1 == 11 / 1.5                              # all units / slowest rank
    assert g0 == g1 == [1.5] * 4


def test_shard_covers_everything_for_any_world():
    from btcdet_b200 import dist as bd
    for n in (0, 1, 7, 16, 24):
        for world in (1, 2, 4, 8):
            parts = [list(bd.shard(n, r, world)) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_cpulist_and_host_binding_never_raise():
    """dist.bind_host_to_gpu must degrade to a no-op (returning its reason) wherever it cannot bind — here: no GPU."""
    from btcdet_b200 import dist
    assert dist._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert dist._cpulist("") == set()
    info = dist.bind_host_to_gpu(0, local_rank=1, local_world=4)
    assert set(info) >= {"numa_node", "cores", "bound"} and info["bound"] in (True, False)
