"""World-size-2 gloo run (CPU) of the host-side multi-GPU logic: fragment sharding and the timing all-gather, the only collective
of the path (fragments are independent: SURVEY.md 8e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imfnet_b200.pipeline import aggregate_throughput, gather_records, shard_indices


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, weights, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_indices(len(weights), rank, world, weights)
        voxels = float(sum(weights[i] for i in mine))
        rec = torch.tensor([rank, voxels, 10.0 * (rank + 1)], dtype=torch.float64)      # rank 1 is the slow one
        allrec = gather_records(rec, world)
        value, ms = aggregate_throughput(allrec)
        ids = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
        pad = torch.full((8,), -1, dtype=torch.int64)
        pad[: len(mine)] = torch.tensor(mine)
        dist.all_gather(ids, pad)
        ret[rank] = (value, ms, allrec.tolist(), [t.tolist() for t in ids])
    finally:
        dist.destroy_process_group()


def test_sharding_and_timing_allgather_world2():
    weights = [50000, 20000, 30000, 50000, 10000, 40000]
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, weights, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    v0, ms0, rec0, ids0 = ret[0]
    v1, ms1, rec1, ids1 = ret[1]
    assert rec0 == rec1 and ids0 == ids1                        # every rank sees the same gathered records
    got = sorted(i for row in ids0 for i in row if i >= 0)
    assert got == list(range(len(weights)))                     # each fragment on exactly one rank
    assert ms0 == ms1 == 20.0                                   # job time = slowest rank
    assert abs(v0 - sum(weights) / 20e-3) < 1e-6 and v0 == v1   # job work = sum over ranks
    loads = [r[1] for r in rec0]
    assert abs(loads[0] - loads[1]) <= max(weights)             # greedy balance


def test_world_size_1_needs_no_process_group():
    rec = torch.tensor([0.0, 123.0, 4.0], dtype=torch.float64)
    allrec = gather_records(rec, 1)
    assert allrec.shape == (1, 3)
    assert aggregate_throughput(allrec) == (123.0 / 4e-3, 4.0)
