"""world_size-2 gloo test of the multi-GPU plumbing: static sharding + the final trace gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stpbenchmark_b200 import distributed as sd


def test_shard_bounds_cover_everything_once():
    for B in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sd.shard_bounds(B, world, r)
                got += list(range(lo, hi))
            assert got == list(range(B))
    assert sd.shard(list("abcdefg"), 3, 2) == ["g"]


def _worker(rank, world, port, B, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sd.shard_bounds(B, world, rank)
    trace = (torch.arange(lo, hi).unsqueeze(1) * 100 + torch.arange(T).unsqueeze(0)).to(torch.int32)
    vals = trace.double() / 7.0
    out = sd.gather_traces(trace, vals, B)
    if rank == 0:
        q.put((out[0].clone(), out[1].clone()))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_traces_world2_ragged():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    B, T = 7, 5                        # 4 + 3 campaigns: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    tr, vals = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = (torch.arange(B).unsqueeze(1) * 100 + torch.arange(T).unsqueeze(0)).to(torch.int32)
    assert torch.equal(tr, want) and torch.allclose(vals, want.double() / 7.0)
