"""world_size-2 gloo test of the multi-GPU plumbing: static sharding + the final trace gather."""
import datetime
import os
import queue
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
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
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


def _first_result(q, procs, seconds=240):
    """What rank 0 put on the queue; a rank that died (or a rendezvous that never completes) fails the test instead of
    blocking the suite."""
    try:
        return q.get(timeout=seconds)
    except queue.Empty:
        for p in procs:
            if p.is_alive():
                p.terminate()
        raise AssertionError(f"no result within {seconds} s (exit codes {[p.exitcode for p in procs]})")


def test_gather_traces_world2_ragged():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    B, T = 7, 5                        # 4 + 3 campaigns: ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    tr, vals = _first_result(q, procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = (torch.arange(B).unsqueeze(1) * 100 + torch.arange(T).unsqueeze(0)).to(torch.int32)
    assert torch.equal(tr, want) and torch.allclose(vals, want.double() / 7.0)


def _fake_executor(pools, campaigns, on_done=None):
    """What the device engine returns, computed on the CPU from the campaign itself (so any rank can be checked)."""
    res = {}
    for i, c in enumerate(campaigns):
        N = len(pools[c.pool][0])
        k = len(c.design) + (2 if c.seed % 5 == 0 else c.n_trials)       # some campaigns stop early with a status
        res[i] = (list(c.design) + [(c.seed * 7 + j) % N for j in range(k - len(c.design))], 3 if c.seed % 5 == 0 else 0)
    return res


def _sweep_worker(rank, world, port, tmp, q):
    """The product path under torch.distributed: runners.run_sweep shards the campaigns of ALL jobs over the ranks,
    gathers the traces on rank 0; a job that ran entirely on one rank is written by that rank (here b.csv by rank 1), the
    others by rank 0."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    from stpbenchmark_b200 import runners
    g = torch.Generator().manual_seed(0)
    jobs = []
    for name, N, d, cls, seeds in (("a", 40, 2, runners.ExactGP, [1, 2, 3, 5, 8]), ("b", 55, 3, runners.VarTGP, [10, 11, 12])):
        X = torch.rand(N, d, dtype=torch.double, generator=g); y = torch.rand(N, dtype=torch.double, generator=g)
        jobs.append(dict(X=X, y=y, model_class=cls, seeds=seeds, output_path=os.path.join(tmp, f"{name}.csv"),
                         n_initial=10, n_trials=6))
    frames = runners.run_sweep(jobs, _executor=_fake_executor)
    dist.barrier()
    if rank == 0:
        q.put("done")
    dist.destroy_process_group()


def test_run_sweep_shards_over_two_ranks_and_owning_ranks_write_the_csvs(tmp_path):
    import pandas as pd
    from stpbenchmark_b200 import runners, utils
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    assert _first_result(q, procs) == "done"
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # the same sweep in one process gives the same files
    g = torch.Generator().manual_seed(0)
    for name, N, d, seeds in (("a", 40, 2, [1, 2, 3, 5, 8]), ("b", 55, 3, [10, 11, 12])):
        X = torch.rand(N, d, dtype=torch.double, generator=g); y = torch.rand(N, dtype=torch.double, generator=g)
        back = pd.read_csv(tmp_path / f"{name}.csv", header=[0, 1], index_col=0)
        designs = utils.initial_designs_batch(X, y, seeds, 10, 50.0)
        for k, s_ in enumerate(seeds):
            want = designs[k].tolist() + [(s_ * 7 + j) % N for j in range(2 if s_ % 5 == 0 else 6)]
            col = back[f"seed_{s_}", "x_index"]
            assert col.dropna().astype(int).tolist() == want and len(col) == 16
