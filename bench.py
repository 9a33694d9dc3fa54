#!/usr/bin/env python
"""bench.py -- BO iterations/s (fit + predict + acquire + update) over batched campaigns.

Workload (BASELINE.json configs[3], the north-star shape): synthetic pools N = 2048, d = 6
(negated Hartmann-6 + 0.05 t_3 noise, 8 pools), 1024 concurrent ExactGP campaigns PER GPU,
n_initial = 10, 80 trials each.  Campaigns are kept at staggered ages (n = 10 ... 89 uniformly,
"continuous batching": a campaign that finishes its 80 trials is replaced by a fresh seed), so
that every step is one BO iteration of every campaign at the campaign-average cost, whatever
--steps is.  One step = one launch of stp_exact_bo_step per campaign group of the rank (4 groups of 256 campaigns
on 4 CUDA streams, so that the tail of one group's launch overlaps the others).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # CUDA arm (N > 1 under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]    # CPU arm: the oracle on host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POOL, DIM, N_POOLS = 2048, 6, 8
N_INITIAL, N_TRIALS = 10, 80
CAP = N_INITIAL + N_TRIALS + 1
METRIC, UNIT = "bo_iterations_per_sec", "it/s"
NCU_DRAM_BYTES_PER_LAUNCH = 6.717696e6 + 0.091904e6   # measured once with ncu (see roofline.traffic_source)
WORKLOAD = ("cfg4: synthetic pools N=2048 d=6 (Hartmann-6 + t3 noise), {B} concurrent ExactGP campaigns per GPU, "
            "n_initial=10, 80 trials, ages staggered over n=10..89 (continuous batching)")


# ------------------------------------------------------------------------------------------
# CPU workers (the oracle).  Forked BEFORE CUDA is initialised; they never touch the GPU.
# ------------------------------------------------------------------------------------------
def _cpu_init():
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)


def _cpu_one_iteration(job):
    """One oracle BO iteration from a given campaign state -> (picked pool index, seconds, nfev)."""
    import numpy as np
    import torch
    from oracle.loop import bo_iteration
    Xp, yp, picked = job
    Xp = torch.from_numpy(Xp); yp = torch.from_numpy(yp)
    mask = torch.ones(len(Xp), dtype=torch.bool); mask[picked] = False
    info = {}
    t0 = time.perf_counter()
    best, _ = bo_iteration("ExactGP", Xp[picked], yp[picked], Xp[mask], info=info)
    dt = time.perf_counter() - t0
    return int(torch.where(mask)[0][best]), dt, int(info["nfev"])


def _cpu_campaign(job):
    """`iters` consecutive oracle BO iterations of one campaign that starts at a given age (training set = Sobol
    initial design + `age` random pool points, then free-running) -> per-iteration seconds."""
    import torch
    from oracle.loop import bo_iteration
    from stpbenchmark_b200 import synthetic
    g, seed, iters, age = job
    X, y = synthetic.cfg4_pool(g)
    init = synthetic.initial_designs(X, y, [seed], N_INITIAL)[0].tolist()
    gen = torch.Generator().manual_seed(seed)
    extra = [i for i in torch.randperm(len(X), generator=gen).tolist() if i not in set(init)][:age]
    picked = init + extra
    mask = torch.ones(len(X), dtype=torch.bool); mask[picked] = False
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        best, _acq = bo_iteration("ExactGP", X[picked], y[picked], X[mask])
        idx = int(torch.where(mask)[0][best])
        picked.append(idx); mask[idx] = False
        times.append(time.perf_counter() - t0)
    return times


def run_reference_arm(args):
    """--impl reference: the reference path's CPU restatement (oracle/) on every host core.
    A step = iteration t of `cores` independent campaigns, one per core (bounded sample of the
    1024-campaign workload); W + K consecutive iterations of each campaign are run, the first W
    are warm-up."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    W, K = args.warmup, args.steps
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        span = max(N_TRIALS - (W + K), 1)          # ages spread over the campaign so that n covers 10..89 like the GPU mix
        ages = [(c * span) // max(cores - 1, 1) if cores > 1 else 0 for c in range(cores)]
        jobs = [(c % N_POOLS, 900000 + c, W + K, ages[c]) for c in range(cores)]
        t0 = time.perf_counter()
        per = pool.map(_cpu_campaign, jobs)
        wall = time.perf_counter() - t0
    # step time = slowest core at that iteration (all cores advance in lock-step conceptually)
    step_s = [max(p[t] for p in per) for t in range(W, W + K)]
    total = sum(step_s)
    value = cores * K / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(B=args.campaigns), "sample": f"{cores} campaigns (one per host core) started at "
                   f"ages 0..{max(ages)} (Sobol design + age random pool points), {W} warm-up + {K} timed consecutive "
                   f"BO iterations each (n = {N_INITIAL}..{N_INITIAL + max(ages) + W + K - 1})"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} campaigns x {K} consecutive BO iterations, oracle/ (torch fp64 + scipy "
                                   f"L-BFGS-B), 1 thread per campaign; wall {wall:.1f}s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def algorithmic_flops(n, nfev, d, N):
    """W = F (n^3 + n^2 (3d+6)) + n^3/3 + Nc (n^2 + n (3d+8) + 60) per campaign-iteration,
    with the campaign's actual closure count F and Nc = N - n remaining candidates."""
    n = n.double(); F = nfev.double()
    return F * (n ** 3 + n ** 2 * (3 * d + 6)) + n ** 3 / 3 + (N - n) * (n ** 2 + n * (3 * d + 8) + 60)


def run_cuda_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cpu_pool = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_pool = mp.get_context("fork").Pool(min(cores, 64), initializer=_cpu_init)  # before CUDA init

    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (CUDA arm) needs a GPU; use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from stpbenchmark_b200 import _lib, synthetic
    from stpbenchmark_b200.engine import CampaignBatch

    B, W, K = args.campaigns, args.warmup, args.steps
    per_pool = B // N_POOLS
    pools = [synthetic.cfg4_pool(g) for g in range(N_POOLS)]
    PX = torch.stack([p[0] for p in pools]); PY = torch.stack([p[1] for p in pools])
    pool_id = [b // per_pool if per_pool else 0 for b in range(B)]
    pool_id = [min(p, N_POOLS - 1) for p in pool_id]
    NG = max(1, min(args.groups, B // 16 if B >= 16 else 1))
    queue = args.schedule == "queue"           # persistent work-queue kernel: K iterations per campaign in ONE launch
    if queue and args.groups == 16:
        NG = 1
    bounds = [(g * B // NG, (g + 1) * B // NG) for g in range(NG)]
    if args.schedule == "cohort":                             # one age per group, groups evenly staggered
        group_of = [next(g for g, (lo, hi) in enumerate(bounds) if lo <= b < hi) for b in range(B)]
        age0 = [(group_of[b] * N_TRIALS) // NG for b in range(B)]
    else:
        age0 = [b % N_TRIALS for b in range(B)]             # trials already done when timing starts
    # seeds: first generation as config #4 (1000 g + k), later generations continue per slot
    n_steps_total = W + K + K                                # device-resident region + e2e region
    gens = 1 + (n_steps_total + N_TRIALS) // N_TRIALS + 1
    rank_off = 100000 * rank
    seed_of = lambda b, gen: rank_off + synthetic.campaign_seed(pool_id[b], b % max(per_pool, 1)) + 10007 * gen
    t_setup = time.perf_counter()
    designs = {}
    for g in range(N_POOLS):
        slots = [b for b in range(B) if pool_id[b] == g]
        seeds = [seed_of(b, gen) for gen in range(gens) for b in slots]
        if seeds:
            D = synthetic.initial_designs(pools[g][0], pools[g][1], seeds, N_INITIAL)
            for i, (gen, b) in enumerate([(gen, b) for gen in range(gens) for b in slots]):
                designs[(b, gen)] = D[i]
    # The campaigns of a rank are split into NG contiguous groups, each with its own CampaignBatch and CUDA stream.
    # One step of the job = one stp_exact_bo_step launch per group; the launches of different groups overlap, so the
    # tail of one group's launch (a step cannot end before its longest fit: 38..449 closures) is filled by the others.
    total_steps = W + K

    class Group:
        def __init__(self, lo, hi):
            self.lo, self.hi, self.nb = lo, hi, hi - lo
            self.stream = torch.cuda.Stream(device=dev)
            self.batch = CampaignBatch(PX, PY, pool_id[lo:hi], [designs[(b, 0)].tolist() for b in range(lo, hi)],
                                       n_cap=CAP, kind="ExactGP", device=dev)
            self.batch.ragged = args.schedule != "cohort"    # ages differ: longest fits are scheduled first
            # the next initial designs of every slot: the kernel recycles finished campaigns itself (stp_turnover)
            self.bank = torch.stack([torch.stack([designs[(b, g)] for g in range(1, gens)]) for b in range(lo, hi)])
            self.host = None

        def preroll(self):
            # untimed: bring campaign b to age0[b]; it starts at step N_TRIALS - age0[b]
            start = torch.tensor([N_TRIALS - age0[b] for b in range(self.lo, self.hi)], device=dev)
            bt = self.batch
            if args.schedule == "cohort":                    # the whole group runs its common age
                for t in range(age0[self.lo]):
                    bt.step()
            elif queue:                                      # one persistent launch: campaign b does age0[b] iterations
                bt.run_queue([age0[b] for b in range(self.lo, self.hi)])
            else:
                bt.status.fill_(-1)                          # not started yet
                for t in range(N_TRIALS):
                    bt.status[start == t] = 0
                    bt.step()
                bt.status[bt.status == -1] = 0               # age-0 campaigns start now
                bt.assume_ages([N_INITIAL + age0[b] for b in range(self.lo, self.hi)])
            bt.enable_turnover(self.bank, n_final=N_INITIAL + N_TRIALS, sink_capacity=self.nb * 4)

        def step(self):
            self.batch.step()                                # argsort (longest first) + ONE kernel; nothing else

        def e2e_step(self):
            bt = self.batch
            for kname, hv in self.host.items():
                getattr(bt, kname).copy_(hv, non_blocking=True)
            bt.step()
            for kname, hv in self.host.items():
                hv.copy_(getattr(bt, kname), non_blocking=True)

    groups = [Group(lo, hi) for lo, hi in bounds]
    torch.cuda.synchronize()                                 # construction ran on the default stream
    for gr in groups:
        with torch.cuda.stream(gr.stream):
            gr.preroll()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_groups(fn):
        for gr in groups:
            with torch.cuda.stream(gr.stream):
                fn(gr)

    # ---- device-resident arm: W warm-up + K timed steps ------------------------------------
    if queue:
        all_groups(lambda gr: gr.batch.run_queue(W))
    else:
        for _ in range(W):
            all_groups(lambda gr: gr.step())
    barrier()
    def totals():
        return (float(sum(gr.batch.work_acc.sum().item() for gr in groups)),
                int(sum(gr.batch.iters_acc.sum().item() for gr in groups)))
    flops0, iters0 = totals()
    sampler = ClockSampler(local_rank); sampler.start()
    launches0 = sum(gr.batch.launches for gr in groups)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev0.record()                                             # default stream, right after the synchronize
    for gr in groups:
        gr.stream.wait_event(ev0)
    if queue:
        all_groups(lambda gr: gr.batch.run_queue(K))
    else:
        for k in range(K):
            all_groups(lambda gr: gr.step())
    ends = []
    for gr in groups:
        e = torch.cuda.Event(enable_timing=True)
        e.record(gr.stream)
        ends.append(e)
    barrier()
    clocks = sampler.stop()
    gpu_launches = sum(gr.batch.launches for gr in groups) - launches0
    elapsed_ms = max(ev0.elapsed_time(e) for e in ends)
    t = torch.tensor([elapsed_ms], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms_max = float(t.item())
    flops1, iters1 = totals()
    status_bad = int(sum(int(((gr.batch.status != 0) & (gr.batch.status != 6)).sum().item()) for gr in groups))
    done = torch.tensor([iters1 - iters0], dtype=torch.int64, device=dev)   # completed iterations, counted by the kernel
    if world > 1:
        dist.all_reduce(done)
    value = float(done.item()) / (elapsed_ms_max * 1e-3)
    launches_all = torch.tensor([gpu_launches], dtype=torch.int64, device=dev)       # whole job, like `value`
    if world > 1:
        dist.all_reduce(launches_all)

    # roofline of the dominant (only) kernel: algorithmic flops (accumulated by the kernel itself from every
    # campaign's n and actual closure count) / the time its launches ran in
    achieved_tflops = (flops1 - flops0) / (elapsed_ms * 1e-3) / 1e12
    fp64_fma_peak = _lib.fp64_probe(4096)
    fp64_mma_peak = _lib.fp64_mma_probe(4096)
    fp64_peak = max(fp64_fma_peak, fp64_mma_peak)
    n_bar = N_INITIAL + (N_TRIALS - 1) / 2.0
    bytes_per_iter = (N_POOL - n_bar) * DIM * 8 + N_POOL + n_bar * (DIM + 1) * 8 + 16 + 24
    hbm_bytes = bytes_per_iter * (iters1 - iters0)
    hbm_gbs = hbm_bytes / (elapsed_ms * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    mean_closures = float(torch.cat([gr.batch.nfev for gr in groups]).double().mean().item())

    # ---- end-to-end arm: same step through HOST buffers (pinned), H2D + launch + D2H per step --
    for gr in groups:
        gr.host = {k_: getattr(gr.batch, k_).cpu().pin_memory() for k_ in ("X", "y", "n", "mask", "trace", "status", "gen")}
    h2d = sum(v.numel() * v.element_size() for gr in groups for v in gr.host.values())
    d2h = h2d

    def e2e_step():
        # every group's state lives in HOST buffers between its steps: before a group's next step the host waits
        # for that group's previous device->host copy (its stream), not for the other groups
        for gr in groups:
            gr.stream.synchronize()
            with torch.cuda.stream(gr.stream):
                gr.e2e_step()
    for _ in range(min(W, 2)):
        e2e_step()
    barrier()
    Ke = K                                         # as many steps as the device-resident arm
    _, it0 = totals()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3      # host-side clock: the region ends with a device synchronize
    _, it1 = totals()
    t = torch.tensor([e2e_ms], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    de = torch.tensor([it1 - it0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(de)
    e2e_value = float(de.item()) / (float(t.item()) * 1e-3)

    sink = torch.cat([gr.batch.sink for gr in groups])
    # ---- the one collective of the path: gather the finished traces on rank 0 ------------------
    if world > 1:
        gathered = [torch.empty_like(sink) for _ in range(world)] if rank == 0 else None
        dist.gather(sink, gathered, dst=0)

    # ---- CPU baseline (rank 0): oracle BO iterations from the GPU's own campaign states ----------
    cpu = None
    parity = None
    if cpu_pool is not None:
        tr = torch.cat([gr.batch.trace for gr in groups]).cpu().long(); nn = torch.cat([gr.batch.n for gr in groups]).cpu().long()
        pid = torch.cat([gr.batch.pool_id for gr in groups]).cpu().long()
        S = min(B, max(16, min(cores, 64) * 3))
        pick = [int(i) for i in torch.linspace(0, B - 1, S).round().long().tolist()]
        # state one step back: the last pick is what the oracle must reproduce
        jobs = []
        for b in pick:
            nb = int(nn[b]) - 1
            if nb < N_INITIAL:
                continue
            jobs.append((PX[pid[b]].numpy(), PY[pid[b]].numpy(), tr[b, :nb].tolist(), int(tr[b, nb])))
        t0 = time.perf_counter()
        res = cpu_pool.map(_cpu_one_iteration, [j[:3] for j in jobs])
        wall = time.perf_counter() - t0
        cpu_pool.close()
        agree = sum(1 for r, j in zip(res, jobs) if r[0] == j[3])
        parity = {"oracle_index_agreement": f"{agree}/{len(jobs)}"}
        ncores = min(cores, 64)
        cpu = {"value": len(jobs) / wall, "unit": UNIT, "cores": ncores, "kind": "port",
               "per_core": len(jobs) / sum(r[1] for r in res),
               "sample": f"{len(jobs)} BO iterations (campaign states taken from this run, n = "
                         f"{min(len(j[2]) for j in jobs)}..{max(len(j[2]) for j in jobs)}), oracle/ on {ncores} "
                         f"processes x 1 thread; wall {wall:.1f}s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(B=B), "campaigns_per_gpu": B, "pool": N_POOL, "d": DIM,
                       "l2": "inputs are L2-resident by design (0.8 MB of pools); the kernel is FP64/latency bound, "
                             "no HBM stream to flush", "threads_per_campaign": os.environ.get("STP_THREADS", "auto"),
                             "groups": NG, "schedule": args.schedule},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA mma.sync.m8n8k4.f64 + DFMA; same peak on B200)",
                         "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp64_peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH * (B / NG) / 1024.0,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of k_exact_bo_step from one ncu --set "
                                           "full capture of a 1024-campaign launch (profiles/r01_ncu_prof_r1_bostep_v7.txt), scaled "
                                           f"to the {B // NG} campaigns of one launch; algorithmic bytes per launch: "
                                           f"{hbm_bytes / max(gpu_launches, 1):.3g}",
                         "kernel": "k_exact_bo_step",
                         "peak_source": "max of stp_fp64_probe (DFMA) and stp_fp64_mma_probe (DMMA) microbenchmarks, this "
                                        "run; MEASURED_PEAKS.json has no FP64 figure",
                         "peak_dfma": fp64_fma_peak, "peak_dmma": fp64_mma_peak,
                         "hbm": {"achieved": hbm_gbs, "peak": hbm_peak,
                                 "unit": "GB/s"}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke},
            "gpu_launches": int(launches_all.item()),
            "clocks": clocks,
            "parity": parity,
            "frozen_campaigns": status_bad,
            "mean_closures_per_fit": mean_closures,
            "streams": NG,
            "setup_s": setup_s,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_var_arm(args):
    """Secondary workload: B concurrent VarGP / VarTGP / VarLGP campaigns (pool 2048, d = 6, ages staggered over
    n = 10..89, 600 epochs, S = 2048 Monte-Carlo samples, in-kernel Philox).  Training sets of age a are the Sobol
    initial design + a random distinct pool points (no BO pre-roll: a pre-roll alone would take minutes); every
    step is a real BO iteration (fresh model -> stp_var_fit -> stp_var_acq with the update)."""
    import torch
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from stpbenchmark_b200 import _lib, synthetic
    from stpbenchmark_b200.engine import CampaignBatch
    B, W, K = args.campaigns, args.warmup, args.steps
    per_pool = max(B // N_POOLS, 1)
    pools = [synthetic.cfg4_pool(g) for g in range(N_POOLS)]
    PX = torch.stack([p[0] for p in pools]); PY = torch.stack([p[1] for p in pools])
    pool_id = [min(b // per_pool, N_POOLS - 1) for b in range(B)]
    gen = torch.Generator().manual_seed(4242 + rank)
    init = []
    for b in range(B):
        g = pool_id[b]
        d0 = synthetic.initial_designs(pools[g][0], pools[g][1], [100000 * rank + synthetic.campaign_seed(g, b % per_pool)], N_INITIAL)[0].tolist()
        age = (b * N_TRIALS // B + b) % (N_TRIALS - W - K) if N_TRIALS - W - K > 0 else 0
        rest = [i for i in torch.randperm(N_POOL, generator=gen).tolist() if i not in set(d0)][:age]
        init.append(d0 + rest)
    from stpbenchmark_b200.engine import CampaignGroups
    batch = CampaignGroups(PX, PY, pool_id, init, n_cap=CAP, kind=args.model, groups=args.groups, device=dev,
                           epochs=args.epochs, learning_rate=0.01, num_samples=2048)
    for m in batch.members:
        m.ragged = True
    cat_n = lambda: torch.cat([m.n for m in batch.members])
    cat_status = lambda: torch.cat([m.status for m in batch.members])
    n0 = cat_n().clone()
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(W):
        batch.step()
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = batch.launches
    n_before = cat_n().clone()
    ok_before = int((cat_status() == 0).sum().item())
    ev0 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for st in batch.streams:
        st.wait_event(ev0)
    for _ in range(K):
        batch.step()
    ends = []
    for st in batch.streams:
        e = torch.cuda.Event(enable_timing=True)
        e.record(st)
        ends.append(e)
    barrier()
    clocks = sampler.stop()
    done = torch.tensor([int((cat_n() - n_before).sum().item())], dtype=torch.int64, device=dev)   # completed iterations
    if world > 1:
        dist.all_reduce(done)
    ms = max(ev0.elapsed_time(e) for e in ends)
    t = torch.tensor([ms], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    bad = int((cat_status() != 0).sum().item())
    E, S, d = args.epochs, 2048, DIM
    nn = n_before.double() + (K - 1) / 2.0
    C = 60.0 if args.model == "VarGP" else S * 80.0
    flops = float((E * (9 * nn ** 3 + nn ** 2 * (14 * d + 12) + 800 * nn) + 3 * nn ** 3 +
                   (N_POOL - nn) * (nn * (3 * d + 4) + 2 * nn ** 2 + C)).sum()) * K
    peak = max(_lib.fp64_probe(4096), _lib.fp64_mma_probe(4096))
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": float(done.item()) / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{B} concurrent {args.model} campaigns per GPU, synthetic pools N=2048 d=6, n staggered over "
                                   f"{int(n0.min())}..{int(n0.max())} (+{W + K} steps), {E} epochs, S=2048 (in-kernel Philox)",
                       "campaigns_per_gpu": B},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA + DFMA)", "achieved": flops / (ms * 1e-3) / 1e12,
                         "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / peak, "traffic": None, "kernel": "k_var_fit + k_var_acq"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": batch.launches - l0, "clocks": clocks,
            "frozen_campaigns": bad, "streams": len(batch.streams)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--campaigns", type=int, default=1024, help="campaigns per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--groups", type=int, default=16, help="campaign groups per GPU, one CUDA stream each (at most B / 64)")
    ap.add_argument("--schedule", default="cohort", choices=["mixed", "cohort", "queue"],
                    help="mixed: every group holds campaigns of all ages; cohort: the campaigns of a group share their "
                         "age (groups staggered over the 80 trials), so each launch is sized for its own n")
    ap.add_argument("--model", default="ExactGP", choices=["ExactGP", "VarGP", "VarTGP", "VarLGP"],
                    help="ExactGP = the headline workload; the others run the secondary variational workload")
    ap.add_argument("--epochs", type=int, default=600)
    ap.add_argument("--trials", type=int, default=N_TRIALS,
                    help="trials per campaign (diagnostic: 80 is the workload; smaller keeps n = 10..9+trials)")
    args = ap.parse_args()
    if args.trials != N_TRIALS:
        globals().update(N_TRIALS=args.trials, CAP=N_INITIAL + args.trials + 1)
    if args.impl == "reference":
        args.steps = max(1, min(args.steps, N_TRIALS - args.warmup))  # one campaign per core has 80 trials
        run_reference_arm(args)
    elif args.model != "ExactGP":
        args.steps = min(args.steps, 10)
        run_var_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
