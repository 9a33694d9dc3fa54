#!/usr/bin/env python
"""bench.py -- BO iterations/s (fit + predict + acquire + update) over batched campaigns.

  python bench.py [--gpus N] [--steps K] [--warmup W]                    # headline: BASELINE configs[3] ("cfg4")
  python bench.py --model VarTGP                                         # north-star VarTGP shape (pool 2048, d = 6)
  python bench.py --config cfg5 | cfg2 | cfg3                            # the other BASELINE configs
  python bench.py --config f3                                            # benchmark_predictions.py study
  python bench.py --impl reference [--config ...] [--steps K] [--warmup W]   # CPU arm: the oracle on the host cores
  (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Workloads:
  cfg4  synthetic pools N = 2048, d = 6 (negated Hartmann-6 + 0.05 t_3 noise, 8 pools), 1024 concurrent ExactGP
        campaigns PER GPU, n_initial = 10, 80 trials.  Device-resident arm: campaigns at staggered ages (continuous
        batching), 24 age cohorts on 24 CUDA streams, one stp_exact_bo_step launch per cohort and step.  e2e arm: the
        SAME 1024 x 80 job through the product API (runners.run_sweep: host tensors in, CSV traces out).
  VarTGP the same pools, 1024 concurrent VarTGP campaigns, n = 10..89 mix, 600 epochs, S = 2048 in-kernel Philox.
  cfg5  pools N = 8192, d = 8, 512 VarTGP campaigns per GPU (4096 over 8 GPUs), n_initial = M = 256, 4 trials.
  cfg2  VarTGP x the 5 materials datasets x the 50 seeds of random_seeds.txt through run_benchmark's path (one device job).
  cfg3  Hartmann-6 fit-and-predict (test_hartmann.py scenario), 256 seeds x {ExactGP, VarGP, VarLGP, VarTGP}.
  f3    benchmark_predictions.py's study: {VarTGP, VarGP, ExactGP} x 5 datasets x 25 seeds, fit on 80 % (n up to 480), MAE
        on the rest; device-resident per (model, dataset) pair + end to end through evaluate_model_across_datasets.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.  The CPU legs (cpu_baseline of the CUDA
arms, the whole --impl reference arm) run oracle/ in a SEPARATE interpreter started with OMP/MKL/OPENBLAS_NUM_THREADS=1
that forks one single-thread worker per core before it imports torch: the same code path for both, so their
per-core rates agree (round 1's in-run leg forked from a process that had already spun up torch's thread pools and
was 6 - 40x slower than the reference arm on the same box).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POOL, DIM, N_POOLS = 2048, 6, 8
N_INITIAL, N_TRIALS = 10, 80
CAP = N_INITIAL + N_TRIALS + 1
DEFAULT_GROUPS = 24   # age cohorts (CUDA streams) of the device-resident arm: 16 / 24 / 32 -> 71.7 / 73.7 / 73.5 k it/s at
                      # --steps 20 --warmup 5 (profiles/r02_ab_groups.log); 40 streams and more serialise (52 - 58 k)
METRIC, UNIT = "bo_iterations_per_sec", "it/s"
WORKLOAD = ("cfg4: synthetic pools N=2048 d=6 (Hartmann-6 + t3 noise), {B} concurrent ExactGP campaigns per GPU, "
            "n_initial=10, 80 trials, ages staggered over n=10..89 (continuous batching)")
# dram__bytes_read.sum + dram__bytes_write.sum of ONE cohort launch (64 campaigns) of k_exact_bo_step, ncu --set full
# (profiles/r02_ncu_cohort_n40.txt, _n60.txt, _n85.txt): keyed by the launch's block size (128 / 160 / 256 threads <-> n_max <= 48 / 66 / 89)
NCU_DRAM_BYTES_COHORT = {128: 1.238016e6 + 256, 160: 1.313024e6 + 5888, 256: 1.40032e6 + 8704}
NCU_DRAM_BYTES_VARFIT = 46.688512e6 + 1.44043e9     # k_var_fit, B = 296, n = 60, 20 epochs (profiles/r02_ncu_final_varfit_n60.txt)


# ------------------------------------------------------------------------------------------
# CPU legs: oracle/ in a separate single-thread-per-core interpreter
# ------------------------------------------------------------------------------------------
def _cpu_init():
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)
    # untimed warm-up of every code path a job can take (imports, scipy's L-BFGS-B, autograd's first backward: ~4 s cold)
    from oracle import loop as oloop, variational as ov
    g = torch.Generator().manual_seed(0)
    X = torch.rand(40, 3, dtype=torch.double, generator=g); y = torch.rand(40, dtype=torch.double, generator=g)
    oloop.bo_iteration("ExactGP", X[:12], y[:12], X[12:])
    st = ov.train_natural(X[:12], y[:12], "VarTGP", 2, 0.01, init_noise=torch.zeros(12, dtype=torch.double))
    ov.acquisition(st, X[12:], y[:12].max(), eps=torch.zeros(28, 8, dtype=torch.double))


def _pool_of(name, g, raw_y=False):
    import torch
    from stpbenchmark_b200 import synthetic
    if name == "cfg4":
        return synthetic.cfg4_pool(g)
    if name == "cfg5":
        return synthetic.cfg5_pool(g)
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))       # the reference's data/*.csv, preprocessed
    X = torch.tensor(z[f"{name}_X"], dtype=torch.double); y = torch.tensor(z[f"{name}_y"], dtype=torch.double)
    return X, (1.0 / y if name in ("AgNP", "Perovskite") and not raw_y else y)            # run_benchmark.py:27-28


def _aged_training_set(X, y, seed, n_initial, age):
    """Sobol initial design + `age` random distinct pool points (what bench uses instead of a BO pre-roll)."""
    import torch
    from stpbenchmark_b200 import utils
    init = utils.initial_designs_batch(X, y, [seed], n_initial, 50.0)[0].tolist()
    gen = torch.Generator().manual_seed(seed)
    extra = [i for i in torch.randperm(len(X), generator=gen).tolist() if i not in set(init)][:age]
    return init + extra


def _cpu_job(job):
    """One unit of oracle work (runs in a forked single-thread worker)."""
    import torch
    from oracle import loop as oloop
    t = job["type"]
    if t == "exact_campaign":       # `iters` consecutive BO iterations of one ExactGP campaign started at `age`
        X, y = _pool_of(job["pool"], job["g"])
        picked = _aged_training_set(X, y, job["seed"], job["n_initial"], job["age"])
        mask = torch.ones(len(X), dtype=torch.bool); mask[picked] = False
        times = []
        for _ in range(job["iters"]):
            t0 = time.perf_counter()
            best, _acq = oloop.bo_iteration("ExactGP", X[picked], y[picked], X[mask])
            idx = int(torch.where(mask)[0][best])
            picked.append(idx); mask[idx] = False
            times.append(time.perf_counter() - t0)
        return times
    if t == "exact_state":          # one BO iteration from a given campaign state -> (pick, seconds, nfev)
        X, y = _pool_of(job["pool"], job["g"])
        picked = job["picked"]
        mask = torch.ones(len(X), dtype=torch.bool); mask[picked] = False
        info = {}
        t0 = time.perf_counter()
        best, _ = oloop.bo_iteration("ExactGP", X[picked], y[picked], X[mask], info=info)
        return [int(torch.where(mask)[0][best]), time.perf_counter() - t0, int(info["nfev"])]
    if t == "var_iteration":        # one variational BO iteration, optionally with fewer epochs (extrapolated by the caller)
        from oracle import variational as ov
        X, y = _pool_of(job["pool"], job["g"])
        picked = job.get("picked") or _aged_training_set(X, y, job["seed"], job["n_initial"], job["age"])
        mask = torch.ones(len(X), dtype=torch.bool); mask[picked] = False
        cand = torch.where(mask)[0]
        if job.get("acq_candidates"):
            cand = cand[: job["acq_candidates"]]
        gen = torch.Generator().manual_seed(job.get("seed", 0) + 17)
        t0 = time.perf_counter()
        st = ov.train_natural(X[picked], y[picked], job["kind"], job["epochs"], 0.01,
                              init_noise=torch.randn(len(picked), dtype=torch.double, generator=gen))
        t1 = time.perf_counter()
        eps = torch.randn(len(cand), job["S"], dtype=torch.double, generator=gen) if job["kind"] != "VarGP" else None
        ov.acquisition(st, X[cand], y[picked].max(), eps=eps)
        t2 = time.perf_counter()
        return [t1 - t0, t2 - t1, len(picked), len(cand), int(mask.sum())]
    if t == "hartmann_fit":         # test_hartmann.py scenario: fit on 100 random points, MAE on 100 more
        from oracle import exact as oexact, variational as ov
        from stpbenchmark_b200 import synthetic, utils
        utils.set_seeds(job["seed"])
        Xtr = torch.rand(100, 6).double(); ytr = synthetic.hartmann6(Xtr); Xte = torch.rand(100, 6).double()
        yte = synthetic.hartmann6(Xte)
        t0 = time.perf_counter()
        if job["kind"] == "ExactGP":
            theta, _res = oexact.fit_exact(Xtr, ytr)
            mean, _ = oexact.posterior_exact(theta, Xtr, ytr, Xte)
        else:
            st = ov.train_natural(Xtr, ytr, job["kind"], job["epochs"], 0.01)
            m, _ = ov.latent_posterior(st, Xte)
            mean = m * (ytr.std(unbiased=False) + 1e-10) + ytr.mean()
        return [time.perf_counter() - t0, float((mean - yte).abs().mean())]
    if t == "predict_fit":          # benchmark_predictions.py:19-75 on one 80/20 split; `epochs` of 600 timed (variational)
        from oracle import exact as oexact, variational as ov
        from stpbenchmark_b200 import benchmark_predictions as bp
        X, y = _pool_of(job["pool"], 0, raw_y=True)
        tr, te = bp.split_indices(len(X), job["seed"])
        mu, sd = y[tr].mean(), y[tr].std(unbiased=False) + 1e-10
        ys = (y[tr] - mu) / sd
        if job["kind"] != "ExactGP":
            ov.train_natural(X[tr], ys, job["kind"], 1, 0.01)     # untimed: first-call costs must not be extrapolated
        t0 = time.perf_counter()
        if job["kind"] == "ExactGP":
            theta, _res = oexact.fit_exact(X[tr], ys)
            mean, _ = oexact.posterior_exact(theta, X[tr], ys, X[te])
            t1 = t2 = time.perf_counter()
        else:
            st = ov.train_natural(X[tr], ys, job["kind"], job["epochs"], 0.01)
            t1 = time.perf_counter()
            mean, _ = ov.latent_posterior(st, X[te])
            t2 = time.perf_counter()
        return [t1 - t0, t2 - t1, len(tr), float(((mean * sd + mu) - y[te]).abs().mean())]
    raise ValueError(t)


def _cpu_ready(_):
    time.sleep(0.05)
    return os.getpid()


def _cpu_worker_main():
    """`python bench.py --cpu-worker`: reads {"cores", "jobs"} from stdin, forks the pool BEFORE torch is imported."""
    spec = json.loads(sys.stdin.read())
    with mp.get_context("fork").Pool(spec["cores"], initializer=_cpu_init) as pool:
        pool.map(_cpu_ready, range(4 * spec["cores"]), chunksize=1)      # every worker has finished its warm-up
        t0 = time.perf_counter()
        res = pool.map(_cpu_job, spec["jobs"], chunksize=1)
        wall = time.perf_counter() - t0
    print(json.dumps({"results": res, "wall": wall}))


def run_cpu_jobs(jobs, cores):
    """Results of `jobs` from a fresh interpreter with one single-thread oracle worker per core."""
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker"], input=json.dumps({"cores": cores, "jobs": jobs}),
                         capture_output=True, text=True, env=env, cwd=ROOT)
    if out.returncode != 0:
        raise RuntimeError("cpu worker failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------
# --impl reference: the reference path's CPU restatement on every host core
# ------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    W, K = args.warmup, args.steps
    cfg = args.config if args.model == "ExactGP" or args.config != "cfg4" else "vartgp"
    note = ""
    if cfg == "f3":
        import numpy as np
        names = ["AgNP", "AutoAM", "CrossedBarrel", "P3HT", "Perovskite"]
        S_ = 25 if args.campaigns == 1024 else args.campaigns
        seeds = [int(s) for s in np.load(os.path.join(ROOT, "tests", "golden", "traces.npz"))["seeds"]][:S_]
        fits = 3 * len(names) * S_
        cpu = f3_cpu_leg(["VarTGP", "VarGP", "ExactGP"], names, seeds, S_, args.epochs, fits, cores)
        print(json.dumps({
            "impl": "reference", "metric": "model_fits_per_sec", "value": cpu["value"], "unit": "fits/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": 1e3 * cpu["study_core_seconds"] / cores, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "the reference's data/*.csv",
            "config": {"workload": F3_WORKLOAD.format(S=S_, E=args.epochs), "sample": cpu["sample"]}, "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return
    if cfg == "cfg4":
        # a step = iteration t of `cores` campaigns (one per core, bounded sample of the 1024-campaign workload); the
        # campaigns start at ages spread over the 80 trials so that n covers 10..89 like the GPU mix; W + K = 80 is
        # one full campaign per core
        K = max(1, min(K, N_TRIALS - W))
        span = max(N_TRIALS - (W + K), 0)
        ages = [(c * span) // max(cores - 1, 1) if cores > 1 else 0 for c in range(cores)]
        jobs = [dict(type="exact_campaign", pool="cfg4", g=c % N_POOLS, seed=900000 + c, n_initial=N_INITIAL, iters=W + K,
                     age=ages[c]) for c in range(cores)]
        r = run_cpu_jobs(jobs, cores)
        per_core_s = [sum(p[W:]) for p in r["results"]]            # every core's own timed seconds
        total = max(per_core_s)                                   # the job ends with its slowest core
        value = cores * K / total
        per_core = cores * K / sum(per_core_s)
        workload = WORKLOAD.format(B=args.campaigns)
        sample = (f"{cores} campaigns (one per host core) started at ages 0..{max(ages)} (Sobol design + age random pool "
                  f"points), {W} warm-up + {K} timed consecutive BO iterations each (n = {N_INITIAL}.."
                  f"{N_INITIAL + max(ages) + W + K - 1}); value = cores x K / slowest core's timed seconds")
    else:
        kind = "VarTGP"
        K = 1
        if cfg == "vartgp":
            jobs = [dict(type="var_iteration", pool="cfg4", g=c % N_POOLS, seed=900000 + c, n_initial=N_INITIAL,
                         age=(c * 79) // max(cores - 1, 1), kind=kind, epochs=args.epochs, S=2048) for c in range(cores)]
            workload = var_workload(args.campaigns, kind, args.epochs)
            scale = lambda res: res[0] + res[1]
        elif cfg == "cfg5":
            E_s = 6                                                # 6 of 600 epochs timed, 512 of ~7936 candidates: extrapolated
            jobs = [dict(type="var_iteration", pool="cfg5", g=c % N_POOLS, seed=900000 + c, n_initial=256, age=0, kind=kind,
                         epochs=E_s, S=2048, acq_candidates=512) for c in range(cores)]
            workload = CFG5_WORKLOAD.format(B=args.campaigns if args.campaigns != 1024 else 512)
            scale = lambda res: res[0] * (600.0 / E_s) + res[1] * (res[4] / res[3])
            note = f"; EXTRAPOLATED from {E_s} of 600 epochs and 512 of {8192 - 256} candidates per iteration"
        elif cfg == "cfg2":
            names = ["AgNP", "AutoAM", "CrossedBarrel", "P3HT", "Perovskite"]
            jobs = [dict(type="var_iteration", pool=names[c % 5], g=0, seed=45 + c, n_initial=10, age=(c * 79) // max(cores - 1, 1),
                         kind=kind, epochs=600, S=2048) for c in range(cores)]
            workload = CFG2_WORKLOAD
            scale = lambda res: res[0] + res[1]
        else:
            raise SystemExit(f"--impl reference is not provided for --config {cfg}")
        r = run_cpu_jobs(jobs, cores)
        per_core_s = [scale(res) for res in r["results"]]
        total = max(per_core_s)
        value = cores * K / total
        per_core = cores / sum(per_core_s)
        sample = (f"{cores} {kind} BO iterations, one per host core, training sets of n = "
                  f"{min(res[2] for res in r['results'])}..{max(res[2] for res in r['results'])}{note}")
        W = 0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic" if cfg != "cfg2" else "the reference's data/*.csv",
        "config": {"workload": workload, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "per_core": per_core,
                         "sample": sample + f"; oracle/ (torch fp64 + scipy L-BFGS-B), 1 thread per campaign; wall {r['wall']:.1f}s"},
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


class Dist:
    """torch.distributed plumbing of one bench process (one process per GPU under torchrun)."""

    def __init__(self):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py (CUDA arm) needs a GPU; use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, v, op="sum"):
        import torch
        t = torch.tensor([v], dtype=torch.double, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.item()

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def fp64_peaks():
    from stpbenchmark_b200 import _lib
    fma, mma = _lib.fp64_probe(4096), _lib.fp64_mma_probe(4096)
    return {"peak": max(fma, mma), "peak_dfma": fma, "peak_dmma": mma,
            "peak_source": "max of stp_fp64_probe (DFMA) and stp_fp64_mma_probe (DMMA) microbenchmarks, this run; "
                           "MEASURED_PEAKS.json has no FP64 figure"}


def measured_hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
    except Exception:
        return 6650.0


# ------------------------------------------------------------------------------------------
# cfg4, ExactGP: the headline workload
# ------------------------------------------------------------------------------------------
def exact_e2e_job(D, B, PXh, PYh, pool_id, seed_of, tmpdir):
    """The cfg4 job through the product API: host tensors -> runners.run_sweep (initial designs, H2D, every BO
    iteration of every campaign, D2H of the traces, CSV files) -> iterations completed, seconds, bytes moved.
    With several ranks the job holds `world` x B campaigns; run_sweep shards it (each rank its own B) and gathers."""
    from stpbenchmark_b200 import models, runners
    jobs = []
    for r in range(D.world):
        for g in range(N_POOLS):
            slots = [b for b in range(B) if pool_id[b] == g]
            if slots:
                jobs.append(dict(X=PXh[g], y=PYh[g], model_class=models.ExactGP,
                                 seeds=[seed_of(b, 0, r) + 7 for b in slots], n_initial=N_INITIAL, n_trials=N_TRIALS,
                                 output_path=os.path.join(tmpdir, f"ExactGP_rank{r}_pool{g}_traces_.csv")))
    D.barrier()
    t0 = time.perf_counter()
    frames = runners.run_sweep(jobs)
    D.barrier()
    dt = time.perf_counter() - t0
    iters = 0
    if D.rank == 0:
        for f in frames:
            x = f.xs("x_index", axis=1, level=1)
            iters += int(x.iloc[N_INITIAL:].notna().sum().sum())
    per_rank = len(jobs) // D.world
    h2d = per_rank * (N_POOL * DIM * 8 + N_POOL * 8) + B * (N_INITIAL * 4 + N_POOL)      # pools, designs, masks
    d2h = B * (CAP * 4 + 4 + 4)                                                               # traces, n, status
    return iters, dt, h2d, d2h


def run_exact_arm(args):
    import torch
    D = Dist()
    dev, rank, world = D.dev, D.rank, D.world
    from stpbenchmark_b200 import synthetic, utils
    from stpbenchmark_b200.engine import CampaignBatch

    B, W, K = args.campaigns, args.warmup, args.steps
    per_pool = B // N_POOLS
    pools = [synthetic.cfg4_pool(g) for g in range(N_POOLS)]
    PX = torch.stack([p[0] for p in pools]); PY = torch.stack([p[1] for p in pools])
    pool_id = [min(b // per_pool, N_POOLS - 1) if per_pool else 0 for b in range(B)]
    NG = max(1, min(args.groups, B // 16 if B >= 16 else 1))
    queue = args.schedule == "queue"           # persistent work-queue kernel: K iterations per campaign in ONE launch
    if queue and args.groups == DEFAULT_GROUPS:
        NG = 1
    bounds = [(g * B // NG, (g + 1) * B // NG) for g in range(NG)]
    if args.schedule == "cohort":                             # one age per group, groups evenly staggered
        group_of = [next(g for g, (lo, hi) in enumerate(bounds) if lo <= b < hi) for b in range(B)]
        age0 = [(group_of[b] * N_TRIALS) // NG for b in range(B)]
    else:
        age0 = [b % N_TRIALS for b in range(B)]             # trials already done when timing starts
    # seeds: first generation as config #4 (1000 g + k), later generations continue per slot
    gens = 1 + (W + K + N_TRIALS) // N_TRIALS + 1
    seed_of = lambda b, gen, r=rank: 100000 * r + synthetic.campaign_seed(pool_id[b], b % max(per_pool, 1)) + 10007 * gen
    t_setup = time.perf_counter()
    designs = {}
    for g in range(N_POOLS):
        slots = [b for b in range(B) if pool_id[b] == g]
        seeds = [seed_of(b, gen) for gen in range(gens) for b in slots]
        if seeds:
            Dg = utils.initial_designs_batch(pools[g][0], pools[g][1], seeds, N_INITIAL, 50.0, device=dev)
            for i, (gen, b) in enumerate([(gen, b) for gen in range(gens) for b in slots]):
                designs[(b, gen)] = Dg[i]
    designs_s = time.perf_counter() - t_setup

    class Group:
        def __init__(self, lo, hi):
            self.lo, self.hi, self.nb = lo, hi, hi - lo
            self.stream = torch.cuda.Stream(device=dev)
            self.batch = CampaignBatch(PX, PY, pool_id[lo:hi], [designs[(b, 0)].tolist() for b in range(lo, hi)],
                                       n_cap=CAP, kind="ExactGP", device=dev)
            self.batch.ragged = args.schedule == "mixed"     # ages differ: longest fits are scheduled first
            # the next initial designs of every slot: the kernel recycles finished campaigns itself (stp_turnover)
            self.bank = torch.stack([torch.stack([designs[(b, g)] for g in range(1, gens)]) for b in range(lo, hi)])

        def preroll(self):
            # untimed: bring campaign b to age0[b] (persistent launches: campaign b does age0[b] iterations)
            bt = self.batch
            ages = [age0[b] for b in range(self.lo, self.hi)]
            if max(ages) > 0:
                bt.run_queue(ages[0] if args.schedule == "cohort" else ages)
            bt.enable_turnover(self.bank, n_final=N_INITIAL + N_TRIALS, sink_capacity=self.nb * 4)

    groups = [Group(lo, hi) for lo, hi in bounds]
    torch.cuda.synchronize()                                 # construction ran on the default stream
    for gr in groups:
        with torch.cuda.stream(gr.stream):
            gr.preroll()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    def all_groups(fn):
        for gr in groups:
            with torch.cuda.stream(gr.stream):
                fn(gr)

    chunk = max(1, int(args.chunk)) if args.schedule == "cohort" else 1

    def run_steps(k):
        if queue:
            all_groups(lambda gr: gr.batch.run_queue(k))
        elif chunk > 1:      # cohorts advanced by persistent work-queue launches of `chunk` iterations (the product path's form)
            for c in range(0, k, chunk):
                all_groups(lambda gr: gr.batch.run_queue(min(chunk, k - c)))
        else:
            for _ in range(k):
                all_groups(lambda gr: gr.batch.step())

    def totals():
        return (float(sum(gr.batch.work_acc.sum().item() for gr in groups)),
                int(sum(gr.batch.iters_acc.sum().item() for gr in groups)))

    # ---- device-resident arm: W warm-up + K timed steps ------------------------------------
    if chunk > 1:
        W = -(-W // chunk) * chunk      # whole chunks: a cohort then never crosses the end of its campaigns inside a launch
    run_steps(W)
    D.barrier()
    flops0, iters0 = totals()
    sampler = ClockSampler(D.local_rank); sampler.start()
    launches0 = sum(gr.batch.launches for gr in groups)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev0.record()                                             # default stream, right after the synchronize
    for gr in groups:
        gr.stream.wait_event(ev0)
    run_steps(K)
    ends = []
    for gr in groups:
        e = torch.cuda.Event(enable_timing=True)
        e.record(gr.stream)
        ends.append(e)
    D.barrier()
    clocks = sampler.stop()
    gpu_launches = sum(gr.batch.launches for gr in groups) - launches0
    elapsed_ms = max(ev0.elapsed_time(e) for e in ends)
    elapsed_ms_max = D.reduce(elapsed_ms, "max")
    flops1, iters1 = totals()
    status_bad = int(sum(int(((gr.batch.status != 0) & (gr.batch.status != 6)).sum().item()) for gr in groups))
    done = D.reduce(iters1 - iters0)                          # completed iterations, counted by the kernel
    value = done / (elapsed_ms_max * 1e-3)
    launches_all = int(D.reduce(gpu_launches))

    # roofline of the dominant (only) kernel: algorithmic flops (accumulated by the kernel itself from every
    # campaign's n and actual closure count) / the time its launches ran in
    achieved_tflops = (flops1 - flops0) / (elapsed_ms * 1e-3) / 1e12
    peaks = fp64_peaks()
    n_bar = N_INITIAL + (N_TRIALS - 1) / 2.0
    bytes_per_iter = (N_POOL - n_bar) * DIM * 8 + N_POOL + n_bar * (DIM + 1) * 8 + 16 + 24
    hbm_bytes = bytes_per_iter * (iters1 - iters0)
    hbm_gbs = hbm_bytes / (elapsed_ms * 1e-3) / 1e9
    mean_closures = float(torch.cat([gr.batch.nfev for gr in groups]).double().mean().item())
    traffic = None
    if all(v is not None for v in NCU_DRAM_BYTES_COHORT.values()) and args.schedule == "cohort" and chunk == 1:
        # launches of the three block sizes in proportion to the ages a cohort spends at n_max <= 48 / 66 / 89
        w = {128: 38 / 80.0, 160: 18 / 80.0, 256: 24 / 80.0}
        traffic = sum(w[t] * NCU_DRAM_BYTES_COHORT[t] for t in w) * (B / NG) / 64.0

    # state one step back of a sample of campaigns (for the CPU leg: the last pick is what the oracle must reproduce)
    tr = torch.cat([gr.batch.trace for gr in groups]).cpu().long(); nn = torch.cat([gr.batch.n for gr in groups]).cpu().long()
    pid = torch.cat([gr.batch.pool_id for gr in groups]).cpu().long()
    del groups
    torch.cuda.empty_cache()

    # ---- end-to-end arm: the same 1024 x 80 job through the product API, host buffers in, CSV traces out ---------
    PXh = [p[0] for p in pools]; PYh = [p[1] for p in pools]
    e2e = None
    with tempfile.TemporaryDirectory() as tmp:
        if not args.no_e2e:
            exact_e2e_job(D, min(B, 64), PXh, PYh, pool_id, seed_of, os.path.join(tmp, "warm"))   # library / pandas warm-up
            it, dt, h2d, d2h = exact_e2e_job(D, B, PXh, PYh, pool_id, seed_of, os.path.join(tmp, "job"))
            it_all = D.reduce(it)
            e2e = {"value": it_all / dt, "unit": UNIT, "h2d_bytes_per_step": h2d / N_TRIALS, "d2h_bytes_per_step": d2h / N_TRIALS,
                   "steps": N_TRIALS, "wall_s": dt,
                   "what": f"runners.run_sweep on HOST tensors: {world} x {B} campaigns x {N_TRIALS} trials from the Sobol initial "
                           "designs (host draw + stp_design_snap) to the CSV files, one call; bytes are per GPU and per BO step"}

    # ---- CPU baseline (rank 0): oracle BO iterations from the GPU's own campaign states ----------
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = min(host_cores(), 64)
        S = min(B, max(16, cores * 3))
        pick = [int(i) for i in torch.linspace(0, B - 1, S).round().long().tolist()]
        jobs, want = [], []
        for b in pick:
            nb = int(nn[b]) - 1
            if nb < N_INITIAL:
                continue
            jobs.append(dict(type="exact_state", pool="cfg4", g=int(pid[b]), picked=tr[b, :nb].tolist()))
            want.append(int(tr[b, nb]))
        r = run_cpu_jobs(jobs, cores)
        res = r["results"]
        agree = sum(1 for x, w_ in zip(res, want) if x[0] == w_)
        parity = {"oracle_index_agreement": f"{agree}/{len(jobs)}"}
        cpu = {"value": len(jobs) / r["wall"], "unit": UNIT, "cores": cores, "kind": "port",
               "per_core": len(jobs) / sum(x[1] for x in res),
               "sample": f"{len(jobs)} BO iterations (campaign states taken from this run, n = "
                         f"{min(len(j['picked']) for j in jobs)}..{max(len(j['picked']) for j in jobs)}), oracle/ on {cores} "
                         f"single-thread processes of a separate interpreter; wall {r['wall']:.1f}s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(B=B), "campaigns_per_gpu": B, "pool": N_POOL, "d": DIM,
                       "l2": "inputs are L2-resident by design (0.8 MB of pools); the kernel is FP64/latency bound, "
                             "no HBM stream to flush", "threads_per_campaign": os.environ.get("STP_THREADS", "auto"),
                       "groups": NG, "schedule": args.schedule, "chunk": chunk},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA mma.sync.m8n8k4.f64 + DFMA; same peak on B200)",
                         "achieved": achieved_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / peaks["peak"],
                         "traffic": traffic,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of single cohort launches of "
                                           "k_exact_bo_step (64 campaigns, block sizes 128 / 160 / 256), ncu --set full, "
                                           "profiles/r02_ncu_cohort_n40/60/85.txt, weighted by the share of the 80 ages each "
                                           f"block size serves; algorithmic bytes per launch: {hbm_bytes / max(gpu_launches, 1):.3g}",
                         "kernel": "k_exact_bo_step", **peaks,
                         "hbm": {"achieved": hbm_gbs, "peak": measured_hbm_peak(), "unit": "GB/s"}},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches_all,
            "clocks": clocks,
            "parity": parity,
            "frozen_campaigns": status_bad,
            "mean_closures_per_fit": mean_closures,
            "streams": NG,
            "setup_s": setup_s,
            "designs_s": designs_s,
        }
        print(json.dumps(line), flush=True)
    D.close()


# ------------------------------------------------------------------------------------------
# variational arms: north-star VarTGP shape (cfg4 pools) and cfg5
# ------------------------------------------------------------------------------------------
def var_workload(B, kind, E):
    return (f"{B} concurrent {kind} campaigns per GPU, synthetic pools N=2048 d=6 (cfg4 pools), training sets n = 10..89 "
            f"(Sobol design + random pool points), {E} epochs NGD+Adam, S=2048 (in-kernel Philox)")


CFG5_WORKLOAD = ("cfg5: synthetic pools N=8192 d=8 (4 Gaussian bumps + t3 noise), {B} VarTGP campaigns per GPU (4096 over 8 GPUs), "
                 "n_initial = M = 256 inducing points, 4 trials, 600 epochs, S=2048 (in-kernel Philox)")
CFG2_WORKLOAD = ("cfg2: VarTGP x 5 materials datasets (N = 94..600, d = 3..5) x the 50 seeds of random_seeds.txt, n_initial=10, "
                 "80 trials, 600 epochs, S=2048: 250 campaigns = 20 000 BO iterations as ONE device job (runners.run_sweep)")


def var_flops(n, d, E, N, C):
    """W = E (9 n^3 + n^2 (14 d + 12) + 800 n) + 3 n^3 + Nc (n (3 d + 4) + 2 n^2 + C)."""
    return E * (9 * n ** 3 + n ** 2 * (14 * d + 12) + 800 * n) + 3 * n ** 3 + (N - n) * (n * (3 * d + 4) + 2 * n ** 2 + C)


def run_var_arm(args):
    """B concurrent variational campaigns; every step is a real BO iteration (fresh model -> stp_var_fit ->
    stp_var_acq with the update).  cfg4 pools: training sets of age a are the Sobol initial design + a random distinct
    pool points (a BO pre-roll alone would take minutes); cfg5: n_initial = M = 256, at most 4 trials."""
    import torch
    D = Dist()
    dev, rank, world = D.dev, D.rank, D.world
    from stpbenchmark_b200 import synthetic, utils
    from stpbenchmark_b200.engine import CampaignGroups
    cfg5 = args.config == "cfg5"
    kind = args.model if args.model != "ExactGP" else "VarTGP"
    B, W, K, E, S = args.campaigns, args.warmup, args.steps, args.epochs, 2048
    if cfg5:
        B = 512 if args.campaigns == 1024 else args.campaigns
        K = max(1, min(K, 4 - W)) if W < 4 else 1
        n_init, N, d, trials = 256, 8192, 8, 4
        pools = [synthetic.cfg5_pool(g) for g in range(N_POOLS)]
    else:
        n_init, N, d, trials = N_INITIAL, N_POOL, DIM, N_TRIALS
        pools = [synthetic.cfg4_pool(g) for g in range(N_POOLS)]
    cap = n_init + trials + 1
    per_pool = max(B // N_POOLS, 1)
    PX = torch.stack([p[0] for p in pools]); PY = torch.stack([p[1] for p in pools])
    pool_id = [min(b // per_pool, N_POOLS - 1) for b in range(B)]
    gen = torch.Generator().manual_seed(4242 + rank)
    t_setup = time.perf_counter()
    init = [None] * B
    for g in range(N_POOLS):
        slots = [b for b in range(B) if pool_id[b] == g]
        if not slots:
            continue
        Dg = utils.initial_designs_batch(pools[g][0], pools[g][1], [100000 * rank + (5000 if cfg5 else 0) +
                                         synthetic.campaign_seed(g, b % per_pool) for b in slots], n_init, 50.0, device=dev)
        for k, b in enumerate(slots):
            d0 = Dg[k].tolist()
            span = trials - W - 2 * K - 3          # room for the warm-up, the timed and the end-to-end steps
            age = 0 if cfg5 else ((b * trials // B + b) % span if span > 0 else 0)
            rest = [i for i in torch.randperm(N, generator=gen).tolist() if i not in set(d0)][:age] if age else []
            init[b] = d0 + rest
    batch = CampaignGroups(PX, PY, pool_id, init, n_cap=cap, kind=kind, groups=args.groups if not cfg5 else min(args.groups, 4),
                           device=dev, epochs=E, learning_rate=0.01, num_samples=S)
    for m in batch.members:
        m.ragged = not cfg5
    setup_s = time.perf_counter() - t_setup
    cat_n = lambda: torch.cat([m.n for m in batch.members])
    cat_status = lambda: torch.cat([m.status for m in batch.members])
    n0 = cat_n().clone()
    for _ in range(W):
        batch.step()
    D.barrier()
    sampler = ClockSampler(D.local_rank); sampler.start()
    l0 = batch.launches
    n_before = cat_n().clone()
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
    D.barrier()
    clocks = sampler.stop()
    done_local = int((cat_n() - n_before).sum().item())
    done = D.reduce(done_local)                              # completed iterations
    ms_local = max(ev0.elapsed_time(e) for e in ends)
    ms = D.reduce(ms_local, "max")
    bad = int((cat_status() != 0).sum().item())
    launches = batch.launches - l0
    C = 60.0 if kind == "VarGP" else S * 80.0
    nn = n_before.double()
    flops = float(sum(var_flops(nn + k, d, E, N, C).sum() for k in range(K)))
    peaks = fp64_peaks()

    # ---- end-to-end: the same step with the campaign state in pinned HOST buffers, H2D + fit + acquire + D2H per step
    e2e = None
    if not args.no_e2e:
        Ke = K if not cfg5 else 1
        names = ("X", "y", "n", "mask", "trace", "status")
        host = [{k_: getattr(m, k_).cpu().pin_memory() for k_ in names} for m in batch.members]
        h2d = sum(v.numel() * v.element_size() for h in host for v in h.values())

        def e2e_step():
            for m, st, h in zip(batch.members, batch.streams, host):
                st.synchronize()
                with torch.cuda.stream(st):
                    for k_, hv in h.items():
                        getattr(m, k_).copy_(hv, non_blocking=True)
                    m.step()
                    for k_, hv in h.items():
                        hv.copy_(getattr(m, k_), non_blocking=True)
        can = all(int(m.n.max()) + Ke < cap for m in batch.members)
        if can:
            if not cfg5:
                e2e_step()
            D.barrier()
            nb0 = int(cat_n().sum().item())
            t0 = time.perf_counter()
            for _ in range(Ke):
                e2e_step()
            D.barrier()
            dt = time.perf_counter() - t0
            it = D.reduce(int(cat_n().sum().item()) - nb0)
            e2e = {"value": it / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d, "steps": Ke,
                   "what": "campaign state (X, y, n, mask, trace, status) lives in pinned host buffers: H2D -> stp_var_fit -> "
                           "stp_var_acq (update) -> D2H every step; a stream group's next step waits for its own D2H only"}

    # ---- CPU baseline: the oracle, one iteration per core (extrapolated for cfg5) ---------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = min(host_cores(), 64)
        if cfg5:
            E_s = 6
            jobs = [dict(type="var_iteration", pool="cfg5", g=c % N_POOLS, seed=800000 + c, n_initial=256, age=0, kind=kind, epochs=E_s,
                         S=S, acq_candidates=512) for c in range(cores)]
            r = run_cpu_jobs(jobs, cores)
            secs = [x[0] * (E / E_s) + x[1] * (x[4] / x[3]) for x in r["results"]]
            how = f"EXTRAPOLATED from {E_s} of {E} epochs and 512 of {N - 256} candidates per iteration (a full iteration is ~40 s per core)"
        else:
            jobs = [dict(type="var_iteration", pool="cfg4", g=c % N_POOLS, seed=800000 + c, n_initial=N_INITIAL,
                         age=(c * 79) // max(cores - 1, 1), kind=kind, epochs=E, S=S) for c in range(cores)]
            r = run_cpu_jobs(jobs, cores)
            secs = [x[0] + x[1] for x in r["results"]]
            how = "full iterations"
        cpu = {"value": cores / max(secs), "unit": UNIT, "cores": cores, "kind": "port", "per_core": cores / sum(secs),
               "sample": f"{cores} {kind} BO iterations, one per core, n = {min(x[2] for x in r['results'])}.."
                         f"{max(x[2] for x in r['results'])}, {how}; oracle/ on {cores} single-thread processes; wall {r['wall']:.1f}s"}

    launches_all = int(D.reduce(launches))
    if rank == 0:
        achieved = flops / (ms_local * 1e-3) / 1e12
        workload = CFG5_WORKLOAD.format(B=B) if cfg5 else var_workload(B, kind, E)
        print(json.dumps({
            "metric": METRIC, "value": done / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "campaigns_per_gpu": B, "n_range": [int(n0.min()), int(n0.max()) + W + K],
                       "l2": "state and pools are L2-resident (<= 0.8 MB pools; per-campaign work squares in L2 for M = 256); "
                             "FP64-bound, no HBM stream to flush"},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA + DFMA)", "achieved": achieved, "unit": "TFLOP/s",
                         "frac": achieved / peaks["peak"], "traffic": NCU_DRAM_BYTES_VARFIT,
                         "traffic_source": ("dram bytes of one k_var_fit launch (B = 296, n = 60, 20 epochs), ncu --set full, "
                                            "profiles/r02_ncu_final_varfit_n60.txt") if NCU_DRAM_BYTES_VARFIT else None,
                         "kernel": "k_var_fit + k_var_acq", **peaks},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_all, "clocks": clocks,
            "frozen_campaigns": bad, "streams": len(batch.streams), "setup_s": setup_s}), flush=True)
    D.close()


# ------------------------------------------------------------------------------------------
# cfg2: VarTGP on the real datasets through run_benchmark's path
# ------------------------------------------------------------------------------------------
def run_cfg2_arm(args):
    import numpy as np
    D = Dist()
    from stpbenchmark_b200 import models, runners
    names = ["AgNP", "AutoAM", "CrossedBarrel", "P3HT", "Perovskite"]
    seeds = [int(s) for s in np.load(os.path.join(ROOT, "tests", "golden", "traces.npz"))["seeds"]]   # random_seeds.txt
    E, trials = args.epochs, args.trials

    def jobs(tmp, seed_list, n_trials, epochs):
        out = []
        for name in names:
            X, y = _pool_of(name, 0)
            out.append(dict(X=X, y=y, model_class=models.VarTGP, seeds=seed_list, n_initial=10, n_trials=n_trials, epochs=epochs,
                            learning_rate=0.01, invert_y=name in ("AgNP", "Perovskite"),
                            output_path=os.path.join(tmp, f"VarTGP_{name}_dataset_traces_.csv")))
        return out
    with tempfile.TemporaryDirectory() as tmp:
        runners.run_sweep(jobs(os.path.join(tmp, "warm"), seeds[:4], 3, 20))        # warm-up: library, allocator, pandas
        D.barrier()
        sampler = ClockSampler(D.local_rank); sampler.start()
        t0 = time.perf_counter()
        frames = runners.run_sweep(jobs(os.path.join(tmp, "job"), seeds, trials, E))
        D.barrier()
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
    iters = sum(int(f.xs("x_index", axis=1, level=1).iloc[10:].notna().sum().sum()) for f in frames) if D.rank == 0 else 0
    bad = sum(int(f.xs("x_index", axis=1, level=1).iloc[-1].isna().sum()) for f in frames) if D.rank == 0 else 0
    sizes = {n: tuple(_pool_of(n, 0)[0].shape) for n in names}
    flops = sum(len(seeds) * sum(var_flops(float(10 + t), sizes[n][1], E, sizes[n][0], 2048 * 80.0) for t in range(trials)) for n in names)
    h2d = sum(sizes[n][0] * (sizes[n][1] + 1) * 8 + len(seeds) * (10 * 4 + sizes[n][0]) for n in names)
    d2h = len(names) * len(seeds) * ((10 + trials + 1) * 4 + 8)
    peaks = fp64_peaks()
    cpu = None
    if D.rank == 0 and not args.no_cpu_baseline:
        cores = min(host_cores(), 64)
        cj = [dict(type="var_iteration", pool=names[c % 5], g=0, seed=45 + c, n_initial=10, age=(c * 79) // max(cores - 1, 1),
                   kind="VarTGP", epochs=E, S=2048) for c in range(cores)]
        r = run_cpu_jobs(cj, cores)
        secs = [x[0] + x[1] for x in r["results"]]
        cpu = {"value": cores / max(secs), "unit": UNIT, "cores": cores, "kind": "port", "per_core": cores / sum(secs),
               "sample": f"{cores} VarTGP BO iterations on the five datasets, one per core, n = 10..89; oracle/; wall {r['wall']:.1f}s"}
    if D.rank == 0:
        val = iters / dt
        print(json.dumps({
            "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": D.world, "steps": trials, "warmup": 3,
            "ms_per_step": 1e3 * dt / trials, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "the reference's data/*.csv (preprocessed: tests/golden/datasets.npz)",
            "config": {"workload": CFG2_WORKLOAD, "campaigns": len(names) * len(seeds), "epochs": E,
                       "timing": "host clock around runners.run_sweep, device synchronised on both sides: value IS the end-to-end "
                                 "number of this config (inputs start on the host, CSV files are written inside the timed region)"},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA + DFMA)", "achieved": flops / dt / 1e12, "unit": "TFLOP/s",
                         "frac": flops / dt / 1e12 / peaks["peak"], "traffic": None, "kernel": "k_var_fit + k_var_acq", **peaks},
            "cpu_baseline": cpu,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": h2d / trials, "d2h_bytes_per_step": d2h / trials, "steps": trials,
                    "wall_s": dt},
            "gpu_launches": 2 * 3 * trials * D.world, "clocks": clocks, "frozen_campaigns": bad}), flush=True)
    D.close()


# ------------------------------------------------------------------------------------------
# cfg3: Hartmann-6 fit-and-predict, 256 batched seeds x 4 models
# ------------------------------------------------------------------------------------------
def run_cfg3_arm(args):
    import torch
    D = Dist()
    dev = D.dev
    from stpbenchmark_b200 import _lib, priors, synthetic, utils, varops
    S_ = 256 if args.campaigns == 1024 else args.campaigns
    n, d, E = 100, 6, args.epochs
    Xtr = torch.zeros(S_, n + 1, d, dtype=torch.double); ytr = torch.zeros(S_, n + 1, dtype=torch.double)
    Xte = torch.zeros(S_, n, d, dtype=torch.double)
    for s in range(S_):                      # test_hartmann.py:13-27 per seed
        utils.set_seeds(s)
        a = torch.rand(n, d).double(); b = torch.rand(n, d).double()
        Xtr[s, :n] = a; ytr[s, :n] = synthetic.hartmann6(a); Xte[s] = b
    yte = synthetic.hartmann6(Xte.reshape(-1, d)).reshape(S_, n).to(dev)
    Xd, yd, Xq = Xtr.to(dev), ytr.to(dev), Xte.to(dev)
    nn = torch.full((S_,), n, dtype=torch.int32, device=dev)
    pid = torch.arange(S_, dtype=torch.int32, device=dev)
    results, total_ms = {}, 0.0
    launches = [0]

    def one(kind):
        if kind == "ExactGP":
            theta = torch.tensor([priors.exact_theta0(d)] * S_, dtype=torch.double, device=dev)
            lo, hi = priors.exact_bounds(d)
            _lib.exact_fit(Xd, yd, theta, lo, hi, priors.exact_prior_vector(d), n=nn)
            out = _lib.exact_acq(Xq, Xd, yd, theta, pool_id=pid, n=nn, want_pointwise=True)
            launches[0] += 2
            return out["mean"]
        lik = _lib.LIK[kind]
        st = varops.alloc_state(S_, n + 1, d, dev)
        status = torch.zeros(S_, dtype=torch.int32, device=dev)
        work = _lib.var_workspace(S_, n + 1, dev)
        _lib.var_fit(Xd, yd, nn, st, lik, E, 0.01, varops.ADAM_LR, varops.var_hyp0(d), varops.var_prior_vector(),
                     varops.gauss_hermite_table(), status, work, fresh=True, seed=7)
        bf = torch.zeros(S_, dtype=torch.double, device=dev)
        out = _lib.var_acq(Xq, st, nn, 0, status, work, best_f=bf, pool_id=pid, S=1, want_pointwise=True)
        launches[0] += 2
        mu = yd[:, :n].mean(1, keepdim=True); sd = yd[:, :n].std(1, unbiased=False, keepdim=True) + 1e-10
        return out["mean"] * sd + mu
    kinds = ["ExactGP", "VarGP", "VarLGP", "VarTGP"]
    for k in kinds:
        one(k)                                 # warm-up
    D.barrier()
    launches[0] = 0
    sampler = ClockSampler(D.local_rank); sampler.start()
    for k in kinds:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mean = one(k); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        total_ms += ms
        results[k] = {"fits_per_sec": S_ / (ms * 1e-3), "test_mae": float((mean - yte).abs().mean().item())}
    clocks = sampler.stop()
    cpu = None
    if D.rank == 0 and not args.no_cpu_baseline:
        cores = min(host_cores(), 64)
        cj = [dict(type="hartmann_fit", seed=c, kind=kinds[c % 4], epochs=E) for c in range(cores)]
        r = run_cpu_jobs(cj, cores)
        cpu = {"value": cores / max(x[0] for x in r["results"]), "unit": "fits/s", "cores": cores, "kind": "port",
               "per_core": cores / sum(x[0] for x in r["results"]),
               "sample": f"{cores} fits (model kinds round-robin), one per core; oracle/; wall {r['wall']:.1f}s",
               "test_mae": {k: sum(x[1] for x, j in zip(r["results"], cj) if j["kind"] == k) / max(1, sum(1 for j in cj if j["kind"] == k))
                            for k in kinds}}
    if D.rank == 0:
        val = 4 * S_ / (total_ms * 1e-3)
        print(json.dumps({
            "metric": "model_fits_per_sec", "value": val, "unit": "fits/s", "n_gpus": D.world, "steps": 1, "warmup": 1,
            "ms_per_step": total_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg3: Hartmann-6 fit-and-predict (test_hartmann.py scenario), {S_} seeds x "
                                   f"{{ExactGP (L-BFGS-B), VarGP, VarLGP, VarTGP ({E} epochs NGD+Adam)}}, 100 train / 100 test points"},
            "per_model": results, "roofline": None, "cpu_baseline": cpu, "e2e": None, "gpu_launches": launches[0], "clocks": clocks}),
            flush=True)
    D.close()


# ------------------------------------------------------------------------------------------
# f3: the fit-and-predict study of benchmark_predictions.py (3 models x 5 datasets x 25 seeds, 80/20 splits, n up to 480)
# ------------------------------------------------------------------------------------------
F3_WORKLOAD = ("f3: benchmark_predictions.py study: {{VarTGP, VarGP, ExactGP}} x 5 materials datasets x {S} seeds, fit on an 80 % "
               "split (n = 76, 80, 132, 143, 480), {E} epochs NGD+Adam / L-BFGS-B, MAE of posterior.mean on the other 20 %")


def f3_cpu_leg(kind_names, names, seeds, S_, E, fits, cores):
    """The study on the host cores with the oracle: one fit + predict per core, (model, dataset) pairs round-robin,
    variational fits extrapolated from E_s of E epochs; value = fits / (summed single-core seconds of the study / cores)."""
    E_s = min(60, E)                                  # 60 of the 600 epochs timed (a 600-epoch fit at n = 480 is ~36 s per core)
    order = [(k, n_) for n_ in names for k in kind_names]
    cj = [dict(type="predict_fit", pool=order[c % len(order)][1], kind=order[c % len(order)][0], seed=seeds[c % S_],
               epochs=E_s) for c in range(max(cores, len(order)))]       # every pair at least once
    r = run_cpu_jobs(cj, cores)
    full = {}
    for j, x in zip(cj, r["results"]):
        t_fit = x[0] if j["kind"] == "ExactGP" else x[0] * (E / E_s)
        full.setdefault((j["kind"], j["pool"]), []).append(t_fit + x[1])
    mean_s = {kd: sum(v) / len(v) for kd, v in full.items()}
    covered = [kd for kd in order if kd in mean_s]
    study_core_s = S_ * sum(mean_s[kd] for kd in covered) * (len(order) / max(len(covered), 1))
    return {"value": fits / (study_core_s / cores), "unit": "fits/s", "cores": cores, "kind": "port",
            "per_core": fits / study_core_s,
            "sample": f"{len(cj)} fits on {cores} cores, (model, dataset) pairs round-robin ({len(covered)} of {len(order)} covered); "
                      f"variational fits EXTRAPOLATED from {E_s} of {E} epochs; value = {fits} fits / (summed single-core seconds of "
                      f"the study / cores); oracle/; wall {r['wall']:.1f}s",
            "seconds_per_fit": {f"{k}/{n_}": mean_s[(k, n_)] for (k, n_) in covered}, "study_core_seconds": study_core_s}


def run_f3_arm(args):
    import numpy as np
    import torch
    D = Dist()
    from stpbenchmark_b200 import benchmark_predictions as bp
    from stpbenchmark_b200.models import ExactGP, VarGP, VarTGP
    names = ["AgNP", "AutoAM", "CrossedBarrel", "P3HT", "Perovskite"]
    S_ = 25 if args.campaigns == 1024 else args.campaigns
    E = args.epochs
    seeds = [int(s) for s in np.load(os.path.join(ROOT, "tests", "golden", "traces.npz"))["seeds"]][:S_]   # random_seeds.txt
    z = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
    kinds = [("VarTGP", VarTGP), ("VarGP", VarGP), ("ExactGP", ExactGP)]
    with tempfile.TemporaryDirectory() as tmp:
        for name in names:                           # the reference's data/*.csv
            tab = z[f"{name}_raw"]
            np.savetxt(os.path.join(tmp, f"{name}_dataset.csv"), tab, delimiter=",",
                       header=",".join(f"c{i}" for i in range(tab.shape[1])), comments="")
        files = [f"{n}_dataset.csv" for n in names]
        for _k, cls in kinds:                        # warm-up: library, allocator, shared-memory grants
            bp.evaluate_model_across_datasets(cls, n_seeds=2, epochs=5, data_dir=tmp, seeds=seeds, verbose=False, dataset_files=files)
        # (1) per (model, dataset) pair alone on the device (diagnostic table): CUDA events around fit + predict
        flops, serial_ms, launches, per, dev_in = 0.0, 0.0, 0, {}, {}
        sampler = ClockSampler(D.local_rank); sampler.start()
        for name in names:
            X, y = _pool_of(name, 0, raw_y=True)
            tr, te = zip(*[bp.split_indices(len(X), s_) for s_ in seeds])
            tr, te = torch.stack(tr), torch.stack(te)
            Xtr, ytr, Xte = X[tr].to(D.dev), y[tr].to(D.dev), X[te].to(D.dev)
            dev_in[name] = (Xtr, ytr, Xte)
            n, d, m = Xtr.shape[1], Xtr.shape[2], Xte.shape[1]
            for k, _cls in kinds:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                res = bp.fit_predict_batch(k, Xtr, ytr, Xte, epochs=E, lr=0.01)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                serial_ms += ms
                if k == "ExactGP":
                    w = sum(float(F) * (n ** 3 + n * n * (3 * d + 6)) + n ** 3 / 3 + m * (n * n + n * (3 * d + 8) + 60)
                            for F in res["nfev"].tolist())
                else:
                    w = S_ * (E * (9 * n ** 3 + n * n * (14 * d + 12) + 800 * n) + 3 * n ** 3 + m * (n * (3 * d + 4) + 2 * n * n + 60))
                flops += w
                per[f"{k}/{name}"] = {"n": n, "ms": ms, "fits_per_sec": S_ / (ms * 1e-3), "tflops": w / (ms * 1e-3) / 1e12,
                                      "bad": int((res["status"] != 0).sum())}
        # (2) the timed step: all pairs in flight together, inputs resident in HBM, one CUDA stream per pair (what the
        #     product path does), longest pairs first; device time from the start event to the last stream's end event
        order = sorted(((k, nm) for k, _c in kinds for nm in names),
                       key=lambda p_: -(dev_in[p_[1]][0].shape[1] ** 3) * (E if p_[0] != "ExactGP" else 100.0))
        torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
        ends, keep = [], []
        for k, nm in order:
            st = torch.cuda.Stream(); st.wait_event(ev0)
            with torch.cuda.stream(st):
                keep.append(bp.launch_fit_predict(k, *dev_in[nm], epochs=E, lr=0.01))
                e = torch.cuda.Event(enable_timing=True); e.record(st); ends.append(e)
            launches += 2
        torch.cuda.synchronize()
        dev_ms = max(ev0.elapsed_time(e) for e in ends)
        del keep
        # (3) end to end through the product API: CSV files -> result dictionaries (host buffers both ways)
        D.barrier()
        t0 = time.perf_counter()
        r = bp.evaluate_models_across_datasets([c for _k, c in kinds], n_seeds=S_, epochs=E, lr=0.01, data_dir=tmp, seeds=seeds,
                                               verbose=False, dataset_files=files)
        maes = {k: {ds.replace("_dataset", ""): v["mean_mae"] for ds, v in r[c.__name__].items()} for k, c in kinds}
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
    fits = len(kinds) * len(names) * S_ * D.world      # under torchrun every rank runs the whole study (replicas only)
    dev_ms = D.reduce(dev_ms, "max"); wall = D.reduce(wall, "max")
    sizes = {n_: z[f"{n_}_X"].shape for n_ in names}
    h2d = len(kinds) * S_ * sum(sizes[n_][0] * (sizes[n_][1] + 1) * 8 for n_ in names)
    d2h = len(kinds) * S_ * sum(int(0.2 * sizes[n_][0]) * 16 + 4 for n_ in names)
    peaks = fp64_peaks()
    cpu = None
    if D.rank == 0 and not args.no_cpu_baseline:
        cores = min(host_cores(), 64)
        cpu = f3_cpu_leg([k for k, _c in kinds], names, seeds, S_, E, fits // D.world, cores)
    if D.rank == 0:
        print(json.dumps({
            "metric": "model_fits_per_sec", "value": fits / (dev_ms * 1e-3), "unit": "fits/s", "n_gpus": D.world, "steps": 1, "warmup": 1,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "the reference's data/*.csv (tests/golden/datasets.npz raw tables written back as CSV)",
            "config": {"workload": F3_WORKLOAD.format(S=S_, E=E), "fits": fits, "multi_gpu": "replicas only (every rank runs the study)",
                       "l2": "each campaign's working squares (1.9 - 15 MB at n = 480) are L2-resident; FP64 / latency bound"},
            "per_pair": per, "serial_ms": serial_ms, "mean_mae": maes,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DMMA + DFMA)", "achieved": flops / (dev_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                         "frac": flops / (dev_ms * 1e-3) / 1e12 / peaks["peak"], "traffic": None,
                         "kernel": "k_var_fit<BIG> (CrossedBarrel, n = 480: 25 CTAs on 148 SMs)", **peaks},
            "cpu_baseline": cpu,
            "e2e": {"value": fits / wall, "unit": "fits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "wall_s": wall,
                    "api": "benchmark_predictions.evaluate_models_across_datasets on CSV files"},
            "gpu_launches": launches, "clocks": clocks}), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5", "cfg2", "cfg3", "f3"])
    ap.add_argument("--campaigns", type=int, default=1024, help="campaigns per GPU (cfg5: 512, cfg3: 256 seeds)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--groups", type=int, default=DEFAULT_GROUPS,
                    help="campaign groups per GPU, one CUDA stream each (at most B / 16; beyond 32 streams launches serialise)")
    ap.add_argument("--chunk", type=int, default=1,
                    help="cohort schedule: BO iterations per launch (1 = stp_exact_bo_step, > 1 = stp_exact_bo_run)")
    ap.add_argument("--schedule", default="cohort", choices=["mixed", "cohort", "queue"],
                    help="cohort: the campaigns of a group share their age (groups staggered over the 80 trials), each launch "
                         "is sized for its own n; mixed: every group holds all ages; queue: ONE persistent work-queue launch "
                         "(stp_exact_bo_run) does all K steps")
    ap.add_argument("--model", default="ExactGP", choices=["ExactGP", "VarGP", "VarTGP", "VarLGP"],
                    help="ExactGP = the headline workload; the others run the variational workload on the cfg4 pools")
    ap.add_argument("--epochs", type=int, default=600)
    ap.add_argument("--trials", type=int, default=N_TRIALS,
                    help="trials per campaign (diagnostic: 80 is the workload; smaller keeps n = 10..9+trials)")
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        return _cpu_worker_main()
    if args.trials != N_TRIALS:
        globals().update(N_TRIALS=args.trials, CAP=N_INITIAL + args.trials + 1)
    variational = args.model != "ExactGP" or args.config == "cfg5"
    if args.steps is None:
        args.steps = 40 if not variational and args.config == "cfg4" else (2 if args.config == "cfg5" else 5)
    if args.warmup is None:
        args.warmup = 1 if args.config == "cfg5" else 3
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config == "cfg2":
        run_cfg2_arm(args)
    elif args.config == "cfg3":
        run_cfg3_arm(args)
    elif args.config == "f3":
        run_f3_arm(args)
    elif variational:
        args.steps = min(args.steps, 10)
        run_var_arm(args)
    else:
        run_exact_arm(args)


if __name__ == "__main__":
    main()
