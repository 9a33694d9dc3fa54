"""Batched campaign engine: B independent BO campaigns advanced together on one GPU.

The reference runs campaigns strictly one after another (run_benchmark.py:20-43,
runners.py:174-200); they share nothing, so here every campaign is one CTA of the same launch
and the whole loop body runners.py:64-109 (fresh model -> fit -> LogEI over the pool -> argmax
-> move the pick from the pool to the training set) is ONE kernel per iteration
(stp_exact_bo_step) or two (stp_var_fit + stp_var_acq_step) for the variational models.

State kept resident in HBM for the whole run (row-major, batch-major, fp64 / int32 / uint8):
    pool   [G, N, d]   candidate inputs of the G distinct pools (shared by their campaigns)
    pool_y [G, N]      their targets
    pool_id[B]         which pool a campaign draws from
    mask   [B, N]      1 = still a candidate (runners.py:53-54)
    X      [B, cap, d] training inputs, rows < n[b] valid (duplicates kept, App. D-4)
    y      [B, cap]    training targets (RAW, App. D-1)
    n      [B]         current training-set size
    trace  [B, cap]    selected pool index per row: initial design first, then one per trial
    status [B]         0 running; non-zero = frozen (the reference's try/except, runners.py:113-115)
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, priors


class CampaignBatch:
    """Device-resident state of B campaigns over G pools, for any of the four model kinds."""

    def __init__(self, pools: torch.Tensor, pool_ys: torch.Tensor, pool_id: Sequence[int],
                 init_indices: Sequence[Sequence[int]], n_cap: int, kind: str = "ExactGP",
                 device: Optional[torch.device] = None, epochs: int = 600, learning_rate: float = 0.01,
                 num_samples: int = 2048, seeds: Optional[Sequence[int]] = None,
                 pool_sizes: Optional[Sequence[int]] = None, lengthscale_prior=None, kernel: str = "rbf",
                 acquisition: str = "logei", beta: float = 0.0):
        if not torch.cuda.is_available():
            raise RuntimeError("CampaignBatch needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        if pools.dim() == 2:
            pools, pool_ys = pools.unsqueeze(0), pool_ys.reshape(1, -1)
        self.kind = kind
        # kernel / acquisition variants (row f4); None = RBF + LogEI, the reference's path
        self.model_opts = _lib.ModelOpts.make(kernel, acquisition, beta)
        if kind != "ExactGP" and str(kernel).lower() != "rbf":
            raise NotImplementedError("the variational models provide the RBF kernel only")
        self.G, self.N, self.d = pools.shape
        self.B = len(pool_id)
        self.cap = int(n_cap)
        if self.d > _lib.MAX_D:
            raise ValueError(f"d <= {_lib.MAX_D} required")
        if kind == "ExactGP":
            # beyond the shared-memory capacity the working squares live in a global workspace (stp_exact_bo_step_large)
            self.large = self.cap > _lib.MAX_N
            if self.cap > _lib.MAX_N_LARGE:
                raise ValueError(f"n_cap <= {_lib.MAX_N_LARGE} required for ExactGP")
        else:
            self.large = False
            self.cap = _lib.var_capacity(self.cap, self.d)      # the smallest launchable capacity (raises beyond the limit)
        dev = self.device
        self.pool = pools.to(dev, torch.double).contiguous()
        self.pool_y = pool_ys.to(dev, torch.double).contiguous()
        self.pool_id = torch.tensor(list(pool_id), dtype=torch.int32, device=dev)
        B, cap, d, N = self.B, self.cap, self.d, self.N
        # host-side assembly of the initial designs, then one H2D copy per array
        mask = torch.ones(B, N, dtype=torch.uint8)
        trace = torch.full((B, cap), -1, dtype=torch.int32)
        n = torch.zeros(B, dtype=torch.int32)
        lengths = {len(idx) for idx in init_indices}
        if max(lengths, default=0) > cap:
            raise ValueError("initial design longer than n_cap")
        if len(lengths) == 1 and B > 0 and max(lengths) > 0:      # the usual case: one vectorised assembly
            k = max(lengths)
            t = torch.as_tensor(np.asarray(init_indices, dtype=np.int64).reshape(B, k))
            mask.scatter_(1, t, 0)
            trace[:, :k] = t.to(torch.int32)
            n[:] = k
        else:
            for b, idx in enumerate(init_indices):
                t = torch.tensor([int(i) for i in idx], dtype=torch.long)
                mask[b, t] = 0
                trace[b, : len(idx)] = t.to(torch.int32)
                n[b] = len(idx)
        if pool_sizes is not None:      # ragged pools, zero-padded to N rows: the padding is never a candidate
            for b, g in enumerate(pool_id):
                mask[b, int(pool_sizes[int(g)]):] = 0
        self.pool_sizes = None if pool_sizes is None else [int(v) for v in pool_sizes]
        self.lengthscale_prior = lengthscale_prior
        self.mask = mask.to(dev)
        self.trace = trace.to(dev)
        self.n = n.to(dev)
        self.X = torch.zeros(B, cap, d, dtype=torch.double, device=dev)
        self.y = torch.zeros(B, cap, dtype=torch.double, device=dev)
        self._gather_training_sets()
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.theta = torch.zeros(B, d + 3, dtype=torch.double, device=dev)
        self.acq = torch.zeros(B, dtype=torch.double, device=dev)
        self.nit = torch.zeros(B, dtype=torch.int32, device=dev)
        self.nfev = torch.zeros(B, dtype=torch.int32, device=dev)
        self.launches = 0
        self.epochs, self.learning_rate, self.num_samples = int(epochs), float(learning_rate), int(num_samples)
        self.seeds = list(seeds) if seeds is not None else list(range(B))
        self.iteration = 0
        self.ragged = len({len(i) for i in init_indices}) > 1   # campaigns of different ages: schedule longest first
        self._turnover = None
        # host mirror of n[b] (an upper bound: frozen campaigns stop growing).  Lets a launch be sized for the current
        # training sets instead of the row capacity (stp_exact_bo_step's n_max); None = unknown.
        self._n_host: Optional[np.ndarray] = np.array([len(i) for i in init_indices], dtype=np.int64)
        self.work_acc = torch.zeros(B, dtype=torch.double, device=dev)   # algorithmic flops done per campaign slot
        self.iters_acc = torch.zeros(B, dtype=torch.int32, device=dev)   # completed BO iterations per campaign slot
        if kind == "ExactGP":
            self._theta0 = priors.exact_theta0(d)
            self._lower, self._upper = priors.exact_bounds(d)
            self._prior = priors.exact_prior_vector(d)
            self._opts = _lib.LbfgsbOpts.scipy_defaults()
            self._work = _lib.exact_large_work(B, cap, dev) if self.large else None
        else:
            from . import varops
            self._var = varops.VarBatchState(self)

    def _gather_training_sets(self):
        """X[b, i] = pool[pool_id[b], trace[b, i]] for i < n[b] (runners.py:56-57)."""
        idx = self.trace.clamp_min(0).long()
        pid = self.pool_id.long().unsqueeze(1).expand_as(idx)
        valid = (torch.arange(self.cap, device=self.device).unsqueeze(0) < self.n.unsqueeze(1))
        self.X.copy_(self.pool[pid, idx] * valid.unsqueeze(-1))
        self.y.copy_(self.pool_y[pid, idx] * valid)

    def step(self) -> None:
        """One BO iteration (runners.py:64-109) for every running campaign."""
        if self.kind == "ExactGP":
            # longest-first: a fit costs ~ closures x n^3, so the biggest training sets go first (no host sync)
            order = torch.argsort(self.n, descending=True, stable=True).to(torch.int32) if self.ragged else None
            n_max = 0 if self._n_host is None else int(min(self._n_host.max(), self.cap - 1))
            if self.large:
                if self._turnover is not None:
                    raise NotImplementedError("continuous batching is provided for the shared-memory step only")
                _lib.exact_bo_step_large(self.pool, self.pool_y, self.pool_id, self.mask, self.X, self.y, self.n, self.trace,
                                         self._theta0, self._lower, self._upper, self._prior, self.status, self._work,
                                         theta_out=self.theta, acq_out=self.acq, nit=self.nit, nfev=self.nfev,
                                         opts=self._opts, order=order, work_acc=self.work_acc, iters_acc=self.iters_acc,
                                         n_max=n_max, model=self.model_opts)
                self.launches += 1
                if self._n_host is not None:
                    self._n_host += 1
                self.iteration += 1
                return
            _lib.exact_bo_step(self.pool, self.pool_y, self.pool_id, self.mask, self.X, self.y, self.n, self.trace,
                               self._theta0, self._lower, self._upper, self._prior, self.status,
                               theta_out=self.theta, acq_out=self.acq, nit=self.nit, nfev=self.nfev,
                               opts=self._opts, order=order, turnover=self._turnover, work_acc=self.work_acc,
                               iters_acc=self.iters_acc, n_max=n_max, model=self.model_opts)
            self.launches += 1
            if self._n_host is not None:          # what the kernel does to n: +1, and the turnover's restart
                self._n_host += 1
                if self._turnover is not None:
                    self._n_host[self._n_host >= self._turnover.n_final] = self._turnover.n_init
        else:
            self.launches += self._var.step()
            if self._n_host is not None:
                self._n_host += 1
        self.iteration += 1

    def run(self, n_trials: int, chunk: Optional[int] = None) -> None:
        """``n_trials`` BO iterations of every campaign.  ExactGP: persistent work-queue launches (stp_exact_bo_run)
        of ``chunk`` iterations each (default 16; each launch is sized for the training sets it can reach, so young
        campaigns run 3 - 4 per SM); campaigns advance independently inside a launch.  Variational kinds: lock-step
        steps (two launches per iteration)."""
        n_trials = int(n_trials)
        if self.kind != "ExactGP":
            for _ in range(n_trials):
                self.step()
            return
        chunk = int(chunk) if chunk else 16
        done = 0
        while done < n_trials:
            k = min(chunk, n_trials - done)
            self.run_queue(k)
            done += k

    def run_queue(self, iters, max_ctas: int = 0) -> None:
        """ONE persistent launch that advances every running campaign by ``iters`` BO iterations (ExactGP); ``iters``
        is an int or one count per campaign."""
        if self.kind != "ExactGP":
            raise NotImplementedError("the persistent work-queue launch is provided for the ExactGP step")
        if self.large:          # no persistent form beyond the shared-memory capacity: one launch per iteration
            if not isinstance(iters, int):
                raise NotImplementedError("per-campaign iteration counts need the shared-memory work-queue kernel")
            for _ in range(iters):
                self.step()
            return
        if not hasattr(self, "_sched"):
            self._sched = torch.zeros(self.B + 1, dtype=torch.int32, device=self.device)
            self._iters_left = torch.zeros(self.B, dtype=torch.int32, device=self.device)
        if isinstance(iters, int):
            self._iters_left.fill_(iters)
            per = None
        else:                                  # per-campaign iteration counts
            per = np.asarray(list(iters), dtype=np.int64)
            self._iters_left.copy_(torch.as_tensor(per, dtype=torch.int32))
            iters = int(per.max()) if per.size else 0
        if self._n_host is None:
            n_max = 0
        elif self._turnover is not None:      # restarted campaigns are young again: the oldest one bounds the launch
            n_max = int(min(max(self._n_host.max() + iters, self._turnover.n_init + iters), self._turnover.n_final - 1,
                            self.cap - 1))
        else:
            n_max = int(min(self._n_host.max() + iters - 1, self.cap - 1))
        _lib.exact_bo_run(self.pool, self.pool_y, self.pool_id, self.mask, self.X, self.y, self.n, self.trace,
                          self._theta0, self._lower, self._upper, self._prior, self.status, self._iters_left, self._sched,
                          theta_out=self.theta, acq_out=self.acq, nit=self.nit, nfev=self.nfev, opts=self._opts,
                          turnover=self._turnover, work_acc=self.work_acc, iters_acc=self.iters_acc, n_max=n_max,
                          max_ctas=max_ctas, model=self.model_opts)
        self.launches += 1
        self.iteration += int(iters)
        if self._n_host is not None:
            self._n_host += int(iters) if per is None else per
            if self._turnover is not None:
                span = self._turnover.n_final - self._turnover.n_init
                over = self._n_host >= self._turnover.n_final
                self._n_host[over] = self._turnover.n_init + (self._n_host[over] - self._turnover.n_final) % span

    def assume_ages(self, n_host) -> None:
        """Tell the engine the current training-set sizes (host array [B], or None = unknown) after state that it
        mirrors on the host (n, status) has been changed from outside."""
        self._n_host = None if n_host is None else np.asarray(n_host, dtype=np.int64).copy()

    def traces(self, sync: bool = True):
        """(indices [B, cap] int64 with -1 beyond the last pick, n [B], status [B]) on the host.  ``sync=False``: the
        caller has already synchronised the stream this batch runs on."""
        if sync:
            torch.cuda.synchronize(self.device)
        return self.trace.cpu().long(), self.n.cpu().long(), self.status.cpu().long()

    def selected_values(self):
        idx, n, _ = self.traces()
        py = self.pool_y.cpu()
        pid = self.pool_id.cpu().long()
        out: List[List[float]] = []
        for b in range(self.B):
            out.append([float(py[pid[b], idx[b, i]]) for i in range(int(n[b]))])
        return out

    # -- continuous batching inside the kernel (ExactGP): stp_turnover --------------------------------
    def enable_turnover(self, bank: torch.Tensor, n_final: int, sink_capacity: int = 0) -> None:
        """Let stp_exact_bo_step recycle finished campaigns itself: ``bank`` is int [B, G, n_init] (the next G initial
        designs of every slot), a campaign is finished when its training set has ``n_final`` rows; finished traces are
        appended to ``self.sink`` (``sink_capacity`` rows).  No host or torch work between steps."""
        if self.kind != "ExactGP":
            raise NotImplementedError("in-kernel turnover is provided for the ExactGP step")
        B, G, ni = bank.shape
        assert B == self.B and n_final <= self.cap
        self.bank = bank.to(self.device, torch.int32).contiguous()
        self.gen = torch.zeros(B, dtype=torch.int32, device=self.device)
        self.sink = torch.full((max(sink_capacity, 1), self.cap), -1, dtype=torch.int32, device=self.device)
        self.sink_count = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._turnover = _lib.Turnover(int(n_final), int(ni), int(G), int(sink_capacity), self.bank.data_ptr(),
                                       self.gen.data_ptr(), self.sink.data_ptr() if sink_capacity else None,
                                       self.sink_count.data_ptr())

    # -- continuous batching: finished campaigns leave, fresh ones take their slots --------------
    def recycle(self, slots: torch.Tensor, new_init: torch.Tensor, sink: Optional[torch.Tensor] = None,
                sink_offset: int = 0) -> None:
        """Replace the campaigns in ``slots`` (int64 [k], device) by fresh ones whose initial
        designs are ``new_init`` (int64 [k, n_initial], device).  The finished traces are first
        copied into ``sink[sink_offset : sink_offset + k]`` when a sink is given.  Plumbing only
        (a handful of tiny torch index kernels); no host synchronisation."""
        if slots.numel() == 0:
            return
        self._n_host = None   # slots live on the device: the host mirror of n is no longer valid
        if sink is not None:
            sink[sink_offset: sink_offset + slots.numel()] = self.trace[slots]
        k, ni = new_init.shape
        pid = self.pool_id[slots].long()
        self.mask[slots] = 1
        self.mask[slots.unsqueeze(1), new_init] = 0
        self.trace[slots] = -1
        self.trace[slots, :ni] = new_init.to(torch.int32)
        self.X[slots] = 0
        self.y[slots] = 0
        self.X[slots, :ni] = self.pool[pid.unsqueeze(1), new_init]
        self.y[slots, :ni] = self.pool_y[pid.unsqueeze(1), new_init]
        self.n[slots] = ni
        self.status[slots] = 0


class Campaign:
    """One campaign of a job: which pool, which model kind, its initial design and its budget."""

    __slots__ = ("key", "kind", "pool", "design", "n_trials", "epochs", "learning_rate", "num_samples",
                 "lengthscale_prior", "rng_state", "seed", "n_initial", "kernel", "acquisition", "beta")

    def __init__(self, key, kind: str, pool: int, design: Optional[Sequence[int]], n_trials: int, epochs: int = 100,
                 learning_rate: float = 0.1, num_samples: int = 2048, lengthscale_prior=None, rng_state=None, seed: int = 0,
                 n_initial: Optional[int] = None, kernel: str = "rbf", acquisition: str = "logei", beta: float = 0.0):
        """``design`` may be None until the rank that owns the campaign fills it in (``n_initial`` then says how long
        it will be).  ``kernel`` / ``acquisition`` / ``beta``: the variants of include/stp_b200.h's stp_model_opts
        (defaults = the reference's RBF + LogEI)."""
        self.kernel, self.acquisition, self.beta = str(kernel).lower(), str(acquisition).lower(), float(beta)
        self.key, self.kind, self.pool = key, kind, int(pool)
        self.design = None if design is None else [int(i) for i in design]
        self.n_initial = int(n_initial) if n_initial is not None else len(self.design)
        self.n_trials, self.epochs, self.learning_rate = int(n_trials), int(epochs), float(learning_rate)
        self.num_samples, self.lengthscale_prior, self.rng_state, self.seed = int(num_samples), lengthscale_prior, rng_state, int(seed)

    def group_key(self, d: int):
        """Campaigns that can share launches: same kernel arguments and the same age at every step."""
        lp = None if self.lengthscale_prior is None else tuple(float(v) for v in self.lengthscale_prior)
        return (self.kind, d, len(self.design), self.n_trials, self.epochs, self.learning_rate, self.num_samples, lp,
                self.kernel, self.acquisition, self.beta)


class CampaignJob:
    """Every campaign of a sweep -- seeds x datasets x model kinds -- on ONE GPU at once (the reference runs them one
    after another: run_benchmark.py:20-43 -> runners.py:174-178).

    Campaigns are grouped by what a launch must have in common (model kind, input dimension, budget); the pools of a
    group may have different sizes (they are zero-padded to the largest one and the padding is masked out), so the
    five datasets of BASELINE config #2 form three groups (d = 3, 4, 5) instead of five half-empty batches.  Every
    group is a CampaignBatch on its own CUDA stream: ExactGP groups advance by persistent work-queue launches
    (stp_exact_bo_run, ``chunk`` iterations each), variational groups in lock-step (stp_var_fit + stp_var_acq), and
    the groups overlap on the device.  ``run`` returns when all are done; ``on_group_done`` is called per group as
    soon as ITS campaigns have finished (the hook the CSV writer uses to save finished seeds early).
    """

    def __init__(self, pools: Sequence, campaigns: Sequence[Campaign], device: Optional[torch.device] = None,
                 chunk: int = 16, max_group: int = 256):
        if not torch.cuda.is_available():
            raise RuntimeError("CampaignJob needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.campaigns = list(campaigns)
        # tuning knobs (defaults measured on cfg #4, profiles/r02_e2e_tuning.md): iterations per persistent launch and
        # the largest ExactGP sub-group
        chunk = int(os.environ.get("STP_JOB_CHUNK", chunk))
        max_group = int(os.environ.get("STP_JOB_MAX_GROUP", max_group))
        self.chunk = int(chunk)
        by_key = {}
        for c_id, c in enumerate(self.campaigns):
            d = pools[c.pool][0].shape[1]
            by_key.setdefault(c.group_key(d), []).append(c_id)
        self.groups = []
        # a launch cannot end before its slowest campaign: large ExactGP groups are cut into sub-groups of at most
        # `max_group` campaigns on their own streams, so that the tail of one launch is filled by the others
        parts = []
        for key, ids in by_key.items():
            if key[0] == "ExactGP" and max_group > 0 and len(ids) > max_group:
                pieces = -(-len(ids) // max_group)
                parts += [(key, ids[p * len(ids) // pieces:(p + 1) * len(ids) // pieces]) for p in range(pieces)]
            else:
                parts.append((key, ids))
        # earlier groups get the higher stream priority: their CTAs win the SM slots that free up, they finish first,
        # and the host-side work on their traces (frames, CSV files) overlaps with the device work of the later ones
        prio_levels = 1
        lo_p = 0
        try:
            if os.environ.get("STP_JOB_PRIORITIES", "1") == "0":
                raise RuntimeError("priorities disabled")
            a_p, b_p = torch.cuda.Stream.priority_range()        # least and greatest priority (lower number = higher)
            lo_p, hi_p = max(a_p, b_p), min(a_p, b_p)
            prio_levels = max(1, lo_p - hi_p + 1)
        except Exception:
            prio_levels, lo_p = 1, 0
        for part_no, (key, ids) in enumerate(parts):
            kind, d, n_init, n_trials, epochs, lr, S, lp, kernel, acquisition, beta = key
            used = sorted({self.campaigns[i].pool for i in ids})
            local = {g: k for k, g in enumerate(used)}
            sizes = [int(pools[g][0].shape[0]) for g in used]
            N = max(sizes)
            PX = torch.zeros(len(used), N, d, dtype=torch.double)
            PY = torch.zeros(len(used), N, dtype=torch.double)
            for g in used:
                PX[local[g], : sizes[local[g]]] = pools[g][0].detach().cpu().double()
                PY[local[g], : sizes[local[g]]] = pools[g][1].detach().cpu().double().flatten()
            trials = min(n_trials, min(sizes) - n_init)      # runners.py:62 (per pool; equal budgets share a group)
            if any(min(n_trials, sz - n_init) != trials for sz in sizes):
                raise ValueError("pools of one group must leave the same number of trials (runners.py:62)")
            rank_p = min(len(parts) - 1 - part_no, prio_levels - 1) if len(parts) > 1 else 0
            stream = torch.cuda.Stream(device=self.device, priority=lo_p - rank_p)
            with torch.cuda.stream(stream):
                batch = CampaignBatch(PX, PY, [local[self.campaigns[i].pool] for i in ids],
                                      [self.campaigns[i].design for i in ids], n_cap=n_init + max(trials, 0) + 1, kind=kind,
                                      device=self.device, epochs=epochs, learning_rate=lr, num_samples=S,
                                      seeds=[self.campaigns[i].seed for i in ids], pool_sizes=sizes, lengthscale_prior=lp,
                                      kernel=kernel, acquisition=acquisition, beta=beta)
                if kind != "ExactGP" and all(self.campaigns[i].rng_state is not None for i in ids):
                    batch._var.set_rng_states([self.campaigns[i].rng_state for i in ids])
            self.groups.append({"key": key, "ids": ids, "batch": batch, "stream": stream, "trials": max(trials, 0)})
        torch.cuda.synchronize(self.device)

    def run(self, on_group_done=None):
        """Advance every campaign to the end of its budget.  Returns {campaign index: (indices, status)}."""
        # enqueue: round-robin over the groups so that their launches interleave in submission order too
        progress = [0] * len(self.groups)
        pending = True
        while pending:
            pending = False
            for gi, gr in enumerate(self.groups):
                left = gr["trials"] - progress[gi]
                if left <= 0:
                    continue
                pending = True
                with torch.cuda.stream(gr["stream"]):
                    if gr["batch"].kind == "ExactGP":
                        k = min(self.chunk, left)
                        gr["batch"].run_queue(k)
                    else:
                        k = 1
                        gr["batch"].step()
                progress[gi] += k
        for gr in self.groups:
            gr["done"] = torch.cuda.Event()
            gr["done"].record(gr["stream"])
        out = {}
        remaining = list(range(len(self.groups)))
        while remaining:                                 # hand the groups over in completion order
            ready = [gi for gi in remaining if self.groups[gi]["done"].query()]
            if not ready:
                self.groups[remaining[0]]["done"].synchronize()
                ready = [remaining[0]]
            for gi in ready:
                remaining.remove(gi)
                gr = self.groups[gi]
                gr["stream"].synchronize()
                idx, n, status = gr["batch"].traces(sync=False)
                rows, counts, stat = idx.tolist(), n.tolist(), status.tolist()     # one conversion each, not one per element
                res = {}
                for row, c_id in enumerate(gr["ids"]):
                    res[c_id] = (rows[row][: counts[row]], int(stat[row]))
                out.update(res)
                if on_group_done is not None:
                    on_group_done(res)
        return out

    @property
    def launches(self) -> int:
        return sum(gr["batch"].launches for gr in self.groups)


class CampaignGroups:
    """The campaigns of one GPU split into contiguous groups, each a CampaignBatch on its own CUDA stream.

    A step of an ExactGP batch cannot end before its longest fit (38...449 L-BFGS-B closures), so a single launch
    over all campaigns leaves the SMs idle ~1/3 of the time; with independent groups the tail of one group's
    launch is filled by the others' (campaigns never interact, so no cross-group synchronisation is needed).
    Same ``step`` / ``run`` / ``traces`` surface as CampaignBatch."""

    def __init__(self, pools: torch.Tensor, pool_ys: torch.Tensor, pool_id: Sequence[int],
                 init_indices: Sequence[Sequence[int]], n_cap: int, kind: str = "ExactGP", groups: int = 4,
                 device: Optional[torch.device] = None, **kwargs):
        B = len(pool_id)
        groups = max(1, min(int(groups), B))
        self.bounds = [(g * B // groups, (g + 1) * B // groups) for g in range(groups)]
        self.members: List[CampaignBatch] = []
        self.streams: List[torch.cuda.Stream] = []
        seeds = kwargs.pop("seeds", None)
        for lo, hi in self.bounds:
            self.members.append(CampaignBatch(pools, pool_ys, list(pool_id[lo:hi]), list(init_indices[lo:hi]), n_cap,
                                              kind=kind, device=device,
                                              seeds=None if seeds is None else list(seeds[lo:hi]), **kwargs))
            self.streams.append(torch.cuda.Stream(device=self.members[-1].device))
        torch.cuda.synchronize(self.members[0].device)   # construction ran on the current stream

    def step(self) -> None:
        for m, s in zip(self.members, self.streams):
            with torch.cuda.stream(s):
                m.step()

    def run(self, n_trials: int) -> None:
        for _ in range(int(n_trials)):
            self.step()

    def synchronize(self) -> None:
        for s in self.streams:
            s.synchronize()

    def traces(self):
        self.synchronize()
        parts = [m.traces() for m in self.members]
        return tuple(torch.cat([p[i] for p in parts]) for i in range(3))

    @property
    def launches(self) -> int:
        return sum(m.launches for m in self.members)
