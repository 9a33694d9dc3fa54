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

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, priors


class CampaignBatch:
    """Device-resident state of B campaigns over G pools, for any of the four model kinds."""

    def __init__(self, pools: torch.Tensor, pool_ys: torch.Tensor, pool_id: Sequence[int],
                 init_indices: Sequence[Sequence[int]], n_cap: int, kind: str = "ExactGP",
                 device: Optional[torch.device] = None, epochs: int = 600, learning_rate: float = 0.01,
                 num_samples: int = 2048, seeds: Optional[Sequence[int]] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("CampaignBatch needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        if pools.dim() == 2:
            pools, pool_ys = pools.unsqueeze(0), pool_ys.reshape(1, -1)
        self.kind = kind
        self.G, self.N, self.d = pools.shape
        self.B = len(pool_id)
        self.cap = int(n_cap)
        max_n = _lib.MAX_N if kind == "ExactGP" else _lib.MAX_N_VAR
        if self.cap > max_n or self.d > _lib.MAX_D:
            raise ValueError(f"n_cap <= {max_n} and d <= {_lib.MAX_D} required for {kind}")
        dev = self.device
        self.pool = pools.to(dev, torch.double).contiguous()
        self.pool_y = pool_ys.to(dev, torch.double).contiguous()
        self.pool_id = torch.tensor(list(pool_id), dtype=torch.int32, device=dev)
        B, cap, d, N = self.B, self.cap, self.d, self.N
        # host-side assembly of the initial designs, then one H2D copy per array
        mask = torch.ones(B, N, dtype=torch.uint8)
        trace = torch.full((B, cap), -1, dtype=torch.int32)
        n = torch.zeros(B, dtype=torch.int32)
        for b, idx in enumerate(init_indices):
            idx = [int(i) for i in idx]
            if len(idx) > cap:
                raise ValueError("initial design longer than n_cap")
            t = torch.tensor(idx, dtype=torch.long)
            mask[b, t] = 0
            trace[b, : len(idx)] = t.to(torch.int32)
            n[b] = len(idx)
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
            _lib.exact_bo_step(self.pool, self.pool_y, self.pool_id, self.mask, self.X, self.y, self.n, self.trace,
                               self._theta0, self._lower, self._upper, self._prior, self.status,
                               theta_out=self.theta, acq_out=self.acq, nit=self.nit, nfev=self.nfev,
                               opts=self._opts, order=order, turnover=self._turnover, work_acc=self.work_acc,
                               iters_acc=self.iters_acc, n_max=n_max)
            self.launches += 1
            if self._n_host is not None:          # what the kernel does to n: +1, and the turnover's restart
                self._n_host += 1
                if self._turnover is not None:
                    self._n_host[self._n_host >= self._turnover.n_final] = self._turnover.n_init
        else:
            self.launches += self._var.step()
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
                          max_ctas=max_ctas)
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

    def traces(self):
        """(indices [B, cap] int64 with -1 beyond the last pick, n [B], status [B]) on the host."""
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
