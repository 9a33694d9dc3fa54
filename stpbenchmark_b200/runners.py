"""Campaign loop with the reference's signatures (runners.py:14-202), batched on the GPU.

``run_single_loop``  one campaign (runners.py:14-115): same arguments, same return value
                     ``(selected_indices, selected_values)``, never raises.
``run_many_loops``   all requested seeds (runners.py:118-202): same CSV layout (two-level
                     columns ``(seed_<s>, x_index|y_value)``, index ``iteration``), same resume rule
                     (seeds already present as columns are skipped), NaN padding of unfinished
                     traces.  The seeds are advanced TOGETHER as one device batch instead of one
                     after another.
``run_sweep``        what run_benchmark.py's nested loops do (run_benchmark.py:20-43): SEVERAL
                     ``run_many_loops`` jobs -- model x dataset -- submitted at once, so that every
                     (model, dataset, seed) campaign of the sweep shares the GPU (engine.CampaignJob) and, under
                     ``torch.distributed``, the GPUs of the box (contiguous shards, one final gather of the traces).
                     A job's CSV is written as soon as ITS campaigns have finished.

Deliberate deviations from the reference (documented in INTEGRATION.md):
  * with ``invert_y=True`` a resumed run no longer re-inverts the y columns of previously saved
    seeds (runners.py:155 x :193); ``x_index`` is unaffected;
  * a seed counts as done only if its column holds at least one row (an all-NaN column written by a
    crashed run is recomputed), and a failure never writes columns for seeds that produced nothing;
  * ``model_class`` must be one of the four stock classes and ``model_kwargs`` may only carry what their
    constructors take (``lengthscale_prior`` is honoured; anything else is an error, as
    ``model_class(**model_kwargs)`` would be at runners.py:74) -- the arithmetic runs in fused kernels, so
    a subclass override could not be honoured and is rejected instead of silently ignored.
"""
from __future__ import annotations

import inspect
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Type, Union

import numpy as np
import pandas as pd
import torch

from . import distributed as sdist
from .engine import Campaign, CampaignJob
from .models import ExactGP, VarGP, VarLGP, VarTGP
from .utils import initial_designs_batch

_KIND = {ExactGP: "ExactGP", VarGP: "VarGP", VarLGP: "VarLGP", VarTGP: "VarTGP"}
_SET_BY_THE_LOOP = {"ExactGP": ("X_train", "y_train", "input_transform"), "VarGP": ("inducing_points",),
                    "VarLGP": ("inducing_points",), "VarTGP": ("inducing_points",)}


def _kind_of(model_class) -> str:
    """Model kind of a stock class (or its name).  Subclasses are rejected: their overrides (forward, priors,
    likelihood ...) cannot reach the fused kernels, and ignoring them silently would change results."""
    if isinstance(model_class, str):
        if model_class in _KIND.values():
            return model_class
        raise TypeError(f"unsupported model kind {model_class!r}")
    for cls, name in _KIND.items():
        if model_class is cls:
            return name
    if isinstance(model_class, type) and any(issubclass(model_class, cls) for cls in _KIND):
        raise NotImplementedError(
            f"{model_class.__name__} subclasses a stock model; the batched engine runs the stock arithmetic in fused "
            "kernels and cannot honour overrides -- pass ExactGP / VarGP / VarLGP / VarTGP themselves")
    raise TypeError(f"unsupported model class {model_class!r}")


def _lengthscale_prior_of(kind: str, model_class, model_kwargs: Optional[dict]):
    """What ``model_class(**model_kwargs)`` (runners.py:66-74) would have taken from ``model_kwargs``: the keys the
    loop sets itself are overwritten there and ignored here; ``lengthscale_prior`` (models.py:30, 125, 217) is
    returned as (loc, scale); any other key raises like the constructor call would."""
    if not model_kwargs:
        return None
    accepted = set()
    if not isinstance(model_class, str):
        accepted = set(inspect.signature(model_class.__init__).parameters) - {"self"}
    prior = None
    for k, v in model_kwargs.items():
        if k in _SET_BY_THE_LOOP[kind]:
            continue
        if k == "lengthscale_prior" and kind != "ExactGP":
            if v is None:
                continue
            prior = (float(v[0]), float(v[1])) if isinstance(v, (tuple, list)) else (float(v.loc), float(v.scale))
            continue
        if k in accepted:
            raise NotImplementedError(f"model_kwargs[{k!r}] is not supported by the batched engine")
        raise TypeError(f"{kind}.__init__() got an unexpected keyword argument {k!r}")
    return prior


def initial_design(X: torch.Tensor, y: torch.Tensor, seed: int, n_initial: int):
    """runners.py:33, 46-48: seed everything, then the Sobol initial design.  Returns the indices
    and the state of the global torch generator right after (what the variational models'
    first-forward noise is drawn from, App. A.6)."""
    idx, states = initial_designs_batch(X, y, [int(seed)], n_initial, 50.0, return_rng_states=True)
    return idx[0].tolist(), states[0]


# ---------------------------------------------------------------------------------------------
# the engine entry point: many campaigns, many pools, one call
# ---------------------------------------------------------------------------------------------
def _execute_on_device(pools, campaigns: List[Campaign], on_done: Optional[Callable] = None):
    """{campaign index: (indices, status)} for the given campaigns on the current CUDA device."""
    if not campaigns:
        return {}
    job = CampaignJob(pools, campaigns)
    return job.run(on_group_done=on_done)


def _shard_order(campaigns: List[Campaign]) -> List[int]:
    """Campaign order used for sharding: what shares launches (kind, pool, budget) stays together, so that a rank's
    contiguous block forms few, full groups."""
    return sorted(range(len(campaigns)), key=lambda i: (campaigns[i].kind, campaigns[i].pool, campaigns[i].n_trials, i))


def run_campaigns(pools, campaigns: List[Campaign], on_done: Optional[Callable] = None, _executor=None,
                  prepare: Optional[Callable] = None, on_local: Optional[Callable] = None):
    """Run every campaign; under an initialised ``torch.distributed`` group the campaign list is sharded in
    contiguous blocks over the ranks (one process per GPU, no data-path collective) and the traces are gathered
    on rank 0 at the end.  Returns {campaign index: (indices, status)} on rank 0 (all campaigns)
    and on the other ranks (their own shard).  ``on_done`` receives finished groups on the rank that ran them, or,
    with several ranks, the complete result once on rank 0.  ``prepare(ids)`` is called with the indices this rank
    owns before they run (run_sweep computes the initial designs of its own shard only).  ``on_local(res)`` (several
    ranks only) is called on EVERY rank with the results of its own shard before the gather: run_sweep uses it to write
    the CSV files of the jobs a rank ran completely on that rank, in parallel, instead of all of them on rank 0."""
    import torch.distributed as dist
    execute = _executor or _execute_on_device
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        if prepare is not None:
            prepare(list(range(len(campaigns))))
        return execute(pools, campaigns, on_done)
    rank = dist.get_rank()
    order = _shard_order(campaigns)
    lo, hi = sdist.shard_bounds(len(order), world, rank)
    mine = order[lo:hi]
    if prepare is not None:
        prepare(mine)
    local = execute(pools, [campaigns[i] for i in mine], None)
    if on_local is not None:
        on_local({mine[row]: local[row] for row in range(len(mine))})
    T = max((c.n_initial + c.n_trials for c in campaigns), default=0) + 1
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    tr = torch.full((len(mine), T), -1, dtype=torch.int32)
    st = torch.full((len(mine), T), float("nan"), dtype=torch.double)   # column 0 carries the status
    for row in range(len(mine)):
        idx, status = local[row]
        tr[row, : len(idx)] = torch.tensor(idx, dtype=torch.int32)
        st[row, 0] = float(status)
    got = sdist.gather_traces(tr.to(dev), st.to(dev), len(order))
    if got is None:
        return {mine[row]: local[row] for row in range(len(mine))}
    gt, gs = got[0].cpu(), got[1].cpu()
    rows, counts, stat = gt.tolist(), (gt >= 0).sum(dim=1).tolist(), gs[:, 0].tolist()   # one conversion each: 8192 campaigns at N = 8
    out = {}
    for pos, c_id in enumerate(order):
        out[c_id] = (rows[pos][: counts[pos]], int(stat[pos]))
    if on_done is not None:
        on_done(out)
    return out


def run_campaign_batch(X: torch.Tensor, y: torch.Tensor, model_class, seeds: Sequence[int], n_initial: int = 10,
                       n_trials: int = 25, epochs: int = 100, learning_rate: float = 0.1,
                       num_samples: int = 2048, model_kwargs: Optional[dict] = None, kernel: str = "rbf",
                       acquisition: str = "logei", beta: float = 0.0
                       ) -> Tuple[List[List[int]], List[List[float]], List[int]]:
    """Advance one campaign per seed on the same pool, all together.  Returns
    (indices per seed, values per seed, status per seed).  ``kernel`` ("rbf" | "matern52", ExactGP only) and
    ``acquisition`` ("logei" | "ucb" | "ei", with UCB's ``beta``) select the variants of SURVEY section 8 row f4;
    the defaults are what the reference's loop does (models.py:342-358, runners.py:88-90)."""
    X = X.double()
    y = y.double().flatten()
    kind = _kind_of(model_class)
    prior = _lengthscale_prior_of(kind, model_class, model_kwargs)
    n_trials = min(int(n_trials), len(X) - int(n_initial))  # runners.py:62
    seeds = [int(s) for s in seeds]
    dev = "cuda" if torch.cuda.is_available() else None
    designs, states = initial_designs_batch(X, y, seeds, n_initial, 50.0, return_rng_states=True, device=dev)
    camps = [Campaign(s, kind, 0, designs[k].tolist(), n_trials, epochs, learning_rate, num_samples, prior,
                      states[k], seed=s, kernel=kernel, acquisition=acquisition, beta=beta)
             for k, s in enumerate(seeds)]
    res = run_campaigns([(X, y)], camps)
    y_host = y.cpu()
    out_idx = [res[k][0] for k in range(len(seeds))]
    out_val = [[float(y_host[i]) for i in row] for row in out_idx]
    return out_idx, out_val, [res[k][1] for k in range(len(seeds))]


def run_single_loop(
    X: torch.Tensor,
    y: torch.Tensor,
    model_class: Type[Union[VarTGP, VarGP, ExactGP]],
    model_kwargs: Optional[dict] = None,
    n_initial: int = 10,
    n_trials: int = 25,
    epochs: int = 100,
    learning_rate: float = 0.1,
    seed: int = 42,
) -> Tuple[List[int], List[float]]:
    """One BO campaign on (X, y); drop-in for runners.py:14-115 (a batch of one).  Like the reference it never
    raises: whatever was selected before a failure is returned (at least the initial design once that exists)."""
    selected_indices: List[int] = []
    selected_values: List[float] = []
    try:
        Xd, yd = X.double(), y.double().flatten()
        init, _ = initial_design(Xd, yd, seed, n_initial)      # runners.py:46-52: recorded before anything can fail
        selected_indices = list(init)
        selected_values = [float(yd[i]) for i in init]
        idx, val, status = run_campaign_batch(Xd, yd, model_class, [seed], n_initial=n_initial, n_trials=n_trials,
                                              epochs=epochs, learning_rate=learning_rate, model_kwargs=model_kwargs)
        selected_indices, selected_values = idx[0], val[0]
        if status[0] != 0:
            print(f"Error occurred at iteration {len(selected_values)}: campaign stopped with status {status[0]}")
        if isinstance(model_kwargs, dict):   # runners.py:67-72 leaves the last training set in the caller's dict
            kind = _kind_of(model_class)
            Xt, yt = Xd[selected_indices[:-1]] if len(idx[0]) > n_initial else Xd[selected_indices], None
            if kind == "ExactGP":
                yt = yd[selected_indices[:-1]] if len(idx[0]) > n_initial else yd[selected_indices]
                model_kwargs.update(X_train=Xt, y_train=yt, input_transform=None)
            else:
                model_kwargs["inducing_points"] = Xt
        return selected_indices, selected_values
    except Exception as e:  # same contract as runners.py:113-115
        print(f"Error occurred at iteration {len(selected_values)}: {str(e)}")
        return selected_indices, selected_values


# ---------------------------------------------------------------------------------------------
# CSV layer (runners.py:141-202)
# ---------------------------------------------------------------------------------------------
def _empty_frame(seeds, total_iterations) -> pd.DataFrame:
    columns = pd.MultiIndex.from_tuples(
        [(f"seed_{seed}", metric) for seed in seeds for metric in ["x_index", "y_value"]])
    frame = pd.DataFrame(index=range(total_iterations), columns=columns)
    frame.index.name = "iteration"
    return frame


def _save(frame: pd.DataFrame, path: str, invert_y: bool) -> None:
    """Write the frame; columns of seeds that hold nothing are left out (a later run recomputes them)."""
    save_df = frame.copy()
    x = save_df.xs("x_index", axis=1, level=1)
    empty = list(x.columns[x.isna().all().values])
    if empty:
        save_df = save_df.drop(columns=empty, level=0)
    if invert_y:
        y_cols = [c for c in save_df.columns if c[1] == "y_value"]
        save_df[y_cols] = 1.0 / save_df[y_cols].astype(float)
    tmp = path + ".tmp"
    text = _csv_text(save_df)
    if text is None:
        save_df.to_csv(tmp)
    else:
        with open(tmp, "w", newline="") as fh:
            fh.write(text)
    os.replace(tmp, path)           # a crash mid-write never truncates what was saved before


def _csv_text(frame: pd.DataFrame) -> Optional[str]:
    """What ``frame.to_csv()`` writes for a trace frame (two-level columns, index ``iteration``, int64 / float64
    columns, NaN as the empty string, floats in their shortest round-trip form), assembled directly: pandas spends
    ~15 ms per file on block management for the 2 x 128 columns of a config #4 job, eight files per job.  Returns
    None for anything else (object columns after a resume from an older file, ...): the caller then uses pandas.
    Byte-for-byte equality with pandas is checked in tests/test_runners_host.py."""
    if frame.columns.nlevels != 2 or frame.index.name != "iteration" or len(frame.columns) == 0:
        return None
    dt = frame.dtypes.to_numpy()
    int_pos = np.flatnonzero(dt == np.dtype(np.int64))
    flt_pos = np.flatnonzero(dt == np.dtype(np.float64))
    if len(int_pos) + len(flt_pos) != len(dt):
        return None
    cols: List[Optional[list]] = [None] * len(dt)
    # one positional take per dtype (a label lookup per column costs 0.1 ms on a two-level index: 25 ms per file)
    # A trace frame repeats few distinct values (pool indices, the y of a pool point): format every distinct value once
    # and gather (92 160 repr() calls per config #4 job otherwise, half of the CSV time).  Floats are told apart by
    # their bit patterns, so that -0.0 / 0.0 keep their own spelling.
    if len(int_pos):
        block = np.ascontiguousarray(frame.iloc[:, int_pos].to_numpy().T)
        lo_v, hi_v = int(block.min()), int(block.max())
        if 0 <= lo_v and hi_v < (1 << 20):
            table = np.array([str(v) for v in range(hi_v + 1)], dtype=object)
            for c, col in zip(int_pos.tolist(), table[block].tolist()):
                cols[c] = col
        else:
            for c, col in zip(int_pos.tolist(), block.tolist()):
                cols[c] = list(map(str, col))
    if len(flt_pos):
        block = np.ascontiguousarray(frame.iloc[:, flt_pos].to_numpy().T)
        uniq, inv = np.unique(block.view(np.int64).ravel(), return_inverse=True)
        table = np.array([repr(x) if x == x else "" for x in uniq.view(np.float64).tolist()], dtype=object)
        for c, col in zip(flt_pos.tolist(), table[inv.reshape(block.shape)].tolist()):
            cols[c] = col
    idx = frame.index.to_numpy()
    if idx.dtype != np.int64:
        return None
    head0 = "," + ",".join(str(c[0]) for c in frame.columns)
    head1 = "," + ",".join(str(c[1]) for c in frame.columns)
    head2 = "iteration" + "," * len(cols)
    rows = [",".join(r) for r in zip([str(i) for i in idx.tolist()], *cols)]
    return "\n".join([head0, head1, head2] + rows) + "\n"


class _CsvJob:
    """One ``run_many_loops`` call inside a sweep: its frame, its resume state and its pending seeds."""

    def __init__(self, X, y, model_class, seeds, output_path, invert_y=False, **kwargs):
        self.X, self.y = X.double(), y.double().flatten()
        self.model_class, self.output_path, self.invert_y = model_class, output_path, bool(invert_y)
        self.kwargs = dict(kwargs)
        self.n_initial = int(kwargs.get("n_initial", 10))
        self.n_trials = int(kwargs.get("n_trials", 25))
        self.total = self.n_initial + self.n_trials
        self.seeds = [int(s) for s in seeds]
        out_dir = os.path.dirname(output_path)
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
        if os.path.exists(output_path):
            self._frame = pd.read_csv(output_path, header=[0, 1], index_col=0)
            x = self.frame.xs("x_index", axis=1, level=1)
            done = {int(col.split("_")[1]) for col in x.columns[~x.isna().all().values]}
            if self.invert_y:  # the file holds 1/y: undo it so that saving is idempotent (App. B.3)
                y_cols = [c for c in self.frame.columns if c[1] == "y_value"]
                self.frame[y_cols] = 1.0 / self.frame[y_cols].astype(float)
        else:
            self._frame = None             # built on demand (an all-NaN frame of 2 x len(seeds) object columns costs ~12 ms)
            done = set()
        self.remaining = [s for s in self.seeds if s not in done]

    @property
    def frame(self) -> pd.DataFrame:
        if self._frame is None:
            self._frame = _empty_frame(self.seeds, self.total)
        return self._frame

    @frame.setter
    def frame(self, value: pd.DataFrame) -> None:
        self._frame = value

    def record_many(self, items: Sequence[Tuple[int, List[int]]]) -> None:
        """Traces of several finished seeds at once (runners.py:181-186, NaN-padded to n_initial + n_trials rows)."""
        y_host = self.y.cpu().numpy()
        if (self._frame is None and len(items) == len(self.seeds) and [s for s, _ in items] == self.seeds
                and all(len(ix) >= self.total for _, ix in items)):
            # a fresh job whose seeds all finished at once with full traces (the batched engine's normal case): two
            # homogeneous blocks (int64 indices, float64 values) instead of 2 x len(seeds) single-column ones
            xi = np.asarray([ix[: self.total] for _, ix in items], dtype=np.int64).T
            S = len(self.seeds)
            names = np.repeat(np.array([f"seed_{s}" for s in self.seeds], dtype=object), 2)
            both = pd.concat([pd.DataFrame(np.ascontiguousarray(xi)), pd.DataFrame(np.ascontiguousarray(y_host[xi]))], axis=1)
            frame = both.iloc[:, np.arange(2 * S).reshape(2, S).T.reshape(-1)]          # x_0, y_0, x_1, y_1, ... (positional)
            # = MultiIndex.from_arrays([names, x_index / y_value alternating]) without its factorisation pass (2 ms per job)
            labels = [f"seed_{s}" for s in self.seeds]
            if len(set(labels)) == S:
                level0 = sorted(labels)
                pos = {lab: k for k, lab in enumerate(level0)}
                frame.columns = pd.MultiIndex(levels=[level0, ["x_index", "y_value"]],
                                              codes=[np.repeat(np.array([pos[lab] for lab in labels]), 2), np.tile(np.array([0, 1]), S)],
                                              verify_integrity=False)
            else:
                frame.columns = pd.MultiIndex.from_arrays([names, np.tile(np.array(["x_index", "y_value"], dtype=object), S)])
            frame.index.name = "iteration"
            self._frame = frame
            return
        cols = {}
        for seed, indices in items:
            k = min(len(indices), self.total)
            picked = np.asarray(indices[:k], dtype=np.int64)
            if k == self.total:
                idx = picked
            else:
                idx = np.full(self.total, np.nan)
                idx[:k] = picked
            val = np.full(self.total, np.nan)
            val[:k] = y_host[picked]
            cols[(f"seed_{seed}", "x_index")] = idx
            cols[(f"seed_{seed}", "y_value")] = val
        new = pd.DataFrame(cols, index=self.frame.index)
        new.columns = pd.MultiIndex.from_tuples(list(cols))
        old = set(self.frame.columns)
        present = [c for c in new.columns if c in old]
        if len(present) == len(old):                 # every column of the frame is replaced (a fresh run finishing at once)
            self.frame = new[list(self.frame.columns)]
        elif present:
            keep = [c for c in self.frame.columns if c not in set(present)]
            order = list(self.frame.columns)
            self.frame = pd.concat([self.frame[keep], new[present]], axis=1)[order]
        extra = [c for c in new.columns if c not in old]
        if extra:
            self.frame = pd.concat([self.frame, new[extra]], axis=1)
        self.frame.index.name = "iteration"

    def record(self, seed: int, indices: List[int]) -> None:
        self.record_many([(seed, indices)])

    def save(self) -> None:
        _save(self.frame, self.output_path, self.invert_y)


def run_sweep(jobs: Sequence[dict], _executor=None) -> List[pd.DataFrame]:
    """Run several ``run_many_loops`` jobs (dicts of its arguments: X, y, model_class, seeds, output_path, invert_y,
    n_initial, n_trials, epochs, learning_rate, model_kwargs) as ONE device job.  Every campaign of every job shares
    the GPU(s); a job's CSV is (re)written whenever a group of its campaigns finishes.  Returns the frames."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    csv_jobs, pools, camps, owner, job_of_pool = [], [], [], [], []
    on_cuda = torch.cuda.is_available() and _executor is None
    for j, spec in enumerate(jobs):
        spec = dict(spec)
        job = _CsvJob(spec.pop("X"), spec.pop("y"), spec.pop("model_class"), spec.pop("seeds"), spec.pop("output_path"),
                      spec.pop("invert_y", False), **spec)
        csv_jobs.append(job)
        if not job.remaining:
            print("     [INFO] All requested seeded runs already completed.")
            continue
        try:
            kind = _kind_of(job.model_class)
            prior = _lengthscale_prior_of(kind, job.model_class, job.kwargs.get("model_kwargs"))
        except Exception as e:
            print(f"Error in batch of seeds {job.remaining}: {str(e)}")
            continue
        pools.append((job.X, job.y))
        job_of_pool.append(j)
        trials = min(job.n_trials, len(job.X) - job.n_initial)
        for s in job.remaining:   # the initial design is filled in by the rank that owns the campaign (`prepare`)
            camps.append(Campaign((j, s), kind, len(pools) - 1, None, trials,
                                  int(job.kwargs.get("epochs", 100)), float(job.kwargs.get("learning_rate", 0.1)),
                                  int(job.kwargs.get("num_samples", 2048)), prior, None, seed=s,
                                  n_initial=min(job.n_initial, len(job.X))))
            owner.append((j, s))

    def prepare(ids: List[int]):
        by_pool: Dict[int, List[int]] = {}
        for i in ids:
            by_pool.setdefault(camps[i].pool, []).append(i)
        for pool_id, members in by_pool.items():
            job = csv_jobs[job_of_pool[pool_id]]
            # the generator state after the draw is what a variational model's first-forward noise continues from
            # (App. A.6); the exact GP never draws again
            need_states = any(camps[i].kind != "ExactGP" for i in members)
            res = initial_designs_batch(job.X, job.y, [camps[i].seed for i in members], job.n_initial, 50.0,
                                        return_rng_states=need_states, device="cuda" if on_cuda else None)
            designs, states = res if need_states else (res, None)
            rows = designs.tolist()
            for k, i in enumerate(members):
                camps[i].design = rows[k]
                camps[i].rng_state = states[k] if need_states else None

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    # With several ranks a job whose pending campaigns all fall into ONE rank's shard is written by that rank (the
    # files of a sweep are then written in parallel); rank 0 still records every trace for the frames it returns.
    whole_on: Dict[int, int] = {}
    if world > 1 and camps:
        order = _shard_order(camps)
        rank_of = {}
        for r in range(world):
            lo, hi = sdist.shard_bounds(len(order), world, r)
            for c_id in order[lo:hi]:
                rank_of[c_id] = r
        ranks_of_job: Dict[int, set] = {}
        for c_id, (j, _) in enumerate(owner):
            ranks_of_job.setdefault(j, set()).add(rank_of[c_id])
        whole_on = {j: next(iter(rs)) for j, rs in ranks_of_job.items() if len(rs) == 1}

    def record(res: Dict[int, tuple], save_if: Callable[[int], bool], verbose: bool):
        per_job: Dict[int, list] = {}
        for c_id, (indices, status) in res.items():
            j, s = owner[c_id]
            if status != 0 and verbose:
                print(f"Error occurred at iteration {len(indices)}: campaign (job {j}, seed {s}) stopped with status {status}")
            per_job.setdefault(j, []).append((s, indices))
        for j, items in per_job.items():
            if world > 1:      # the gather returns the campaigns in shard order: back to the job's seed order
                pos = {s: k for k, s in enumerate(csv_jobs[j].remaining)}
                items.sort(key=lambda it: pos.get(it[0], len(pos)))
            csv_jobs[j].record_many(items)
            if save_if(j):
                csv_jobs[j].save()

    def on_local(res: Dict[int, tuple]):
        if rank == 0:
            return             # rank 0 records its own shard with everybody else's in on_done
        mine = {c_id: v for c_id, v in res.items() if whole_on.get(owner[c_id][0]) == rank}
        record(mine, lambda j: True, verbose=False)

    def on_done(res: Dict[int, tuple]):
        if rank != 0:
            return
        record(res, lambda j: whole_on.get(j, 0) == 0, verbose=True)

    if camps:
        try:
            run_campaigns(pools, camps, on_done=on_done, _executor=_executor, prepare=prepare,
                          on_local=on_local if world > 1 else None)
        except Exception as e:
            print(f"Error in sweep: {str(e)}")
            if rank == 0:
                for job in csv_jobs:     # what finished before the failure is already saved; keep files consistent
                    if os.path.exists(job.output_path) or not job.frame.isna().all().all():
                        job.save()
    return [job.frame for job in csv_jobs]


def run_many_loops(
    X: torch.Tensor,
    y: torch.Tensor,
    model_class: Type[Union[VarTGP, VarGP, ExactGP]],
    seeds: List[int],
    output_path: str,
    invert_y: bool = False,
    **kwargs,
) -> pd.DataFrame:
    """Run one campaign per seed and write the traces CSV; drop-in for runners.py:118-202."""
    return run_sweep([dict(X=X, y=y, model_class=model_class, seeds=seeds, output_path=output_path, invert_y=invert_y,
                           **kwargs)])[0]
