"""Campaign loop with the reference's signatures (runners.py:14-202), batched on the GPU.

``run_single_loop``  one campaign (runners.py:14-115): same arguments, same return value
                     ``(selected_indices, selected_values)``, never raises.
``run_many_loops``   all requested seeds (runners.py:118-202): same CSV layout (two-level
                     columns ``(seed_<s>, x_index|y_value)``, index ``iteration``), same resume rule
                     (seeds already present as columns are skipped), NaN padding of unfinished
                     traces.  The seeds are advanced TOGETHER as one device batch instead of one
                     after another.  One deliberate deviation: with ``invert_y=True`` a resumed run
                     no longer re-inverts the y columns of previously saved seeds
                     (runners.py:155 x :193); ``x_index`` is unaffected.
``run_campaign_batch`` the engine entry point both of them use.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple, Type, Union

import numpy as np
import pandas as pd
import torch

from .engine import CampaignBatch
from .models import ExactGP, VarGP, VarLGP, VarTGP
from .utils import get_initial_samples_sobol, set_seeds

_KIND = {ExactGP: "ExactGP", VarGP: "VarGP", VarLGP: "VarLGP", VarTGP: "VarTGP"}


def _kind_of(model_class) -> str:
    if isinstance(model_class, str):
        return model_class
    for cls, name in _KIND.items():
        if model_class is cls or (isinstance(model_class, type) and issubclass(model_class, cls)):
            return name
    raise TypeError(f"unsupported model class {model_class!r}")


def initial_design(X: torch.Tensor, y: torch.Tensor, seed: int, n_initial: int):
    """runners.py:33, 46-48: seed everything, then the Sobol initial design.  Returns the indices
    and the state of the global torch generator right after (what the variational models'
    first-forward noise is drawn from, App. A.6)."""
    set_seeds(seed)
    idx = get_initial_samples_sobol(X, y, n_samples=n_initial, percentile=50)
    return idx.tolist(), torch.get_rng_state()


def run_campaign_batch(X: torch.Tensor, y: torch.Tensor, model_class, seeds: Sequence[int], n_initial: int = 10,
                       n_trials: int = 25, epochs: int = 100, learning_rate: float = 0.1,
                       num_samples: int = 2048) -> Tuple[List[List[int]], List[List[float]], List[int]]:
    """Advance one campaign per seed on the same pool, all together.  Returns
    (indices per seed, values per seed, status per seed)."""
    X = X.double()
    y = y.double().flatten()
    kind = _kind_of(model_class)
    n_trials = min(int(n_trials), len(X) - int(n_initial))  # runners.py:62
    designs, rng_states = [], []
    for s in seeds:
        idx, state = initial_design(X, y, int(s), n_initial)
        designs.append(idx)
        rng_states.append(state)
    cap = max(len(d_) for d_ in designs) + n_trials + 1
    batch = CampaignBatch(X, y, [0] * len(designs), designs, n_cap=cap, kind=kind, epochs=epochs,
                          learning_rate=learning_rate, num_samples=num_samples, seeds=[int(s) for s in seeds])
    if kind != "ExactGP":
        batch._var.set_rng_states(rng_states)
    batch.run(n_trials)
    idx, n, status = batch.traces()
    y_host = y.cpu()
    out_idx = [[int(v) for v in idx[b, : int(n[b])]] for b in range(len(designs))]
    out_val = [[float(y_host[i]) for i in row] for row in out_idx]
    return out_idx, out_val, [int(s) for s in status]


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
    """One BO campaign on (X, y); drop-in for runners.py:14-115 (a batch of one)."""
    try:
        idx, val, status = run_campaign_batch(X, y, model_class, [seed], n_initial=n_initial, n_trials=n_trials,
                                              epochs=epochs, learning_rate=learning_rate)
        if status[0] != 0:
            print(f"Error occurred at iteration {len(val[0])}: campaign stopped with status {status[0]}")
        return idx[0], val[0]
    except Exception as e:  # same contract as runners.py:113-115
        print(f"Error occurred at iteration 0: {str(e)}")
        return [], []


def _empty_frame(seeds, total_iterations) -> pd.DataFrame:
    columns = pd.MultiIndex.from_tuples(
        [(f"seed_{seed}", metric) for seed in seeds for metric in ["x_index", "y_value"]])
    frame = pd.DataFrame(index=range(total_iterations), columns=columns)
    frame.index.name = "iteration"
    return frame


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
    n_initial = kwargs.get("n_initial", 10)
    n_trials = kwargs.get("n_trials", 25)
    total_iterations = n_initial + n_trials
    out_dir = os.path.dirname(output_path)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    seeds = [int(s) for s in seeds]

    if os.path.exists(output_path):
        results_df = pd.read_csv(output_path, header=[0, 1], index_col=0)
        done = {int(col.split("_")[1]) for col in results_df.columns.get_level_values(0)}
        remaining = [s for s in seeds if s not in done]
        if invert_y:  # the file holds 1/y: undo it so that saving is idempotent (App. B.3)
            y_cols = [c for c in results_df.columns if c[1] == "y_value"]
            results_df[y_cols] = 1.0 / results_df[y_cols]
    else:
        remaining = seeds
        results_df = _empty_frame(seeds, total_iterations)

    if len(remaining) == 0:
        print("     [INFO] All requested seeded runs already completed.")
        return results_df

    try:
        idx, val, _ = run_campaign_batch(
            X, y, model_class, remaining, n_initial=n_initial, n_trials=n_trials,
            epochs=kwargs.get("epochs", 100), learning_rate=kwargs.get("learning_rate", 0.1))
    except Exception as e:
        print(f"Error in batch of seeds {remaining}: {str(e)}")
        results_df.to_csv(output_path)
        return results_df

    for seed, indices, values in zip(remaining, idx, val):
        indices = list(indices) + [np.nan] * (total_iterations - len(indices))  # runners.py:181-182
        values = list(values) + [np.nan] * (total_iterations - len(values))
        results_df[f"seed_{seed}", "x_index"] = indices[:total_iterations]
        results_df[f"seed_{seed}", "y_value"] = values[:total_iterations]
    save_df = results_df.copy()
    if invert_y:
        y_cols = [c for c in save_df.columns if c[1] == "y_value"]
        save_df[y_cols] = 1.0 / save_df[y_cols].astype(float)
    save_df.to_csv(output_path)
    return results_df
