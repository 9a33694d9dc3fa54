"""Oracle (CPU, fp64): one BO campaign, restating ``run_single_loop`` (runners.py:14-115).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Also the timed CPU baseline of bench.py.
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Tuple

import torch

from stpbenchmark_b200.utils import get_initial_samples_sobol, set_seeds  # host-only helpers

from . import acq as oacq
from . import exact as oexact
from . import variational as ovar


def bo_iteration(kind: str, X_train: torch.Tensor, y_train: torch.Tensor, X_cand: torch.Tensor,
                 epochs: int = 600, lr: float = 0.01, eps: Optional[torch.Tensor] = None,
                 init_noise: Optional[torch.Tensor] = None, info: Optional[dict] = None) -> Tuple[int, torch.Tensor]:
    """The loop body runners.py:64-95: fresh model -> fit -> LogEI over the pool -> argmax.
    Returns (index into X_cand, acquisition values)."""
    if kind == "ExactGP":
        theta, res = oexact.fit_exact(X_train, y_train)
        mean, var = oexact.posterior_exact(theta, X_train, y_train, X_cand)
        acq = oacq.log_ei(mean, var, y_train.max())
        if info is not None:
            info.update(theta=theta, nit=res.nit, nfev=res.nfev, status=res.status, fun=res.fun,
                        mean=mean, var=var)
    else:
        st = ovar.train_natural(X_train, y_train, kind, epochs, lr, init_noise=init_noise)
        acq = ovar.acquisition(st, X_cand, y_train.max(), eps=eps)  # best_f RAW (App. D-2)
        if info is not None:
            info.update(state=st, loss=st.losses[-1] if st.losses else None)
    return oacq.first_argmax(acq), acq


def run_single_loop(X: torch.Tensor, y: torch.Tensor, kind: str, n_initial: int = 10, n_trials: int = 25,
                    epochs: int = 100, learning_rate: float = 0.1, seed: int = 42,
                    on_iteration: Optional[Callable] = None) -> Tuple[List[int], List[float]]:
    """One campaign: seed, Sobol initial design, then ``n_trials`` BO iterations.

    ``kind`` is one of ExactGP / VarGP / VarTGP / VarLGP.  Returns the selected pool indices
    (initial design first, duplicates kept) and their y values, like the reference."""
    set_seeds(seed)
    X = X.double()
    y = y.double().flatten()
    init = get_initial_samples_sobol(X, y, n_samples=n_initial, percentile=50)
    picked = init.tolist()
    values = y[init].tolist()
    mask = torch.ones(len(X), dtype=torch.bool)
    mask[init] = False
    X_train, y_train = X[picked], y[picked]
    n_trials = min(n_trials, len(X) - n_initial)
    for trial in range(n_trials):
        t0 = time.perf_counter()
        info = {} if on_iteration is not None else None
        X_cand = X[mask]
        best, _ = bo_iteration(kind, X_train, y_train, X_cand, epochs=epochs, lr=learning_rate, info=info)
        pool_idx = int(torch.where(mask)[0][best].item())
        picked.append(pool_idx)
        values.append(float(y[pool_idx].item()))
        X_train = torch.cat([X_train, X[pool_idx].unsqueeze(0)])
        y_train = torch.cat([y_train, y[pool_idx].unsqueeze(0)])
        mask[pool_idx] = False
        if on_iteration is not None:
            on_iteration(trial, pool_idx, time.perf_counter() - t0, info)
    return picked, values
