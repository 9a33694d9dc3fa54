"""Host-side utilities of the BO hot path (seeding, scalers, data prep, initial design).

Mirrors the public names and the numerics of the reference's ``utils.py`` (row a11 of
SURVEY.md section 8): ``set_seeds`` (utils.py:11-16), ``TorchStandardScaler`` (utils.py:19-78),
``TorchNormalizer`` (utils.py:81-133), ``preprocess_data`` (utils.py:154-202),
``get_initial_samples`` (utils.py:205-236), ``get_initial_samples_sobol`` (utils.py:239-281)
and ``EI`` (utils.py:299-309).  These stay on the host on purpose: the scrambled Sobol draw
consumes the *global* torch CPU generator and the golden traces pin its output bit for bit
(tests/test_host_utils.py checks all 250 golden initial designs).

Nothing here touches CUDA, so the module imports on a CPU-only box.
"""
from __future__ import annotations

import math
import random
from typing import Tuple

import numpy as np
import torch
from torch.quasirandom import SobolEngine

__all__ = [
    "set_seeds",
    "TorchStandardScaler",
    "TorchNormalizer",
    "preprocess_data",
    "get_initial_samples",
    "get_initial_samples_sobol",
    "get_device",
    "EI",
]


def set_seeds(seed) -> None:
    """Seed numpy, torch (CPU and every CUDA device) and ``random`` (utils.py:11-16).

    The order is part of the contract: the torch CPU generator state after this call is
    what the Sobol scramble and the variational initial noise draw from.
    """
    s = int(seed)
    np.random.seed(s)
    torch.manual_seed(s)
    torch.cuda.manual_seed_all(s)
    random.seed(s)


class TorchStandardScaler:
    """z = (x - mean) / (population_std + 1e-10) along dim 0 (utils.py:19-78)."""

    _EPS = 1e-10

    def fit(self, x: torch.Tensor) -> None:
        self.mean = x.mean(0, keepdim=True)
        self.std = x.std(0, unbiased=False, keepdim=True)

    def transform(self, x: torch.Tensor) -> torch.Tensor:
        out = x.clone()
        out -= self.mean
        out /= self.std + self._EPS
        return out

    def fit_transform(self, x: torch.Tensor) -> torch.Tensor:
        self.fit(x)
        return self.transform(x)

    def inverse_transform(self, x: torch.Tensor) -> torch.Tensor:
        out = x.clone()
        out *= self.std + self._EPS
        out += self.mean
        return out


class TorchNormalizer:
    """Min-max scaling of every column to [0, 1] (utils.py:81-133)."""

    def fit(self, x: torch.Tensor) -> None:
        self.max = torch.max(x, dim=0).values
        self.min = torch.min(x, dim=0).values

    def transform(self, x: torch.Tensor) -> torch.Tensor:
        return (x.clone() - self.min) / (self.max - self.min)

    def fit_transform(self, x: torch.Tensor) -> torch.Tensor:
        self.fit(x)
        return self.transform(x)

    def inverse_transform(self, x: torch.Tensor) -> torch.Tensor:
        out = x.clone()
        out *= self.max - self.min
        out += self.min
        return out


def preprocess_array(data: np.ndarray) -> Tuple[torch.Tensor, torch.Tensor]:
    """Array form of :func:`preprocess_data`: last column is the target.

    Rows that share a feature vector are merged and their targets averaged
    (utils.py:187-190); features are then min-max normalised (utils.py:200).
    ``np.unique(axis=0)`` sorts the rows lexicographically, which fixes the pool order the
    golden ``x_index`` columns refer to.
    """
    if data.size == 0:
        raise ValueError("Empty dataset")
    feats, target = data[:, :-1], data[:, -1]
    uniq, inverse = np.unique(feats, axis=0, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    mean_target = np.bincount(inverse, weights=target) / np.bincount(inverse)
    X = torch.tensor(uniq, dtype=torch.double)
    y = torch.tensor(mean_target, dtype=torch.double).flatten()
    return TorchNormalizer().fit_transform(X), y


def preprocess_data(filepath: str, skip_rows: int = 1, delimiter: str = ",") -> Tuple[torch.Tensor, torch.Tensor]:
    """Load a benchmark CSV and return (X in [0,1]^d, y) as fp64 tensors (utils.py:154-202)."""
    try:
        data = np.loadtxt(filepath, delimiter=delimiter, skiprows=skip_rows)
    except FileNotFoundError:
        raise FileNotFoundError(f"Data file not found: {filepath}")
    except ValueError:
        raise ValueError(f"Failed to parse data from: {filepath}")
    return preprocess_array(data)


def get_initial_samples(y: torch.Tensor, n_samples: int, percentile: float = 95) -> torch.Tensor:
    """Random initial design among the points at or below a percentile of y (utils.py:205-236).

    Used by the older ``gold_runs/benchmark_0`` traces; draws one ``randperm`` from the global
    torch generator.
    """
    cut = torch.quantile(y, percentile / 100)
    eligible = torch.where(y <= cut)[0]
    if n_samples > len(eligible):
        raise ValueError(
            f"Cannot sample {n_samples} points. Only {len(eligible)} points below {percentile}th percentile."
        )
    return eligible[torch.randperm(len(eligible))[:n_samples]]


def get_initial_samples_sobol(
    X: torch.Tensor, y: torch.Tensor, n_samples: int = 10, percentile: float = 95.0
) -> torch.Tensor:
    """Scrambled-Sobol initial design snapped to the nearest below-percentile pool point.

    Follows utils.py:239-281: the Sobol points are fp32, are scaled to the
    bounding box of the eligible points, and each is replaced by the first nearest eligible
    point, so repeated indices are possible and are kept.
    """
    cut = torch.quantile(y, percentile / 100)
    eligible = (y <= cut).nonzero(as_tuple=True)[0]
    if len(eligible) < n_samples:
        return eligible
    Xe = X[eligible]
    lo = Xe.min(dim=0)[0]
    span = Xe.max(dim=0)[0] - lo
    span[span == 0] = 1.0
    pts = SobolEngine(dimension=X.shape[1], scramble=True).draw(n_samples)
    pts = pts * span + lo
    picks = []
    for p in pts:
        dist = torch.norm(Xe - p, dim=1)
        picks.append(eligible[torch.argmin(dist)].item())
    return torch.tensor(picks, dtype=torch.long)


def get_device() -> torch.device:
    """CUDA device of this process; the B200 build has no MPS/CPU path (utils.py:285-295)."""
    if not torch.cuda.is_available():
        raise RuntimeError("stpbenchmark_b200 needs a CUDA device (sm_100a); none is visible")
    return torch.device("cuda", torch.cuda.current_device())


def EI(mean, std, best_observed, minimize: bool = False):
    """Plain expected improvement (utils.py:299-309)."""
    gain = (best_observed - mean) if minimize else (mean - best_observed)
    z = gain / std
    unit = torch.distributions.Normal(0, 1)
    return gain * unit.cdf(z) + std * unit.log_prob(z).exp()


# ---------------------------------------------------------------------------------------------
# Post-processing helpers of the reference's utils.py (none of them is on the BO hot path; host only).
# ---------------------------------------------------------------------------------------------
def initial_points(x: torch.Tensor, y: torch.Tensor, num_initial_points: int):
    """Random initial design away from the lowest 5 % of the targets (utils.py:136-151): a `torch.randperm` draw
    (global generator) over the points with y above the 5 % quantile; returns (x_initial, y_initial)."""
    keep = torch.nonzero(y > torch.quantile(y, 0.05), as_tuple=True)[0]
    chosen = keep[torch.randperm(len(keep))[:num_initial_points]]
    return x[chosen], y[chosen]


def get_top_samples(train_y: torch.Tensor, top_percentage: float = 5):
    """The ceil(top_percentage %) largest target values as a list, largest first, first occurrence first on ties
    (utils.py:312-334; the reference also prints them)."""
    values = train_y.detach().cpu().numpy().reshape(-1)
    count = int(math.ceil(len(values) * (top_percentage / 100)))
    order = np.argsort(-values, kind="stable")[:count]
    top = values[order].tolist()
    print(f"Number of top {top_percentage}% samples: {len(top)}")
    print(f"Top {top_percentage}% samples: {top}")
    return top


def num_top_samples(y, top_samples) -> float:
    """Fraction of the top samples that a campaign has found (utils.py:337-338)."""
    return sum(1 for v in y if v in top_samples) / len(top_samples)


def collect_arrays(prefix: str, num_arrays: int, namespace: dict | None = None):
    """The arrays named f"{prefix}{i}", i < num_arrays, that exist in `namespace` (utils.py:342-349 reads its own
    module globals; pass the caller's `globals()` to get the same effect from another module)."""
    scope = globals() if namespace is None else namespace
    return [scope[f"{prefix}{i}"] for i in range(num_arrays) if scope.get(f"{prefix}{i}") is not None]


def pad_array(array, max_length: int):
    """Right-pad a 1-D array to `max_length` by repeating its last value (utils.py:352-355)."""
    array = np.asarray(array)
    return np.concatenate([array, np.full(max_length - len(array), array[-1], dtype=array.dtype)])


def find_max_length(prefix: str, num_arrays: int, namespace: dict | None = None) -> int:
    """Length of the longest collected array (utils.py:358-360)."""
    return max(len(a) for a in collect_arrays(prefix, num_arrays, namespace))


def process_arrays(prefix: str, num_arrays: int, max_length: int, namespace: dict | None = None):
    """Mean and (population) standard deviation over the collected arrays after padding (utils.py:364-370)."""
    stack = np.stack([pad_array(a, max_length) for a in collect_arrays(prefix, num_arrays, namespace)])
    return stack.mean(axis=0), stack.std(axis=0)
