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
    "sobol_scrambled_draw",
    "sobol_scrambled_draw_batch",
    "initial_designs_batch",
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


# ---------------------------------------------------------------------------------------------
# Batched initial designs: what `set_seeds(seed); get_initial_samples_sobol(X, y, n, p)`
# returns for MANY seeds on one pool, without paying the per-seed Python / SobolEngine construction cost
# (1.2 ms per design serially; 1024 designs of BASELINE config #4 took longer than the GPU needs for all their
# 81 920 BO iterations).  The scrambled-Sobol draw stays on the host -- it is defined by torch's CPU generator --
# but goes straight to the three ATen ops SobolEngine itself uses, with a private generator seeded like the
# global one; the O(S n N d) nearest-eligible-point snap is one vectorised pass (host) or one kernel
# (stp_design_snap, when the pool lives on a CUDA device).
# ---------------------------------------------------------------------------------------------
_SOBOL_BASE = {}


def _sobol_base(d: int):
    if d not in _SOBOL_BASE:
        state = torch.zeros(d, SobolEngine.MAXBIT, dtype=torch.long)
        torch._sobol_engine_initialize_state_(state, d)
        _SOBOL_BASE[d] = (state, torch.pow(2, torch.arange(0, SobolEngine.MAXBIT)))
    return _SOBOL_BASE[d]


def sobol_scrambled_draw(seed: int, d: int, n: int, generator: torch.Generator = None):
    """(points [n, d] fp32, generator) == ``torch.manual_seed(seed); SobolEngine(d, scramble=True).draw(n)`` and the
    state the global generator is left in (utils.py:262-269 after set_seeds, utils.py:11-16): same ATen calls in
    the same order (torch/quasirandom.py: randint shift, randint lower-triangular matrices, _sobol_engine_scramble_,
    _sobol_engine_draw), on a private generator, so nothing global is touched."""
    base, pow2 = _sobol_base(d)
    g = generator if generator is not None else torch.Generator()
    g.manual_seed(int(seed))
    maxbit = SobolEngine.MAXBIT
    shift = torch.mv(torch.randint(2, (d, maxbit), generator=g), pow2)
    ltm = torch.randint(2, (d, maxbit, maxbit), generator=g).tril()
    state = base.clone()
    torch._sobol_engine_scramble_(state, ltm, d)
    first = (shift / 2 ** maxbit).reshape(1, -1).to(torch.float32)
    if n == 1:
        return first, g
    rest, _ = torch._sobol_engine_draw(shift.clone(), n - 1, state, d, 0, dtype=torch.float32)
    return torch.cat((first, rest), dim=-2), g


def sobol_scrambled_draw_batch(seeds, d: int, n: int, return_rng_states: bool = False):
    """``sobol_scrambled_draw`` for many seeds at once: [len(seeds), n, d] fp32 (+ the generator states).

    Only the two ``randint`` calls per seed touch torch's generator (they define the scramble); the linear scramble
    of the direction numbers (ATen `_sobol_engine_scramble_`: new_v[d][j] = sum_p parity(popcount(row_p(LTM_d) & v[d][j]))
    2^(MAXBIT-1-p), rows read MSB-first with a unit diagonal) and the Gray-code draw (`_sobol_engine_draw`: point i =
    point i-1 xor the direction number at the lowest zero bit of i-1, scaled by 2^-MAXBIT in fp32) are vectorised
    over the seeds.  Bit-exact against SobolEngine (tests/test_host_utils.py)."""
    maxbit = SobolEngine.MAXBIT
    base, pow2 = _sobol_base(d)
    S = len(seeds)
    shift_bits = torch.empty(S, d, maxbit, dtype=torch.long)
    ltm = torch.empty(S, d, maxbit, maxbit, dtype=torch.float32)
    states = [] if return_rng_states else None
    g = torch.Generator()
    for i, seed in enumerate(seeds):
        g.manual_seed(int(seed))
        shift_bits[i] = torch.randint(2, (d, maxbit), generator=g)
        ltm[i] = torch.randint(2, (d, maxbit, maxbit), generator=g)
        if return_rng_states:
            states.append(g.get_state())
    shift = (shift_bits @ pow2).numpy()                                  # [S, d]
    ltm = ltm.tril()
    ltm.diagonal(dim1=-2, dim2=-1).fill_(1.0)
    # parity(popcount(row_p & v_j)) = (sum_c LTM[p][c] * bit_{MAXBIT-1-c}(v_j)) mod 2: small exact fp32 contractions
    rev = torch.arange(maxbit - 1, -1, -1)
    vbits = ((base.unsqueeze(-1) >> rev) & 1).to(torch.float32)          # [d, j, c]
    par = torch.einsum("sdpc,djc->sdjp", ltm, vbits).remainder_(2.0)
    state = (par.double() @ torch.pow(2.0, rev.double())).long().numpy()  # [S, d, j]
    pts = np.empty((S, n, d), dtype=np.int64)
    quasi = shift.copy()
    pts[:, 0] = quasi
    for i in range(n - 1):
        l = 0
        while (i >> l) & 1:                                              # lowest zero bit of i
            l += 1
        quasi = quasi ^ state[:, :, l]
        pts[:, i + 1] = quasi
    out = torch.from_numpy(pts).to(torch.float32) * torch.tensor(2.0 ** -maxbit, dtype=torch.float32)
    return (out, states) if return_rng_states else out


def initial_designs_batch(X: torch.Tensor, y: torch.Tensor, seeds, n_samples: int = 10, percentile: float = 50.0,
                          return_rng_states: bool = False, device=None):
    """Initial designs of many campaigns on one pool: row s == ``set_seeds(seeds[s]);
    get_initial_samples_sobol(X, y, n_samples, percentile)`` (runners.py:33, 46-48), bit for bit.

    Returns an int64 tensor [len(seeds), n_samples] (on the host) and, with ``return_rng_states``, the list of
    torch CPU generator states after each draw (what the variational models' first-forward noise continues from,
    App. A.6).  ``device``: a CUDA device runs the nearest-point snap in stp_design_snap; None = vectorised host pass.
    """
    Xh = X.detach().cpu().double()
    yh = y.detach().cpu().double().flatten()
    seeds = [int(s) for s in seeds]
    cut = torch.quantile(yh, percentile / 100)
    eligible = (yh <= cut).nonzero(as_tuple=True)[0]
    if len(eligible) < n_samples:                      # utils.py:254-255: not enough eligible points -> all of them
        out = eligible.unsqueeze(0).expand(len(seeds), -1).clone()
        states = None
        if return_rng_states:
            states = []
            for s in seeds:
                g = torch.Generator(); g.manual_seed(s); states.append(g.get_state())
        return (out, states) if return_rng_states else out
    Xe = Xh[eligible].contiguous()
    lo = Xe.min(dim=0)[0]
    span = Xe.max(dim=0)[0] - lo
    span[span == 0] = 1.0
    d = Xh.shape[1]
    if return_rng_states:
        pts, states = sobol_scrambled_draw_batch(seeds, d, n_samples, return_rng_states=True)
    else:
        pts, states = sobol_scrambled_draw_batch(seeds, d, n_samples), None
    P = pts * span + lo                                 # fp32 * fp64 -> fp64, two roundings (utils.py:271)
    if device is not None and torch.device(device).type == "cuda":
        from . import _lib
        idx = _lib.design_snap(P.reshape(-1, d).to(device), Xe.to(device)).cpu().long().reshape(len(seeds), n_samples)
    else:
        idx = torch.empty(len(seeds), n_samples, dtype=torch.long)
        chunk = max(1, (1 << 22) // max(1, n_samples * Xe.shape[0] * d))
        for a in range(0, len(seeds), chunk):
            diff = Xe.reshape(1, 1, -1, d) - P[a:a + chunk].unsqueeze(2)
            idx[a:a + chunk] = torch.argmin(torch.linalg.vector_norm(diff, dim=-1), dim=-1)
    out = eligible[idx]
    return (out, states) if return_rng_states else out


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
