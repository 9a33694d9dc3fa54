"""Shared loaders for the committed golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATASETS = ["AgNP", "AutoAM", "CrossedBarrel", "P3HT", "Perovskite"]
INVERTED = ("AgNP", "Perovskite")  # run_benchmark.py:27-28 maximises 1/y on these

_cache = {}


def _npz(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name))
    return _cache[name]


def dataset(name, invert=True):
    """(X, y) as the reference's run_benchmark.py hands them to run_many_loops."""
    z = _npz("datasets.npz")
    X = torch.tensor(z[f"{name}_X"], dtype=torch.double)
    y = torch.tensor(z[f"{name}_y"], dtype=torch.double)
    if invert and name in INVERTED:
        y = 1.0 / y
    return X, y


def raw_table(name):
    return _npz("datasets.npz")[f"{name}_raw"]


def seeds():
    return [int(s) for s in _npz("traces.npz")["seeds"]]


def trace(which, model, name, seed):
    """x_index column of one golden trace; which in {'gold1', 'smoke'}."""
    arr = _npz("traces.npz")[f"{which}_{model}_{name}"]
    return arr[seeds().index(int(seed))].astype(int).tolist()


# Seeds whose whole 90-row ExactGP gold_runs/benchmark_1 trace the oracle reproduces
EXACT_KAT = {
    "AgNP": [45, 359, 102, 123, 5676, 909, 604, 766, 824, 8623, 1946, 169, 577, 369, 9342, 980, 288, 715, 2672,
             184, 404, 849, 177, 8, 245, 412, 764, 882, 362, 4143],
    "AutoAM": [45, 102, 123, 368, 604, 824, 8623, 135, 169, 577, 369, 9342, 288, 2672, 931, 177, 340, 8, 759,
               245, 479, 753, 882, 362, 4143],
    "CrossedBarrel": [45, 359, 123, 369, 310, 288, 2672, 412, 4143],
    "Perovskite": [s for s in [45, 359, 6185, 102, 421, 123, 5676, 368, 945, 909, 604, 31, 7736, 766, 824, 8623,
                               788, 781, 118, 1946, 135, 169, 577, 369, 9342, 980, 310, 690, 288, 715, 2672, 184,
                               148, 262, 931, 404, 849, 177, 340, 8, 3794, 759, 245, 412, 479, 764, 753, 882, 362,
                               4143] if s not in (5676, 9342, 177, 8, 3794)],
}
