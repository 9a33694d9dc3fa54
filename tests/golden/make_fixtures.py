"""Regenerates tests/golden/*.npz from the read-only reference checkout.  Run in the BUILD
container only (``python tests/golden/make_fixtures.py``); /root/reference does not exist on
the GPU box, which is why the outputs are committed.

Outputs
  datasets.npz  raw CSV tables of reference data/*.csv (float64), and X / y after the
                reference's own ``utils.preprocess_data`` (imported from /root/reference).
  traces.npz    the ``x_index`` columns (int16) of results/gold_runs/benchmark_1/*_traces_.csv
                (4 models x 5 datasets x 50 seeds x 90 rows), of results/*_SMOKE.csv
                (4 x 5 x 3 seeds x 40 rows), the first 10 rows of gold_runs/benchmark_0
                (the older random-percentile initial design) and random_seeds.txt.
  acqf_ref.npz  inputs / outputs of the reference's own acqf.py (imports here: numpy, torch,
                scipy only) against a stub posterior.
Only indices are taken from the traces: the y_value columns of the inverted datasets are
unreliable.
"""
import importlib.util
import os
import sys

import numpy as np
import pandas as pd
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
DATASETS = ["AgNP", "AutoAM", "CrossedBarrel", "P3HT", "Perovskite"]
MODELS = ["ExactGP", "VarGP", "VarLGP", "VarTGP"]


def _load(name):
    spec = importlib.util.spec_from_file_location(f"ref_{name}", f"{REF}/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    sys.dont_write_bytecode = True
    ref_utils = _load("utils")
    seeds = np.loadtxt(f"{REF}/random_seeds.txt", dtype=int)

    ds = {}
    for name in DATASETS:
        path = f"{REF}/data/{name}_dataset.csv"
        ds[f"{name}_raw"] = np.loadtxt(path, delimiter=",", skiprows=1)
        X, y = ref_utils.preprocess_data(path)
        ds[f"{name}_X"] = X.numpy()
        ds[f"{name}_y"] = y.numpy()
    np.savez_compressed(f"{OUT}/datasets.npz", **ds)

    tr = {"seeds": seeds.astype(np.int64)}
    for model in MODELS:
        for name in DATASETS:
            df = pd.read_csv(f"{REF}/results/gold_runs/benchmark_1/{model}_{name}_dataset_traces_.csv",
                             header=[0, 1], index_col=0)
            tr[f"gold1_{model}_{name}"] = np.stack(
                [df[f"seed_{s}", "x_index"].values for s in seeds]).astype(np.int16)
            df = pd.read_csv(f"{REF}/results/{model}_{name}_dataset_traces_SMOKE.csv", header=[0, 1], index_col=0)
            tr[f"smoke_{model}_{name}"] = np.stack(
                [df[f"seed_{s}", "x_index"].values for s in seeds[:3]]).astype(np.int16)
    for name in DATASETS:
        df = pd.read_csv(f"{REF}/results/gold_runs/benchmark_0/ExactGP_{name}_dataset_traces.csv",
                         header=[0, 1], index_col=0)
        first = np.stack([df[f"seed_{s}", "x_index"].values[:10] for s in seeds]).astype(np.float64)
        if name in ("AgNP", "Perovskite"):  # that run stored 1/index for the inverted datasets
            with np.errstate(divide="ignore"):
                first = np.where(np.isinf(1.0 / first), 0.0, np.rint(1.0 / first))
        tr[f"gold0_init_{name}"] = first.astype(np.int16)
    np.savez_compressed(f"{OUT}/traces.npz", **tr)

    # reference acqf.py against a stub model (row a10)
    ref_acqf = _load("acqf")
    g = torch.Generator().manual_seed(7)
    mean = torch.randn(64, generator=g, dtype=torch.double)
    var = torch.rand(64, generator=g, dtype=torch.double) * 0.5 + 1e-3
    var[5] = 0.0
    df_t = torch.full((64,), 4.5, dtype=torch.double)
    mc_mean = mean.unsqueeze(0) + 0.1 * torch.randn(16, 64, generator=g, dtype=torch.double)

    class _Post:
        def __init__(self, m, v, df=None):
            self.mean, self.variance, self.df = m, v, df

    class _Stub:
        def __init__(self, post):
            self._p = post

        def posterior(self, X, **kw):
            return self._p

    Xq = torch.zeros(64, 3, dtype=torch.double)
    best = torch.tensor(0.3, dtype=torch.double)
    out = {
        "mean": mean.numpy(), "var": var.numpy(), "df": df_t.numpy(), "mc_mean": mc_mean.numpy(), "best_f": best.numpy(),
        "logei": ref_acqf.log_expected_improvement(_Stub(_Post(mean, var)), Xq, best).numpy(),
        "logei_mc": ref_acqf.log_expected_improvement(
            _Stub(_Post(mc_mean, var.expand(16, 64))), Xq, best).numpy(),
        "logei_t": ref_acqf.log_expected_improvement_t(_Stub(_Post(mean, var, df_t)), Xq, best).numpy(),
    }
    np.savez_compressed(f"{OUT}/acqf_ref.npz", **out)
    for f in ("datasets.npz", "traces.npz", "acqf_ref.npz"):
        print(f, os.path.getsize(f"{OUT}/{f}"), "bytes")


if __name__ == "__main__":
    main()
