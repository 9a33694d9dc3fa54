"""The fit-and-predict study of benchmark_predictions.py (3 models x 5 datasets x 25 seeds, 80/20 splits, 600 epochs)
through stpbenchmark_b200.benchmark_predictions, timed per (model, dataset) pair: one batched fit launch + one batched
predict launch each.  Datasets come from the committed fixtures (tests/golden/datasets.npz), written as the CSV files
the reference reads.  Prints one JSON document (profiles/r02_predictions_study.json)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _fixtures import DATASETS, raw_table, seeds  # noqa: E402
from stpbenchmark_b200 import benchmark_predictions as bp  # noqa: E402
from stpbenchmark_b200.models import ExactGP, VarGP, VarTGP  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 25
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 600
d = tempfile.mkdtemp()
for name in DATASETS:
    tab = raw_table(name)
    np.savetxt(os.path.join(d, f"{name}_dataset.csv"), tab, delimiter=",",
               header=",".join(f"c{i}" for i in range(tab.shape[1])), comments="")
sd = seeds()[:n_seeds]
out = {"n_seeds": n_seeds, "epochs": epochs, "models": {}}
bp.evaluate_model_across_datasets(ExactGP, n_seeds=2, data_dir=d, seeds=sd, verbose=False, dataset_files=["AutoAM_dataset.csv"])  # warm-up
t_all = time.perf_counter()
for cls in (VarTGP, VarGP, ExactGP):
    per = {}
    for name in DATASETS:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = bp.evaluate_model_across_datasets(cls, n_seeds=n_seeds, epochs=epochs, lr=0.01, data_dir=d, seeds=sd, verbose=False,
                                              dataset_files=[f"{name}_dataset.csv"])[f"{name}_dataset"]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        per[name] = {"seconds": dt, "fits_per_sec": n_seeds / dt,
                     "mean_mae": r["mean_mae"], "std_mae": r["std_mae"], "bad": int(sum(1 for s in r["status"] if s != 0))}
    out["models"][cls.__name__] = per
out["total_seconds"] = time.perf_counter() - t_all
out["fits"] = 3 * len(DATASETS) * n_seeds
print(json.dumps(out))
