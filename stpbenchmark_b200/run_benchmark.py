"""The sweep of reference run_benchmark.py: model x dataset x seeds -> results/*.csv.

Same constants, same file-name pattern ``results/{name}_{dataset}_traces_{'SMOKE' or ''}.csv``
(run_benchmark.py:41).  Two driver defects of the reference are NOT reproduced (SURVEY App.
D-12): datasets are iterated in sorted order rather than filesystem order, and every model
name is paired with its own class (the reference zips two names with four classes, so it runs
VarTGP under the name "VarLGP" and VarGP under "ExactGP").

usage: python -m stpbenchmark_b200.run_benchmark [--data DIR] [--seeds FILE] [--results DIR]
                                                [--smoke] [--models ExactGP,VarTGP,...]
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from .models import ExactGP, VarGP, VarLGP, VarTGP
from .runners import run_many_loops
from .utils import preprocess_data

MODELS = {"VarTGP": VarTGP, "VarGP": VarGP, "VarLGP": VarLGP, "ExactGP": ExactGP}
INVERTED = ("Perovskite_dataset.csv", "AgNP_dataset.csv")  # minimisation problems (run_benchmark.py:27)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="data")
    ap.add_argument("--seeds", default="random_seeds.txt")
    ap.add_argument("--results", default="results")
    ap.add_argument("--smoke", action="store_true")
    ap.add_argument("--models", default="VarTGP,VarGP,VarLGP,ExactGP")
    args = ap.parse_args(argv)
    seed_list = np.loadtxt(args.seeds, dtype=int)
    smoke = args.smoke
    if smoke:
        print("\nRUNNING SMOKE TEST")
    for model_name in args.models.split(","):
        model_class = MODELS[model_name]
        for dataset in sorted(os.listdir(args.data)):
            X, y = preprocess_data(filepath=os.path.join(args.data, dataset))
            if dataset in INVERTED:
                y = 1.0 / y
            print(f"\nRunning {model_name} on {dataset} dataset")
            run_many_loops(
                X, y, model_class=model_class,
                seeds=seed_list[:3] if smoke else seed_list,
                n_initial=10, n_trials=30 if smoke else 80, epochs=200 if smoke else 600, learning_rate=0.01,
                output_path=os.path.join(args.results, f"{model_name}_{dataset[:-4]}_traces_{'SMOKE' if smoke else ''}.csv"),
                invert_y=dataset in INVERTED,
            )


if __name__ == "__main__":
    main()
