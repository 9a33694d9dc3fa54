"""The sweep of reference run_benchmark.py: model x dataset x seeds -> results/*.csv.

Same constants, same file-name pattern ``results/{name}_{dataset}_traces_{'SMOKE' or ''}.csv``
(run_benchmark.py:41).  Two driver defects of the reference are NOT reproduced (SURVEY App.
D-12): datasets are iterated in sorted order rather than filesystem order, and every model
name is paired with its own class (the reference zips two names with four classes, so it runs
VarTGP under the name "VarLGP" and VarGP under "ExactGP").

Unlike the reference, which runs (model, dataset, seed) strictly one after another, the whole sweep is ONE device job
(runners.run_sweep -> engine.CampaignJob): every campaign shares the GPU, and the GPUs of the box under torchrun.

usage: python -m stpbenchmark_b200.run_benchmark [--data DIR] [--seeds FILE] [--results DIR]
                                                [--smoke] [--models ExactGP,VarTGP,...] [--sequential]
       python -m torch.distributed.run --nproc-per-node 8 -m stpbenchmark_b200.run_benchmark ...
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from .models import ExactGP, VarGP, VarLGP, VarTGP
from .runners import run_sweep
from .utils import preprocess_data

MODELS = {"VarTGP": VarTGP, "VarGP": VarGP, "VarLGP": VarLGP, "ExactGP": ExactGP}
INVERTED = ("Perovskite_dataset.csv", "AgNP_dataset.csv")  # minimisation problems (run_benchmark.py:27)


def build_jobs(data_dir, seed_list, results_dir, models, smoke):
    """One run_many_loops job per (model, dataset), in the order of run_benchmark.py:20-43."""
    jobs = []
    for model_name in models:
        model_class = MODELS[model_name]
        for dataset in sorted(os.listdir(data_dir)):
            if not dataset.endswith(".csv"):
                continue
            X, y = preprocess_data(filepath=os.path.join(data_dir, dataset))
            if dataset in INVERTED:
                y = 1.0 / y
            jobs.append(dict(
                X=X, y=y, model_class=model_class,
                seeds=seed_list[:3] if smoke else seed_list,
                n_initial=10, n_trials=30 if smoke else 80, epochs=200 if smoke else 600, learning_rate=0.01,
                output_path=os.path.join(results_dir, f"{model_name}_{dataset[:-4]}_traces_{'SMOKE' if smoke else ''}.csv"),
                invert_y=dataset in INVERTED,
            ))
    return jobs


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="data")
    ap.add_argument("--seeds", default="random_seeds.txt")
    ap.add_argument("--results", default="results")
    ap.add_argument("--smoke", action="store_true")
    ap.add_argument("--models", default="VarTGP,VarGP,VarLGP,ExactGP")
    ap.add_argument("--sequential", action="store_true",
                    help="one run_many_loops call per (model, dataset) like the reference, instead of one device job")
    args = ap.parse_args(argv)
    seed_list = [int(s) for s in np.loadtxt(args.seeds, dtype=int)]
    if args.smoke:
        print("\nRUNNING SMOKE TEST")
    # One process per GPU under torchrun: the campaigns of the whole sweep are sharded over the ranks and the traces
    # gathered on rank 0 (runners.run_campaigns); a plain `python -m ...` run uses the current device.
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    jobs = build_jobs(args.data, seed_list, args.results, args.models.split(","), args.smoke)
    if args.sequential:
        for job in jobs:
            print(f"\nRunning {os.path.basename(job['output_path'])}")
            run_sweep([job])
    else:
        print(f"\nRunning {len(jobs)} (model, dataset) jobs x {len(jobs[0]['seeds']) if jobs else 0} seeds as one device job")
        run_sweep(jobs)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
