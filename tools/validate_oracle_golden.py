"""Re-establishes SURVEY App. C.2: runs the CPU oracle free-running on all 250 ExactGP
campaigns of results/gold_runs/benchmark_1 (from the committed fixtures) and records, per
(dataset, seed), the first row at which the oracle's trace leaves the golden one.
Writes tests/golden/oracle_kat_report.json.  ~15 min on 7 cores.

usage: python tools/validate_oracle_golden.py [--procs 7] [--model ExactGP]
"""
import argparse
import json
import os
import sys
import time
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _one(job):
    import torch
    torch.set_num_threads(1)
    from _fixtures import dataset, trace
    from oracle.loop import run_single_loop
    ds, seed, model, which = job
    X, y = dataset(ds)
    gold = trace(which, model, ds, seed)
    n_trials, epochs = (80, 600) if which == "gold1" else (30, 200)
    t0 = time.time()
    sel, _ = run_single_loop(X, y, model, n_initial=10, n_trials=n_trials, epochs=epochs,
                             learning_rate=0.01, seed=seed)
    first = next((i for i, (a, b) in enumerate(zip(sel, gold)) if a != b), len(sel))
    return ds, seed, first, len(sel), time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=7)
    ap.add_argument("--model", default="ExactGP")
    ap.add_argument("--which", default="gold1")
    ap.add_argument("--nseeds", type=int, default=50)
    args = ap.parse_args()
    from _fixtures import DATASETS, seeds
    jobs = [(ds, s, args.model, args.which) for ds in DATASETS for s in seeds()[: args.nseeds]]
    report = {}
    with Pool(args.procs) as pool:
        for ds, seed, first, total, dt in pool.imap_unordered(_one, jobs):
            report.setdefault(ds, {})[str(seed)] = first
            print(f"{ds} {seed}: first divergence {first}/{total} ({dt:.0f}s)", flush=True)
    summary = {ds: sum(1 for v in r.values() if v >= (90 if args.which == "gold1" else 40)) for ds, r in report.items()}
    out = {"model": args.model, "which": args.which, "exact_traces": summary,
           "first_divergence": {ds: dict(sorted(r.items(), key=lambda kv: int(kv[0]))) for ds, r in report.items()}}
    name = f"oracle_kat_report_{args.model}_{args.which}.json"
    with open(os.path.join(ROOT, "tests", "golden", name), "w") as fh:
        json.dump(out, fh, indent=1)
    print(summary)


if __name__ == "__main__":
    main()
