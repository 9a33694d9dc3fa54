"""Full-length parity sweep on the GPU: every (model, dataset, seed) campaign of the reference's
results/gold_runs/benchmark_1 (4 models x 5 datasets x 50 seeds, 10 + 80 rows, 600 epochs) is run through
stpbenchmark_b200.runners.run_campaign_batch and compared with the golden x_index traces (tests/golden/traces.npz).

ExactGP: index-exactness per campaign (the CPU oracle reproduces 109 / 250 exactly; P3HT is
bimodal and the reference does not reproduce itself there).  Variational models: the fits are chaotic in round-off
beyond ~100 epochs and the MC stream of the reference is unrecoverable (App. C.5, C.7), so the comparison is
statistical: best-so-far value after 20 / 40 / 90 rows and overlap of the visited sets.

usage: python tools/gpu_parity_sweep.py [--models ExactGP,VarGP,VarTGP,VarLGP] [--out profiles/r01_parity_gold1.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _fixtures import DATASETS, dataset, seeds, trace  # noqa: E402
from stpbenchmark_b200.runners import run_campaign_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="ExactGP,VarGP,VarTGP,VarLGP")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_parity_gold1.json"))
    ap.add_argument("--nseeds", type=int, default=50)
    args = ap.parse_args()
    report = {}
    for model in args.models.split(","):
        report[model] = {}
        for name in DATASETS:
            X, y = dataset(name)
            sd = seeds()[: args.nseeds]
            torch.cuda.synchronize(); t0 = time.perf_counter()
            idx, val, status = run_campaign_batch(X, y, model, sd, n_initial=10, n_trials=80, epochs=600, learning_rate=0.01)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            gold = [trace("gold1", model, name, s) for s in sd]
            first = [next((i for i in range(min(len(a), 90)) if a[i] != g[i]), min(len(a), 90)) for a, g in zip(idx, gold)]
            yv = y.numpy()
            row = {"campaigns": len(sd), "seconds": round(dt, 2), "iterations_per_s": round(len(sd) * 80 / dt, 1),
                   "failed": int(sum(1 for s in status if s != 0)),
                   "init_design_equal": int(sum(a[:10] == g[:10] for a, g in zip(idx, gold))),
                   "index_exact_90_rows": int(sum(f >= 90 for f in first)),
                   "exact_seeds": [int(s) for s, f in zip(sd, first) if f >= 90],
                   "index_exact_40_rows": int(sum(f >= 40 for f in first)),
                   "median_first_divergence": float(np.median(first)),
                   "step_agreement_free_running": round(float(np.mean([np.mean(np.array(a[10:90]) == np.array(g[10:90]))
                                                                        for a, g in zip(idx, gold) if len(a) >= 90])), 4)}
            for upto in (20, 40, 90):
                ours = np.mean([yv[a[:upto]].max() for a in idx if len(a) >= upto])
                ref = np.mean([yv[g[:upto]].max() for g in gold])
                row[f"best_so_far_{upto}"] = {"ours": float(ours), "golden": float(ref)}
            row["visited_overlap"] = round(float(np.mean([len(set(a[10:]) & set(g[10:])) / 80.0 for a, g in zip(idx, gold)])), 4)
            report[model][name] = row
            print(model, name, row, flush=True)
    summary = {m: {"index_exact_90_rows": sum(r["index_exact_90_rows"] for r in d.values()),
                   "campaigns": sum(r["campaigns"] for r in d.values())} for m, d in report.items()}
    report["summary"] = summary
    with open(args.out, "w") as fh:
        json.dump(report, fh, indent=1)
    print(summary)


if __name__ == "__main__":
    main()
