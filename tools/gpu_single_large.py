"""One batched ExactGP fit + predict of CrossedBarrel's 80 % splits (n = 480, 25 seeds) through the large entry points:
the ncu target for k_exact_fit<LARGE> / k_exact_acq_large (tools/gpu_r2_large_evidence.sh)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _fixtures import dataset, seeds  # noqa: E402
from stpbenchmark_b200 import benchmark_predictions as bp  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "ExactGP"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 600
X, y = dataset("CrossedBarrel", invert=False)
tr, te = zip(*[bp.split_indices(len(X), s) for s in seeds()[:25]])
tr, te = torch.stack(tr), torch.stack(te)
res = bp.fit_predict_batch(kind, X[tr], y[tr], X[te], epochs=epochs)
print(kind, "status", res["status"].tolist(), "mae", float((res["mean"] - y[te]).abs().mean()))
