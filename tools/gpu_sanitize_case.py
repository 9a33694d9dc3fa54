"""Smallest end-to-end case for compute-sanitizer: 2 ExactGP campaigns x 2 BO steps and 2 VarTGP campaigns x 1 step
(5 epochs, 64 MC samples) on a 64-point pool."""
import sys, torch
sys.path.insert(0, '/root/repo')
from stpbenchmark_b200 import synthetic
from stpbenchmark_b200.engine import CampaignBatch
X, y = synthetic.cfg4_pool(0, N=64)
init = [list(range(0, 12)), list(range(5, 25))]
b = CampaignBatch(X, y, [0, 0], init, n_cap=24, kind="ExactGP"); b.ragged = True
b.run(2)
print('exact', b.traces()[0][:, :23].tolist(), b.status.tolist())
v = CampaignBatch(X, y, [0, 0], init, n_cap=24, kind="VarTGP", epochs=5, num_samples=64); v.ragged = True
v.run(1)
print('var', v.traces()[1].tolist(), v.status.tolist())
