"""Two stp_var_acq launches (VarTGP, B = 296, n = 50, pool 2048, S = 2048 in-kernel Philox samples) for ncu."""
import sys, torch
sys.path.insert(0, '/root/repo')
from stpbenchmark_b200 import _lib, synthetic, varops
dev = "cuda"; X, y = synthetic.cfg4_pool(0); d = 6; N = 2048; B, n = 296, (int(sys.argv[1]) if len(sys.argv) > 1 else 50)
cap = n + 1; g = torch.Generator().manual_seed(n)
Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double); mask = torch.ones(B, N, dtype=torch.uint8)
for b in range(B):
    idx = torch.randperm(N, generator=g)[:n]; Xb[b, :n] = X[idx]; yb[b, :n] = y[idx]; mask[b, idx] = 0
Xb, yb, mask = Xb.to(dev), yb.to(dev), mask.to(dev); nb = torch.full((B,), n, dtype=torch.int32, device=dev)
st = varops.alloc_state(B, cap, d, dev); status = torch.zeros(B, dtype=torch.int32, device=dev); work = _lib.var_workspace(B, cap, dev)
_lib.var_fit(Xb, yb, nb, st, 1, 20, 0.01, 0.01, varops.var_hyp0(d), varops.var_prior_vector(), varops.gauss_hermite_table(), status, work, fresh=True, seed=1)
pool = X.to(dev).unsqueeze(0).contiguous(); best = yb.max(dim=1).values.contiguous()
for rep in range(2):
    _lib.var_acq(pool, st, nb, 1, status, work, best_f=best, mask=mask, S=2048, seed=2)
torch.cuda.synchronize(); print('ok', int((status != 0).sum()))
