"""One stp_var_fit launch for ncu: VarTGP, B = 296 campaigns of n points on a cfg #4 pool (E = 20 epochs), or with
`cfg5` B = 148 campaigns of n = 256 on a cfg #5 pool (pool 8192, d = 8; E = 3).  usage: gpu_prof_var.py [n] [cfg5]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from stpbenchmark_b200 import _lib, synthetic, varops
CFG5 = "cfg5" in sys.argv
dev = 'cuda'
X, y = synthetic.cfg5_pool(0) if CFG5 else synthetic.cfg4_pool(0)
d = X.shape[1]; N = X.shape[0]
nums = [a for a in sys.argv[1:] if a.isdigit()]
B, n, E = (148, 256, 3) if CFG5 else (296, int(nums[0]) if nums else 50, 20)
cap = n + 1; g = torch.Generator().manual_seed(n)
Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double)
for b in range(B):
    idx = torch.randperm(N, generator=g)[:n]; Xb[b, :n] = X[idx]; yb[b, :n] = y[idx]
Xb, yb = Xb.to(dev), yb.to(dev); nb = torch.full((B,), n, dtype=torch.int32, device=dev)
st = varops.alloc_state(B, cap, d, dev); status = torch.zeros(B, dtype=torch.int32, device=dev); work = _lib.var_workspace(B, cap, dev)
for rep in range(2):
    _lib.var_fit(Xb, yb, nb, st, 1, E, 0.01, 0.01, varops.var_hyp0(d), varops.var_prior_vector(), varops.gauss_hermite_table(), status, work, fresh=True, seed=1)
torch.cuda.synchronize(); print('ok', int((status != 0).sum()))
