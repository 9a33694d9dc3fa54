"""How much of k_var_acq is the Monte-Carlo LogEI?  Times stp_var_acq on the same trained state (B = 1024, n = 50,
pool 2048, d = 6) as VarGP (analytic acquisition: triangular products + one LogEI per candidate) and as VarTGP with
S = 16 / 256 / 2048 in-kernel Philox samples per candidate.  usage (gpurun): python tools/gpu_time_acq_split.py"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from stpbenchmark_b200 import _lib, synthetic, varops
dev = 'cuda'
X, y = synthetic.cfg4_pool(0); d = 6; N = 2048
for B, n in ((1024, 50), (1024, 89)):
    cap = n + 1
    g = torch.Generator().manual_seed(n)
    Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double)
    mask = torch.ones(B, N, dtype=torch.uint8)
    for b in range(B):
        idx = torch.randperm(N, generator=g)[:n]
        Xb[b, :n] = X[idx]; yb[b, :n] = y[idx]; mask[b, idx] = 0
    Xb, yb, mask = Xb.to(dev), yb.to(dev), mask.to(dev)
    nb = torch.full((B,), n, dtype=torch.int32, device=dev)
    pool = X.to(dev).unsqueeze(0).contiguous()
    best = yb.max(dim=1).values.contiguous()
    for lik, name in ((0, "VarGP"), (1, "VarTGP")):
        st = varops.alloc_state(B, cap, d, dev); status = torch.zeros(B, dtype=torch.int32, device=dev)
        work = _lib.var_workspace(B, cap, dev)
        _lib.var_fit(Xb, yb, nb, st, lik, 20, 0.01, 0.01, varops.var_hyp0(d), varops.var_prior_vector(), varops.gauss_hermite_table(),
                     status, work, fresh=True, seed=1)
        for S in ((1,) if lik == 0 else (16, 256, 2048)):
            for rep in range(2):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                _lib.var_acq(pool, st, nb, lik, status, work, best_f=best, mask=mask, S=S, seed=2)
                torch.cuda.synchronize(); t1 = time.perf_counter()
            print(f"B={B} n={n} {name:6s} S={S:5d}: acq {1e3 * (t1 - t0):7.2f} ms", flush=True)
