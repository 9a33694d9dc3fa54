"""Times stp_var_fit / stp_var_acq (VarTGP) on synthetic cfg4-shaped campaigns: quick device numbers."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from stpbenchmark_b200 import _lib, synthetic, varops
dev = 'cuda'
CFG5 = len(sys.argv) > 2 and sys.argv[2] == "cfg5"     # config #5 shape: pool 8192, d = 8, M = n = 256
X, y = synthetic.cfg5_pool(0) if CFG5 else synthetic.cfg4_pool(0)
d = 8 if CFG5 else 6; N = 8192 if CFG5 else 2048
E = int(sys.argv[1]) if len(sys.argv) > 1 else 600
for B, n in ([(148, 256), (296, 256)] if CFG5 else [(148, 20), (296, 20), (296, 50), (296, 89), (592, 50), (1024, 50)]):
    cap = n + 1
    g = torch.Generator().manual_seed(n)
    Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double)
    mask = torch.ones(B, N, dtype=torch.uint8)
    for b in range(B):
        idx = torch.randperm(N, generator=g)[:n]
        Xb[b, :n] = X[idx]; yb[b, :n] = y[idx]; mask[b, idx] = 0
    Xb, yb, mask = Xb.to(dev), yb.to(dev), mask.to(dev)
    nb = torch.full((B,), n, dtype=torch.int32, device=dev)
    st = varops.alloc_state(B, cap, d, dev); status = torch.zeros(B, dtype=torch.int32, device=dev)
    work = _lib.var_workspace(B, cap, dev)
    pool = X.to(dev).unsqueeze(0).contiguous()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.var_fit(Xb, yb, nb, st, 1, E, 0.01, 0.01, varops.var_hyp0(d), varops.var_prior_vector(), varops.gauss_hermite_table(),
                     status, work, fresh=True, seed=1)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        out = _lib.var_acq(pool, st, nb, 1, status, work, best_f=yb.max(dim=1).values.contiguous(), mask=mask, S=2048, seed=2)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    flops_fit = B * E * (9 * n**3 + n * n * (14 * d + 12) + 800 * n)
    print(f'B={B:5d} n={n:3d} E={E}: fit {1e3*(t1-t0):8.1f} ms ({flops_fit/(t1-t0)/1e12:5.2f} TFLOP/s alg), acq {1e3*(t2-t1):7.1f} ms, '
          f'it/s {B/(t2-t0):8.1f}, bad {int((status!=0).sum())}', flush=True)
