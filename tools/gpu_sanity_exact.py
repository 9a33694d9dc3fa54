# first GPU sanity: CUDA path vs oracle (mll f/g, fit, acq, bo_step) + a tiny timing
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from _fixtures import dataset, trace
from stpbenchmark_b200 import _lib as L, priors
from oracle import exact as oe, acq as oa
dev = 'cuda'
print('fp64 probe TFLOP/s', L.fp64_probe(2048))
X, y = dataset('AgNP'); d = X.shape[1]; N = len(X)
tr = trace('gold1', 'ExactGP', 'AgNP', 45)
prior = priors.exact_prior_vector(d); th0 = priors.exact_theta0(d); lo, hi = priors.exact_bounds(d)
ns = [10, 17, 30, 55, 80]; B = len(ns); cap = 90
Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double)
for b, n in enumerate(ns):
    Xb[b, :n] = X[tr[:n]]; yb[b, :n] = y[tr[:n]]
nb = torch.tensor(ns, dtype=torch.int32)
theta = torch.tensor([th0] * B, dtype=torch.double)
f, g, st = L.exact_mll_fg(Xb.to(dev), yb.to(dev), theta.to(dev), prior, n=nb.to(dev))
torch.cuda.synchronize()
for b, n in enumerate(ns):
    fo, go = oe.mll_value_and_grad(np.array(th0), X[tr[:n]], y[tr[:n]])
    print('mll n', n, 'f err', abs(f[b].item() - fo), 'g relerr', np.abs(g[b].cpu().numpy() - go).max() / np.abs(go).max(), 'st', st[b].item())
thd = theta.to(dev).clone()
t0 = time.time()
ff, nit, nfev, st = L.exact_fit(Xb.to(dev), yb.to(dev), thd, lo, hi, prior, n=nb.to(dev))
torch.cuda.synchronize(); print('fit wall', time.time() - t0)
for b, n in enumerate(ns):
    ths, res = oe.fit_exact(X[tr[:n]], y[tr[:n]])
    print('fit n', n, 'gpu', nit[b].item(), nfev[b].item(), st[b].item(), ff[b].item(), 'scipy', res.nit, res.nfev, res.fun,
          'dtheta', np.abs(thd[b].cpu().numpy() - ths).max() / np.abs(ths).max())
    mask = torch.ones(N, dtype=torch.uint8); mask[tr[:n]] = 0
    out = L.exact_acq(X.to(dev).unsqueeze(0).contiguous(), Xb[b:b+1, :].to(dev).contiguous(), yb[b:b+1].to(dev).contiguous(),
                      torch.tensor(ths).to(dev).unsqueeze(0).contiguous(), best_f=y[tr[:n]].max().reshape(1).to(dev),
                      mask=mask.to(dev).unsqueeze(0).contiguous(), n=nb[b:b+1].to(dev), want_pointwise=True)
    mo, vo = oe.posterior_exact(ths, X[tr[:n]], y[tr[:n]], X); ao = oa.log_ei(mo, vo, y[tr[:n]].max())
    am = ao.clone(); am[mask == 0] = -np.inf
    print('   acq idx', out['idx'].item(), int(torch.argmax(am)), 'gold', tr[n], 'mean err', (out['mean'][0].cpu() - mo).abs().max().item() / mo.abs().max().item(),
          'var relerr', ((out['var'][0].cpu() - vo).abs() / vo.abs()).max().item(), 'acq err', (out['acq'][0].cpu() - ao).abs().max().item())
# free-running campaign via bo_step
B = 3; seeds = [45, 359, 6185]
Xs = torch.zeros(B, cap, d, dtype=torch.double); ys = torch.zeros(B, cap, dtype=torch.double)
mask = torch.ones(B, N, dtype=torch.uint8); trc = torch.full((B, cap), -1, dtype=torch.int32)
for b, s in enumerate(seeds):
    init = trace('gold1', 'ExactGP', 'AgNP', s)[:10]
    Xs[b, :10] = X[init]; ys[b, :10] = y[init]; mask[b, init] = 0; trc[b, :10] = torch.tensor(init, dtype=torch.int32)
Xs, ys, mask, trc = Xs.to(dev), ys.to(dev), mask.to(dev), trc.to(dev)
nn = torch.full((B,), 10, dtype=torch.int32, device=dev); status = torch.zeros(B, dtype=torch.int32, device=dev)
pool = X.to(dev).unsqueeze(0).contiguous(); pool_y = y.to(dev).unsqueeze(0).contiguous(); pid = torch.zeros(B, dtype=torch.int32, device=dev)
torch.cuda.synchronize(); t0 = time.time()
for t in range(80):
    L.exact_bo_step(pool, pool_y, pid, mask, Xs, ys, nn, trc, th0, lo, hi, prior, status)
torch.cuda.synchronize(); print('80 bo steps x3 campaigns wall', time.time() - t0, 'status', status.tolist())
for b, s in enumerate(seeds):
    g_ = trace('gold1', 'ExactGP', 'AgNP', s); mine = trc[b].cpu().tolist()
    fd = next((i for i in range(90) if mine[i] != g_[i]), 90)
    print('seed', s, 'first divergence vs gold', fd)
