"""Trainers with the reference's signatures (optim.py:11-181), running on the sm_100a library.

``train_exact_model_botorch``        optim.py:11-32   botorch fit_gpytorch_mll = L-BFGS-B on the exact MLL
``train_natural_variational_model``  optim.py:105-181 NGD (lr * n) on q(u) + Adam(0.01) on the rest
``train_exact_model``                optim.py:35-58   plain Adam on the exact MLL
``train_variational_model``          optim.py:61-102  plain Adam on the ELBO (only caller is the broken
                                                      test_hartmann.py; provided for completeness)
All train the model in place and return None.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .utils import TorchStandardScaler


def _exact_batch(model):
    X = model.train_inputs[0].double().unsqueeze(0).contiguous()
    y = model.train_targets.double().reshape(1, -1).contiguous()
    return X, y


def fit_exact_scipy_lockstep(X, y, n, theta0, lower, upper, prior, maxiter=15000, maxfun=15000):
    """Parity mode: scipy's own L-BFGS-B (one ``setulb`` state machine per campaign, advanced in
    lock-step) driving the CUDA value/gradient kernel -- one batched launch per closure round.
    Same optimiser binary as the reference's botorch fit, so trajectories differ from the
    oracle's only through the last bits of f and g.  X [B, cap, d], y [B, cap] on the device.
    Returns (theta [B, d+3] cpu, nit, nfev, status)."""
    from scipy.optimize import _lbfgsb  # the compiled routine scipy.optimize.minimize drives

    B, cap, d = X.shape
    npar = d + 3
    m, maxls = 10, 20
    factr = 2.220446049250313e-09 / np.finfo(float).eps
    from scipy.linalg.lapack import HAS_ILP64
    int_dtype = np.int64 if HAS_ILP64 else np.int32
    lo = np.array([0.0 if v is None else v for v in lower]); hi = np.array([0.0 if v is None else v for v in upper])
    nbd = np.array([(1 if l is not None else 0) + (2 if u is not None else 0) for l, u in zip(lower, upper)])
    nbd = np.where(nbd == 3, 2, np.where(nbd == 2, 3, nbd)).astype(int_dtype)  # 1 lower, 2 both, 3 upper
    st = []
    for b in range(B):
        x = np.array(theta0, dtype=np.float64)
        x = np.where(nbd == 1, np.maximum(x, lo), x)
        st.append(dict(x=x, f=np.array(0.0), g=np.zeros(npar), wa=np.zeros(2 * m * npar + 5 * npar + 11 * m * m + 8 * m),
                       iwa=np.zeros(3 * npar, dtype=int_dtype), task=np.zeros(2, dtype=int_dtype),
                       ln_task=np.zeros(2, dtype=int_dtype), lsave=np.zeros(4, dtype=int_dtype),
                       isave=np.zeros(44, dtype=int_dtype), dsave=np.zeros(29), nit=0, nfev=0, done=False, status=0,
                       last_x=None))
    theta_dev = torch.empty(B, npar, dtype=torch.double, device=X.device)
    while True:
        need = []
        for b, s in enumerate(st):
            while not s["done"]:
                _lbfgsb.setulb(m, s["x"], lo, hi, nbd, s["f"], s["g"], factr, 1e-5, s["wa"], s["iwa"], s["task"],
                               s["lsave"], s["isave"], s["dsave"], maxls, s["ln_task"])
                if s["task"][0] == 3:
                    if s["last_x"] is not None and np.array_equal(s["last_x"], s["x"]):
                        continue  # scipy's ScalarFunction cache: same point, same f and g
                    need.append(b)
                    break
                if s["task"][0] == 1:
                    s["nit"] += 1
                    if s["nit"] >= maxiter or s["nfev"] > maxfun:
                        s["task"][0], s["task"][1] = 5, 504
                        s["status"] = 4
                    continue
                s["done"] = True
                if s["task"][0] != 4 and s["status"] == 0:
                    s["status"] = 5
        if not need:
            break
        host = np.stack([st[b]["x"] if b in need else np.array(theta0) for b in range(B)])
        theta_dev.copy_(torch.from_numpy(host))
        f, g, status = _lib.exact_mll_fg(X, y, theta_dev, prior, n=n)
        f, g = f.cpu().numpy(), g.cpu().numpy()
        for b in need:
            s = st[b]
            s["f"] = np.array(f[b]); s["g"] = g[b].astype(np.float64).copy(); s["nfev"] += 1
            s["last_x"] = s["x"].copy()
    theta = torch.tensor(np.stack([s["x"] for s in st]), dtype=torch.double)
    return theta, [s["nit"] for s in st], [s["nfev"] for s in st], [s["status"] for s in st]


def train_exact_model_botorch(model, X, y, y_standardize=True, optimizer: str = "device"):
    """Fit ExactGP hyper-parameters by L-BFGS-B on -(MLL + log-priors)/n (optim.py:11-32).

    As in the reference the standardisation of ``y`` is dead code: the marginal likelihood is
    evaluated on ``model.train_targets``, which hold the RAW targets.
    ``optimizer``: "device" = the CUDA-resident L-BFGS-B (stp_exact_fit); "scipy" = scipy's
    L-BFGS-B on the host driving the CUDA value/gradient kernel (parity mode)."""
    if y_standardize:
        y = TorchStandardScaler().fit_transform(y)  # unused, like optim.py:26-27
    model.train()
    model.likelihood.train()
    Xd, yd = _exact_batch(model)
    lower, upper = model.bounds
    theta = model.theta().reshape(1, -1).contiguous()
    if optimizer == "device":
        f, nit, nfev, status = _lib.exact_fit(Xd, yd, theta, lower, upper, model.prior_vector)
        info = dict(fun=float(f.item()), nit=int(nit.item()), nfev=int(nfev.item()), status=int(status.item()))
    elif optimizer == "scipy":
        th, nit, nfev, status = fit_exact_scipy_lockstep(Xd, yd, None, theta[0].cpu().tolist(), lower, upper,
                                                         model.prior_vector)
        theta = th.to(Xd.device)
        info = dict(nit=nit[0], nfev=nfev[0], status=status[0])
    else:
        raise ValueError("optimizer must be 'device' or 'scipy'")
    if info["status"] in (1, 2, 3):
        raise RuntimeError("exact GP fit failed: " + _lib.STATUS_TEXT[info["status"]])
    model.set_theta(theta[0])
    model.fit_info = info
    return None


def train_exact_model(model, X, y, epochs=100, lr=0.01, verbose=False, y_standardize=True):
    """Plain Adam on the exact MLL (optim.py:35-58): CUDA value/gradient + torch.optim.Adam."""
    if y_standardize:
        y = TorchStandardScaler().fit_transform(y)
    model.train()
    model.likelihood.train()
    # optim.py:49-51: the loss is -mll(model(X), y) with the y handed in -- STANDARDISED when y_standardize is set
    # (unlike the botorch path, whose closure reads model.train_targets, i.e. raw y: SURVEY App. D-1)
    dev = model.train_inputs[0].device
    Xd = X.detach().to(dev).double().unsqueeze(0).contiguous()
    yd = y.detach().to(dev).double().reshape(1, -1).contiguous()
    params = [model.likelihood.noise_covar.raw_noise, model.mean_module.raw_constant,
              model.covar_module.raw_outputscale, model.covar_module.base_kernel.raw_lengthscale]
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    history = []
    for _ in range(epochs):
        opt.zero_grad()
        f, g, status = _lib.exact_mll_fg(Xd, yd, model.theta().reshape(1, -1).contiguous(), model.prior_vector)
        if int(status.item()) != 0:
            raise RuntimeError("exact MLL evaluation failed: " + _lib.STATUS_TEXT[int(status.item())])
        history.append(float(f.item()))
        g = g[0].to(params[0].dtype)
        params[0].grad = g[0:1].clone()
        params[1].grad = g[1].clone()
        params[2].grad = g[2].clone()
        params[3].grad = g[3:].reshape(1, -1).clone()
        opt.step()
    if verbose:
        return history


def train_natural_variational_model(model, X, y, epochs=500, lr=0.05, verbose=False, y_standardize=True,
                                    init_noise=None):
    """NGD + Adam training of a variational model (optim.py:105-181), all ``epochs`` in one
    persistent CUDA launch (stp_var_fit).  ``lr`` is the NGD step (scaled by n, App. D-10);
    Adam's learning rate is the reference's hard-coded 0.01.  The 'best state' bookkeeping of
    optim.py:159-176 is a no-op upstream (App. D-6): the model after the last epoch is kept."""
    from . import varops
    losses = varops.train_natural(model, X, y, epochs=epochs, lr=lr, y_standardize=y_standardize,
                                  init_noise=init_noise)
    if verbose:
        print(int(np.argmin(losses)) if len(losses) else 0)
    return None


def train_variational_model(model, X, y, epochs=100, lr=0.01, verbose=False, y_standardize=True, init_noise=None):
    """Plain Adam on every parameter incl. the natural parameters of q(u) (optim.py:61-102): one CUDA
    value-and-gradient launch (stp_var_elbo_fg) per epoch + torch.optim.Adam on the packed state.  Not on the BO hot
    path (the loop uses train_natural_variational_model); kept so that the module surface is complete."""
    from . import varops
    losses = varops.train_adam(model, X, y, epochs=epochs, lr=lr, y_standardize=y_standardize, init_noise=init_noise)
    if verbose:
        return losses
    return None
