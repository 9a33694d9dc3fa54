"""Oracle (CPU, fp64): exact GP marginal likelihood, its L-BFGS-B fit and the predictive.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates what the reference reaches through
``models.ExactGP`` (models.py:300-395), ``train_exact_model_botorch`` (optim.py:11-32) and
``ExactGP.posterior`` (models.py:387-395); formulas: SURVEY.md App. A.0-A.3, A.5.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
from scipy.optimize import minimize

LOG_2PI = math.log(2.0 * math.pi)


def _f32(v: float) -> float:
    """Round a Python float through fp32 -- prior buffers are built in fp32 and widened by
    ``model.double()`` (runners.py:74)."""
    return float(torch.tensor(v, dtype=torch.float32).item())


def _lognormal_mode_f32(loc: float, scale: float) -> float:
    """``LogNormalPrior.mode`` evaluated in fp32, as at construction time (models.py:319-347)."""
    lo = torch.tensor(loc, dtype=torch.float32)
    sc = torch.tensor(scale, dtype=torch.float32)
    return float((lo - sc.square()).exp().item())


@dataclass(frozen=True)
class ExactPriors:
    """LogNormal prior parameters and bounds of ExactGP for input dimension d (models.py:319-358)."""

    noise_loc: float
    noise_scale: float
    os_loc: float
    os_scale: float
    ls_loc: float
    ls_scale: float
    noise_lb: float = 1e-4
    os_lb: float = 1e-4
    ls_lb: float = 2.5e-2

    @staticmethod
    def for_dim(d: int) -> "ExactPriors":
        return ExactPriors(
            noise_loc=_f32(-4.0), noise_scale=_f32(1.0),
            os_loc=_f32(0.0), os_scale=_f32(1.0),
            ls_loc=_f32(math.sqrt(2) + math.log(d) * 0.5), ls_scale=_f32(math.sqrt(3)),
        )


def initial_theta(d: int) -> np.ndarray:
    """theta0 = (noise, constant, outputscale, lengthscale_1..d): prior modes, mean 0 (App. A.2)."""
    return np.array(
        [_lognormal_mode_f32(-4.0, 1.0), 0.0, _lognormal_mode_f32(0.0, 1.0)]
        + [_lognormal_mode_f32(math.sqrt(2) + math.log(d) * 0.5, math.sqrt(3))] * d,
        dtype=np.float64,
    )


def theta_bounds(d: int):
    """Box handed to L-BFGS-B: the ``GreaterThan`` constraints (transform=None) of models.py:322-357."""
    return [(1e-4, None), (None, None), (1e-4, None)] + [(2.5e-2, None)] * d


def lognormal_logpdf(x: torch.Tensor, loc: float, scale: float) -> torch.Tensor:
    lx = torch.log(x)
    return -lx - math.log(scale) - 0.5 * LOG_2PI - (lx - loc) ** 2 / (2.0 * scale * scale)


def sq_dist(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """Squared Euclidean distance the way gpytorch's ``Kernel.covar_dist`` forms it (App. A.1):
    centre on the row mean of x1, one augmented matmul, clamp at zero."""
    shift = x1.mean(-2, keepdim=True)
    a = x1 - shift
    b = x2 - shift
    a2 = a.pow(2).sum(-1, keepdim=True)
    b2 = b.pow(2).sum(-1, keepdim=True)
    lhs = torch.cat([-2.0 * a, a2, torch.ones_like(a2)], dim=-1)
    rhs = torch.cat([b, torch.ones_like(b2), b2], dim=-1)
    return lhs.matmul(rhs.transpose(-1, -2)).clamp_min(0)


def rbf(x1: torch.Tensor, x2: torch.Tensor, ls: torch.Tensor, os: torch.Tensor) -> torch.Tensor:
    """os * exp(-0.5 * ||(x1 - x2) / ls||^2): ScaleKernel(RBFKernel(ard)) (models.py:342-358)."""
    return os * torch.exp(-0.5 * sq_dist(x1 / ls, x2 / ls))


def matern52(x1: torch.Tensor, x2: torch.Tensor, ls: torch.Tensor, os: torch.Tensor) -> torch.Tensor:
    """[UPSTREAM] ScaleKernel(MaternKernel(nu=2.5, ard)) as gpytorch evaluates it: r = sqrt(max(sq_dist, 1e-30)) of
    the lengthscale-scaled inputs, os (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r).  The reference instantiates RBF only
    (models.py:342-358); Matern is the alternative BASELINE.json's north_star names."""
    r = sq_dist(x1 / ls, x2 / ls).clamp_min(1e-30).sqrt()
    s5 = math.sqrt(5.0) * r
    return os * (1.0 + s5 + (5.0 / 3.0) * r * r) * torch.exp(-s5)


KERNELS = {"rbf": rbf, "matern52": matern52}

PSD_SAFE_JITTERS = (0.0, 1e-8, 1e-7, 1e-6)   # linear_operator.utils.cholesky.psd_safe_cholesky in fp64: 3 tries, x10


def psd_safe_cholesky(K: torch.Tensor) -> torch.Tensor:
    """[UPSTREAM] what gpytorch factorises with (every Cholesky of the path goes through it): a plain Cholesky; if
    that fails, the same matrix with 1e-8, then 1e-7, then 1e-6 added to its diagonal; then NotPSDError."""
    eye = torch.eye(K.shape[-1], dtype=K.dtype)
    for k, jit in enumerate(PSD_SAFE_JITTERS):
        L, info = torch.linalg.cholesky_ex(K if jit == 0.0 else K + jit * eye)
        if int(info) == 0:
            return L
    raise torch.linalg.LinAlgError("matrix not positive definite after jitter 1e-6")


def neg_mll(theta: torch.Tensor, X: torch.Tensor, y: torch.Tensor, pri: ExactPriors | None = None,
            kernel: str = "rbf") -> torch.Tensor:
    """-(log N(y | c, K + noise I) + sum log-priors) / n, y RAW (App. A.3, App. D-1)."""
    n, d = X.shape
    pri = pri or ExactPriors.for_dim(d)
    noise, c, os, ls = theta[0], theta[1], theta[2], theta[3:]
    K = KERNELS[kernel](X, X, ls, os) + noise * torch.eye(n, dtype=X.dtype)
    L = psd_safe_cholesky(K)
    r = (y - c).unsqueeze(-1)
    a = torch.cholesky_solve(r, L)
    quad = (r * a).sum()
    logdet = 2.0 * torch.log(torch.diagonal(L)).sum()
    total = -0.5 * (quad + logdet + n * LOG_2PI)
    total = total + lognormal_logpdf(noise, pri.noise_loc, pri.noise_scale)
    total = total + lognormal_logpdf(os, pri.os_loc, pri.os_scale)
    total = total + lognormal_logpdf(ls, pri.ls_loc, pri.ls_scale).sum()
    return -(total / n)


def mll_value_and_grad(theta_np: np.ndarray, X: torch.Tensor, y: torch.Tensor, kernel: str = "rbf"):
    """One L-BFGS-B closure call: loss and autograd gradient (optim.py:32 via botorch)."""
    t = torch.tensor(np.asarray(theta_np, dtype=np.float64), dtype=torch.double, requires_grad=True)
    try:
        loss = neg_mll(t, X, y, kernel=kernel)
        loss.backward()
        return float(loss.item()), t.grad.numpy().astype(np.float64)
    except Exception:  # non-PD matrix: report NaN like a failed closure
        return float("nan"), np.full(t.shape, np.nan)


def mll_value_and_grad_analytic(theta_np: np.ndarray, X: torch.Tensor, y: torch.Tensor):
    """Same closure with the hand-derived gradient of App. A.3 (no autograd), direct-difference
    kernel.  Second, independent statement used to cross-check the CUDA formula."""
    Xn = X.numpy()
    yn = y.numpy()
    n, d = Xn.shape
    pri = ExactPriors.for_dim(d)
    noise, c, os = theta_np[0], theta_np[1], theta_np[2]
    ls = np.asarray(theta_np[3:], dtype=np.float64)
    diff = Xn[:, None, :] - Xn[None, :, :]
    D2 = (diff / ls) ** 2
    K0 = os * np.exp(-0.5 * D2.sum(-1))
    K = K0 + noise * np.eye(n)
    L = np.linalg.cholesky(K)
    Kinv = np.linalg.inv(K)
    r = yn - c
    alpha = Kinv @ r
    logN = -0.5 * (r @ alpha + 2.0 * np.log(np.diag(L)).sum() + n * LOG_2PI)
    W = np.outer(alpha, alpha) - Kinv

    def lp(x, lo, sc):
        return -np.log(x) - math.log(sc) - 0.5 * LOG_2PI - (np.log(x) - lo) ** 2 / (2 * sc * sc)

    def dlp(x, lo, sc):
        return -(1.0 + (np.log(x) - lo) / (sc * sc)) / x

    total = logN + lp(noise, pri.noise_loc, pri.noise_scale) + lp(os, pri.os_loc, pri.os_scale) \
        + lp(ls, pri.ls_loc, pri.ls_scale).sum()
    g = np.empty(d + 3)
    g[0] = 0.5 * np.trace(W) + dlp(noise, pri.noise_loc, pri.noise_scale)
    g[1] = alpha.sum()
    g[2] = 0.5 * (W * K0).sum() / os + dlp(os, pri.os_loc, pri.os_scale)
    for k in range(d):
        g[3 + k] = 0.5 * (W * K0 * diff[:, :, k] ** 2).sum() / ls[k] ** 3 + dlp(ls[k], pri.ls_loc, pri.ls_scale)
    return -total / n, -g / n


def fit_exact(X: torch.Tensor, y: torch.Tensor, theta0: np.ndarray | None = None, kernel: str = "rbf"):
    """scipy L-BFGS-B with scipy's defaults over (noise, c, os, ls) (App. A.3).  Returns
    (theta*, OptimizeResult)."""
    d = X.shape[1]
    x0 = initial_theta(d) if theta0 is None else np.asarray(theta0, dtype=np.float64)
    res = minimize(lambda t: mll_value_and_grad(t, X, y, kernel), x0, jac=True, method="L-BFGS-B",
                   bounds=theta_bounds(d))
    return res.x.copy(), res


def posterior_exact(theta: np.ndarray, X: torch.Tensor, y: torch.Tensor, Xc: torch.Tensor, kernel: str = "rbf"):
    """Noisy predictive at every candidate (models.py:387-395, App. A.5):
    mean = c + k^T alpha, var = os - ||L^-1 k||^2 + noise."""
    t = torch.as_tensor(theta, dtype=torch.double)
    noise, c, os, ls = t[0], t[1], t[2], t[3:]
    n = X.shape[0]
    kfn = KERNELS[kernel]
    K = kfn(X, X, ls, os) + noise * torch.eye(n, dtype=X.dtype)
    L = torch.linalg.cholesky(K)
    alpha = torch.cholesky_solve((y - c).unsqueeze(-1), L).squeeze(-1)
    Kc = kfn(Xc, X, ls, os)
    mean = c + Kc @ alpha
    V = torch.linalg.solve_triangular(L, Kc.T, upper=False)
    var = os - (V * V).sum(0) + noise
    return mean, var
