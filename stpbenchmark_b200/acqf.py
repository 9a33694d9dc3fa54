"""Acquisition functions.

``LogExpectedImprovement``            replaces ``botorch.acquisition.LogExpectedImprovement`` as the
                                      campaign loop uses it (runners.py:8, 88-90): analytic LogEI of
                                      ``model.posterior(X)``; evaluated by the CUDA library.
``log_expected_improvement``          the reference's hand-rolled Gaussian LogEI (acqf.py:55-109)
``log_expected_improvement_t``        the reference's Student-t LogEI (acqf.py:7-52)
The last two are never called by the reference's loop but are part of its
public surface; they keep their exact numerics (eps = 1e-9, ``log(Phi + eps)``, zeros / -inf at
sigma <= eps) and are checked against the reference's own file in tests/test_acqf_host.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch

__all__ = ["LogExpectedImprovement", "UpperConfidenceBound", "ExpectedImprovement", "log_expected_improvement",
           "log_expected_improvement_t"]


class LogExpectedImprovement:
    """LogEI = log_h((mean - best_f) / sigma) + log sigma, sigma = sqrt(max(var, 1e-12)).

    ``forward(X[b, 1, d])`` returns ``[b]`` for analytic models and ``[S, b]`` for the Monte-Carlo
    models (VarTGP / VarLGP), whose ``posterior().mean`` is an [S, b, 1] sample tensor and whose
    variance is the constant likelihood variance."""

    def __init__(self, model, best_f, posterior_transform=None, maximize: bool = True):
        if getattr(model, "num_outputs", 1) != 1:
            raise ValueError("LogExpectedImprovement needs a single-output model")
        self.model = model
        self.best_f = torch.as_tensor(best_f, dtype=torch.double)
        self.maximize = maximize

    _kind, _beta = "logei", 0.0

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        from . import _lib
        post = self.model.posterior(X=X, posterior_transform=None)
        mean = post.mean
        if mean.dim() >= 2 and mean.shape[-1] == 1 and X.dim() >= 3:
            mean = mean.squeeze(-1)
        if mean.dim() >= 2 and mean.shape[-1] == 1 and X.dim() >= 3:
            mean = mean.squeeze(-1)
        var = post.variance
        var = var.expand_as(post.mean).reshape(mean.shape) if var.numel() != mean.numel() else var.reshape(mean.shape)
        best = self.best_f.to(mean.device)
        if not self.maximize:
            mean, best = -mean, -best
        if self._kind == "logei":
            return _lib.log_ei(mean.contiguous().double(), var.contiguous().double(), float(best))
        return _lib.acq_values(mean.contiguous().double(), var.contiguous().double(), float(best), self._kind, self._beta)

    __call__ = forward


class ExpectedImprovement(LogExpectedImprovement):
    """botorch ``ExpectedImprovement`` = utils.EI (reference utils.py:299-309) over ``model.posterior``:
    sigma (phi(u) + u Phi(u)), evaluated by the CUDA library (stp_acq_values)."""

    _kind = "ei"


class UpperConfidenceBound(LogExpectedImprovement):
    """botorch ``UpperConfidenceBound``: (mean if maximize else -mean) + sqrt(beta) sigma,
    sigma = sqrt(max(var, 1e-12)); evaluated by the CUDA library (stp_acq_values)."""

    _kind = "ucb"

    def __init__(self, model, beta, posterior_transform=None, maximize: bool = True):
        super().__init__(model, 0.0, posterior_transform, maximize)
        self._beta = float(beta)


def log_expected_improvement(model, test_X: torch.Tensor, best_f: torch.Tensor) -> torch.Tensor:
    """Reference acqf.py:55-109, same numerics, evaluated on the device the posterior lives on."""
    with torch.no_grad():
        post = model.posterior(test_X)
        mean = post.mean
        std = post.variance.sqrt()
    if len(mean.shape) > 1:  # MC likelihood samples: average over the sample axis (acqf.py:77-79)
        mean = torch.mean(mean, dim=0)
        std = torch.mean(std.expand(post.mean.shape) if std.shape != post.mean.shape else std, dim=0)
    eps = 1e-9
    std = std + eps
    best_f = torch.as_tensor(best_f).to(mean.device)
    gain = mean - best_f
    z = gain / std
    positive = gain > 0
    safe_gain = torch.where(positive, gain, torch.ones_like(gain))
    log_phi = -0.5 * z * z - math.log(math.sqrt(2.0 * math.pi))
    Phi = 0.5 * (1.0 + torch.erf(z / math.sqrt(2.0)))
    out = torch.where(positive, torch.log(safe_gain) + torch.log(Phi + eps), log_phi + torch.log(std))
    return torch.where(std > eps, out, torch.zeros_like(gain))


def log_expected_improvement_t(model, test_X: torch.Tensor, best_f: torch.Tensor) -> torch.Tensor:
    """Reference acqf.py:7-52: Student-t LogEI through scipy.stats.t on the host (by
    construction a CPU/numpy function in the reference as well)."""
    from scipy.stats import t as student_t

    with torch.no_grad():
        post = model.posterior(test_X)
        mean, std, df = post.mean, post.variance.sqrt(), post.df
    if len(mean.shape) > 1:
        std = std.expand(mean.shape) if std.shape != mean.shape else std
        df = df.expand(mean.shape) if df.shape != mean.shape else df
        mean, std, df = torch.mean(mean, dim=0), torch.mean(std, dim=0), torch.mean(df, dim=0)
    if mean.is_cuda:   # device form (stp_log_ei_t): the same formula without the round trip through scipy
        from . import _lib
        out = _lib.log_ei_t(mean.double().contiguous(), std.double().contiguous(), df.double().contiguous(),
                            float(torch.as_tensor(best_f)))
        return out.to(device=test_X.device, dtype=test_X.dtype)
    mean_np, std_np = mean.detach().cpu().numpy(), std.detach().cpu().numpy()
    df_np = df.detach().cpu().numpy()
    best_np = torch.as_tensor(best_f).detach().cpu().numpy()
    gain = mean_np - best_np
    with np.errstate(divide="ignore", invalid="ignore"):
        z = gain / std_np
        dist = student_t(df=df_np)
        out = np.where(gain > 0, np.log(gain) + dist.logcdf(z), dist.logpdf(z) + np.log(std_np))
    out = np.where(std_np > 1e-9, out, -np.inf)
    return torch.tensor(out, device=test_X.device, dtype=test_X.dtype)
