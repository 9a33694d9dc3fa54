"""Oracle (CPU, fp64): botorch's analytic LogExpectedImprovement, restated.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  The reference calls
``LogExpectedImprovement(model, best_f=y_train.max(), maximize=True).forward(X[Nc,1,d])``
(runners.py:88-90); botorch is not installed, so its ``_log_ei_helper`` is restated from
SURVEY.md App. A.5.  The acquisition over MC models is the mean over the sample axis
(runners.py:92-93); argmax takes the first maximal index and lets NaN win (runners.py:95).
"""
from __future__ import annotations

import math

import torch

_LOG_2PI = math.log(2.0 * math.pi)
_INV_SQRT2 = 1.0 / math.sqrt(2.0)
_NEG_INV_SQRT2 = -_INV_SQRT2
_LOG_SQRT_PI_OVER_2 = math.log(math.sqrt(math.pi / 2.0))
_INV_SQRT_2PI = 1.0 / math.sqrt(2.0 * math.pi)


def _phi(u):
    return _INV_SQRT_2PI * torch.exp(-0.5 * u * u)


def _Phi(u):
    return 0.5 * torch.erfc(_NEG_INV_SQRT2 * u)


def log1mexp(x: torch.Tensor) -> torch.Tensor:
    """log(1 - exp(x)) for x < 0: log(-expm1(x)) above -ln 2, log1p(-exp(x)) below."""
    return torch.where(-math.log(2.0) < x, (-torch.expm1(x)).log(), (-torch.exp(x)).log1p())


def log_ei_helper(u: torch.Tensor) -> torch.Tensor:
    """log(phi(u) + u Phi(u)) in three branches (u > -1; -1e6 < u <= -1; u <= -1e6)."""
    bound = -1.0
    u_hi = u.masked_fill(u < bound, bound)
    val_hi = (_phi(u_hi) + u_hi * _Phi(u_hi)).log()
    neg_inv_sqrt_eps = -1e6  # fp64 branch point
    u_lo = u.masked_fill(u > bound, bound)
    u_eps = u_lo.masked_fill(u < neg_inv_sqrt_eps, neg_inv_sqrt_eps)
    w = torch.log(torch.special.erfcx(_NEG_INV_SQRT2 * u_eps) * u_eps.abs()) + _LOG_SQRT_PI_OVER_2
    val_lo = -0.5 * (u * u + _LOG_2PI) + torch.where(
        u > neg_inv_sqrt_eps, log1mexp(w), -2.0 * u_lo.abs().log()
    )
    return torch.where(u > bound, val_hi, val_lo)


def log_ei(mean: torch.Tensor, var: torch.Tensor, best_f) -> torch.Tensor:
    """LogEI = log_h((mean - best_f) / sigma) + log sigma, sigma = sqrt(max(var, 1e-12))."""
    sigma = var.clamp_min(1e-12).sqrt()
    sigma = sigma.expand_as(mean) if sigma.shape != mean.shape else sigma
    u = (mean - best_f) / sigma
    return log_ei_helper(u) + sigma.log()


def ucb(mean: torch.Tensor, var: torch.Tensor, beta: float) -> torch.Tensor:
    """[UPSTREAM] botorch UpperConfidenceBound(maximize=True): mean + sqrt(beta) sigma, sigma = sqrt(max(var, 1e-12))
    (botorch's AnalyticAcquisitionFunction._mean_and_sigma clamps the variance at 1e-12).  north_star names UCB; the
    reference's loop itself uses LogEI only (runners.py:88-90)."""
    return mean + math.sqrt(beta) * var.clamp_min(1e-12).sqrt()


def ei(mean: torch.Tensor, var: torch.Tensor, best_f) -> torch.Tensor:
    """utils.EI (utils.py:299-309) = botorch ExpectedImprovement: improvement Phi(z) + sigma phi(z), z = improvement /
    sigma, with botorch's sigma clamp so that it is defined on the whole pool."""
    sigma = var.clamp_min(1e-12).sqrt()
    gain = mean - best_f
    z = gain / sigma
    return gain * _Phi(z) + sigma * _phi(z)


def log_ei_t(mean, sd, df, best_f):
    """acqf.log_expected_improvement_t's elementwise formula (acqf.py:34-50) with scipy.stats.t, as the reference
    evaluates it on the host: log(gain) + logcdf(Z) where gain > 0, logpdf(Z) + log(sd) elsewhere, -inf at sd <= 1e-9."""
    import numpy as np
    from scipy.stats import t as student_t
    mean, sd, df = (np.asarray(a, dtype=np.float64) for a in (mean, sd, df))
    gain = mean - float(best_f)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = gain / sd
        dist = student_t(df=df)
        out = np.where(gain > 0, np.log(gain) + dist.logcdf(z), dist.logpdf(z) + np.log(sd))
    return np.where(sd > 1e-9, out, -np.inf)


def first_argmax(acq: torch.Tensor) -> int:
    """torch.argmax semantics: first maximal index, NaN wins (runners.py:95)."""
    return int(torch.argmax(acq).item())
