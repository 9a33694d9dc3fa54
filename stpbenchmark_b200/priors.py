"""Prior, bound and initial-value constants of the four reference models (models.py).

Every model is built in fp32 and then cast with ``.double()`` (runners.py:74), so prior
parameters -- and ``prior.mode``, which becomes the initial parameter value through
``GreaterThan(..., transform=None, initial_value=prior.mode)`` -- are fp32 numbers widened to
fp64.  These helpers reproduce that rounding with torch fp32 arithmetic.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch


def f32(v: float) -> float:
    return float(torch.tensor(v, dtype=torch.float32).item())


def lognormal_mode(loc: float, scale: float) -> float:
    """LogNormalPrior(loc, scale).mode = exp(loc - scale^2), evaluated in fp32."""
    lo = torch.tensor(loc, dtype=torch.float32)
    sc = torch.tensor(scale, dtype=torch.float32)
    return float((lo - sc.square()).exp().item())


NOISE_PRIOR = (-4.0, 1.0)          # LogNormalPrior(loc=-4, scale=1)       models.py:319, 74, 166, 261
OUTPUTSCALE_PRIOR = (0.0, 1.0)     # LogNormalPrior(0, 1)                  models.py:351, 65, 157, 251
VAR_LENGTHSCALE_PRIOR = (0.0, 1.0)  # lengthscale_prior=None default        models.py:50, 144, 237
NOISE_LB = 1e-4
OUTPUTSCALE_LB = 1e-4
LENGTHSCALE_LB = 2.5e-2
GAMMA_DF_PRIOR = (2.0, 0.1)        # GammaPrior(concentration, rate) on nu  models.py:75
DF_INIT = 7.0                      # gpytorch StudentTLikelihood's initial deg_free
VAR_JITTER = 1e-6                  # gpytorch variational Cholesky jitter, fp64


def exact_lengthscale_prior(d: int) -> Tuple[float, float]:
    """LogNormalPrior(sqrt(2) + 0.5 log d, sqrt(3)) of models.py:338-340."""
    return (math.sqrt(2) + math.log(d) * 0.5, math.sqrt(3))


def exact_prior_vector(d: int) -> List[float]:
    """(noise_loc, noise_scale, os_loc, os_scale, ls_loc, ls_scale), fp32-rounded -- the
    `prior` argument of stp_exact_mll_fg / stp_exact_fit / stp_exact_bo_step."""
    ls = exact_lengthscale_prior(d)
    return [f32(NOISE_PRIOR[0]), f32(NOISE_PRIOR[1]), f32(OUTPUTSCALE_PRIOR[0]), f32(OUTPUTSCALE_PRIOR[1]),
            f32(ls[0]), f32(ls[1])]


def exact_theta0(d: int) -> List[float]:
    """(noise, constant, outputscale, lengthscale x d) at construction: prior modes, mean 0."""
    ls = exact_lengthscale_prior(d)
    return [lognormal_mode(*NOISE_PRIOR), 0.0, lognormal_mode(*OUTPUTSCALE_PRIOR)] + [lognormal_mode(*ls)] * d


def exact_bounds(d: int):
    """L-BFGS-B box from the GreaterThan constraints (models.py:322-326, 345-347, 355-357)."""
    lower = [NOISE_LB, None, OUTPUTSCALE_LB] + [LENGTHSCALE_LB] * d
    upper = [None] * (d + 3)
    return lower, upper
