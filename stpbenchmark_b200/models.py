"""Drop-in model classes: ExactGP, VarGP, VarLGP, VarTGP (reference models.py:18-395).

Same constructor signatures, attribute tree and methods the reference's callers touch
(SURVEY.md section 8(b1)), with no gpytorch underneath: each class is a thin
``torch.nn.Module`` parameter holder whose fit / predict arithmetic runs in the sm_100a
library (stpbenchmark_b200/_lib.py -> libstp_b200.so).  Parameter names follow gpytorch's so
that ``state_dict()`` keys, ``named_parameters()`` order (the order botorch flattens them in:
noise, constant, outputscale, lengthscale) and attribute reads such as
``model.likelihood.noise`` or ``model.covar_module.base_kernel.lengthscale`` keep working.

All constraints in the reference are ``GreaterThan(lb, transform=None)``: the raw parameter IS
the value; the bound is only the L-BFGS-B box for ExactGP and is not enforced under Adam
(SURVEY App. A.2, D-5).  The Student-T degrees of freedom use gpytorch's default
``GreaterThan(2)`` with softplus: nu = 2 + softplus(raw_deg_free), initialised to nu = 7.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
from torch import nn

from . import priors

__all__ = ["ExactGP", "VarGP", "VarLGP", "VarTGP", "Posterior"]


def _compute_device(ref: torch.Tensor) -> torch.device:
    if ref.is_cuda:
        return ref.device
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return ref.device  # CPU-only box: parameters can be inspected, every compute call raises


class Posterior:
    """What ``model.posterior(X)`` / ``model(x)`` return: ``.mean``, ``.variance`` (and ``.df``
    for the Student-T likelihood), plus ``confidence_region()`` (mean -/+ 2 stddev)."""

    def __init__(self, mean, variance, df=None, covariance=None):
        self.mean = mean
        self.variance = variance
        self.df = df
        self._covariance = covariance

    @property
    def stddev(self):
        return self.variance.clamp_min(1e-12).sqrt() if self.variance is not None else None

    @property
    def covariance_matrix(self):
        return self._covariance

    def confidence_region(self):
        s2 = self.stddev * 2.0
        return self.mean - s2, self.mean + s2


class _ConstantMean(nn.Module):
    def __init__(self, device):
        super().__init__()
        self.raw_constant = nn.Parameter(torch.zeros((), device=device))

    @property
    def constant(self):
        return self.raw_constant

    def forward(self, x):
        return self.raw_constant.expand(x.shape[:-1])


class _RBFKernel(nn.Module):
    def __init__(self, d, init, device):
        super().__init__()
        self.ard_num_dims = d
        self.raw_lengthscale = nn.Parameter(torch.full((1, d), init, device=device))

    @property
    def lengthscale(self):
        return self.raw_lengthscale


class _ScaleKernel(nn.Module):
    def __init__(self, d, ls_init, os_init, device):
        super().__init__()
        self.raw_outputscale = nn.Parameter(torch.tensor(os_init, device=device))
        self.base_kernel = _RBFKernel(d, ls_init, device)

    @property
    def outputscale(self):
        return self.raw_outputscale

    def forward(self, x1, x2=None):
        """Dense prior covariance os * exp(-1/2 ||(x1-x2)/l||^2) (plain torch; not a hot path)."""
        x2 = x1 if x2 is None else x2
        l = self.base_kernel.raw_lengthscale
        diff = (x1.unsqueeze(-2) - x2.unsqueeze(-3)) / l
        return self.raw_outputscale * torch.exp(-0.5 * diff.pow(2).sum(-1))


class _HomoskedasticNoise(nn.Module):
    def __init__(self, init, device):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.tensor([init], device=device))


class _GaussianLikelihoodExact(nn.Module):
    """gpytorch.likelihoods.GaussianLikelihood as ExactGP owns it (noise_covar.raw_noise)."""

    kind = "gaussian"

    def __init__(self, init, device):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(init, device)

    @property
    def noise(self):
        return self.noise_covar.raw_noise


class _VarLikelihood(nn.Module):
    """Gaussian / Student-T / Laplace likelihood of the variational models (raw_noise[, raw_deg_free])."""

    def __init__(self, kind, init, device):
        super().__init__()
        self.kind = kind
        if kind == "gaussian":
            self.noise_covar = _HomoskedasticNoise(init, device)
        else:
            self.raw_noise = nn.Parameter(torch.tensor([init], device=device))
        if kind == "student_t":
            raw = priors.DF_INIT - 2.0
            raw = raw + math.log(-math.expm1(-raw))  # inverse softplus
            self.raw_deg_free = nn.Parameter(torch.tensor([raw], device=device))

    @property
    def noise(self):
        return self.noise_covar.raw_noise if self.kind == "gaussian" else self.raw_noise

    @property
    def deg_free(self):
        return torch.nn.functional.softplus(self.raw_deg_free) + 2.0


class _NaturalVariationalDistribution(nn.Module):
    def __init__(self, m, device):
        super().__init__()
        self.natural_vec = nn.Parameter(torch.zeros(m, device=device))
        self.natural_mat = nn.Parameter(-0.5 * torch.eye(m, device=device))
        self.mean_init_std = 1e-3


class _VariationalStrategy(nn.Module):
    def __init__(self, inducing_points, device):
        super().__init__()
        self.inducing_points = nn.Parameter(inducing_points.detach().clone().to(device))
        self._variational_distribution = _NaturalVariationalDistribution(inducing_points.size(0), device)
        self.register_buffer("variational_params_initialized", torch.tensor(0))
        self.jitter_val = priors.VAR_JITTER


class _GPBase(nn.Module):
    num_outputs = 1

    def hyperparameters(self):
        for name, p in self.named_parameters():
            if "_variational_distribution" not in name:
                yield p

    def variational_parameters(self):
        for name, p in self.named_parameters():
            if "_variational_distribution" in name:
                yield p

    @property
    def device(self):
        return next(self.parameters()).device


class ExactGP(_GPBase):
    """Exact GP regression: ConstantMean + ScaleKernel(RBF, ARD), Gaussian noise, LogNormal
    priors incl. the dimension-scaled lengthscale prior (reference models.py:300-395)."""

    def __init__(self, X_train, y_train, input_transform=None):
        super().__init__()
        dev = _compute_device(X_train)
        d = X_train.shape[1]
        self.train_inputs = (X_train.detach().to(dev),)
        self.train_targets = y_train.detach().to(dev)
        self.input_transform = input_transform
        ls_prior = priors.exact_lengthscale_prior(d)
        self.likelihood = _GaussianLikelihoodExact(priors.lognormal_mode(*priors.NOISE_PRIOR), dev)
        self.mean_module = _ConstantMean(dev)
        self.covar_module = _ScaleKernel(d, priors.lognormal_mode(*ls_prior),
                                         priors.lognormal_mode(*priors.OUTPUTSCALE_PRIOR), dev)
        self.prior_vector = priors.exact_prior_vector(d)
        self.bounds = priors.exact_bounds(d)
        self.fit_info = None

    # -- flat parameter vector in botorch's order ------------------------------------------
    def theta(self) -> torch.Tensor:
        return torch.cat([self.likelihood.noise.detach().reshape(1), self.mean_module.raw_constant.detach().reshape(1),
                          self.covar_module.raw_outputscale.detach().reshape(1),
                          self.covar_module.base_kernel.raw_lengthscale.detach().reshape(-1)]).double()

    def set_theta(self, theta: torch.Tensor) -> None:
        with torch.no_grad():
            t = theta.to(self.device)
            self.likelihood.noise_covar.raw_noise.copy_(t[0:1])
            self.mean_module.raw_constant.copy_(t[1])
            self.covar_module.raw_outputscale.copy_(t[2])
            self.covar_module.base_kernel.raw_lengthscale.copy_(t[3:].reshape(1, -1))

    def transform_inputs(self, X=None, input_transform=None):
        if X is None:
            raise ValueError("X must be provided")
        tf = input_transform if input_transform is not None else self.input_transform
        return tf(X) if tf is not None else X

    def forward(self, x):
        """Prior at x (models.py:379-385)."""
        if self.training:
            x = self.transform_inputs(x)
        x = x.to(self.device)
        return Posterior(self.mean_module(x), None, covariance=self.covar_module(x))

    def _predict(self, X, observation_noise):
        from . import _lib
        Xq = X.detach().to(self.device).double()
        pts = Xq.reshape(-1, Xq.shape[-1]).contiguous()
        Xt = self.train_inputs[0].double().unsqueeze(0).contiguous()
        yt = self.train_targets.double().reshape(1, -1).contiguous()
        out = _lib.exact_acq(pts.unsqueeze(0), Xt, yt, self.theta().reshape(1, -1).contiguous(), want_pointwise=True)
        if int(out["status"].item()) != 0:
            raise RuntimeError("ExactGP prediction failed: " + _lib.STATUS_TEXT[int(out["status"].item())])
        mean, var = out["mean"][0], out["var"][0]
        if not observation_noise:
            var = var - self.likelihood.noise.detach().double()
        shape = Xq.shape[:-1]
        return Posterior(mean.reshape(shape), var.reshape(shape))

    def __call__(self, x, **kwargs):
        if self.training:
            return self.forward(x)
        return self._predict(x, observation_noise=False)

    def posterior(self, X, observation_noise=True, num_samples=2048, **kwargs):
        """Predictive at X ([b,1,d] or [Nt,d]); mean/variance per point; the likelihood noise is
        included by default, as at models.py:387-395."""
        self.eval()
        return self._predict(X, observation_noise)


class _VarBase(_GPBase):
    _kind = "gaussian"

    def _setup(self, inducing_points, lengthscale_prior):
        dev = _compute_device(inducing_points)
        d = inducing_points.size(1)
        if lengthscale_prior is None:
            self.lengthscale_prior = priors.VAR_LENGTHSCALE_PRIOR
        elif isinstance(lengthscale_prior, (tuple, list)):
            self.lengthscale_prior = (float(lengthscale_prior[0]), float(lengthscale_prior[1]))
        else:  # a gpytorch-style LogNormalPrior object
            self.lengthscale_prior = (float(lengthscale_prior.loc), float(lengthscale_prior.scale))
        self.variational_strategy = _VariationalStrategy(inducing_points, dev)
        self.mean_module = _ConstantMean(dev)
        self.covar_module = _ScaleKernel(d, priors.lognormal_mode(*self.lengthscale_prior),
                                         priors.lognormal_mode(*priors.OUTPUTSCALE_PRIOR), dev)
        self.likelihood = _VarLikelihood(self._kind, priors.lognormal_mode(*priors.NOISE_PRIOR), dev)
        self.num_likelihood_samples = 2048
        self.fit_info = None
        self._y_scaler = None

    def forward(self, x):
        x = x.to(self.device)
        return Posterior(self.mean_module(x), None, covariance=self.covar_module(x))

    def _latent(self, X):
        from . import varops
        return varops.latent_posterior(self, X)

    def __call__(self, x, **kwargs):
        mean, var = self._latent(x)
        return Posterior(mean, var)

    def posterior(self, X, observation_noise=True, num_samples=2048, **kwargs):
        """models.py:95-111 / 187-202 / 280-297.  VarGP: N(m, v + noise).  VarTGP / VarLGP:
        ``num_samples`` draws f = m + sqrt(v) eps wrapped in the likelihood: ``.mean`` is the
        [S, ...] sample tensor, ``.variance`` the (constant) likelihood variance (App. A.5)."""
        from . import varops
        self.num_likelihood_samples = num_samples
        self.eval()
        return varops.posterior(self, X, observation_noise, num_samples, eps=kwargs.get("eps"))


class VarGP(_VarBase):
    """Whitened SVGP, Gaussian likelihood, learned inducing locations (reference models.py:114-202)."""

    _kind = "gaussian"

    def __init__(self, inducing_points, lengthscale_prior=None):
        super().__init__()
        self._setup(inducing_points, lengthscale_prior)


class VarTGP(_VarBase):
    """Whitened SVGP with Student-T likelihood (reference models.py:18-111)."""

    _kind = "student_t"

    def __init__(self, inducing_points, lengthscale_prior=None):
        super().__init__()
        self._setup(inducing_points, lengthscale_prior)


class VarLGP(_VarBase):
    """Whitened SVGP with Laplace likelihood (reference models.py:205-297)."""

    _kind = "laplace"

    def __init__(self, inducing_points, lengthscale_prior=None):
        super().__init__()
        self._setup(inducing_points, lengthscale_prior)
