"""Oracle (CPU, fp64): whitened SVGP with natural-gradient + Adam training, and its acquisition.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Restates what the reference reaches through
``models.VarGP / VarTGP / VarLGP`` (models.py:114-202, 18-111, 205-297),
``train_natural_variational_model`` (optim.py:105-181) and ``Var*.posterior``
(models.py:95-111, 187-202, 280-297); formulas: SURVEY.md App. A.4, A.5.  The forward pass
is written the way gpytorch evaluates it (explicit S and chol S) and differentiated with
autograd, so it is independent of the hand-derived backward used by the CUDA path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from .acq import log_ei_helper
from .exact import LOG_2PI, _f32, _lognormal_mode_f32, lognormal_logpdf, psd_safe_cholesky, rbf

JITTER = 1e-6  # gpytorch's variational Cholesky jitter in fp64 (App. A.4 step 2)
KINDS = ("VarGP", "VarTGP", "VarLGP")

_gh_x, _gh_w = np.polynomial.hermite.hermgauss(20)
# Gauss-Hermite nodes/weights pass through fp32 upstream (App. A.4 step 4, App. D-9).
GH_X = torch.tensor(_gh_x, dtype=torch.float32).double()
GH_W = torch.tensor(_gh_w, dtype=torch.float32).double()

GAMMA_CONC = _f32(2.0)   # GammaPrior(2.0, 0.1) on the Student-T degrees of freedom (models.py:75)
GAMMA_RATE = _f32(0.1)


def inv_softplus(x: float) -> float:
    return x + math.log(-math.expm1(-x))


def standardise(y: torch.Tensor) -> torch.Tensor:
    """optim.py:126-127 via utils.TorchStandardScaler: population std, +1e-10."""
    return (y - y.mean(0, keepdim=True)) / (y.std(0, unbiased=False, keepdim=True) + 1e-10)


@dataclass
class VarState:
    """Everything a variational model owns (App. A.2 'Var*' rows)."""

    kind: str
    Z: torch.Tensor          # [n, d] inducing locations, learned
    c: torch.Tensor          # [1]
    ls: torch.Tensor         # [1, d]
    os: torch.Tensor         # []
    noise: torch.Tensor      # [1]
    rawdf: torch.Tensor | None  # [1], VarTGP only; nu = 2 + softplus(rawdf)
    nat1: torch.Tensor       # [n]
    nat2: torch.Tensor       # [n, n]
    adam: torch.optim.Adam | None = None
    losses: list = field(default_factory=list)

    def hyper(self):
        h = [self.Z, self.c, self.ls, self.os, self.noise]
        if self.rawdf is not None:
            h.append(self.rawdf)
        return h

    def df(self):
        return torch.nn.functional.softplus(self.rawdf) + 2.0


def init_state(X: torch.Tensor, kind: str, init_noise: torch.Tensor | None = None) -> VarState:
    """Fresh model as built at runners.py:66-74 (+ the first-forward init of q(u), App. A.2).

    ``init_noise`` is the N(0,1) vector scaled by 1e-3 into nat1; by default it is drawn from
    the global torch generator exactly where upstream draws it (App. A.6)."""
    assert kind in KINDS
    n, d = X.shape
    mode_ls = _lognormal_mode_f32(0.0, 1.0)
    st = VarState(
        kind=kind,
        Z=X.clone().double().requires_grad_(),
        c=torch.zeros(1, dtype=torch.double, requires_grad=True),
        ls=torch.full((1, d), mode_ls, dtype=torch.double, requires_grad=True),
        os=torch.tensor(_lognormal_mode_f32(0.0, 1.0), dtype=torch.double, requires_grad=True),
        noise=torch.tensor([_lognormal_mode_f32(-4.0, 1.0)], dtype=torch.double, requires_grad=True),
        rawdf=(torch.tensor([_f32(inv_softplus(5.0))], dtype=torch.double, requires_grad=True)
               if kind == "VarTGP" else None),
        nat1=torch.zeros(n, dtype=torch.double),
        nat2=-0.5 * torch.eye(n, dtype=torch.double),
    )
    eps = torch.randn(n, dtype=torch.double) if init_noise is None else init_noise.double()
    st.nat1 = eps.clone().mul_(1e-3)
    st.adam = torch.optim.Adam(st.hyper(), lr=0.01)  # optim.py:136-141: lr hard-coded
    return st


def _interp(st: VarState, X: torch.Tensor) -> torch.Tensor:
    """A = L_K^-1 K_ZX with K_ZZ + 1e-6 I = L_K L_K^T (App. A.4 step 2)."""
    m = st.Z.shape[0]
    Kzz = rbf(st.Z, st.Z, st.ls, st.os) + JITTER * torch.eye(m, dtype=X.dtype)
    Lk = psd_safe_cholesky(Kzz)          # optim.py:153-154 -> variational strategy -> psd_safe_cholesky (jitter retry)
    Kzx = rbf(st.Z, X, st.ls, st.os)
    return torch.linalg.solve_triangular(Lk, Kzx, upper=False)


def studentt_logpdf(y, df, loc, scale):
    z = (y - loc) / scale
    norm = scale.log() + 0.5 * df.log() + 0.5 * math.log(math.pi) + torch.lgamma(0.5 * df) - torch.lgamma(0.5 * (df + 1.0))
    return -0.5 * (df + 1.0) * torch.log1p(z ** 2 / df) - norm


def laplace_logpdf(y, loc, scale):
    return -torch.log(2.0 * scale) - torch.abs(y - loc) / scale


def natural_to_moments(nat1, nat2):
    """(theta1, Theta2) -> (mu, S) the way gpytorch's NaturalVariationalDistribution does."""
    n = nat1.shape[0]
    eye = torch.eye(n, dtype=nat1.dtype)
    Linv = torch.linalg.cholesky(-2.0 * nat2)
    L = torch.linalg.solve_triangular(Linv, eye, upper=False)
    S = L.T @ L
    mu = (S @ nat1.unsqueeze(-1)).squeeze(-1)
    return mu, S


def elbo_loss(st: VarState, X: torch.Tensor, ys: torch.Tensor, mu: torch.Tensor, S: torch.Tensor):
    """loss = -(E_q log p(y|f) - KL + log-prior) / n (App. A.4 steps 2-7). mu, S may carry grad."""
    n = X.shape[0]
    eye = torch.eye(n, dtype=X.dtype)
    Ls = torch.linalg.cholesky(S)
    A = _interp(st, X)
    fmean = (A.T @ mu.unsqueeze(-1)).squeeze(-1) + st.c
    fvar = st.os + JITTER + (A * ((Ls @ Ls.T - eye) @ A)).sum(0)
    if st.kind == "VarGP":
        ell = (-0.5 * (((ys - fmean) ** 2 + fvar) / st.noise + st.noise.log() + LOG_2PI)).sum()
    else:
        locs = torch.sqrt(2.0 * fvar).unsqueeze(0) * GH_X.unsqueeze(-1) + fmean.unsqueeze(0)
        if st.kind == "VarTGP":
            lp = studentt_logpdf(ys.unsqueeze(0), st.df(), locs, st.noise.sqrt())
        else:
            lp = laplace_logpdf(ys.unsqueeze(0), locs, st.noise.sqrt())
        ell = ((1.0 / math.sqrt(math.pi)) * (lp * GH_W.unsqueeze(-1))).sum(0).sum()
    kl = 0.5 * (-(2.0 * torch.log(torch.diagonal(Ls)).sum()) + (Ls ** 2).sum() + (mu ** 2).sum() - n)
    lprior = (lognormal_logpdf(st.ls, _f32(0.0), _f32(1.0)).sum()
              + lognormal_logpdf(st.os, _f32(0.0), _f32(1.0))
              + lognormal_logpdf(st.noise, _f32(-4.0), _f32(1.0)).sum())
    if st.kind == "VarTGP":
        df = st.df()
        lprior = lprior + (GAMMA_CONC * math.log(GAMMA_RATE) + (GAMMA_CONC - 1.0) * torch.log(df)
                           - GAMMA_RATE * df - math.lgamma(GAMMA_CONC)).sum()
    return -(ell / n - kl / n + lprior / n)


def epoch(st: VarState, X: torch.Tensor, ys: torch.Tensor, lr: float):
    """One pass of the loop body optim.py:150-170.  Returns (loss, grads dict) and updates
    ``st`` in place: NGD step with lr*n on (nat1, nat2), Adam(0.01) on the hyper-parameters."""
    n = X.shape[0]
    st.adam.zero_grad()
    mu, S = natural_to_moments(st.nat1, st.nat2)
    mu_ = mu.detach().requires_grad_()
    S_ = S.detach().requires_grad_()
    loss = elbo_loss(st, X, ys, mu_, S_)
    loss.backward()
    gS = S_.grad
    g1 = mu_.grad - 2.0 * (gS @ mu.unsqueeze(-1)).squeeze(-1)
    grads = {"nat1": g1.clone(), "nat2": gS.clone(), "mu": mu_.grad.clone(),
             "Z": st.Z.grad.clone(), "c": st.c.grad.clone(), "ls": st.ls.grad.clone(),
             "os": st.os.grad.clone(), "noise": st.noise.grad.clone()}
    if st.rawdf is not None:
        grads["rawdf"] = st.rawdf.grad.clone()
    st.nat1 = st.nat1 - lr * n * g1
    st.nat2 = st.nat2 - lr * n * gS
    st.adam.step()
    val = float(loss.item())
    st.losses.append(val)
    return val, grads


def train_natural(X: torch.Tensor, y: torch.Tensor, kind: str, epochs: int, lr: float,
                  init_noise: torch.Tensor | None = None) -> VarState:
    """optim.py:105-181 on a fresh model; the 'best state' restore is a no-op (App. D-6)."""
    ys = standardise(y.double())
    st = init_state(X, kind, init_noise)
    for _ in range(epochs):
        epoch(st, X, ys, lr)
    return st


def train_adam(X: torch.Tensor, y: torch.Tensor, kind: str, epochs: int, lr: float,
               init_noise: torch.Tensor | None = None) -> VarState:
    """train_variational_model (optim.py:61-102) on a fresh model: ONE torch.optim.Adam(model.parameters(), lr)
    over every parameter, the natural parameters of q(u) included.  gpytorch's natural parameterisation hands the
    optimiser the gradient with respect to the EXPECTATION parameters as `.grad` of (natural_vec, natural_mat)
    (that is what makes a plain SGD step a natural-gradient step), so Adam is fed the same two arrays the NGD
    step of `epoch` uses.  Parity unpinned: no reference artefact exercises this trainer."""
    ys = standardise(y.double())
    st = init_state(X, kind, init_noise)
    nat1 = st.nat1.clone().requires_grad_()
    nat2 = st.nat2.clone().requires_grad_()
    opt = torch.optim.Adam(st.hyper() + [nat1, nat2], lr=lr)
    for _ in range(epochs):
        opt.zero_grad()
        mu, S = natural_to_moments(nat1.detach(), nat2.detach())
        mu_ = mu.detach().requires_grad_()
        S_ = S.detach().requires_grad_()
        st.nat1, st.nat2 = nat1.detach(), nat2.detach()
        loss = elbo_loss(st, X, ys, mu_, S_)
        loss.backward()
        gS = S_.grad
        nat1.grad = mu_.grad - 2.0 * (gS @ mu.unsqueeze(-1)).squeeze(-1)
        nat2.grad = gS.clone()
        opt.step()
        st.losses.append(float(loss.item()))
    st.nat1, st.nat2 = nat1.detach().clone(), nat2.detach().clone()
    return st


def latent_posterior(st: VarState, Xc: torch.Tensor):
    """q(f) at the candidates: mean m_*, variance v_* (App. A.4 step 3), no likelihood noise."""
    with torch.no_grad():
        n = st.Z.shape[0]
        mu, S = natural_to_moments(st.nat1, st.nat2)
        A = _interp(st, Xc)
        mean = (A.T @ mu.unsqueeze(-1)).squeeze(-1) + st.c
        var = st.os + JITTER + (A * ((S - torch.eye(n, dtype=torch.double)) @ A)).sum(0)
    return mean, var


def likelihood_variance(st: VarState) -> torch.Tensor:
    """Variance of the likelihood wrapped around the samples (App. A.5): nu/(nu-2) noise or 2 noise."""
    with torch.no_grad():
        if st.kind == "VarTGP":
            df = st.df()
            return st.noise * df / (df - 2.0)
        if st.kind == "VarLGP":
            return 2.0 * st.noise
        return st.noise.clone()


def acquisition(st: VarState, Xc: torch.Tensor, best_f, eps: torch.Tensor | None = None,
                num_samples: int = 2048) -> torch.Tensor:
    """LogEI over the pool as runners.py:88-93 sees it.

    VarGP: analytic on N(m, v + noise).  VarTGP / VarLGP: mean over S samples
    f = m + sqrt(v) eps of LogEI with the constant likelihood sigma (App. A.5, App. D-8);
    ``eps`` is [Nc, S]; by default drawn as randn(Nc, 1, S) from the global generator."""
    mean, var = latent_posterior(st, Xc)
    with torch.no_grad():
        if st.kind == "VarGP":
            sigma = (var + st.noise).clamp_min(1e-12).sqrt()
            return log_ei_helper((mean - best_f) / sigma) + sigma.log()
        if eps is None:
            eps = torch.randn(Xc.shape[0], 1, num_samples, dtype=torch.double).squeeze(1)
        f = mean.unsqueeze(0) + var.sqrt().unsqueeze(0) * eps.T  # [S, Nc]
        sigma = likelihood_variance(st).clamp_min(1e-12).sqrt()
        acq = log_ei_helper((f - best_f) / sigma) + sigma.log()
        return acq.mean(0)
