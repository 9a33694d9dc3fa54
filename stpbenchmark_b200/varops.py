"""Host glue of the variational models (VarGP / VarTGP / VarLGP): packs model parameters into the
batch-major state arrays the CUDA library works on, and back.

Two users:
  * ``VarBatchState``  -- B campaigns advanced together by ``engine.CampaignBatch`` (two launches per BO
    iteration: stp_var_fit with all epochs inside, then stp_var_acq with the argmax and the update).
  * ``train_natural`` / ``latent_posterior`` / ``posterior`` -- the per-model calls behind
    ``optim.train_natural_variational_model`` and ``Var*.posterior`` (a batch of one).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, priors
from .models import Posterior

_LIK_OF_KIND = {"VarGP": 0, "VarTGP": 1, "VarLGP": 2, "gaussian": 0, "student_t": 1, "laplace": 2}
ADAM_LR = 0.01  # hard-coded at optim.py:140, whatever `learning_rate` says (App. D-10)


def gauss_hermite_table():
    """20 Gauss-Hermite nodes then weights, rounded through fp32 as upstream does (App. D-9)."""
    x, w = np.polynomial.hermite.hermgauss(20)
    x = torch.tensor(x, dtype=torch.float32).double().tolist()
    w = torch.tensor(w, dtype=torch.float32).double().tolist()
    return x + w


def var_prior_vector(ls_prior=priors.VAR_LENGTHSCALE_PRIOR):
    f = priors.f32
    return [f(ls_prior[0]), f(ls_prior[1]), f(priors.OUTPUTSCALE_PRIOR[0]), f(priors.OUTPUTSCALE_PRIOR[1]),
            f(priors.NOISE_PRIOR[0]), f(priors.NOISE_PRIOR[1]), f(priors.GAMMA_DF_PRIOR[0]), f(priors.GAMMA_DF_PRIOR[1])]


def var_hyp0(d: int, ls_prior=priors.VAR_LENGTHSCALE_PRIOR):
    """(constant, lengthscale x d, outputscale, noise, raw_deg_free) at construction (App. A.2)."""
    import math
    raw = priors.DF_INIT - 2.0
    raw = priors.f32(raw + math.log(-math.expm1(-raw)))
    return [0.0] + [priors.lognormal_mode(*ls_prior)] * d + [priors.lognormal_mode(*priors.OUTPUTSCALE_PRIOR),
                                                             priors.lognormal_mode(*priors.NOISE_PRIOR), raw]


def alloc_state(B: int, cap: int, d: int, device) -> dict:
    z = lambda *s: torch.zeros(*s, dtype=torch.double, device=device)
    return {"Z": z(B, cap, d), "hyp": z(B, d + 4), "nat1": z(B, cap), "nat2": z(B, cap, cap),
            "adam_m": z(B, cap * d + d + 4), "adam_v": z(B, cap * d + d + 4)}


class VarBatchState:
    """Variational side of a CampaignBatch: model state + workspace for B campaigns."""

    def __init__(self, batch):
        self.b = batch
        self.lik = _LIK_OF_KIND[batch.kind]
        B, cap, d = batch.B, batch.cap, batch.d
        self.state = alloc_state(B, cap, d, batch.device)
        self.work = _lib.var_workspace(B, cap, batch.device)
        ls_prior = getattr(batch, "lengthscale_prior", None) or priors.VAR_LENGTHSCALE_PRIOR
        self.prior = var_prior_vector(ls_prior)     # model_kwargs["lengthscale_prior"] (models.py:30, 125, 217)
        self.hyp0 = var_hyp0(d, ls_prior)
        self.gh = gauss_hermite_table()
        self.generators = None        # per-campaign CPU generators (reference RNG order), optional
        self.philox_seed = 0x5742_1F3D
        self.mc_eps = "philox"        # "philox": in-kernel draws; "host": randn(N, 1, S) from the campaign's CPU generator

    def set_rng_states(self, states):
        """Give every campaign its own CPU generator positioned where the reference's global generator
        stands after set_seeds + the Sobol draw (App. A.6): the n normals initialising theta1 are then
        drawn per trial in the reference's order."""
        self.generators = []
        for st in states:
            g = torch.Generator()
            g.set_state(st)
            self.generators.append(g)

    def _init_noise(self):
        b = self.b
        if self.generators is None:
            return None
        # host mirror of n when the engine has one (no device round trip per step); frozen campaigns are skipped by
        # the kernel, so what is drawn for them does not matter
        n_host = b._n_host.tolist() if b._n_host is not None else b.n.cpu().tolist()
        noise = torch.zeros(b.B, b.cap, dtype=torch.double)
        for i, (g, n) in enumerate(zip(self.generators, n_host)):
            noise[i, :n] = torch.randn(n, dtype=torch.double, generator=g)
        return noise.to(b.device)

    def step(self) -> int:
        b = self.b
        noise = self._init_noise()
        order = torch.argsort(b.n, descending=True, stable=True).to(torch.int32) if b.ragged else None
        _lib.var_fit(b.X, b.y, b.n, self.state, self.lik, b.epochs, b.learning_rate, ADAM_LR, self.hyp0, self.prior,
                     self.gh, b.status, self.work, fresh=True, init_noise=noise, seed=self.philox_seed + 7919 * b.iteration,
                     order=order)
        eps = None
        if self.lik != 0 and self.mc_eps == "host" and self.generators is not None:
            eps = torch.stack([torch.randn(b.N, 1, b.num_samples, dtype=torch.double, generator=g).squeeze(1)
                               for g in self.generators]).to(b.device)
        out = _lib.var_acq(b.pool, self.state, b.n, self.lik, b.status, self.work, best_f=None, mask=b.mask,
                           pool_id=b.pool_id, S=b.num_samples, eps=eps, seed=self.philox_seed + 104729 * b.iteration,
                           update={"pool_y": b.pool_y, "X": b.X, "y": b.y, "trace": b.trace}, order=order,
                           model=getattr(b, "model_opts", None))
        b.acq = out["acq_best"]
        return 2


# ---------------------------------------------------------------------------------------------
# single-model entry points
# ---------------------------------------------------------------------------------------------
def _pack(model, cap=None):
    """Model parameters -> state dict (batch of one) on the model's device."""
    Z = model.variational_strategy.inducing_points.detach().double()
    n, d = Z.shape
    cap = cap or _lib.var_capacity(n, d)
    dev = Z.device
    st = alloc_state(1, cap, d, dev)
    st["Z"][0, :n] = Z
    lk = model.likelihood
    hyp = [model.mean_module.raw_constant.detach().double().reshape(1),
           model.covar_module.base_kernel.raw_lengthscale.detach().double().reshape(-1),
           model.covar_module.raw_outputscale.detach().double().reshape(1), lk.noise.detach().double().reshape(1),
           (lk.raw_deg_free.detach().double().reshape(1) if lk.kind == "student_t" else torch.zeros(1, dtype=torch.double, device=dev))]
    st["hyp"][0] = torch.cat(hyp)
    vd = model.variational_strategy._variational_distribution
    st["nat1"][0, :n] = vd.natural_vec.detach().double()
    st["nat2"][0, :n, :n] = vd.natural_mat.detach().double()
    return st, n, d, cap


def _unpack(model, st, n, d):
    with torch.no_grad():
        model.variational_strategy.inducing_points.copy_(st["Z"][0, :n])
        model.mean_module.raw_constant.copy_(st["hyp"][0, 0])
        model.covar_module.base_kernel.raw_lengthscale.copy_(st["hyp"][0, 1:1 + d].reshape(1, d))
        model.covar_module.raw_outputscale.copy_(st["hyp"][0, 1 + d])
        lk = model.likelihood
        (lk.noise_covar.raw_noise if lk.kind == "gaussian" else lk.raw_noise).copy_(st["hyp"][0, 2 + d:3 + d])
        if lk.kind == "student_t":
            lk.raw_deg_free.copy_(st["hyp"][0, 3 + d:4 + d])
        vd = model.variational_strategy._variational_distribution
        vd.natural_vec.copy_(st["nat1"][0, :n])
        vd.natural_mat.copy_(st["nat2"][0, :n, :n])


def train_natural(model, X, y, epochs=500, lr=0.05, y_standardize=True, init_noise=None):
    """optim.py:105-181 for one model, in place.  Returns the loss history (list of floats).

    On the first call of a fresh model theta1 is initialised to 1e-3 * N(0, I) drawn from the GLOBAL torch
    generator, exactly where upstream draws it (first forward, App. A.2 / A.6), unless ``init_noise`` is given.
    The model keeps ``_y_scaler`` = (mean, std + 1e-10) of the raw targets."""
    model.train()
    model.likelihood.train()
    dev = model.device
    if not y_standardize:
        raise NotImplementedError("the reference always standardises (optim.py:126-127); y_standardize=False is not provided")
    vs = model.variational_strategy
    st, n, d, cap = _pack(model)
    Xd = X.detach().to(dev).double().reshape(1, -1, d).contiguous()
    yd = y.detach().to(dev).double().reshape(1, -1).contiguous()
    if Xd.shape[1] != n:
        # more data than inducing points is legal upstream, but the campaign loop never does it
        raise NotImplementedError("train_natural_variational_model expects as many inducing points as training points")
    if int(vs.variational_params_initialized) == 0:
        eps = torch.randn(n, dtype=torch.double) if init_noise is None else init_noise.double().cpu()
        st["nat1"][0, :n] = (eps * vs._variational_distribution.mean_init_std).to(dev)
        vs.variational_params_initialized.fill_(1)
    ls_prior = model.lengthscale_prior
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    work = _lib.var_workspace(1, cap, dev)
    hist = torch.zeros(1, max(int(epochs), 1), dtype=torch.double, device=dev)
    nn = torch.tensor([n], dtype=torch.int32, device=dev)
    # optim.py:131-143 builds a fresh NGD and a fresh Adam optimiser on EVERY call: step count and moments start at
    # zero even when the same model is trained again (the packed state's moments are zero-initialised)
    t0 = 0
    _lib.var_fit(Xd, yd, nn, st, _LIK_OF_KIND[model.likelihood.kind], int(epochs), float(lr), ADAM_LR, var_hyp0(d, ls_prior),
                 var_prior_vector(ls_prior), gauss_hermite_table(), status, work, fresh=False, t0=t0, loss_hist=hist)
    code = int(status.item())
    if code != 0:
        raise RuntimeError("variational fit failed: " + _lib.STATUS_TEXT.get(code, str(code)))
    _unpack(model, st, n, d)
    mean = yd.mean()
    model._y_scaler = (float(mean), float(yd.std(unbiased=False)) + 1e-10)
    losses = hist[0, : int(epochs)].cpu().tolist()
    model.fit_info = {"losses": losses}
    return losses


def train_adam(model, X, y, epochs=100, lr=0.01, y_standardize=True, init_noise=None):
    """optim.py:61-102 for one model, in place: torch.optim.Adam(lr) on EVERY parameter, the natural parameters
    included (they receive the gradient with respect to the expectation parameters, as in gpytorch).  One
    stp_var_elbo_fg launch per epoch gives the loss and all gradients; the Adam update itself is torch's, applied to
    the packed state on the device.  Returns the loss history."""
    model.train()
    model.likelihood.train()
    dev = model.device
    if not y_standardize:
        raise NotImplementedError("the reference always standardises (optim.py:80-81); y_standardize=False is not provided")
    vs = model.variational_strategy
    st, n, d, cap = _pack(model)
    Xd = X.detach().to(dev).double().reshape(1, -1, d).contiguous()
    yd = y.detach().to(dev).double().reshape(1, -1).contiguous()
    if Xd.shape[1] != n:
        raise NotImplementedError("train_variational_model expects as many inducing points as training points")
    if int(vs.variational_params_initialized) == 0:
        eps = torch.randn(n, dtype=torch.double) if init_noise is None else init_noise.double().cpu()
        st["nat1"][0, :n] = (eps * vs._variational_distribution.mean_init_std).to(dev)
        vs.variational_params_initialized.fill_(1)
    ls_prior = model.lengthscale_prior
    prior, gh = var_prior_vector(ls_prior), gauss_hermite_table()
    lik = _LIK_OF_KIND[model.likelihood.kind]
    work = _lib.var_workspace(1, cap, dev)
    nn = torch.tensor([n], dtype=torch.int32, device=dev)
    params = [st[k].requires_grad_(True) for k in ("nat1", "nat2", "Z", "hyp")]
    opt = torch.optim.Adam(params, lr=lr)
    o1, o2, o3 = cap, cap + cap * cap, cap + cap * cap + cap * d
    losses = []
    for _ in range(int(epochs)):
        loss, grads, status = _lib.var_elbo_fg(Xd, yd, nn, {k: st[k].detach() for k in ("Z", "hyp", "nat1", "nat2")},
                                               lik, prior, gh, work)
        code = int(status.item())
        if code != 0:
            raise RuntimeError("variational fit failed: " + _lib.STATUS_TEXT.get(code, str(code)))
        g = grads[0]
        st["nat1"].grad = g[:o1].reshape(1, cap).clone()
        st["nat2"].grad = g[o1:o2].reshape(1, cap, cap).clone()
        st["Z"].grad = g[o2:o3].reshape(1, cap, d).clone()
        st["hyp"].grad = g[o3:].reshape(1, d + 4).clone()
        opt.step()
        losses.append(float(loss.item()))
    for k in ("nat1", "nat2", "Z", "hyp"):
        st[k] = st[k].detach()
    _unpack(model, st, n, d)
    model._y_scaler = (float(yd.mean()), float(yd.std(unbiased=False)) + 1e-10)
    model.fit_info = {"losses": losses}
    return losses


def _pointwise(model, X, best_f=float("nan"), S=1, eps=None):
    dev = model.device
    st, n, d, cap = _pack(model)
    pts = X.detach().to(dev).double().reshape(-1, d).contiguous()
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    work = _lib.var_workspace(1, cap, dev)
    nn = torch.tensor([n], dtype=torch.int32, device=dev)
    bf = torch.tensor([best_f], dtype=torch.double, device=dev)
    out = _lib.var_acq(pts.unsqueeze(0), st, nn, 0, status, work, best_f=bf, S=1, want_pointwise=True)
    code = int(status.item())
    if code != 0:
        raise RuntimeError("variational prediction failed: " + _lib.STATUS_TEXT.get(code, str(code)))
    return out["mean"][0], out["var"][0]


def latent_posterior(model, X):
    """q(f) mean and variance at X (any leading shape), no likelihood noise."""
    shape = X.shape[:-1]
    mean, var = _pointwise(model, X)
    return mean.reshape(shape), var.reshape(shape)


def posterior(model, X, observation_noise=True, num_samples=2048, eps=None):
    """Var*.posterior (models.py:95-111, 187-202, 280-297).  X: [b, 1, d] or [Nt, d]."""
    mean, var = latent_posterior(model, X)
    lk = model.likelihood
    if not observation_noise:
        return Posterior(mean, var)
    noise = lk.noise.detach().double()
    if lk.kind == "gaussian":
        return Posterior(mean, var + noise)
    # Monte-Carlo marginal: S samples of the latent f wrapped in the likelihood (App. A.5)
    flat_m, flat_v = mean.reshape(-1), var.reshape(-1)
    if eps is None:
        eps = torch.randn(flat_m.numel(), 1, num_samples, dtype=torch.double).squeeze(1)  # global CPU generator, (Nc, 1, S)
    eps = eps.to(mean.device)
    samples = flat_m.unsqueeze(0) + flat_v.sqrt().unsqueeze(0) * eps.T           # [S, Nc]
    samples = samples.reshape((eps.shape[1],) + tuple(mean.shape))
    if lk.kind == "student_t":
        df = lk.deg_free.detach().double()
        lik_var = noise * df / (df - 2.0)
        return Posterior(samples, lik_var.expand(samples.shape), df=df.expand(samples.shape))
    return Posterior(samples, (2.0 * noise).expand(samples.shape))
