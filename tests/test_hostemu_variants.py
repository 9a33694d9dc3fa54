"""Kernel / acquisition variants of SURVEY section 8 row f4 (Matern-5/2, UCB, EI, Student-t LogEI): the device
routines compiled for the host (tests/hostemu) against the oracle restatements.  GPU-free; tests/test_gpu_variants.py
runs the same comparisons through the C ABI on the device."""
import numpy as np
import pytest
import torch

import _emu
import hostemu
from _fixtures import dataset, trace
from oracle import acq as oacq, exact as oexact
from stpbenchmark_b200 import priors

torch.set_num_threads(1)
KERN = {"rbf": 0, "matern52": 1}
ACQ = {"logei": 0, "ucb": 1, "ei": 2}


@pytest.fixture
def model():
    lib = hostemu.load()
    yield lambda kernel="rbf", acquisition="logei", beta=0.0: lib.emu_set_model(KERN[kernel], ACQ[acquisition], beta)
    lib.emu_set_model(0, 0, 0.0)


@pytest.mark.parametrize("name", ["AgNP", "P3HT"])
def test_matern_mll_value_and_gradient(name, model):
    X, y = dataset(name)
    sel = trace("gold1", "ExactGP", name, 45)[:31]
    d = X.shape[1]
    rng = np.random.default_rng(3)
    thetas = np.stack([np.array(priors.exact_theta0(d)) * rng.uniform(0.4, 2.5, d + 3) + np.r_[0, rng.normal(), np.zeros(d + 1)]
                       for _ in range(4)])
    Xb = np.broadcast_to(X[sel].numpy(), (4, 31, d)); yb = np.broadcast_to(y[sel].numpy(), (4, 31))
    model("matern52")
    f, g, st = _emu.exact_mll_fg(Xb, yb, thetas, priors.exact_prior_vector(d))
    assert st.tolist() == [0] * 4
    for b in range(4):
        fo, go = oexact.mll_value_and_grad(thetas[b], X[sel], y[sel], kernel="matern52")
        assert abs(f[b] - fo) <= 1e-10 * max(1.0, abs(fo))
        assert np.abs(g[b] - go).max() <= 1e-8 * np.abs(go).max()
        f_rbf, _ = oexact.mll_value_and_grad(thetas[b], X[sel], y[sel])
        assert abs(f[b] - f_rbf) > 1e-6          # the variant really is a different kernel


def test_matern_fit_and_pool_pass_follow_the_oracle(model):
    X, y = dataset("AgNP")
    sel = trace("gold1", "ExactGP", "AgNP", 359)[:24]
    d = X.shape[1]
    model("matern52", "ucb", 4.0)
    lo, hi = priors.exact_bounds(d)
    th, f, nit, nfev, st = _emu.exact_fit(X[sel].numpy()[None], y[sel].numpy()[None], priors.exact_theta0(d), lo, hi,
                                          priors.exact_prior_vector(d))
    theta_o, res = oexact.fit_exact(X[sel], y[sel], kernel="matern52")
    assert st[0] == 0 and (nit[0], nfev[0]) == (res.nit, res.nfev)
    assert np.abs(th[0] - theta_o).max() <= 1e-6 * np.abs(theta_o).max()
    mask = np.ones(len(X), dtype=np.uint8); mask[sel] = 0
    out = _emu.exact_acq(X.numpy(), X[sel].numpy()[None], y[sel].numpy()[None], th, mask=mask[None])
    mean, var = oexact.posterior_exact(th[0], X[sel], y[sel], X, kernel="matern52")
    assert np.abs(out["mean"][0] - mean.numpy()).max() <= 1e-9 * np.abs(mean.numpy()).max()
    assert np.abs(out["var"][0] - var.numpy()).max() <= 1e-8 * np.abs(var.numpy()).max()
    a = oacq.ucb(mean, var, 4.0)
    assert np.abs(out["acq"][0] - a.numpy()).max() <= 1e-9 * np.abs(a.numpy()).max()
    a[torch.as_tensor(mask == 0)] = -np.inf
    assert out["idx"][0] == oacq.first_argmax(a)


def test_ucb_and_ei_pointwise():
    lib = hostemu.load()
    rng = np.random.default_rng(0)
    mean = rng.normal(size=400); var = np.r_[rng.uniform(1e-14, 2.0, 396), 0.0, -1e-3, 1e-12, 5e-13]
    best = 0.3
    m, v = torch.tensor(mean), torch.tensor(var)
    u, e, le = oacq.ucb(m, v, 2.5).numpy(), oacq.ei(m, v, best).numpy(), oacq.log_ei(m, v, best).numpy()
    for i in range(400):
        assert abs(lib.emu_acq_value(1, 2.5, mean[i], var[i], best) - u[i]) <= 1e-14 * max(1, abs(u[i]))
        assert abs(lib.emu_acq_value(2, 0.0, mean[i], var[i], best) - e[i]) <= 1e-13 * max(1e-3, abs(e[i]))
        assert abs(lib.emu_acq_value(0, 0.0, mean[i], var[i], best) - le[i]) <= 1e-9 * max(1, abs(le[i]))
    # EI and LogEI agree where EI does not underflow (the reason the reference uses LogEI in its loop)
    big = e > 1e-200
    assert np.abs(np.log(e[big]) - le[big]).max() < 1e-6


def test_student_t_logei_against_scipy():
    lib = hostemu.load()
    rng = np.random.default_rng(1)
    mean = rng.normal(size=600) * 3; sd = np.exp(rng.normal(size=600)); df = rng.uniform(2.05, 60.0, 600)
    mean[:4] = [0.3, 40.0, -40.0, 0.3 + 1e-12]; sd[4:6] = [1e-9, 0.0]; df[6:9] = [2.0001, 1000.0, 4.0]
    ref = oacq.log_ei_t(mean, sd, df, 0.3)
    for i in range(600):
        got = lib.emu_log_ei_t(mean[i], sd[i], df[i], 0.3)
        if np.isinf(ref[i]):
            assert got == ref[i]
        else:
            assert abs(got - ref[i]) <= 1e-9 * max(1.0, abs(ref[i])), (i, mean[i], sd[i], df[i], got, ref[i])


def test_log_h_of_the_monte_carlo_path_equals_botorchs_form():
    """log_h_mc (log1p(-t) for every u < 0) against the oracle's restatement of botorch's three-branch _log_ei_helper:
    the same function to round-off over the whole range the sampled values can reach."""
    lib = hostemu.load()
    u = np.concatenate([-np.logspace(-9, 7, 3000), np.linspace(-3, 6, 2000), [0.0, -1.0, -1e6, -1.0000001e6, 30.0]])
    ref = oacq.log_ei_helper(torch.tensor(u)).numpy()
    got = np.array([lib.emu_log_h_mc(float(x)) for x in u])
    assert np.isfinite(got).all()
    # 1 - t cancels ~2 log10|u| digits for |u| >> 1 in BOTH forms (log t near 0 in botorch's): 1e-13 at |u| = 35
    assert (np.abs(got - ref) <= 2e-12 * np.maximum(1.0, np.abs(ref))).all()
    mid = np.abs(u) < 5
    assert (np.abs(got[mid] - ref[mid]) <= 1e-14 * np.maximum(1.0, np.abs(ref[mid]))).all()
    assert np.isnan(lib.emu_log_h_mc(float("nan")))
