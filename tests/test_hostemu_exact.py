"""Algebra of the per-campaign device routines, compiled for the host (tests/hostemu), against
the oracle: exact MLL value + analytic gradient, the device L-BFGS-B vs scipy, the fused
predictive / LogEI / argmax pass.  GPU-free; the same sources run on the device in -m gpu tests."""
import numpy as np
import pytest
import torch
from scipy.optimize import minimize

import _emu
from _fixtures import DATASETS, dataset, trace
from oracle import acq as oacq, exact as oexact
from stpbenchmark_b200 import priors

torch.set_num_threads(1)


def _train(name, seed, n):
    X, y = dataset(name)
    sel = trace("gold1", "ExactGP", name, seed)[:n]
    return X, y, sel


@pytest.mark.parametrize("name", DATASETS)
def test_mll_value_and_gradient(name):
    X, y, sel = _train(name, 45, 27)
    d = X.shape[1]
    rng = np.random.default_rng(1)
    thetas = np.stack([np.array(priors.exact_theta0(d)) * rng.uniform(0.4, 2.5, d + 3) + np.r_[0, rng.normal(), np.zeros(d + 1)]
                       for _ in range(4)])
    Xb = np.broadcast_to(X[sel].numpy(), (4, 27, d)); yb = np.broadcast_to(y[sel].numpy(), (4, 27))
    f, g, st = _emu.exact_mll_fg(Xb, yb, thetas, priors.exact_prior_vector(d))
    assert st.tolist() == [0] * 4
    for b in range(4):
        fo, go = oexact.mll_value_and_grad(thetas[b], X[sel], y[sel])
        assert abs(f[b] - fo) <= 1e-11 * max(1.0, abs(fo))
        assert np.abs(g[b] - go).max() <= 1e-9 * np.abs(go).max()


def test_mll_ragged_batch_and_failure_codes():
    X, y, sel = _train("AgNP", 359, 30)
    d = X.shape[1]
    cap = 32
    ns = [1, 2, 10, 30]
    Xb = np.zeros((4, cap, d)); yb = np.zeros((4, cap))
    for b, n in enumerate(ns):
        Xb[b, :n] = X[sel[:n]].numpy(); yb[b, :n] = y[sel[:n]].numpy()
    th = np.tile(priors.exact_theta0(d), (4, 1))
    f, g, st = _emu.exact_mll_fg(Xb, yb, th, priors.exact_prior_vector(d), n=ns)
    for b, n in enumerate(ns):
        fo, go = oexact.mll_value_and_grad(th[b], X[sel[:n]], y[sel[:n]])
        assert st[b] == 0 and abs(f[b] - fo) < 1e-11 * max(1, abs(fo)) and np.abs(g[b] - go).max() < 1e-9 * max(1, np.abs(go).max())
    bad = th.copy(); bad[0, 0] = -1.0; bad[1, 3] = float("nan"); bad[2, 2] = 0.0
    f, g, st = _emu.exact_mll_fg(Xb, yb, bad, priors.exact_prior_vector(d), n=ns)
    assert st.tolist()[:3] == [2, 2, 2] and st[3] == 0 and np.isnan(f[:3]).all()


def test_lbfgsb_matches_scipy_on_bounded_rosenbrock():
    def rosen(x):
        f = sum(100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2)
        g = np.zeros_like(x)
        g[:-1] = -400 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
        g[1:] += 200 * (x[1:] - x[:-1] ** 2)
        return f, g
    for lower, upper in [([None] * 5, [None] * 5), ([-0.5, 0.2, None, None, 0.3], [0.8, None, 0.9, None, 2.0]),
                         ([1.5] * 5, [3.0] * 5)]:
        x0 = np.array([-1.2, 1.0, 0.5, 2.0, 1.7])
        ref = minimize(rosen, x0, jac=True, method="L-BFGS-B", bounds=list(zip(lower, upper)))
        got = _emu.lbfgsb_minimize(rosen, x0, lower, upper)
        assert got["status"] == ref.status == 0
        assert (got["nit"], got["nfev"]) == (ref.nit, ref.nfev)
        assert np.allclose(got["x"], ref.x, rtol=1e-9, atol=1e-12) and abs(got["fun"] - ref.fun) <= 1e-12 * max(1, abs(ref.fun))


def test_lbfgsb_iteration_cap_and_flat_start():
    quad = lambda x: (float(x @ x), 2 * x)
    got = _emu.lbfgsb_minimize(quad, np.zeros(3), [None] * 3, [None] * 3)
    assert got["status"] == 0 and got["nfev"] == 1 and got["nit"] == 0      # projected gradient already below pgtol
    ref = minimize(lambda x: (float(np.sum(x ** 4)), 4 * x ** 3), np.full(4, 3.0), jac=True, method="L-BFGS-B",
                   options=dict(maxiter=5))
    got = _emu.lbfgsb_minimize(lambda x: (float(np.sum(x ** 4)), 4 * x ** 3), np.full(4, 3.0), [None] * 4, [None] * 4, maxiter=5)
    assert got["status"] == 1 and ref.status == 1 and got["nit"] == ref.nit == 5 and np.allclose(got["x"], ref.x, rtol=1e-10)


@pytest.mark.parametrize("name,seed,n", [("AgNP", 45, 35), ("Perovskite", 45, 17), ("CrossedBarrel", 45, 55),
                                          ("CrossedBarrel", 359, 17), ("Perovskite", 359, 80)])
def test_exact_fit_follows_scipy(name, seed, n):
    """Device L-BFGS-B on the device MLL vs scipy L-BFGS-B on the oracle MLL: same iteration and
    evaluation counts and the same optimum on fits that are not round-off chaotic."""
    X, y, sel = _train(name, seed, n)
    d = X.shape[1]
    lower, upper = priors.exact_bounds(d)
    theta, f, nit, nfev, st = _emu.exact_fit(X[sel].numpy()[None], y[sel].numpy()[None], priors.exact_theta0(d), lower,
                                            upper, priors.exact_prior_vector(d))
    ths, res = oexact.fit_exact(X[sel], y[sel])
    assert st[0] == 0 and res.status == 0
    assert (nit[0], nfev[0]) == (res.nit, res.nfev)
    assert np.abs(theta[0] - ths).max() <= 1e-6 * np.abs(ths).max() and abs(f[0] - res.fun) <= 1e-10 * max(1, abs(res.fun))


@pytest.mark.parametrize("name", DATASETS)
def test_fused_acquisition(name):
    X, y, sel = _train(name, 359, 40)
    d = X.shape[1]
    ths, _ = oexact.fit_exact(X[sel], y[sel])
    mask = np.ones(len(X), dtype=np.uint8); mask[sel] = 0
    out = _emu.exact_acq(X.numpy()[None], X[sel].numpy()[None], y[sel].numpy()[None], ths[None], best_f=[float(y[sel].max())],
                         mask=mask[None])
    mean, var = oexact.posterior_exact(ths, X[sel], y[sel], X)
    acq = oacq.log_ei(mean, var, y[sel].max())
    assert np.abs(out["mean"][0] - mean.numpy()).max() <= 1e-10 * np.abs(mean.numpy()).max()
    assert (np.abs(out["var"][0] - var.numpy()) / var.numpy()).max() <= 1e-8
    assert (np.abs(out["acq"][0] - acq.numpy()) / np.maximum(1.0, np.abs(acq.numpy()))).max() <= 1e-6  # LogEI amplifies var error by u^2
    ref = acq.clone(); ref[torch.from_numpy(mask) == 0] = -float("inf")
    assert out["idx"][0] == int(torch.argmax(ref)) == trace("gold1", "ExactGP", name, 359)[40] or name == "P3HT"
    assert out["idx"][0] == int(torch.argmax(ref))


def test_acquisition_edge_cases():
    X, y, sel = _train("AutoAM", 45, 12)
    d = X.shape[1]
    th = np.array(priors.exact_theta0(d))
    none = np.zeros(len(X), dtype=np.uint8)
    out = _emu.exact_acq(X.numpy()[None], X[sel].numpy()[None], y[sel].numpy()[None], th[None], best_f=[0.0], mask=none[None])
    assert out["idx"][0] == -1                                   # empty candidate set
    one = none.copy(); one[77] = 1
    out = _emu.exact_acq(X.numpy()[None], X[sel].numpy()[None], y[sel].numpy()[None], th[None], best_f=[0.0], mask=one[None])
    assert out["idx"][0] == 77
    out = _emu.exact_acq(X.numpy()[None], X[sel].numpy()[None], y[sel].numpy()[None], th[None], best_f=[float("nan")])
    assert out["idx"][0] == 0 and np.isnan(out["acq"][0]).all()  # NaN wins, first NaN first (torch.argmax)


def test_log_h_matches_oracle_everywhere():
    import hostemu
    lib = hostemu.load()
    us = np.concatenate([-np.logspace(7, -3, 200), np.linspace(-1.5, 8, 200), [-1.0, -1e6, 0.0]])
    ref = oacq.log_ei_helper(torch.tensor(us, dtype=torch.double)).numpy()
    got = np.array([lib.emu_log_h(float(u)) for u in us])
    assert np.abs(got - ref).max() <= 1e-10 * np.maximum(1, np.abs(ref)).max() and np.all(np.abs(got - ref) <= 1e-9 * np.maximum(1, np.abs(ref)))
    assert np.isnan(lib.emu_log_h(float("nan")))
    assert lib.emu_log_ei(0.3, -5.0, 0.1) == lib.emu_log_ei(0.3, 1e-12, 0.1)  # variance clamp
