"""Row f3 (benchmark_predictions.py: fit on 80 % of a dataset, n up to 480), GPU-free part:
* the carve-up of the stp_exact_*_large entry points (working square in its own buffer) through the host emulation:
  identical arithmetic to the shared-memory carve-up, and against the oracle at n = 200;
* the split / seeding logic of the batched study (benchmark_predictions.py:125-137);
* the product refuses to run without a CUDA device."""
import hostemu
import numpy as np
import pytest
import torch

import _emu
from _fixtures import dataset
from oracle import exact as oexact
from stpbenchmark_b200 import benchmark_predictions as bp
from stpbenchmark_b200 import priors, utils

torch.set_num_threads(1)


@pytest.fixture
def large_mode():
    lib = hostemu.load()
    lib.emu_set_large(1)
    yield lib
    lib.emu_set_large(0)


def _split(name, seed, n):
    X, y = dataset(name, invert=False)
    tr, te = bp.split_indices(len(X), seed)
    y = (y - y[tr[:n]].mean()) / (y[tr[:n]].std(unbiased=False) + 1e-10)
    return X[tr[:n]], y[tr[:n]], X[te], y[te]


def test_large_carve_up_is_the_same_arithmetic(large_mode):
    Xt, yt, Xs, _ = _split("CrossedBarrel", 45, 60)
    d = Xt.shape[1]
    th = np.array(priors.exact_theta0(d))[None] * np.array([[1.0] * (d + 3), [2.0, 1.0, 0.7] + [1.5] * d])
    Xb = np.broadcast_to(Xt.numpy(), (2, 60, d)); yb = np.broadcast_to(yt.numpy(), (2, 60))
    f1, g1, s1 = _emu.exact_mll_fg(Xb, yb, th, priors.exact_prior_vector(d))
    a1 = _emu.exact_acq(Xs.numpy(), Xb, yb, th, best_f=np.array([1.0, 1.0]))
    large_mode.emu_set_large(0)
    f0, g0, s0 = _emu.exact_mll_fg(Xb, yb, th, priors.exact_prior_vector(d))
    a0 = _emu.exact_acq(Xs.numpy(), Xb, yb, th, best_f=np.array([1.0, 1.0]))
    assert s1.tolist() == s0.tolist() == [0, 0]
    assert np.array_equal(f1, f0) and np.array_equal(g1, g0)
    for k in ("mean", "var", "acq", "idx"):
        assert np.array_equal(a1[k], a0[k])


def test_large_carve_up_fit_is_the_same_trajectory(large_mode):
    """The device L-BFGS-B on the large carve-up (optimiser state behind the vectors, square elsewhere): same theta,
    objective and (nit, nfev) as on the shared-memory carve-up, bit for bit."""
    Xt, yt, _, _ = _split("P3HT", 45, 40)
    d = Xt.shape[1]
    lo, hi = priors.exact_bounds(d)
    args = (Xt.numpy()[None], yt.numpy()[None], priors.exact_theta0(d), lo, hi, priors.exact_prior_vector(d))
    th1, f1, nit1, nfev1, st1 = _emu.exact_fit(*args)
    large_mode.emu_set_large(0)
    th0, f0, nit0, nfev0, st0 = _emu.exact_fit(*args)
    assert st1.tolist() == st0.tolist() == [0] and nit1.tolist() == nit0.tolist() and nfev1.tolist() == nfev0.tolist()
    assert np.array_equal(th1, th0) and np.array_equal(f1, f0) and nit0[0] > 3


def test_large_carve_up_against_the_oracle_at_n200(large_mode):
    Xt, yt, Xs, _ = _split("CrossedBarrel", 359, 200)
    d = Xt.shape[1]
    th = np.array(priors.exact_theta0(d)) * np.array([3.0, 1.0, 1.2] + [0.8] * d) + np.r_[0.0, 0.1, np.zeros(d + 1)]
    f, g, st = _emu.exact_mll_fg(Xt.numpy()[None], yt.numpy()[None], th[None], priors.exact_prior_vector(d))
    fo, go = oexact.mll_value_and_grad(th, Xt, yt)
    assert st[0] == 0 and abs(f[0] - fo) <= 1e-10 * max(1.0, abs(fo))
    assert np.abs(g[0] - go).max() <= 1e-8 * np.abs(go).max()
    out = _emu.exact_acq(Xs.numpy(), Xt.numpy()[None], yt.numpy()[None], th[None])
    mo, vo = oexact.posterior_exact(th, Xt, yt, Xs)
    assert np.abs(out["mean"][0] - mo.numpy()).max() <= 1e-8 * np.abs(mo.numpy()).max()
    assert (np.abs(out["var"][0] - vo.numpy()) / vo.numpy()).max() <= 1e-7


def test_split_indices_follow_the_reference_recipe():
    # benchmark_predictions.py:127-137: set_seeds(seed); indices = torch.randperm(len(X)); test = first int(0.2 N)
    for N, seed in ((600, 45), (178, 359), (94, 6185)):
        tr, te = bp.split_indices(N, seed)
        utils.set_seeds(seed)
        ref = torch.randperm(N)
        k = int(0.2 * N)
        assert torch.equal(te, ref[:k]) and torch.equal(tr, ref[k:])
        assert len(tr) + len(te) == N and len(set(tr.tolist()) & set(te.tolist())) == 0
    assert len(bp.split_indices(600, 1)[0]) == 480      # the size that needs the large entry points
    assert len(bp.split_indices(178, 1)[0]) == 143      # P3HT: one more than stp_exact_fit holds in shared memory


def test_mae_of_the_monte_carlo_models_averages_over_the_samples():
    res = {"mean": torch.tensor([[1.0, 2.0]], dtype=torch.double), "var": torch.tensor([[0.25, 0.0]], dtype=torch.double),
           "scale": torch.tensor([2.0], dtype=torch.double)}
    y = torch.tensor([[1.0, 2.5]], dtype=torch.double)
    exact = bp._mae("VarGP", res, y, 0)
    assert abs(exact.item() - 0.25) < 1e-15
    mc = bp._mae("VarTGP", res, y, 200000, torch.Generator().manual_seed(0))
    # E|sigma eps| = sigma sqrt(2/pi) with sigma = sqrt(0.25) * 2 = 1 for the first point, |2 - 2.5| for the second
    assert abs(mc.item() - 0.5 * (np.sqrt(2 / np.pi) + 0.5)) < 5e-3


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    X = torch.rand(2, 20, 3, dtype=torch.double)
    with pytest.raises(RuntimeError):
        bp.fit_predict_batch("ExactGP", X, torch.rand(2, 20, dtype=torch.double), X[:, :5])
    with pytest.raises(NotImplementedError):
        bp.evaluate_model_across_datasets(dict, n_seeds=1)
