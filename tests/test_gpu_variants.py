"""-m gpu: the kernel / acquisition variants of SURVEY section 8 row f4 through the C ABI against the oracle --
Matern-5/2 MLL value + gradient, device fit, fused pool pass with UCB / EI, a whole Matern + UCB campaign, the
elementwise stp_acq_values / stp_log_ei_t passes."""
import numpy as np
import pytest
import torch

from _fixtures import dataset, trace
from oracle import acq as oacq, exact as oexact
from stpbenchmark_b200 import _lib, priors

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def test_matern_mll_fit_and_pool_pass():
    X, y = dataset("AgNP")
    d = X.shape[1]
    sel = trace("gold1", "ExactGP", "AgNP", 359)[:28]
    rng = np.random.default_rng(5)
    thetas = np.stack([np.array(priors.exact_theta0(d)) * rng.uniform(0.5, 2.0, d + 3) for _ in range(3)])
    Xb, yb = _dev(np.broadcast_to(X[sel].numpy(), (3, 28, d))), _dev(np.broadcast_to(y[sel].numpy(), (3, 28)))
    m = _lib.ModelOpts.make("matern52", "ei")
    f, g, st = _lib.exact_mll_fg(Xb, yb, _dev(thetas), priors.exact_prior_vector(d), model=m)
    assert st.tolist() == [0] * 3
    for b in range(3):
        fo, go = oexact.mll_value_and_grad(thetas[b], X[sel], y[sel], kernel="matern52")
        assert abs(f[b].item() - fo) <= 1e-9 * max(1.0, abs(fo))
        assert np.abs(g[b].cpu().numpy() - go).max() <= 1e-8 * np.abs(go).max()
    th = _dev(np.array(priors.exact_theta0(d))[None]).clone()
    lo, hi = priors.exact_bounds(d)
    fo, nit, nfev, st = _lib.exact_fit(Xb[:1], yb[:1], th, lo, hi, priors.exact_prior_vector(d), model=m)
    theta_o, res = oexact.fit_exact(X[sel], y[sel], kernel="matern52")
    assert st.item() == 0 and (nit.item(), nfev.item()) == (res.nit, res.nfev)
    assert np.abs(th[0].cpu().numpy() - theta_o).max() <= 1e-6 * np.abs(theta_o).max()
    mask = torch.ones(1, len(X), dtype=torch.uint8); mask[0, sel] = 0
    best = torch.tensor([float(y[sel].max())], dtype=torch.double).cuda()
    out = _lib.exact_acq(_dev(X.numpy())[None], Xb[:1], yb[:1], th, best_f=best, mask=mask.cuda(), want_pointwise=True, model=m)
    mean, var = oexact.posterior_exact(th[0].cpu().numpy(), X[sel], y[sel], X, kernel="matern52")
    assert np.abs(out["mean"][0].cpu().numpy() - mean.numpy()).max() <= 1e-9 * np.abs(mean.numpy()).max()
    assert np.abs(out["var"][0].cpu().numpy() - var.numpy()).max() <= 1e-8 * np.abs(var.numpy()).max()
    a = oacq.ei(mean, var, float(y[sel].max()))
    assert np.abs(out["acq"][0].cpu().numpy() - a.numpy()).max() <= 1e-9 * np.abs(a.numpy()).max()
    a[torch.as_tensor(mask[0] == 0)] = -np.inf
    assert int(out["idx"][0]) == oacq.first_argmax(a)


@pytest.mark.parametrize("kernel,acquisition,beta", [("matern52", "ucb", 4.0), ("rbf", "ei", 0.0)])
def test_variant_campaign_follows_the_oracle_loop(kernel, acquisition, beta):
    """A whole campaign through run_campaign_batch with a non-default kernel / acquisition: every pick equals the
    oracle's pick when the oracle is teacher-forced along the device trace (fit -> posterior -> acquisition -> argmax)."""
    from stpbenchmark_b200 import runners
    from stpbenchmark_b200.models import ExactGP
    X, y = dataset("AutoAM")
    idx, val, status = runners.run_campaign_batch(X, y, ExactGP, [45, 359], n_initial=10, n_trials=8, kernel=kernel,
                                                  acquisition=acquisition, beta=beta)
    assert status == [0, 0]
    for row in idx:
        assert len(row) == 18 and all(row[t] not in row[:t] for t in range(10, 18))   # the Sobol design itself may repeat
        for t in range(10, 18):
            sel = row[:t]
            theta, _ = oexact.fit_exact(X[sel], y[sel], kernel=kernel)
            mean, var = oexact.posterior_exact(theta, X[sel], y[sel], X, kernel=kernel)
            a = oacq.ucb(mean, var, beta) if acquisition == "ucb" else oacq.ei(mean, var, float(y[sel].max()))
            a[sel] = -np.inf
            top = torch.topk(a, 2).values
            if float(top[0] - top[1]) > 1e-7 * max(1.0, abs(float(top[0]))):   # a near-tie may go either way in fp64
                assert row[t] == oacq.first_argmax(a), (t, row[t], oacq.first_argmax(a))


def test_the_variational_entry_points_reject_matern():
    from stpbenchmark_b200.engine import CampaignBatch
    X, y = dataset("AgNP")
    with pytest.raises(NotImplementedError, match="RBF"):
        CampaignBatch(X, y, [0], [list(range(10))], n_cap=14, kind="VarGP", kernel="matern52")


def test_elementwise_ucb_ei_and_student_t_logei():
    rng = np.random.default_rng(2)
    mean = rng.normal(size=5000); var = rng.uniform(1e-14, 2.0, 5000)
    m, v = torch.tensor(mean), torch.tensor(var)
    u = _lib.acq_values(m.cuda(), v.cuda(), 0.2, "ucb", 2.5).cpu()
    e = _lib.acq_values(m.cuda(), v.cuda(), 0.2, "ei").cpu()
    assert (u - oacq.ucb(m, v, 2.5)).abs().max() <= 1e-14
    eo = oacq.ei(m, v, 0.2)
    assert ((e - eo).abs() <= 1e-13 * eo.abs().clamp_min(1e-3)).all()
    sd = np.exp(rng.normal(size=5000)); df = rng.uniform(2.05, 60.0, 5000)
    sd[:2] = [1e-9, 0.0]
    got = _lib.log_ei_t(m.cuda() * 3, torch.tensor(sd).cuda(), torch.tensor(df).cuda(), 0.2).cpu().numpy()
    ref = oacq.log_ei_t(mean * 3, sd, df, 0.2)
    assert np.isneginf(got[:2]).all() and np.isneginf(ref[:2]).all()
    assert (np.abs(got[2:] - ref[2:]) <= 1e-9 * np.maximum(1.0, np.abs(ref[2:]))).all()


def test_acqf_classes_on_a_device_model():
    from stpbenchmark_b200 import acqf, optim
    from stpbenchmark_b200.models import ExactGP
    X, y = dataset("P3HT")
    sel = trace("gold1", "ExactGP", "P3HT", 45)[:20]
    model = ExactGP(X[sel], y[sel]).double()
    optim.train_exact_model_botorch(model, X[sel], y[sel])
    post = model.posterior(X)
    mean, var = post.mean.cpu().flatten(), post.variance.cpu().flatten()
    u = acqf.UpperConfidenceBound(model, beta=3.0)(X.unsqueeze(1)).cpu()
    e = acqf.ExpectedImprovement(model, best_f=y[sel].max())(X.unsqueeze(1)).cpu()
    assert (u - oacq.ucb(mean, var, 3.0)).abs().max() <= 1e-12
    assert (e - oacq.ei(mean, var, float(y[sel].max()))).abs().max() <= 1e-12
