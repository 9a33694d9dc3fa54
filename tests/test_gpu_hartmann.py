"""BASELINE config #3 scenario (test_hartmann.py:13-36): fit on 100 random Hartmann-6 points, predict 100 fresh
points, compare train / test MAE with the CPU oracle -- all four model classes through the drop-in API.
The reference script itself is broken (imports a class that does not exist); only its scenario is taken
(SURVEY section 4, section 8(d) cfg #3)."""
import numpy as np
import pytest
import torch

from oracle import exact as oexact, variational as ov

pytestmark = pytest.mark.gpu


def _data(seed):
    from stpbenchmark_b200 import synthetic, utils
    utils.set_seeds(seed)
    X_train = torch.rand(100, 6).double()          # fp32 draws, as in the script (test_hartmann.py:18), widened
    y_train = synthetic.hartmann6(X_train)
    X_test = torch.rand(100, 6).double()
    return X_train, y_train, X_test, synthetic.hartmann6(X_test)


def test_hartmann6_constants():
    from stpbenchmark_b200 import synthetic
    x = torch.tensor([[0.20169, 0.150011, 0.476874, 0.275332, 0.311652, 0.6573]], dtype=torch.double)
    assert abs(synthetic.hartmann6(x, negate=False).item() + 3.322368011) < 1e-6


def test_exactgp_fit_and_predict_matches_oracle():
    from stpbenchmark_b200 import models, optim
    Xt, yt, Xs, ys = _data(0)
    m = models.ExactGP(X_train=Xt, y_train=yt).double()
    optim.train_exact_model_botorch(m, Xt, yt)
    th, res = oexact.fit_exact(Xt, yt)
    assert np.abs(m.theta().cpu().numpy() - th).max() <= 1e-4 * np.abs(th).max()
    m.eval()
    mean_o, _ = oexact.posterior_exact(th, Xt, yt, Xs)
    pred = m(Xs).mean.cpu()
    mae, mae_o = (pred - ys).abs().mean().item(), (mean_o - ys).abs().mean().item()
    assert abs(mae - mae_o) <= 1e-3 * max(mae_o, 1e-3) and mae < 0.5
    train_mae = (m(Xt).mean.cpu() - yt).abs().mean().item()
    assert train_mae < mae


@pytest.mark.parametrize("cls_name", ["VarGP", "VarTGP", "VarLGP"])
def test_variational_fit_and_predict_matches_oracle(cls_name):
    from stpbenchmark_b200 import models, optim
    Xt, yt, Xs, ys = _data(1)
    E = 50                                           # short horizon: end-of-fit states are comparable (App. C.7)
    torch.manual_seed(3)
    m = getattr(models, cls_name)(inducing_points=Xt).double()
    optim.train_natural_variational_model(m, Xt, yt, epochs=E, lr=0.01)
    torch.manual_seed(3)
    ost = ov.train_natural(Xt, yt, cls_name, E, 0.01)
    m.eval()
    mo, _ = ov.latent_posterior(ost, Xs)
    pred = m(Xs).mean.cpu()
    assert (pred - mo).abs().max() <= 1e-5 * mo.abs().max()
    mean, scale = m._y_scaler                       # the model predicts standardised targets (App. D-2)
    mae = ((pred * scale + mean) - ys).abs().mean().item()
    mae_o = ((mo * scale + mean) - ys).abs().mean().item()
    assert abs(mae - mae_o) <= 1e-4 * mae_o
