"""Drop-in surface of the model classes, checked without a GPU."""
import math

import torch

from stpbenchmark_b200 import models, priors
from stpbenchmark_b200.models import ExactGP, VarGP, VarLGP, VarTGP


def test_variational_detection_rule_of_the_reference_loop():
    # runners.py:66 -- "inducing_points" in model_class.__init__.__code__.co_varnames
    for cls in (VarGP, VarLGP, VarTGP):
        assert "inducing_points" in cls.__init__.__code__.co_varnames
    assert "inducing_points" not in ExactGP.__init__.__code__.co_varnames


def test_exact_gp_parameters_names_order_and_initial_values():
    X = torch.rand(10, 5, dtype=torch.double); y = torch.rand(10, dtype=torch.double)
    m = ExactGP(X_train=X, y_train=y, input_transform=None)
    assert m.double() is m and m.num_outputs == 1
    names = [n for n, _ in m.named_parameters()]
    assert names == ["likelihood.noise_covar.raw_noise", "mean_module.raw_constant", "covar_module.raw_outputscale",
                     "covar_module.base_kernel.raw_lengthscale"]   # botorch's flattening order
    th = m.theta().cpu()
    assert th.dtype == torch.double and th.shape == (8,)
    assert abs(th[0].item() - math.exp(-5)) < 1e-9 and th[0].item() != math.exp(-5)   # fp32-rounded mode (App. A.0)
    assert th[1].item() == 0.0 and abs(th[2].item() - math.exp(-1)) < 1e-7
    assert abs(th[3].item() - math.exp(math.sqrt(2) + 0.5 * math.log(5) - 3)) < 1e-6
    assert m.likelihood.noise.shape == (1,) and m.covar_module.base_kernel.lengthscale.shape == (1, 5)
    assert float(m.covar_module.outputscale.detach()) == th[2].item()
    sd = {k: v.clone() for k, v in m.state_dict().items()}  # state_dict() aliases live storage (App. D-6)
    m.set_theta(torch.arange(1, 9, dtype=torch.double))
    assert float(m.mean_module.constant.detach()) == 2.0
    m.load_state_dict(sd)
    assert torch.equal(m.theta().cpu(), th)
    prior = m.forward(X)
    assert prior.mean.shape == (10,) and prior.covariance_matrix.shape == (10, 10)
    assert torch.allclose(torch.diagonal(prior.covariance_matrix).cpu(), th[2].expand(10))


def test_variational_models_parameters():
    Z = torch.rand(12, 4, dtype=torch.double)
    for cls, kind in ((VarGP, "gaussian"), (VarTGP, "student_t"), (VarLGP, "laplace")):
        m = cls(inducing_points=Z, lengthscale_prior=None).double()
        hyper = list(m.hyperparameters()); var = list(m.variational_parameters())
        assert len(var) == 2 and var[0].shape == (12,) and var[1].shape == (12, 12)
        assert torch.equal(var[1].cpu(), -0.5 * torch.eye(12, dtype=torch.double))
        assert len(hyper) == (6 if kind == "student_t" else 5)
        assert torch.equal(m.variational_strategy.inducing_points.detach().cpu(), Z)
        assert abs(float(m.covar_module.base_kernel.lengthscale[0, 0]) - math.exp(-1)) < 1e-7
        assert abs(float(m.likelihood.noise) - math.exp(-5)) < 1e-9
        assert m.likelihood.kind == kind and m.num_outputs == 1
    t = VarTGP(Z).double()
    assert abs(float(t.likelihood.deg_free.detach()) - 7.0) < 1e-6                 # nu = 2 + softplus(raw) = 7
    custom = VarGP(Z, lengthscale_prior=(0.5, 2.0))
    assert custom.lengthscale_prior == (0.5, 2.0)


def test_prior_constants_are_fp32_rounded():
    v = priors.exact_prior_vector(5)
    assert v[:4] == [-4.0, 1.0, 0.0, 1.0]
    assert v[4] == float(torch.tensor(math.sqrt(2) + 0.5 * math.log(5), dtype=torch.float32)) != math.sqrt(2) + 0.5 * math.log(5)
    lo, hi = priors.exact_bounds(3)
    assert lo == [1e-4, None, 1e-4, 0.025, 0.025, 0.025] and hi == [None] * 6
