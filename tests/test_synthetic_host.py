"""Synthetic workloads of BASELINE configs #3-#5: deterministic, documented draw order."""
import torch

from stpbenchmark_b200 import synthetic


def test_hartmann6_global_minimum():
    x = torch.tensor([[0.20169, 0.150011, 0.476874, 0.275332, 0.311652, 0.6573]], dtype=torch.double)
    assert abs(synthetic.hartmann6(x, negate=False).item() + 3.322368011) < 1e-6
    assert abs(synthetic.hartmann6(x).item() - 3.322368011) < 1e-6


def test_cfg4_pool_is_reproducible_and_matches_the_survey_probe():
    X, y = synthetic.cfg4_pool(0)
    X2, y2 = synthetic.cfg4_pool(0)
    assert X.shape == (2048, 6) and X.dtype == torch.double and torch.equal(X, X2) and torch.equal(y, y2)
    assert not torch.equal(X, synthetic.cfg4_pool(1)[0])
    assert -0.5 < y.min().item() < -0.3 and 2.6 < y.max().item() < 2.8      # SURVEY: pool 0 has y in [-0.38, 2.70]
    gen = torch.Generator().manual_seed(1234)
    assert torch.equal(X, torch.rand(2048, 6, generator=gen, dtype=torch.double))  # X drawn first, then the noise


def test_cfg5_pool_shape_and_range():
    X, y = synthetic.cfg5_pool(0, N=1024)
    assert X.shape == (1024, 8) and y.shape == (1024,) and torch.isfinite(y).all()


def test_initial_designs_follow_the_campaign_loop():
    from stpbenchmark_b200 import utils
    X, y = synthetic.cfg4_pool(2, N=256)
    D = synthetic.initial_designs(X, y, [synthetic.campaign_seed(2, 5), 7])
    assert D.shape == (2, 10) and D.dtype == torch.long
    utils.set_seeds(2005)
    assert D[0].tolist() == utils.get_initial_samples_sobol(X, y, 10, percentile=50).tolist()
    assert (y[D] <= torch.quantile(y, 0.5)).all()
