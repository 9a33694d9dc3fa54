"""Host-side utilities (row a11): bit-exact against the reference's own outputs."""
import numpy as np
import pytest
import torch

from _fixtures import DATASETS, dataset, raw_table, seeds, trace, _npz
from stpbenchmark_b200 import utils


@pytest.mark.parametrize("name", DATASETS)
def test_preprocess_matches_reference(name):
    """preprocess_array == the reference's utils.preprocess_data output stored in the fixture."""
    X, y = utils.preprocess_array(raw_table(name))
    Xr, yr = dataset(name, invert=False)
    assert X.dtype == torch.double and torch.equal(X, Xr) and torch.equal(y, yr)


@pytest.mark.parametrize("name", DATASETS)
def test_sobol_initial_designs_match_all_golden_traces(name):
    """set_seeds + get_initial_samples_sobol reproduce rows 0-9 of gold_runs/benchmark_1 for 50/50 seeds."""
    X, y = dataset(name)
    for s in seeds():
        utils.set_seeds(s)
        init = utils.get_initial_samples_sobol(X.double(), y.double().flatten(), n_samples=10, percentile=50)
        assert init.tolist() == trace("gold1", "ExactGP", name, s)[:10]
        assert init.tolist() == trace("gold1", "VarTGP", name, s)[:10]


@pytest.mark.parametrize("name", DATASETS)
def test_random_percentile_initial_designs_match_gold0(name):
    """get_initial_samples (older runs, utils.py:205-236) reproduces gold_runs/benchmark_0 rows 0-9."""
    X, y = dataset(name)
    gold = _npz("traces.npz")[f"gold0_init_{name}"]
    hits = 0
    for i, s in enumerate(seeds()):
        utils.set_seeds(s)
        init = utils.get_initial_samples(y.double().flatten(), n_samples=10, percentile=50)
        hits += init.tolist() == gold[i].astype(int).tolist()
    assert hits == 50


def test_duplicate_initial_points_are_kept():
    X, y = dataset("P3HT")
    utils.set_seeds(45)
    init = utils.get_initial_samples_sobol(X, y, 10, percentile=50).tolist()
    assert init == [11, 11, 107, 12, 8, 59, 11, 112, 11, 99]


def test_standard_scaler_population_std():
    x = torch.tensor([1.0, 2.0, 4.0, 9.0], dtype=torch.double)
    sc = utils.TorchStandardScaler()
    z = sc.fit_transform(x)
    ref = (x - x.mean()) / (x.std(unbiased=False) + 1e-10)
    assert torch.equal(z, ref) and torch.allclose(sc.inverse_transform(z), x)
    nz = utils.TorchNormalizer()
    u = nz.fit_transform(x.reshape(-1, 1))
    assert u.min() == 0 and u.max() == 1 and torch.allclose(nz.inverse_transform(u), x.reshape(-1, 1))


def test_few_eligible_points_returns_them_all():
    X = torch.rand(6, 2, dtype=torch.double)
    y = torch.arange(6, dtype=torch.double)
    out = utils.get_initial_samples_sobol(X, y, n_samples=10, percentile=50)
    assert out.tolist() == [0, 1, 2]
    with pytest.raises(ValueError):
        utils.get_initial_samples(y, n_samples=10, percentile=50)


def test_get_device_raises_without_cuda():
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            utils.get_device()


def test_ei_formula():
    m = torch.tensor([0.0, 1.0], dtype=torch.double); s = torch.tensor([1.0, 2.0], dtype=torch.double)
    ei = utils.EI(m, s, torch.tensor(0.5, dtype=torch.double))
    z = (m - 0.5) / s
    n = torch.distributions.Normal(0, 1)
    assert torch.allclose(ei, (m - 0.5) * n.cdf(z) + s * n.log_prob(z).exp())
    assert torch.allclose(utils.EI(m, s, torch.tensor(0.5), minimize=True), (0.5 - m) * n.cdf(-z) + s * n.log_prob(-z).exp())
