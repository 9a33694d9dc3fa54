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


@pytest.mark.parametrize("name", DATASETS)
def test_batched_initial_designs_equal_the_serial_path(name):
    """utils.initial_designs_batch (row f2: one pass for all seeds, private generator, vectorised scramble + snap) ==
    set_seeds(seed); get_initial_samples_sobol(...) per seed: all 50 golden designs of every dataset, and the state
    the global generator is left in (what the variational models' theta1 noise continues from)."""
    import torch
    from stpbenchmark_b200 import utils
    X, y = dataset(name)
    S = seeds()
    got, states = utils.initial_designs_batch(X, y, S, 10, 50.0, return_rng_states=True)
    for k, s in enumerate(S):
        assert got[k].tolist() == trace("gold1", "ExactGP", name, s)[:10]
    for k in (0, 17, 49):
        utils.set_seeds(S[k])
        ref = utils.get_initial_samples_sobol(X, y, n_samples=10, percentile=50)
        assert torch.equal(ref, got[k]) and torch.equal(torch.get_rng_state(), states[k])


def test_batched_scrambled_sobol_is_bit_exact():
    import torch
    from torch.quasirandom import SobolEngine
    from stpbenchmark_b200 import utils
    S = [0, 1, 45, 6185, 2 ** 31 - 1]
    for d, n in [(3, 10), (6, 10), (8, 256), (5, 1), (4, 33)]:
        got = utils.sobol_scrambled_draw_batch(S, d, n)
        for k, s in enumerate(S):
            torch.manual_seed(s)
            assert torch.equal(SobolEngine(d, scramble=True).draw(n), got[k])
            one, _ = utils.sobol_scrambled_draw(s, d, n)
            assert torch.equal(one, got[k])


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


def test_post_processing_helpers():
    """utils.py:136-151, 312-370: same results as the reference's own helpers (imported from the read-only checkout
    when it is present, otherwise checked against hand-computed values)."""
    import numpy as np
    import torch
    from stpbenchmark_b200 import utils as U
    y = torch.tensor([0.3, 2.0, 1.5, 2.0, -1.0, 0.7, 0.1, 1.9, 0.0, 0.2, 1.1, 0.9, 0.8, 0.6, 0.5, 0.4, 1.2, 1.3, 1.4, 1.6, 1.7])
    top = U.get_top_samples(y, top_percentage=10)
    assert top == pytest.approx([2.0, 2.0, 1.9]) and U.num_top_samples([top[0], 0.3, top[2]], top) == pytest.approx(2 / 3)
    assert U.pad_array(np.array([1.0, 3.0]), 4).tolist() == [1.0, 3.0, 3.0, 3.0]
    ns = {"run0": np.array([1.0, 2.0]), "run1": np.array([3.0, 4.0, 5.0]), "run3": np.array([9.0])}
    assert len(U.collect_arrays("run", 3, ns)) == 2 and U.find_max_length("run", 3, ns) == 3
    mean, std = U.process_arrays("run", 3, 3, ns)
    assert mean.tolist() == [2.0, 3.0, 3.5] and std.tolist() == pytest.approx([1.0, 1.0, 1.5])
    torch.manual_seed(7)
    x = torch.arange(len(y), dtype=torch.double).unsqueeze(1)
    xi, yi = U.initial_points(x, y, 5)
    assert xi.shape == (5, 1) and (yi > torch.quantile(y, 0.05)).all() and len(set(xi.flatten().tolist())) == 5
    torch.manual_seed(7)
    keep = torch.nonzero(y > torch.quantile(y, 0.05), as_tuple=True)[0]
    assert torch.equal(xi.flatten().long(), keep[torch.randperm(len(keep))[:5]])
