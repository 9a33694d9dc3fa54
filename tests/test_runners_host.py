"""Campaign-loop host logic (CSV layout, resume, NaN padding, invert_y) with the device engine
replaced by a stub, plus the scipy lock-step driver on the host-emulated MLL."""
import numpy as np
import pandas as pd
import torch

import _emu
from _fixtures import dataset, trace
from oracle import exact as oexact
from stpbenchmark_b200 import optim, priors, runners


def _stub_engine(monkeypatch, fail_seed=None, boom_on=None):
    """Replaces the device engine (runners._execute_on_device) by a deterministic fake: campaign of seed s selects
    (s + i) mod N; ``fail_seed`` stops after 3 trials with status 1; ``boom_on`` = a seed whose presence makes the
    whole device job raise."""
    calls = []

    def fake(pools, campaigns, on_done=None):
        calls.append([c.seed for c in campaigns])
        if boom_on is not None and any(c.seed == boom_on for c in campaigns):
            raise RuntimeError("device lost")
        res = {}
        for i, c in enumerate(campaigns):
            N = len(pools[c.pool][0])
            k = len(c.design) + (3 if c.seed == fail_seed else c.n_trials)
            res[i] = ([(c.seed + j) % N for j in range(k)], 1 if c.seed == fail_seed else 0)
        if on_done is not None:
            on_done(res)
        return res

    monkeypatch.setattr(runners, "_execute_on_device", fake)
    return calls


def test_csv_layout_resume_and_padding(tmp_path, monkeypatch):
    calls = _stub_engine(monkeypatch, fail_seed=7)
    X = torch.rand(50, 3, dtype=torch.double); y = torch.rand(50, dtype=torch.double) + 0.5
    out = tmp_path / "res" / "ExactGP_toy_traces_.csv"
    df = runners.run_many_loops(X, y, runners.ExactGP, seeds=[3, 7], output_path=str(out), n_initial=10, n_trials=5)
    assert calls == [[3, 7]]
    back = pd.read_csv(out, header=[0, 1], index_col=0)
    assert back.index.name == "iteration" and len(back) == 15
    assert list(back.columns) == [("seed_3", "x_index"), ("seed_3", "y_value"), ("seed_7", "x_index"), ("seed_7", "y_value")]
    assert back["seed_7", "x_index"].isna().sum() == 2 and back["seed_3", "x_index"].isna().sum() == 0  # NaN padding
    # resume: seeds already present are skipped, the new one is appended
    runners.run_many_loops(X, y, runners.ExactGP, seeds=[3, 7, 11], output_path=str(out), n_initial=10, n_trials=5)
    assert calls[-1] == [11]
    back2 = pd.read_csv(out, header=[0, 1], index_col=0)
    assert ("seed_11", "x_index") in back2.columns
    assert np.allclose(back2["seed_3", "y_value"], back["seed_3", "y_value"])
    runners.run_many_loops(X, y, runners.ExactGP, seeds=[3, 7, 11], output_path=str(out), n_initial=10, n_trials=5)
    assert len(calls) == 2                                       # nothing left to do


def test_invert_y_is_idempotent_across_resume(tmp_path, monkeypatch):
    _stub_engine(monkeypatch)
    X = torch.rand(40, 2, dtype=torch.double); y = torch.rand(40, dtype=torch.double) + 0.5
    out = tmp_path / "inv.csv"
    runners.run_many_loops(X, y, runners.VarTGP, seeds=[1], output_path=str(out), invert_y=True, n_initial=10, n_trials=4)
    a = pd.read_csv(out, header=[0, 1], index_col=0)
    runners.run_many_loops(X, y, runners.VarTGP, seeds=[1, 2], output_path=str(out), invert_y=True, n_initial=10, n_trials=4)
    b = pd.read_csv(out, header=[0, 1], index_col=0)
    first = [float(y[(1 + i) % 40]) for i in range(14)]
    assert np.allclose(a["seed_1", "y_value"], 1.0 / np.array(first))
    assert np.allclose(b["seed_1", "y_value"], a["seed_1", "y_value"])   # the reference would have re-inverted (App. B.3)
    assert np.allclose(b["seed_1", "x_index"], a["seed_1", "x_index"])


def test_a_failed_device_job_does_not_poison_the_resume(tmp_path, monkeypatch):
    """an exception inside the batch must not leave all-NaN columns that a later run mistakes for finished
    seeds, and with invert_y=True it must not write the previously saved seeds back un-inverted."""
    X = torch.rand(40, 2, dtype=torch.double); y = torch.rand(40, dtype=torch.double) + 0.5
    out = tmp_path / "inv.csv"
    _stub_engine(monkeypatch)
    runners.run_many_loops(X, y, runners.ExactGP, seeds=[1], output_path=str(out), invert_y=True, n_initial=10, n_trials=4)
    a = pd.read_csv(out, header=[0, 1], index_col=0)
    _stub_engine(monkeypatch, boom_on=2)
    runners.run_many_loops(X, y, runners.ExactGP, seeds=[1, 2, 3], output_path=str(out), invert_y=True, n_initial=10, n_trials=4)
    b = pd.read_csv(out, header=[0, 1], index_col=0)
    assert set(b.columns.get_level_values(0)) == {"seed_1"}                  # nothing written for seeds 2, 3
    assert np.allclose(b["seed_1", "y_value"], a["seed_1", "y_value"])       # still 1/y, once
    calls = _stub_engine(monkeypatch)
    runners.run_many_loops(X, y, runners.ExactGP, seeds=[1, 2, 3], output_path=str(out), invert_y=True, n_initial=10, n_trials=4)
    assert calls == [[2, 3]]
    c = pd.read_csv(out, header=[0, 1], index_col=0)
    assert set(c.columns.get_level_values(0)) == {"seed_1", "seed_2", "seed_3"}
    assert np.allclose(c["seed_1", "y_value"], a["seed_1", "y_value"])
    # a file that carries an all-NaN column (written by an older crashed run) does not count that seed as done
    c[("seed_9", "x_index")] = np.nan; c[("seed_9", "y_value")] = np.nan
    c.to_csv(out)
    calls = _stub_engine(monkeypatch)
    runners.run_many_loops(X, y, runners.ExactGP, seeds=[1, 9], output_path=str(out), invert_y=True, n_initial=10, n_trials=4)
    assert calls == [[9]]


def test_sweep_submits_every_job_at_once_and_writes_one_csv_per_job(tmp_path, monkeypatch):
    """run_benchmark's nested loops (run_benchmark.py:20-43) as ONE device job: model x dataset x seeds."""
    calls = _stub_engine(monkeypatch)
    jobs = []
    for name, N, d, cls in (("a", 30, 2, runners.ExactGP), ("b", 45, 3, runners.VarTGP), ("c", 60, 3, runners.VarTGP)):
        X = torch.rand(N, d, dtype=torch.double); y = torch.rand(N, dtype=torch.double) + 0.5
        jobs.append(dict(X=X, y=y, model_class=cls, seeds=[5, 6], output_path=str(tmp_path / f"{name}.csv"),
                         n_initial=10, n_trials=6, epochs=7, learning_rate=0.01))
    frames = runners.run_sweep(jobs)
    assert calls == [[5, 6, 5, 6, 5, 6]] and len(frames) == 3
    for spec in jobs:
        back = pd.read_csv(spec["output_path"], header=[0, 1], index_col=0)
        N = len(spec["X"])
        assert back["seed_6", "x_index"].astype(int).tolist() == [(6 + i) % N for i in range(16)]
        assert np.allclose(back["seed_6", "y_value"], spec["y"][[(6 + i) % N for i in range(16)]].numpy())


def test_model_kwargs_are_honoured_or_rejected(monkeypatch):
    """runners.py:66-74 builds model_class(**model_kwargs): a lengthscale prior reaches the engine, keys the loop
    sets itself are ignored, anything else (or a subclass) is an error -- never silently dropped."""
    seen = []

    def fake(pools, campaigns, on_done=None):
        seen.extend(c.lengthscale_prior for c in campaigns)
        return {i: (list(c.design), 0) for i, c in enumerate(campaigns)}

    monkeypatch.setattr(runners, "_execute_on_device", fake)
    X = torch.rand(30, 2, dtype=torch.double); y = torch.rand(30, dtype=torch.double)
    kw = {"lengthscale_prior": (0.5, 2.0), "inducing_points": None}
    idx, val = runners.run_single_loop(X, y, runners.VarGP, model_kwargs=kw, n_initial=10, n_trials=0, seed=4)
    assert seen == [(0.5, 2.0)] and len(idx) == 10 and "inducing_points" in kw and kw["inducing_points"].shape == (10, 2)

    class Prior:                      # a gpytorch-style prior object
        loc, scale = 1.0, 0.25
    runners.run_single_loop(X, y, runners.VarTGP, model_kwargs={"lengthscale_prior": Prior()}, n_initial=10, n_trials=0)
    assert seen[-1] == (1.0, 0.25)
    n_seen = len(seen)
    # unknown keyword: the constructor call of the reference raises inside the try -> the initial design is returned
    idx, val = runners.run_single_loop(X, y, runners.ExactGP, model_kwargs={"nu": 2.5}, n_initial=10, n_trials=3, seed=4)
    assert len(idx) == 10 and len(val) == 10 and len(seen) == n_seen

    class MyGP(runners.ExactGP):
        pass
    idx, val = runners.run_single_loop(X, y, MyGP, n_initial=10, n_trials=3, seed=4)
    assert len(idx) == 10 and len(seen) == n_seen
    import pytest
    with pytest.raises(NotImplementedError):
        runners._kind_of(MyGP)


def test_run_single_loop_never_raises(monkeypatch):
    def boom(*a, **k):
        raise RuntimeError("no device")
    monkeypatch.setattr(runners, "_execute_on_device", boom)
    idx, val = runners.run_single_loop(torch.rand(20, 2), torch.rand(20), runners.ExactGP)
    assert len(idx) == 10 and len(val) == 10          # the initial design, like runners.py:46-52 + :113-115
    assert runners.run_single_loop(torch.rand(20, 2), torch.rand(20), runners.ExactGP, n_initial="x") == ([], [])


def test_kind_detection():
    assert runners._kind_of(runners.ExactGP) == "ExactGP" and runners._kind_of(runners.VarTGP) == "VarTGP"
    assert runners._kind_of("VarLGP") == "VarLGP"


def test_scipy_lockstep_driver_reproduces_scipy_minimize(monkeypatch):
    """Parity-mode optimiser: B setulb state machines in lock-step == scipy.optimize.minimize per campaign
    (the MLL kernel is replaced by its host emulation here; the GPU test repeats this on the device)."""
    from stpbenchmark_b200 import _lib

    def emu_fg(X, y, theta, prior, n=None):
        f, g, st = _emu.exact_mll_fg(X.numpy(), y.numpy(), theta.numpy(), prior, n=None if n is None else n.numpy())
        return torch.from_numpy(f), torch.from_numpy(g), torch.from_numpy(st)

    monkeypatch.setattr(_lib, "exact_mll_fg", emu_fg)
    X, y = dataset("AgNP")
    d = X.shape[1]
    ns = [17, 30, 55]  # fits that are not round-off chaotic
    tr = trace("gold1", "ExactGP", "AgNP", 45)
    Xb = torch.zeros(3, 56, d, dtype=torch.double); yb = torch.zeros(3, 56, dtype=torch.double)
    for b, n in enumerate(ns):
        Xb[b, :n] = X[tr[:n]]; yb[b, :n] = y[tr[:n]]
    lower, upper = priors.exact_bounds(d)
    theta, nit, nfev, status = optim.fit_exact_scipy_lockstep(Xb, yb, torch.tensor(ns, dtype=torch.int32), priors.exact_theta0(d),
                                                              lower, upper, priors.exact_prior_vector(d))
    for b, n in enumerate(ns):
        ths, res = oexact.fit_exact(X[tr[:n]], y[tr[:n]])
        assert status[b] == 0 and (nit[b], nfev[b]) == (res.nit, res.nfev)
        assert np.abs(theta[b].numpy() - ths).max() <= 1e-7 * np.abs(ths).max()


def test_fast_csv_writer_is_byte_identical_to_pandas(tmp_path):
    """runners._csv_text replaces DataFrame.to_csv for plain trace frames: same bytes (ints, shortest-repr floats, NaN
    as the empty cell, the three header rows), and anything unusual falls back to pandas."""
    import numpy as np
    import pandas as pd
    from stpbenchmark_b200 import runners
    rng = np.random.default_rng(0)
    cols = {}
    for k, sd in enumerate([3, 17, 45]):
        x = rng.integers(0, 2000, 12).astype(np.int64) if k != 1 else np.r_[rng.integers(0, 2000, 7).astype(float), [np.nan] * 5]
        y = rng.normal(size=12) * 10.0 ** rng.integers(-8, 8, 12)
        if k == 1:
            y[7:] = np.nan
        y[:5] = [1e-5, 123456789.123456789, -0.0, 1e22, 5.0]
        cols[(f"seed_{sd}", "x_index")] = x
        cols[(f"seed_{sd}", "y_value")] = y
    f = pd.DataFrame(cols, index=range(12))
    f.columns = pd.MultiIndex.from_tuples(list(cols))
    f.index.name = "iteration"
    assert runners._csv_text(f) == f.to_csv()
    path = str(tmp_path / "t.csv")
    runners._save(f, path, invert_y=False)
    assert open(path).read() == f.to_csv()
    back = pd.read_csv(path, header=[0, 1], index_col=0)
    assert back.shape == f.shape and np.allclose(back.to_numpy(dtype=float), f.to_numpy(dtype=float), equal_nan=True)
    g = f.astype(object)
    assert runners._csv_text(g) is None          # object columns: pandas writes them


def test_block_built_frame_equals_column_built_frame(tmp_path):
    """A fresh job whose seeds all finish at once is assembled from two homogeneous blocks; a job that gets its seeds
    one by one goes through the column path.  Same frame, same file."""
    import numpy as np
    import pandas as pd
    import torch
    from stpbenchmark_b200 import runners
    from stpbenchmark_b200.models import ExactGP
    rng = np.random.default_rng(1)
    X = torch.rand(60, 3, dtype=torch.double); y = torch.rand(60, dtype=torch.double)
    seeds = [5, 9, 2]
    traces = [(s, [int(v) for v in rng.permutation(60)[:14]]) for s in seeds]
    a = runners._CsvJob(X, y, ExactGP, seeds, str(tmp_path / "a.csv"), n_initial=4, n_trials=10)
    a.record_many(traces); a.save()
    b = runners._CsvJob(X, y, ExactGP, seeds, str(tmp_path / "b.csv"), n_initial=4, n_trials=10)
    for item in traces:
        b.record(*item)
    b.save()
    assert open(tmp_path / "a.csv").read() == open(tmp_path / "b.csv").read()
    assert a.frame.shape == b.frame.shape == (14, 6)
    assert list(a.frame.columns) == list(b.frame.columns)
    assert (a.frame.to_numpy(dtype=float) == b.frame.to_numpy(dtype=float)).all()
    assert pd.read_csv(tmp_path / "a.csv", header=[0, 1], index_col=0).shape == (14, 6)


def test_csv_text_corner_values_equal_pandas():
    """The value-table form of the CSV writer: signed zeros, NaN, infinities, denormals, large indices."""
    import numpy as np
    import pandas as pd
    from stpbenchmark_b200 import runners
    cols = pd.MultiIndex.from_tuples([("seed_3", "x_index"), ("seed_3", "y_value"), ("seed_1", "x_index"), ("seed_1", "y_value")])
    f = pd.DataFrame({cols[0]: np.array([0, 5, 1048575, 7], dtype=np.int64), cols[1]: np.array([-0.0, 0.0, np.nan, 1e22]),
                      cols[2]: np.array([3, 2, 1, 2000000], dtype=np.int64), cols[3]: np.array([1e-5, 0.1 + 0.2, -np.inf, 5e-324])})
    f.columns = cols
    f.index.name = "iteration"
    assert runners._csv_text(f) == f.to_csv()
    g = f.copy()
    g[cols[0]] = np.array([-1, 5, 2, 7], dtype=np.int64)          # negative indices: the plain path
    assert runners._csv_text(g) == g.to_csv()
