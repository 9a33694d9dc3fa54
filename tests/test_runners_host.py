"""Campaign-loop host logic (CSV layout, resume, NaN padding, invert_y) with the device engine
replaced by a stub, plus the scipy lock-step driver on the host-emulated MLL."""
import numpy as np
import pandas as pd
import torch

import _emu
from _fixtures import dataset, trace
from oracle import exact as oexact
from stpbenchmark_b200 import optim, priors, runners


def _stub_engine(monkeypatch, fail_seed=None):
    calls = []

    def fake(X, y, model_class, seeds, n_initial=10, n_trials=25, epochs=100, learning_rate=0.1, num_samples=2048):
        calls.append(list(seeds))
        idx, val, st = [], [], []
        for s in seeds:
            k = n_initial + (3 if s == fail_seed else n_trials)
            row = [(s + i) % len(X) for i in range(k)]
            idx.append(row); val.append([float(y[i]) for i in row]); st.append(1 if s == fail_seed else 0)
        return idx, val, st

    monkeypatch.setattr(runners, "run_campaign_batch", fake)
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


def test_run_single_loop_never_raises(monkeypatch):
    def boom(*a, **k):
        raise RuntimeError("no device")
    monkeypatch.setattr(runners, "run_campaign_batch", boom)
    assert runners.run_single_loop(torch.rand(20, 2), torch.rand(20), runners.ExactGP) == ([], [])


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
