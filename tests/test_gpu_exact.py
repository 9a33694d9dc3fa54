"""Parity of the CUDA exact-GP path with the CPU oracle, through the C ABI (libstp_b200.so).

Tolerances (fp64): MLL value/gradient and posterior mean 1e-9 relative, variance 1e-8 relative,
LogEI 1e-6 relative to max(1, |LogEI|) (the tail of log_h amplifies a relative variance error by u^2),
selected indices bit-exact.  north_star asks for 1e-5.
"""
import numpy as np
import pytest
import torch

from _fixtures import DATASETS, EXACT_KAT, dataset, trace
from oracle import acq as oacq, exact as oexact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from stpbenchmark_b200 import _lib
    assert _lib.lib().stp_version() == 3
    return _lib


def _batch(name, seed, ns, cap, dev):
    X, y = dataset(name)
    d = X.shape[1]
    tr = trace("gold1", "ExactGP", name, seed)
    Xb = torch.zeros(len(ns), cap, d, dtype=torch.double); yb = torch.zeros(len(ns), cap, dtype=torch.double)
    for b, n in enumerate(ns):
        Xb[b, :n] = X[tr[:n]]; yb[b, :n] = y[tr[:n]]
    return X, y, tr, Xb.to(dev), yb.to(dev), torch.tensor(ns, dtype=torch.int32, device=dev)


@pytest.mark.parametrize("name", DATASETS)
def test_mll_value_and_gradient(L, name):
    from stpbenchmark_b200 import priors
    ns = [1, 2, 10, 23, 47, 89]
    X, y, tr, Xb, yb, nb = _batch(name, 45, ns, 90, "cuda")
    d = X.shape[1]
    rng = np.random.default_rng(3)
    th = np.stack([np.array(priors.exact_theta0(d)) * rng.uniform(0.5, 2.0, d + 3) + np.r_[0, rng.normal(), np.zeros(d + 1)]
                   for _ in ns])
    f, g, st = L.exact_mll_fg(Xb, yb, torch.tensor(th, device="cuda"), priors.exact_prior_vector(d), n=nb)
    assert st.tolist() == [0] * len(ns)
    for b, n in enumerate(ns):
        fo, go = oexact.mll_value_and_grad(th[b], X[tr[:n]], y[tr[:n]])
        assert abs(f[b].item() - fo) <= 1e-9 * max(1.0, abs(fo))
        assert np.abs(g[b].cpu().numpy() - go).max() <= 1e-9 * max(1.0, np.abs(go).max())


def test_mll_failure_codes(L):
    from stpbenchmark_b200 import priors
    X, y, tr, Xb, yb, nb = _batch("AgNP", 359, [20, 20, 20, 20], 24, "cuda")
    d = X.shape[1]
    th = torch.tensor([priors.exact_theta0(d)] * 4, dtype=torch.double)
    th[0, 0] = -1.0; th[1, 4] = float("nan"); th[2, 2] = 0.0
    f, g, st = L.exact_mll_fg(Xb, yb, th.cuda(), priors.exact_prior_vector(d), n=nb)
    assert st.tolist() == [2, 2, 2, 0] and torch.isnan(f[:3]).all() and torch.isfinite(f[3])
    # duplicated rows + zero noise bound cannot be factorised: status 1, not a crash
    Xd = Xb.clone(); Xd[3, 1] = Xd[3, 0]
    th2 = torch.tensor([priors.exact_theta0(d)] * 4, dtype=torch.double); th2[3, 0] = 1e-300
    f, g, st = L.exact_mll_fg(Xd, yb, th2.cuda(), priors.exact_prior_vector(d), n=nb)
    assert st[3].item() in (0, 1)


@pytest.mark.parametrize("name,seed,ns", [("AgNP", 45, [17, 30, 55, 80]), ("Perovskite", 45, [10, 17, 30, 55, 80]),
                                           ("CrossedBarrel", 45, [10, 17, 30, 55, 80])])
def test_device_lbfgsb_follows_scipy(L, name, seed, ns):
    """stp_exact_fit (device L-BFGS-B) vs scipy L-BFGS-B on the oracle: same counts and optimum."""
    from stpbenchmark_b200 import priors
    X, y, tr, Xb, yb, nb = _batch(name, seed, ns, 90, "cuda")
    d = X.shape[1]
    lo, hi = priors.exact_bounds(d)
    theta = torch.tensor([priors.exact_theta0(d)] * len(ns), dtype=torch.double, device="cuda")
    f, nit, nfev, st = L.exact_fit(Xb, yb, theta, lo, hi, priors.exact_prior_vector(d), n=nb)
    assert st.tolist() == [0] * len(ns)
    same = 0
    for b, n in enumerate(ns):
        ths, res = oexact.fit_exact(X[tr[:n]], y[tr[:n]])
        assert abs(f[b].item() - res.fun) <= 1e-6 * max(1.0, abs(res.fun))
        if (nit[b].item(), nfev[b].item()) == (res.nit, res.nfev):
            same += 1
            assert np.abs(theta[b].cpu().numpy() - ths).max() <= 1e-6 * np.abs(ths).max()
    assert same >= len(ns) - 1


def test_scipy_lockstep_parity_mode(L):
    """Host scipy L-BFGS-B driving the CUDA closure == scipy on the oracle closure."""
    from stpbenchmark_b200 import optim, priors
    ns = [17, 30, 55]
    X, y, tr, Xb, yb, nb = _batch("AgNP", 45, ns, 56, "cuda")
    d = X.shape[1]
    lo, hi = priors.exact_bounds(d)
    theta, nit, nfev, status = optim.fit_exact_scipy_lockstep(Xb, yb, nb, priors.exact_theta0(d), lo, hi,
                                                              priors.exact_prior_vector(d))
    for b, n in enumerate(ns):
        ths, res = oexact.fit_exact(X[tr[:n]], y[tr[:n]])
        assert status[b] == 0 and (nit[b], nfev[b]) == (res.nit, res.nfev)
        assert np.abs(theta[b].numpy() - ths).max() <= 1e-7 * np.abs(ths).max()


@pytest.mark.parametrize("name", DATASETS)
def test_fused_acquisition_pointwise_and_argmax(L, name):
    ns = [10, 25, 40, 77]
    X, y, tr, Xb, yb, nb = _batch(name, 359, ns, 80, "cuda")
    thetas = np.stack([oexact.fit_exact(X[tr[:n]], y[tr[:n]])[0] for n in ns])
    mask = torch.ones(len(ns), len(X), dtype=torch.uint8)
    for b, n in enumerate(ns):
        mask[b, tr[:n]] = 0
    best = torch.tensor([float(y[tr[:n]].max()) for n in ns], dtype=torch.double)
    out = L.exact_acq(X.cuda().unsqueeze(0).contiguous(), Xb, yb, torch.tensor(thetas).cuda(), best_f=best.cuda(),
                      mask=mask.cuda(), n=nb, want_pointwise=True)
    assert out["status"].tolist() == [0] * len(ns)
    for b, n in enumerate(ns):
        mean, var = oexact.posterior_exact(thetas[b], X[tr[:n]], y[tr[:n]], X)
        acq = oacq.log_ei(mean, var, y[tr[:n]].max())
        assert (out["mean"][b].cpu() - mean).abs().max() <= 1e-9 * mean.abs().max()
        assert ((out["var"][b].cpu() - var).abs() / var.abs()).max() <= 1e-8
        assert ((out["acq"][b].cpu() - acq).abs() / acq.abs().clamp_min(1.0)).max() <= 1e-6
        ref = acq.clone(); ref[mask[b] == 0] = -float("inf")
        assert out["idx"][b].item() == int(torch.argmax(ref))
        assert out["acq_best"][b].item() == out["acq"][b, out["idx"][b].item()].item()


def test_acquisition_edge_cases(L):
    from stpbenchmark_b200 import priors
    X, y, tr, Xb, yb, nb = _batch("AutoAM", 45, [12, 12, 12], 16, "cuda")
    d = X.shape[1]
    th = torch.tensor([priors.exact_theta0(d)] * 3, dtype=torch.double, device="cuda")
    mask = torch.zeros(3, len(X), dtype=torch.uint8); mask[1, 77] = 1; mask[2] = 1
    best = torch.tensor([0.0, 0.0, float("nan")], dtype=torch.double)
    out = L.exact_acq(X.cuda().unsqueeze(0).contiguous(), Xb, yb, th, best_f=best.cuda(), mask=mask.cuda(), n=nb)
    assert out["idx"].tolist() == [-1, 77, 0]        # empty set; single candidate; NaN wins at the first index


@pytest.mark.parametrize("name,seeds", [("AgNP", [45, 359, 102]), ("Perovskite", [45, 359, 6185]), ("AutoAM", [359, 123, 5676]),
                                         ("CrossedBarrel", [45, 359, 123])])
def test_free_running_campaigns_reproduce_golden_traces(name, seeds):
    """L3 parity: whole 80-trial campaigns (device fit + fused step) equal the reference's
    gold_runs/benchmark_1 traces on known-answer seeds; at least 2 of 3 campaigns must be
    index-exact over all 90 rows and every campaign over the first 20."""
    from stpbenchmark_b200.runners import run_campaign_batch
    X, y = dataset(name)
    idx, val, status = run_campaign_batch(X, y, "ExactGP", seeds, n_initial=10, n_trials=80)
    assert status == [0] * len(seeds)
    exact = 0
    for s, row, v in zip(seeds, idx, val):
        gold = trace("gold1", "ExactGP", name, s)
        assert len(row) == 90 and row[:20] == gold[:20]
        assert v == [float(y[i]) for i in row]
        exact += row == gold
    assert exact >= 2


def test_cfg1_smoke_trace_via_drop_in_api():
    """BASELINE config #1 through the reference-shaped API: run_single_loop(ExactGP, AgNP, seed 45, 30 trials)."""
    from stpbenchmark_b200 import models, runners
    X, y = dataset("AgNP")
    sel, vals = runners.run_single_loop(X, y, models.ExactGP, n_initial=10, n_trials=30, seed=45)
    assert sel == trace("smoke", "ExactGP", "AgNP", 45) and len(vals) == 40


def test_model_level_api(L):
    """ExactGP object: train_exact_model_botorch (device and scipy modes) + posterior + LogExpectedImprovement."""
    from stpbenchmark_b200 import acqf, models, optim
    X, y = dataset("AgNP")
    tr = trace("gold1", "ExactGP", "AgNP", 45)
    n = 30
    Xt, yt = X[tr[:n]], y[tr[:n]]
    ths, res = oexact.fit_exact(Xt, yt)
    for mode in ("device", "scipy"):
        m = models.ExactGP(X_train=Xt, y_train=yt, input_transform=None).double()
        optim.train_exact_model_botorch(m, Xt, yt, optimizer=mode)
        assert np.abs(m.theta().cpu().numpy() - ths).max() <= 1e-6 * np.abs(ths).max()
        assert m.fit_info["nit"] == res.nit
        m.eval()
        mask = torch.ones(len(X), dtype=torch.bool); mask[tr[:n]] = False
        acq = acqf.LogExpectedImprovement(m, best_f=yt.max(), maximize=True).forward(X[mask].unsqueeze(1))
        assert acq.shape == (int(mask.sum()),)
        mean, var = oexact.posterior_exact(ths, Xt, yt, X[mask])
        ref = oacq.log_ei(mean, var, yt.max())
        assert ((acq.cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max() <= 1e-6
        assert int(torch.where(mask)[0][torch.argmax(acq)]) == tr[n]
        post = m.posterior(X[:7])                       # [Nt, d] layout
        assert post.mean.shape == (7,) and post.variance.shape == (7,)
        lat = m(X[:7])
        assert torch.allclose(lat.variance + m.likelihood.noise.detach().double(), post.variance)
        lo, hi = lat.confidence_region()
        assert (hi > lo).all()
        assert torch.allclose(acqf.log_expected_improvement(m, X[:7], yt.max()).cpu(),
                              _ref_log_ei(post.mean.cpu(), post.variance.cpu(), yt.max()), atol=1e-12)


def _ref_log_ei(mean, var, best_f):
    """The reference's acqf.log_expected_improvement formula (acqf.py:81-107)."""
    std = var.sqrt() + 1e-9
    gain = mean - best_f
    z = gain / std
    n = torch.distributions.Normal(0, 1)
    pos = gain > 0
    out = torch.where(pos, torch.log(torch.where(pos, gain, torch.ones_like(gain))) + torch.log(n.cdf(z) + 1e-9),
                      n.log_prob(z) + torch.log(std))
    return torch.where(std > 1e-9, out, torch.zeros_like(gain))


def test_synthetic_cfg4_batch_matches_oracle_step():
    """Config #4 shape at reduced batch: 16 campaigns on 2 synthetic pools (N = 2048, d = 6), 6 iterations;
    every pick equals the oracle's pick from the same state (teacher-forced L2 parity)."""
    from stpbenchmark_b200 import synthetic
    from stpbenchmark_b200.engine import CampaignBatch
    pools = [synthetic.cfg4_pool(g) for g in range(2)]
    PX = torch.stack([p[0] for p in pools]); PY = torch.stack([p[1] for p in pools])
    pid = [b // 8 for b in range(16)]
    designs = [synthetic.initial_designs(pools[g][0], pools[g][1], [synthetic.campaign_seed(g, k) for k in range(8)]) for g in range(2)]
    init = [designs[pid[b]][b % 8].tolist() for b in range(16)]
    batch = CampaignBatch(PX, PY, pid, init, n_cap=20, kind="ExactGP")
    batch.run(6)
    idx, n, status = batch.traces()
    assert status.tolist() == [0] * 16 and n.tolist() == [16] * 16
    agree = total = 0
    for b in (0, 5, 9, 15):
        X, y = pools[pid[b]]
        for t in (10, 13, 15):
            picked = idx[b, :t].tolist()
            mask = torch.ones(len(X), dtype=torch.bool); mask[picked] = False
            th, _ = oexact.fit_exact(X[picked], y[picked])
            mean, var = oexact.posterior_exact(th, X[picked], y[picked], X[mask])
            pick = int(torch.where(mask)[0][oacq.first_argmax(oacq.log_ei(mean, var, y[picked].max()))])
            agree += pick == int(idx[b, t]); total += 1
    assert agree == total


def test_in_kernel_turnover_continuous_batching():
    """stp_turnover: finished campaigns hand their trace to the sink and their slot to the next design of the bank,
    inside the kernel.  Every finished trace equals the trace of a stand-alone campaign from the same design."""
    from stpbenchmark_b200.engine import CampaignBatch
    from stpbenchmark_b200.runners import initial_design
    X, y = dataset("AgNP")
    seeds = [[45, 359, 6185], [102, 421, 123]]                     # slot -> generations 0, 1, 2
    designs = [[initial_design(X, y, s, 10)[0] for s in row] for row in seeds]
    n_trials = 5
    batch = CampaignBatch(X, y, [0, 0], [designs[0][0], designs[1][0]], n_cap=16, kind="ExactGP")
    bank = torch.tensor([[designs[0][1], designs[0][2]], [designs[1][1], designs[1][2]]], dtype=torch.int32)
    batch.enable_turnover(bank, n_final=10 + n_trials, sink_capacity=8)
    batch.run(12)                                                  # 5 + 5 + 2 iterations per slot
    idx, n, status = batch.traces()
    assert status.tolist() == [0, 0] and n.tolist() == [12, 12]
    assert int(batch.sink_count.item()) == 4 and batch.gen.tolist() == [2, 2]
    assert batch.iters_acc.tolist() == [12, 12] and (batch.work_acc > 0).all()
    finished = {tuple(r[:10]): r[:15] for r in batch.sink[:4].cpu().tolist()}
    for slot in range(2):
        for g in range(2):
            ref = CampaignBatch(X, y, [0], [designs[slot][g]], n_cap=16, kind="ExactGP")
            ref.run(n_trials)
            want = ref.traces()[0][0, :15].tolist()
            assert finished[tuple(designs[slot][g])] == want
        cur = CampaignBatch(X, y, [0], [designs[slot][2]], n_cap=16, kind="ExactGP")
        cur.run(2)
        assert idx[slot, :12].tolist() == cur.traces()[0][0, :12].tolist()
    # bank exhausted -> the campaign stops with status 6 after its last trial
    batch.run(3)
    assert batch.traces()[2].tolist() == [6, 6] and int(batch.sink_count.item()) == 6


def test_campaign_groups_on_streams_equal_single_batch():
    """CampaignGroups (4 groups, 4 streams) produces exactly the traces of one CampaignBatch."""
    from stpbenchmark_b200 import synthetic
    from stpbenchmark_b200.engine import CampaignBatch, CampaignGroups
    X, y = synthetic.cfg4_pool(3, N=512)
    seeds = list(range(20, 32))
    init = synthetic.initial_designs(X, y, seeds).tolist()
    one = CampaignBatch(X, y, [0] * 12, init, n_cap=20, kind="ExactGP")
    one.run(8)
    many = CampaignGroups(X, y, [0] * 12, init, n_cap=20, kind="ExactGP", groups=4)
    many.run(8)
    a, b = one.traces(), many.traces()
    assert all(torch.equal(p, q) for p, q in zip(a, b)) and many.launches == 32


def test_one_device_job_for_the_whole_sweep_equals_per_dataset_batches(tmp_path):
    """run_benchmark.py:20-43 as ONE device job (runners.run_sweep -> engine.CampaignJob): datasets of equal input
    dimension share launches although their pools differ in size (AgNP 164 and P3HT 178 rows, zero-padded + masked),
    model kinds run side by side on their own streams -- and every trace equals what the same job gives when it is
    submitted per (model, dataset) like the reference does."""
    import pandas as pd
    from stpbenchmark_b200 import models, runners
    from stpbenchmark_b200.engine import CampaignJob
    spec = [("ExactGP", "AgNP", models.ExactGP), ("ExactGP", "P3HT", models.ExactGP), ("ExactGP", "Perovskite", models.ExactGP),
            ("VarGP", "AutoAM", models.VarGP), ("VarGP", "CrossedBarrel", models.VarGP)]
    def jobs(sub):
        out = []
        for mname, dname, cls in spec:
            X, y = dataset(dname)
            out.append(dict(X=X, y=y, model_class=cls, seeds=[45, 359, 6185], n_initial=10, n_trials=7, epochs=25,
                            learning_rate=0.01, output_path=str(tmp_path / sub / f"{mname}_{dname}.csv"),
                            invert_y=dname in ("AgNP", "Perovskite")))
        return out
    seen = []
    orig = CampaignJob.__init__
    def spy(self, *a, **k):
        orig(self, *a, **k)
        seen.append(sorted((g["key"][0], g["key"][1], len(g["ids"]), g["batch"].N) for g in self.groups))
    CampaignJob.__init__ = spy
    try:
        runners.run_sweep(jobs("all"))
    finally:
        CampaignJob.__init__ = orig
    assert seen == [[("ExactGP", 3, 3, 94), ("ExactGP", 5, 6, 178), ("VarGP", 4, 6, 600)]]
    for job in jobs("each"):
        runners.run_sweep([job])
    for mname, dname, _ in spec:
        a = pd.read_csv(tmp_path / "all" / f"{mname}_{dname}.csv", header=[0, 1], index_col=0)
        b = pd.read_csv(tmp_path / "each" / f"{mname}_{dname}.csv", header=[0, 1], index_col=0)
        assert a.shape == (17, 6) and not a.isna().any().any()
        assert a.equals(b), (mname, dname)
        for s_ in (45, 359, 6185):
            assert a[f"seed_{s_}", "x_index"].astype(int).tolist()[:10] == trace("gold1", mname, dname, s_)[:10]
    assert pd.read_csv(tmp_path / "all" / "ExactGP_AgNP.csv", header=[0, 1], index_col=0)["seed_45", "x_index"].astype(
        int).tolist() == trace("gold1", "ExactGP", "AgNP", 45)[:17]


def test_persistent_work_queue_matches_lockstep_steps():
    """stp_exact_bo_run (persistent CTAs serving campaign-iterations from a ticket ring) advances every campaign
    exactly like iters x stp_exact_bo_step: same picks, same fitted theta -- with as many CTAs as fit, with 3 CTAs
    for 12 campaigns (every CTA serves several campaigns, iterations of a campaign hop between SMs), with
    per-campaign iteration counts, and with in-kernel turnover."""
    from stpbenchmark_b200 import synthetic
    from stpbenchmark_b200.engine import CampaignBatch
    X, y = dataset("AgNP")
    seeds = list(range(40, 52))
    init = synthetic.initial_designs(X, y, seeds).tolist()
    ref = CampaignBatch(X, y, [0] * 12, init, n_cap=24, kind="ExactGP")
    for _ in range(9):
        ref.step()
    a = ref.traces()
    for max_ctas in (0, 3):
        q = CampaignBatch(X, y, [0] * 12, init, n_cap=24, kind="ExactGP")
        q.run_queue(9, max_ctas=max_ctas)
        b = q.traces()
        assert all(torch.equal(u, v) for u, v in zip(a, b)), max_ctas
        assert torch.equal(ref.theta, q.theta) and torch.equal(ref.nfev, q.nfev)
        assert q.launches == 1 and int(q._iters_left.sum()) == 0 and int(q._sched[:12].sum()) == 0
    # chunked run() and ragged iteration counts
    q = CampaignBatch(X, y, [0] * 12, init, n_cap=24, kind="ExactGP")
    q.run(9, chunk=4)
    assert all(torch.equal(u, v) for u, v in zip(a, q.traces())) and q.launches == 3
    per = [b % 5 for b in range(12)]
    q = CampaignBatch(X, y, [0] * 12, init, n_cap=24, kind="ExactGP")
    q.run_queue(per, max_ctas=5)
    idx, n, st = q.traces()
    assert n.tolist() == [len(init[b]) + per[b] for b in range(12)] and st.tolist() == [0] * 12
    for b in range(12):
        assert idx[b, : int(n[b])].tolist() == a[0][b, : int(n[b])].tolist()
    # turnover inside the persistent launch: 3 generations of 4-trial campaigns, finished traces land in the sink
    bank = synthetic.initial_designs(X, y, list(range(100, 124))).reshape(12, 2, -1)
    outs = []
    for mode in ("step", "queue"):
        t = CampaignBatch(X, y, [0] * 12, init, n_cap=24, kind="ExactGP")
        t.enable_turnover(bank, n_final=len(init[0]) + 4, sink_capacity=40)
        if mode == "step":
            for _ in range(10):
                t.step()
        else:
            t.run_queue(10, max_ctas=4)
        torch.cuda.synchronize()
        sink = t.sink[: int(t.sink_count)].cpu()
        outs.append((sorted(map(tuple, sink.tolist())), t.traces()))
    assert outs[0][0] == outs[1][0] and len(outs[0][0]) == 24
    assert all(torch.equal(u, v) for u, v in zip(outs[0][1], outs[1][1]))


def test_launch_sized_for_current_n_gives_the_same_campaigns():
    """stp_exact_bo_step's n_max only sizes the launch (shared memory, threads): a lockstep batch run with the
    engine's host mirror of n (n_max = 12 ... 27: 128 threads, 4 CTAs/SM) picks what it picks when every launch is
    sized for the row capacity (256 threads)."""
    from stpbenchmark_b200.engine import CampaignBatch
    X, y = dataset("AgNP")
    designs = [list(trace("gold1", "ExactGP", "AgNP", s)[:12]) for s in (45, 359, 102, 123)]
    runs = []
    for hint in (True, False):
        bt = CampaignBatch(X, y, [0] * 4, designs, n_cap=80, kind="ExactGP")
        if not hint:
            bt.assume_ages(None)
        bt.run(16)
        idx, n, st = bt.traces()
        assert st.tolist() == [0] * 4 and n.tolist() == [28] * 4
        runs.append((idx[:, :28].tolist(), bt.theta.cpu()))
    assert runs[0][0] == runs[1][0]
    assert torch.allclose(runs[0][1], runs[1][1], rtol=1e-4, atol=1e-7)


def test_n_max_smaller_than_a_training_set_stops_that_campaign(L):
    """A campaign whose n exceeds the caller's n_max does not fit the launch's shared memory: status 2, nothing
    written; the others run."""
    from stpbenchmark_b200 import priors
    X, y, tr, Xb, yb, nb = _batch("AgNP", 45, [12, 20], 32, "cuda")
    d, N = X.shape[1], len(X)
    mask = torch.ones(2, N, dtype=torch.uint8)
    trace_t = torch.full((2, 32), -1, dtype=torch.int32)
    for b, n in enumerate([12, 20]):
        mask[b, tr[:n]] = 0
        trace_t[b, :n] = torch.tensor(tr[:n], dtype=torch.int32)
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    lo, hi = priors.exact_bounds(d)
    L.exact_bo_step(X.cuda().unsqueeze(0).contiguous(), y.cuda().unsqueeze(0).contiguous(),
                    torch.zeros(2, dtype=torch.int32, device="cuda"), mask.cuda(), Xb, yb, nb, trace_t.cuda(),
                    priors.exact_theta0(d), lo, hi, priors.exact_prior_vector(d), status, n_max=15)
    assert status.tolist() == [0, 2] and nb.tolist() == [13, 20]
    with pytest.raises(RuntimeError):
        L.exact_bo_step(X.cuda().unsqueeze(0).contiguous(), y.cuda().unsqueeze(0).contiguous(),
                        torch.zeros(2, dtype=torch.int32, device="cuda"), mask.cuda(), Xb, yb, nb, trace_t.cuda(),
                        priors.exact_theta0(d), lo, hi, priors.exact_prior_vector(d), status, n_max=32)


def test_largest_supported_exact_campaigns(L):
    """n_cap = 142 is the largest capacity whose shared-memory carve-up with the optimiser state fits in 227 KB
    (144 without it).  Closure parity at n = 96 ... 141 (7 - 9 tile rows in the blocked sweep, partial last blocks),
    then fit + fused acquisition of the n = 141 campaign against the oracle."""
    from stpbenchmark_b200 import priors, synthetic
    X, y = synthetic.cfg4_pool(1)
    d, cap, ns = X.shape[1], 142, [96, 97, 120, 141]
    g = torch.Generator().manual_seed(11)
    sels = [torch.randperm(len(X), generator=g)[:n] for n in ns]
    Xb = torch.zeros(len(ns), cap, d, dtype=torch.double); yb = torch.zeros(len(ns), cap, dtype=torch.double)
    for b, sel in enumerate(sels):
        Xb[b, :ns[b]] = X[sel]; yb[b, :ns[b]] = y[sel]
    Xb, yb = Xb.cuda(), yb.cuda()
    nb = torch.tensor(ns, dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(5)
    th = np.stack([np.array(priors.exact_theta0(d)) * rng.uniform(0.7, 1.5, d + 3) for _ in ns])
    f, gr, st = L.exact_mll_fg(Xb, yb, torch.tensor(th, device="cuda"), priors.exact_prior_vector(d), n=nb)
    assert st.tolist() == [0] * len(ns)
    for b, (n, sel) in enumerate(zip(ns, sels)):
        fo, go = oexact.mll_value_and_grad(th[b], X[sel], y[sel])
        assert abs(f[b].item() - fo) <= 1e-8 * max(1.0, abs(fo))
        assert np.abs(gr[b].cpu().numpy() - go).max() <= 1e-8 * max(1.0, np.abs(go).max())
    # device fit of the largest one, then the fused pool pass at the device optimum
    lo, hi = priors.exact_bounds(d)
    theta = torch.tensor([priors.exact_theta0(d)], dtype=torch.double, device="cuda")
    fit = L.exact_fit(Xb[3:4].contiguous(), yb[3:4].contiguous(), theta, lo, hi, priors.exact_prior_vector(d), n=nb[3:4].contiguous())
    assert int(fit[3].item()) == 0
    sel = sels[3]
    mask = torch.ones(1, len(X), dtype=torch.uint8); mask[0, sel] = 0
    best = torch.tensor([float(y[sel].max())], dtype=torch.double)
    out = L.exact_acq(X.cuda().unsqueeze(0).contiguous(), Xb[3:4].contiguous(), yb[3:4].contiguous(), theta, best_f=best.cuda(),
                      mask=mask.cuda(), n=nb[3:4].contiguous(), want_pointwise=True)
    assert out["status"].tolist() == [0]
    mean, var = oexact.posterior_exact(theta[0].cpu().numpy(), X[sel], y[sel], X)
    acq = oacq.log_ei(mean, var, y[sel].max())
    assert (out["mean"][0].cpu() - mean).abs().max() <= 1e-8 * mean.abs().max()
    assert ((out["var"][0].cpu() - var).abs() / var.abs()).max() <= 1e-7
    ref = acq.clone(); ref[mask[0] == 0] = -float("inf")
    assert out["idx"][0].item() == int(torch.argmax(ref))


def test_train_exact_model_adam_uses_the_standardised_targets():
    """optim.train_exact_model (reference optim.py:35-58): plain Adam on -mll(model(X), y) with the STANDARDISED y.
    Loss history against torch.optim.Adam on the oracle's neg_mll (autograd) with the same targets."""
    from stpbenchmark_b200 import optim, priors, utils
    from stpbenchmark_b200.models import ExactGP
    X, y = dataset("AgNP")
    sel = trace("gold1", "ExactGP", "AgNP", 45)[:22]
    Xt, yt = X[sel], y[sel]
    model = ExactGP(Xt, yt).double()
    hist = optim.train_exact_model(model, Xt, yt, epochs=6, lr=0.01, verbose=True)
    ys = utils.TorchStandardScaler().fit_transform(yt)
    theta = torch.tensor(priors.exact_theta0(X.shape[1]), dtype=torch.double, requires_grad=True)
    opt = torch.optim.Adam([theta], lr=0.01)
    want = []
    for _ in range(6):
        opt.zero_grad()
        loss = oexact.neg_mll(theta, Xt, ys)
        want.append(float(loss.detach()))
        loss.backward()
        opt.step()
    assert np.abs(np.array(hist) - np.array(want)).max() <= 1e-9 * np.abs(want).max()
    raw = float(oexact.neg_mll(torch.tensor(priors.exact_theta0(X.shape[1]), dtype=torch.double), Xt, yt))
    assert abs(hist[0] - raw) > 1e-3          # and it is NOT the raw-y loss
