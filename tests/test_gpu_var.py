"""Parity of the CUDA variational path (VarGP / VarTGP / VarLGP) with the CPU oracle through the C ABI.

Per SURVEY section 8(c) the 1e-5 tolerance is meaningful per kernel evaluation on identical state and over
short horizons (<= 50 epochs); 600-epoch end-of-fit states are chaotic in round-off and are compared
statistically only (teacher-forced rank of the oracle's pick)."""
import numpy as np
import pytest
import torch

from _fixtures import dataset, trace
from _varhelp import GH, LIK, PRIOR, check_grads, pack, perturbed_state, rel, split_grads
from oracle import acq as oacq, variational as ov

pytestmark = pytest.mark.gpu
KINDS = ["VarGP", "VarTGP", "VarLGP"]


def _dev_state(packs, cap, d):
    from stpbenchmark_b200 import varops
    st = varops.alloc_state(len(packs), cap, d, "cuda")
    for b, (Z, hyp, nat1, nat2) in enumerate(packs):
        st["Z"][b] = torch.tensor(Z); st["hyp"][b] = torch.tensor(hyp)
        st["nat1"][b] = torch.tensor(nat1); st["nat2"][b] = torch.tensor(nat2)
    return st


def _sets(ns, name="AutoAM", seed=45, cap=None):
    X, y = dataset(name)
    tr = trace("gold1", "ExactGP", name, seed)
    cap = cap or max(ns) + 2
    d = X.shape[1]
    Xb = torch.zeros(len(ns), cap, d, dtype=torch.double); yb = torch.zeros(len(ns), cap, dtype=torch.double)
    sels = []
    for b, n in enumerate(ns):
        sel = tr[:n]
        if n > 5:
            sel[3] = sel[1]
        sels.append(sel)
        Xb[b, :n] = X[sel]; yb[b, :n] = y[sel]
    return X, y, sels, Xb.cuda(), yb.cuda(), torch.tensor(ns, dtype=torch.int32, device="cuda"), cap


@pytest.mark.parametrize("kind", KINDS)
def test_elbo_value_and_gradients(kind):
    from stpbenchmark_b200 import _lib
    ns = [3, 14, 33, 60]
    X, y, sels, Xb, yb, nb, cap = _sets(ns)
    d = X.shape[1]
    states = [perturbed_state(X[s], kind, seed=b + 1) for b, s in enumerate(sels)]
    dev = _dev_state([pack(st, cap) for st in states], cap, d)
    work = _lib.var_workspace(len(ns), cap, "cuda")
    loss, grads, status = _lib.var_elbo_fg(Xb, yb, nb, dev, LIK[kind], PRIOR, GH, work)
    assert status.tolist() == [0] * len(ns)
    for b, (n, sel, st) in enumerate(zip(ns, sels, states)):
        lo, ref = ov.epoch(st, X[sel], ov.standardise(y[sel]), lr=0.0)
        assert abs(loss[b].item() - lo) <= 1e-10 * max(1.0, abs(lo))
        check_grads(split_grads(grads[b].cpu().numpy(), n, d, cap), ref, d, 1e-8)


@pytest.mark.parametrize("kind", KINDS)
def test_short_horizon_training_and_pool_pass(kind):
    from stpbenchmark_b200 import _lib, varops
    ns, E = [12, 25, 40], 40
    X, y, sels, Xb, yb, nb, cap = _sets(ns)
    d, B, N, S = X.shape[1], len(ns), len(X), 128
    noise = torch.zeros(B, cap, dtype=torch.double)
    for b, n in enumerate(ns):
        noise[b, :n] = torch.randn(n, dtype=torch.double, generator=torch.Generator().manual_seed(10 + b))
    st = varops.alloc_state(B, cap, d, "cuda")
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    work = _lib.var_workspace(B, cap, "cuda")
    hist = torch.zeros(B, E, dtype=torch.double, device="cuda")
    _lib.var_fit(Xb, yb, nb, st, LIK[kind], E, 0.01, 0.01, varops.var_hyp0(d), PRIOR, GH, status, work, fresh=True,
                 init_noise=noise.cuda(), loss_hist=hist)
    assert status.tolist() == [0] * B
    eps = torch.randn(B, N, S, dtype=torch.double, generator=torch.Generator().manual_seed(99))
    mask = torch.ones(B, N, dtype=torch.uint8)
    for b, sel in enumerate(sels):
        mask[b, sel] = 0
    best = torch.tensor([float(y[s].max()) for s in sels], dtype=torch.double)
    out = _lib.var_acq(X.cuda().unsqueeze(0).contiguous(), st, nb, LIK[kind], status, work, best_f=best.cuda(), mask=mask.cuda(),
                       S=S, eps=eps.cuda(), want_pointwise=True)
    assert status.tolist() == [0] * B
    for b, (n, sel) in enumerate(zip(ns, sels)):
        ost = ov.train_natural(X[sel], y[sel], kind, E, 0.01, init_noise=noise[b, :n])
        assert rel(hist[b].cpu().numpy(), ost.losses) <= 1e-7
        assert rel(st["Z"][b, :n].cpu().numpy(), ost.Z.detach().numpy()) <= 1e-5
        assert rel(st["nat2"][b, :n, :n].cpu().numpy(), ost.nat2.numpy()) <= 1e-5
        mo, vo = ov.latent_posterior(ost, X)
        ao = ov.acquisition(ost, X, float(y[sel].max()), eps=eps[b])
        assert rel(out["mean"][b].cpu().numpy(), mo.numpy()) <= 1e-5
        assert ((out["var"][b].cpu() - vo).abs() / vo).max() <= 1e-5
        assert ((out["acq"][b].cpu() - ao).abs() / ao.abs().clamp_min(1.0)).max() <= 1e-5
        ref = ao.clone(); ref[mask[b] == 0] = -float("inf")
        assert out["idx"][b].item() == int(torch.argmax(ref))


def test_philox_stream_is_what_the_kernel_draws():
    """In-kernel Philox eps: the oracle fed with stp_philox_normal's stream reproduces the MC acquisition."""
    from stpbenchmark_b200 import _lib, varops
    ns, E, S = [15], 30, 256
    X, y, sels, Xb, yb, nb, cap = _sets(ns)
    d, N = X.shape[1], len(X)
    noise = torch.zeros(1, cap, dtype=torch.double); noise[0, :15] = torch.randn(15, dtype=torch.double, generator=torch.Generator().manual_seed(4))
    st = varops.alloc_state(1, cap, d, "cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda"); work = _lib.var_workspace(1, cap, "cuda")
    _lib.var_fit(Xb, yb, nb, st, 1, E, 0.01, 0.01, varops.var_hyp0(d), PRIOR, GH, status, work, fresh=True, init_noise=noise.cuda())
    best = torch.tensor([float(y[sels[0]].max())], dtype=torch.double, device="cuda")
    out = _lib.var_acq(X.cuda().unsqueeze(0).contiguous(), st, nb, 1, status, work, best_f=best, S=S, seed=777, want_pointwise=True)
    eps = _lib.philox_normal(777, 0, 0, N * S).reshape(N, S).cpu()       # stream = campaign 0, counter = cand * S/2 + pair
    ost = ov.train_natural(X[sels[0]], y[sels[0]], "VarTGP", E, 0.01, init_noise=noise[0, :15])
    ao = ov.acquisition(ost, X, float(best.item()), eps=eps)
    assert ((out["acq"][0].cpu() - ao).abs() / ao.abs().clamp_min(1.0)).max() <= 1e-5
    assert out["idx"][0].item() == int(torch.argmax(ao))
    assert abs(eps.mean()) < 0.01 and abs(eps.std() - 1) < 0.01


def test_failure_codes_do_not_crash():
    from stpbenchmark_b200 import _lib, varops
    X, y, sels, Xb, yb, nb, cap = _sets([10, 10])
    d = X.shape[1]
    st = varops.alloc_state(2, cap, d, "cuda")
    status = torch.zeros(2, dtype=torch.int32, device="cuda"); work = _lib.var_workspace(2, cap, "cuda")
    _lib.var_fit(Xb, yb, nb, st, 0, 1, 0.01, 0.01, varops.var_hyp0(d), PRIOR, GH, status, work, fresh=True,
                 init_noise=torch.zeros(2, cap, dtype=torch.double, device="cuda"))
    st["nat2"][1] = torch.eye(cap, dtype=torch.double, device="cuda")       # P = -2 Theta2 not positive definite
    st["hyp"][0, 1] = float("nan")
    loss, grads, code = _lib.var_elbo_fg(Xb, yb, nb, st, 0, PRIOR, GH, work)
    assert code.tolist() == [2, 1]


@pytest.mark.parametrize("cls_name", ["VarGP", "VarTGP", "VarLGP"])
def test_model_level_api(cls_name):
    """models.Var* + optim.train_natural_variational_model + posterior + LogExpectedImprovement, as the
    reference's loop body uses them (runners.py:66-95)."""
    from stpbenchmark_b200 import acqf, models, optim
    X, y = dataset("AutoAM")
    tr = trace("gold1", "ExactGP", "AutoAM", 45)
    n, E = 18, 40
    Xt, yt = X[tr[:n]], y[tr[:n]]
    cls = getattr(models, cls_name)
    torch.manual_seed(5)
    model = cls(inducing_points=Xt).double()
    optim.train_natural_variational_model(model, Xt, yt, epochs=E, lr=0.01)
    torch.manual_seed(5)
    ost = ov.train_natural(Xt, yt, cls_name, E, 0.01)            # draws the same n normals from the global generator
    assert rel(model.variational_strategy.inducing_points.detach().cpu().numpy(), ost.Z.detach().numpy()) <= 1e-5
    assert rel(model.covar_module.base_kernel.lengthscale.detach().cpu().numpy().ravel(), ost.ls.detach().numpy().ravel()) <= 1e-6
    assert abs(float(model.likelihood.noise.detach()) - ost.noise.item()) <= 1e-6 * ost.noise.item()
    if cls_name == "VarTGP":
        assert abs(float(model.likelihood.deg_free.detach()) - ost.df().item()) <= 1e-6 * ost.df().item()
    model.eval()
    mask = torch.ones(len(X), dtype=torch.bool); mask[tr[:n]] = False
    Xc = X[mask]
    mo, vo = ov.latent_posterior(ost, Xc)
    lat = model(Xc)
    assert rel(lat.mean.cpu().numpy(), mo.numpy()) <= 1e-5 and ((lat.variance.cpu() - vo).abs() / vo).max() <= 1e-5
    if cls_name == "VarGP":
        acq = acqf.LogExpectedImprovement(model, best_f=yt.max(), maximize=True).forward(Xc.unsqueeze(1))
        ref = ov.acquisition(ost, Xc, yt.max())
        assert acq.shape == (len(Xc),) and ((acq.cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max() <= 1e-5
        assert int(torch.argmax(acq)) == int(torch.argmax(ref))
    else:
        S = 64
        eps = torch.randn(len(Xc), S, dtype=torch.double, generator=torch.Generator().manual_seed(8))
        post = model.posterior(Xc.unsqueeze(1), num_samples=S, eps=eps)
        assert post.mean.shape == (S, len(Xc), 1) and post.variance.shape == post.mean.shape
        acq = acqf.LogExpectedImprovement(model, best_f=yt.max(), maximize=True)
        model.posterior = lambda X, **kw: post                 # reuse the same draws inside forward()
        vals = acq.forward(Xc.unsqueeze(1))
        assert vals.shape == (S, len(Xc))
        ref = ov.acquisition(ost, Xc, yt.max(), eps=eps)
        got = vals.mean(0).cpu()
        assert ((got - ref).abs() / ref.abs().clamp_min(1.0)).max() <= 1e-5 and int(torch.argmax(got)) == int(torch.argmax(ref))
        if cls_name == "VarTGP":
            assert post.df is not None and torch.allclose(post.variance[0, 0, 0].cpu(), (ost.noise * ost.df() / (ost.df() - 2)).detach()[0], rtol=1e-6)


@pytest.mark.parametrize("kind,name,seed", [("VarGP", "AutoAM", 45), ("VarGP", "Perovskite", 359), ("VarGP", "P3HT", 45)])
def test_vargp_smoke_campaign_follows_golden_trace(kind, name, seed):
    """Free-running VarGP campaigns at SMOKE length (200 epochs, 30 trials) through the batched engine with the
    reference's RNG order for the theta1 noise.  The oracle reproduces these three golden traces exactly
    (SURVEY App. C.4); 200-epoch fits are already round-off sensitive, so the CUDA path must match the golden
    trace on a long prefix (>= 20 of 40 rows) rather than everywhere."""
    from stpbenchmark_b200.runners import run_campaign_batch
    X, y = dataset(name)
    idx, val, status = run_campaign_batch(X, y, kind, [seed], n_initial=10, n_trials=30, epochs=200, learning_rate=0.01)
    gold = trace("smoke", kind, name, seed)
    assert status == [0] and len(idx[0]) == 40 and idx[0][:10] == gold[:10]
    first = next((i for i in range(40) if idx[0][i] != gold[i]), 40)
    assert first >= 20, (first, idx[0], gold)


def test_vargp_known_answer_smoke_campaigns_on_the_cuda_path():
    """GPU twin of tests/test_oracle_var_golden.py: the five VarGP SMOKE campaigns the oracle is pinned on
    (results/VarGP_*_traces_SMOKE.csv), free-running through the drop-in run_single_loop (device fit, 200 epochs,
    30 trials, the reference's RNG order for the theta1 noise).  The 200-epoch fit is round-off sensitive (the CPU
    oracle itself leaves one of the five at row 32), so the bar is the oracle's own result: every campaign follows the
    reference's picks for at least 30 of 40 rows, at least 4 of the 5 over all 40."""
    import _varpin
    from stpbenchmark_b200 import models, runners
    firsts = {}
    for name, seed in _varpin.VARGP_KAT:
        X, y = dataset(name)
        idx, val = runners.run_single_loop(X, y, models.VarGP, seed=seed, **_varpin.SMOKE)
        gold = trace("smoke", "VarGP", name, seed)
        assert len(idx) == 40 and idx[:10] == gold[:10]
        firsts[(name, seed)] = next((i for i in range(40) if idx[i] != gold[i]), 40)
    print("VarGP SMOKE first divergence (CUDA):", firsts)
    # measured on B200: exactly what the CPU oracle does (4 campaigns index-exact over 40 rows, AutoAM / 45 leaves at row 32)
    assert min(firsts.values()) >= 30, firsts
    assert sum(v == 40 for v in firsts.values()) >= 4, firsts


@pytest.mark.parametrize("kind,name,seed,min_first,worst", [("VarTGP", "AutoAM", 45, 18, 3), ("VarLGP", "AgNP", 45, 26, 3)])
def test_mc_models_teacher_forced_along_golden_smoke_trace_on_the_cuda_path(kind, name, seed, min_first, worst):
    """GPU twin of the oracle's teacher-forced pin: the 30 steps of the golden SMOKE trace are 30
    campaigns of ONE batch (training set = the first 10 + t golden rows); after stp_var_fit (200 epochs) the golden
    pick must be the argmax of the Monte-Carlo LogEI (S = 2048, in-kernel Philox) at the surveyed rate (>= 18 / >= 26
    of 30) and never rank below fourth (the CPU oracle with torch's stream: never below third; the Philox stream moves
    the Monte-Carlo near-ties, App. C.5)."""
    from stpbenchmark_b200 import _lib, varops
    from stpbenchmark_b200.engine import CampaignBatch
    X, y = dataset(name)
    gold = trace("smoke", kind, name, seed)
    T = 30
    batch = CampaignBatch(X, y, [0] * T, [gold[: 10 + t] for t in range(T)], n_cap=41, kind=kind, epochs=200,
                          learning_rate=0.01, num_samples=2048)
    v = batch._var
    _lib.var_fit(batch.X, batch.y, batch.n, v.state, v.lik, 200, 0.01, varops.ADAM_LR, v.hyp0, v.prior, v.gh, batch.status,
                 v.work, fresh=True, seed=seed)
    best_f = torch.stack([y[gold[: 10 + t]].max() for t in range(T)]).to(batch.device)
    out = _lib.var_acq(batch.pool, v.state, batch.n, v.lik, batch.status, v.work, best_f=best_f, mask=batch.mask,
                       pool_id=batch.pool_id, S=2048, seed=seed + 1, want_pointwise=True)
    assert batch.status.tolist() == [0] * T
    acq = out["acq"].cpu()
    ranks = []
    for t in range(T):
        avail = batch.mask[t].cpu().bool()
        ranks.append(int((acq[t][avail] > acq[t, gold[10 + t]]).sum()))
        assert int(out["idx"][t]) == int(torch.where(avail)[0][torch.argmax(acq[t][avail])])
    print(kind, "teacher-forced ranks (CUDA):", ranks)
    assert sum(r == 0 for r in ranks) >= min_first and max(ranks) <= worst, ranks


def test_vartgp_batched_campaigns_teacher_forced_rank():
    """VarTGP (the headline model): 4 campaigns x 3 iterations through the batched engine (in-kernel Philox);
    at every step the oracle, trained from the same state with the same theta1 noise, ranks the engine's pick
    within its top 3 under an independent eps draw (MC near-ties may flip, App. C.5)."""
    from stpbenchmark_b200.engine import CampaignBatch
    from stpbenchmark_b200.runners import initial_design
    X, y = dataset("AutoAM")
    seeds = [45, 359, 6185, 102]
    designs, states = zip(*[initial_design(X, y, s, 10) for s in seeds])
    batch = CampaignBatch(X, y, [0] * 4, designs, n_cap=14, kind="VarTGP", epochs=60, learning_rate=0.01, num_samples=1024)
    batch._var.set_rng_states(states)
    batch.run(3)
    idx, n, status = batch.traces()
    assert status.tolist() == [0] * 4 and n.tolist() == [13] * 4
    hits = 0
    for b, s in enumerate(seeds):
        g = torch.Generator(); g.set_state(states[b])
        picked = idx[b, :10].tolist()
        noise = torch.randn(10, dtype=torch.double, generator=g)
        ost = ov.train_natural(X[picked], y[picked], "VarTGP", 60, 0.01, init_noise=noise)
        mask = torch.ones(len(X), dtype=torch.bool); mask[picked] = False
        acq = ov.acquisition(ost, X[mask], y[picked].max(), eps=torch.randn(int(mask.sum()), 1024, dtype=torch.double,
                                                                           generator=torch.Generator().manual_seed(b)))
        order = torch.where(mask)[0][torch.argsort(acq, descending=True)].tolist()
        hits += int(idx[b, 10]) in order[:3]
    assert hits >= 3


def test_vartgp_cfg2_smoke_statistics_via_run_many_loops(tmp_path):
    """BASELINE config #2 at SMOKE length on one dataset: VarTGP, 3 seeds, 30 trials, 200 epochs, through
    run_many_loops (CSV written, all seeds advanced as one device batch, in-kernel Philox).  The MC stream of the
    reference is not recoverable, so the comparison with results/VarTGP_AutoAM_*_SMOKE.csv is
    statistical: same initial designs, comparable best-so-far value, substantial overlap of the visited points."""
    import pandas as pd
    from stpbenchmark_b200 import models, runners
    X, y = dataset("AutoAM")
    seeds = [45, 359, 6185]
    out = tmp_path / "VarTGP_AutoAM_dataset_traces_SMOKE.csv"
    runners.run_many_loops(X, y, model_class=models.VarTGP, seeds=seeds, n_initial=10, n_trials=30, epochs=200,
                           learning_rate=0.01, output_path=str(out), invert_y=False)
    df = pd.read_csv(out, header=[0, 1], index_col=0)
    ours_best, gold_best, overlap = [], [], []
    for s in seeds:
        idx = df[f"seed_{s}", "x_index"].astype(int).tolist()
        gold = trace("smoke", "VarTGP", "AutoAM", s)
        assert len(idx) == 40 and idx[:10] == gold[:10] and len(set(idx[10:])) == 30
        assert np.allclose(df[f"seed_{s}", "y_value"].values, y[idx].numpy())
        ours_best.append(float(y[idx].max())); gold_best.append(float(y[gold].max()))
        overlap.append(len(set(idx[10:]) & set(gold[10:])) / 30.0)
    spread = float(y.max() - y.min())
    assert np.mean(ours_best) >= np.mean(gold_best) - 0.15 * spread, (ours_best, gold_best)
    assert np.mean(overlap) >= 0.3, overlap


@pytest.mark.gpu
def test_cfg5_shape_m256_value_gradients_training_and_pool_pass():
    """BASELINE config #5 shape (VarTGP, d = 8, M = n = 256 inducing points, pool 8192): above n_cap = 144 the
    working square of a campaign moves from shared memory to its global workspace.  ELBO value and every gradient
    at n = 256, then 5 NGD + Adam epochs from a fresh model and the fused pool pass with explicit eps, all against
    the oracle."""
    from stpbenchmark_b200 import _lib, synthetic, varops
    kind, d, N, S, E = "VarTGP", 8, 8192, 8, 5
    X, y = synthetic.cfg5_pool(0)
    ns = [256, 200]
    cap = 260
    B = len(ns)
    g = torch.Generator().manual_seed(5)
    sels = [torch.randperm(N, generator=g)[:n] for n in ns]
    Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double)
    for b, sel in enumerate(sels):
        Xb[b, :ns[b]] = X[sel]; yb[b, :ns[b]] = y[sel]
    Xb, yb = Xb.cuda(), yb.cuda()
    nb = torch.tensor(ns, dtype=torch.int32, device="cuda")
    work = _lib.var_workspace(B, cap, "cuda")
    # (1) value + gradients at a perturbed state
    states = [perturbed_state(X[s], kind, seed=b + 1) for b, s in enumerate(sels)]
    dev = _dev_state([pack(st, cap) for st in states], cap, d)
    loss, grads, status = _lib.var_elbo_fg(Xb, yb, nb, dev, LIK[kind], PRIOR, GH, work)
    assert status.tolist() == [0] * B
    for b, (n, sel, st) in enumerate(zip(ns, sels, states)):
        lo, ref = ov.epoch(st, X[sel], ov.standardise(y[sel]), lr=0.0)
        assert abs(loss[b].item() - lo) <= 1e-9 * max(1.0, abs(lo))
        check_grads(split_grads(grads[b].cpu().numpy(), n, d, cap), ref, d, 1e-7)
    # (2) a few epochs from a fresh model, then the pool pass
    noise = torch.zeros(B, cap, dtype=torch.double)
    for b, n in enumerate(ns):
        noise[b, :n] = torch.randn(n, dtype=torch.double, generator=torch.Generator().manual_seed(10 + b))
    st = varops.alloc_state(B, cap, d, "cuda")
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    hist = torch.zeros(B, E, dtype=torch.double, device="cuda")
    _lib.var_fit(Xb, yb, nb, st, LIK[kind], E, 0.01, 0.01, varops.var_hyp0(d), PRIOR, GH, status, work, fresh=True,
                 init_noise=noise.cuda(), loss_hist=hist)
    assert status.tolist() == [0] * B
    eps = torch.randn(B, N, S, dtype=torch.double, generator=torch.Generator().manual_seed(99))
    mask = torch.ones(B, N, dtype=torch.uint8)
    for b, sel in enumerate(sels):
        mask[b, sel] = 0
    best = torch.tensor([float(y[s].max()) for s in sels], dtype=torch.double)
    out = _lib.var_acq(X.cuda().unsqueeze(0).contiguous(), st, nb, LIK[kind], status, work, best_f=best.cuda(),
                       mask=mask.cuda(), S=S, eps=eps.cuda(), want_pointwise=True)
    assert status.tolist() == [0] * B
    for b, (n, sel) in enumerate(zip(ns, sels)):
        ost = ov.train_natural(X[sel], y[sel], kind, E, 0.01, init_noise=noise[b, :n])
        assert rel(hist[b].cpu().numpy(), ost.losses) <= 1e-7
        assert rel(st["nat2"][b, :n, :n].cpu().numpy(), ost.nat2.numpy()) <= 1e-5
        mo, vo = ov.latent_posterior(ost, X)
        ao = ov.acquisition(ost, X, float(y[sel].max()), eps=eps[b])
        assert rel(out["mean"][b].cpu().numpy(), mo.numpy()) <= 1e-5
        assert ((out["var"][b].cpu() - vo).abs() / vo).max() <= 1e-5
        assert ((out["acq"][b].cpu() - ao).abs() / ao.abs().clamp_min(1.0)).max() <= 1e-5
        ref = ao.clone(); ref[mask[b] == 0] = -float("inf")
        assert out["idx"][b].item() == int(torch.argmax(ref))


def test_cfg5_campaigns_through_the_batched_runner():
    """Config #5 end to end (reduced epochs / trials): 256-point Sobol initial designs on a pool of 8192, VarTGP with
    M = n = 256 ... 258 inducing points, two campaigns advanced together."""
    from stpbenchmark_b200 import synthetic
    from stpbenchmark_b200.models import VarTGP
    from stpbenchmark_b200.runners import run_campaign_batch
    X, y = synthetic.cfg5_pool(1)
    idx, val, status = run_campaign_batch(X, y, VarTGP, [5120, 5121], n_initial=256, n_trials=2, epochs=10,
                                          learning_rate=0.01, num_samples=64)
    assert status == [0, 0]
    for row, v in zip(idx, val):
        assert len(row) == 258 and len(set(row[256:])) == 2 and not (set(row[256:]) & set(row[:256]))
        assert v == [float(y[i]) for i in row]


@pytest.mark.parametrize("cls_name", ["VarGP", "VarTGP"])
def test_train_variational_model_plain_adam(cls_name):
    """optim.train_variational_model (optim.py:61-102: Adam on every parameter, the natural parameters included)
    against the oracle's restatement: loss history and the trained state after 30 epochs."""
    from stpbenchmark_b200 import models, optim
    X, y = dataset("AutoAM")
    tr = trace("gold1", "ExactGP", "AutoAM", 45)
    n, E = 16, 30
    Xt, yt = X[tr[:n]], y[tr[:n]]
    noise = torch.randn(n, dtype=torch.double, generator=torch.Generator().manual_seed(21))
    model = getattr(models, cls_name)(inducing_points=Xt).double()
    hist = optim.train_variational_model(model, Xt, yt, epochs=E, lr=0.01, verbose=True, init_noise=noise)
    ost = ov.train_adam(Xt, yt, cls_name, E, 0.01, init_noise=noise)
    assert len(hist) == E and rel(hist, ost.losses) <= 1e-7
    vd = model.variational_strategy._variational_distribution
    assert rel(vd.natural_vec.detach().cpu().numpy(), ost.nat1.numpy()) <= 1e-5
    assert rel(vd.natural_mat.detach().cpu().numpy(), ost.nat2.numpy()) <= 1e-5
    assert rel(model.variational_strategy.inducing_points.detach().cpu().numpy(), ost.Z.detach().numpy()) <= 1e-5
    assert rel(model.covar_module.base_kernel.lengthscale.detach().cpu().numpy().ravel(), ost.ls.detach().numpy().ravel()) <= 1e-6
    assert abs(float(model.likelihood.noise.detach()) - ost.noise.item()) <= 1e-6 * ost.noise.item()


@pytest.mark.parametrize("kind", ["VarGP", "VarTGP"])
def test_pool_pass_is_the_same_with_256_and_512_thread_ctas(kind):
    """stp_var_acq picks its CTA shape from the shared-memory footprint: 256 threads (three CTAs per SM) for small
    row capacities, 512 threads once only one CTA fits (n_cap > ~75).  The same campaigns at n_cap = 42 and
    n_cap = 100 -- identical data, identical fits, the in-kernel Philox stream -- give the same posterior and acquisition
    (partial sums are grouped differently: 1e-11) and the same pick."""
    from stpbenchmark_b200 import _lib, varops
    ns = [25, 40]
    out = {}
    for cap in (42, 100):
        X, y, sels, Xb, yb, nb, cap = _sets(ns, cap=cap)
        d = X.shape[1]
        st = varops.alloc_state(len(ns), cap, d, "cuda")
        status = torch.zeros(len(ns), dtype=torch.int32, device="cuda"); work = _lib.var_workspace(len(ns), cap, "cuda")
        noise = torch.zeros(len(ns), cap, dtype=torch.double)
        noise[:, :40] = torch.randn(len(ns), 40, dtype=torch.double, generator=torch.Generator().manual_seed(9))
        _lib.var_fit(Xb, yb, nb, st, LIK[kind], 25, 0.01, 0.01, varops.var_hyp0(d), PRIOR, GH, status, work, fresh=True,
                     init_noise=noise.cuda())
        best = torch.stack([y[s].max() for s in sels]).cuda()
        out[cap] = _lib.var_acq(X.cuda().unsqueeze(0).contiguous(), st, nb, LIK[kind], status, work, best_f=best, S=512, seed=31,
                                want_pointwise=True)
        assert status.tolist() == [0, 0]
    a, b = out[42], out[100]
    for key, tol in (("mean", 1e-11), ("var", 1e-10), ("acq", 1e-9)):
        ref = a[key].cpu()
        assert ((b[key].cpu() - ref).abs() <= tol * ref.abs().clamp_min(1.0)).all(), key
    assert a["idx"].tolist() == b["idx"].tolist()


def test_odd_sizes_just_above_the_shared_memory_capacity():
    """n_cap = 153 > 144 with odd n, odd d (AgNP: d = 5): the working square is global with an EVEN leading dimension
    (16-byte operand copies of the staged contractions must cope with pairs that straddle n), blocked sweeps with the
    Cholesky export on global memory, 512-thread CTAs.  ELBO value and gradients, then a short fit, against the oracle."""
    from stpbenchmark_b200 import _lib, varops
    kind, E = "VarTGP", 4
    X, y = dataset("AgNP")
    d, N = X.shape[1], len(X)
    ns, cap = [151, 147], 153
    B = len(ns)
    g = torch.Generator().manual_seed(3)
    sels = [torch.randperm(N, generator=g)[:n] for n in ns]
    Xb = torch.zeros(B, cap, d, dtype=torch.double); yb = torch.zeros(B, cap, dtype=torch.double)
    for b, sel in enumerate(sels):
        Xb[b, :ns[b]] = X[sel]; yb[b, :ns[b]] = y[sel]
    Xb, yb = Xb.cuda(), yb.cuda()
    nb = torch.tensor(ns, dtype=torch.int32, device="cuda")
    work = _lib.var_workspace(B, cap, "cuda")
    states = [perturbed_state(X[s], kind, seed=b + 1) for b, s in enumerate(sels)]
    dev = _dev_state([pack(st, cap) for st in states], cap, d)
    loss, grads, status = _lib.var_elbo_fg(Xb, yb, nb, dev, LIK[kind], PRIOR, GH, work)
    assert status.tolist() == [0] * B
    for b, (n, sel, st) in enumerate(zip(ns, sels, states)):
        lo, ref = ov.epoch(st, X[sel], ov.standardise(y[sel]), lr=0.0)
        assert abs(loss[b].item() - lo) <= 1e-9 * max(1.0, abs(lo))
        check_grads(split_grads(grads[b].cpu().numpy(), n, d, cap), ref, d, 1e-7)
    noise = torch.zeros(B, cap, dtype=torch.double)
    for b, n in enumerate(ns):
        noise[b, :n] = torch.randn(n, dtype=torch.double, generator=torch.Generator().manual_seed(20 + b))
    st = varops.alloc_state(B, cap, d, "cuda")
    status = torch.zeros(B, dtype=torch.int32, device="cuda")
    hist = torch.zeros(B, E, dtype=torch.double, device="cuda")
    _lib.var_fit(Xb, yb, nb, st, LIK[kind], E, 0.01, 0.01, varops.var_hyp0(d), PRIOR, GH, status, work, fresh=True,
                 init_noise=noise.cuda(), loss_hist=hist)
    assert status.tolist() == [0] * B
    for b, (n, sel) in enumerate(zip(ns, sels)):
        ost = ov.train_natural(X[sel], y[sel], kind, E, 0.01, init_noise=noise[b, :n])
        assert rel(hist[b].cpu().numpy(), ost.losses) <= 1e-7
        assert rel(st["nat2"][b, :n, :n].cpu().numpy(), ost.nat2.numpy()) <= 1e-5
