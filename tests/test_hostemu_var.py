"""Variational device routines compiled for the host (tests/hostemu) against the autograd oracle:
ELBO value, the hand-derived backward, NGD + Adam trajectories, the pool pass."""
import ctypes

import numpy as np
import pytest
import torch

import hostemu
from _emu import P
from _fixtures import dataset, trace
from _varhelp import GH, LIK, PRIOR, check_grads, pack, perturbed_state, rel, split_grads
from oracle import variational as ov
from stpbenchmark_b200 import varops

torch.set_num_threads(1)
KINDS = ["VarGP", "VarTGP", "VarLGP"]


def _train_set(n=14, dup=True):
    X, y = dataset("AutoAM")
    sel = trace("gold1", "ExactGP", "AutoAM", 45)[:n]
    if dup:
        sel[3] = sel[1]          # duplicated inducing point, as the Sobol designs produce (App. B.4)
    return X, y, sel, X[sel].contiguous(), y[sel].contiguous()


@pytest.mark.parametrize("kind", KINDS)
def test_elbo_value_and_backward(kind):
    lib = hostemu.load()
    X, y, sel, Xt, yt = _train_set()
    n, d, cap = len(sel), X.shape[1], 16
    st = perturbed_state(Xt, kind)
    Z, hyp, nat1, nat2 = pack(st, cap)
    Xp = np.zeros((cap, d)); Xp[:n] = Xt.numpy(); yp = np.zeros(cap); yp[:n] = yt.numpy()
    loss = np.zeros(1); grads = np.zeros(cap + cap * cap + cap * d + d + 4)
    rc = lib.emu_var_elbo_fg(cap, d, LIK[kind], n, P(Xp), P(yp), 1, P(Z), P(hyp), P(nat1), P(nat2), P(np.array(PRIOR)),
                             P(np.array(GH)), P(loss), P(grads))
    lo, ref = ov.epoch(st, Xt, ov.standardise(yt), lr=0.0)
    assert rc == 0 and abs(loss[0] - lo) <= 1e-12 * max(1.0, abs(lo))
    check_grads(split_grads(grads, n, d, cap), ref, d, 1e-10)


@pytest.mark.parametrize("kind", KINDS)
def test_training_trajectory_and_pool_pass(kind):
    """40 epochs from a fresh model (L1b short-horizon parity) then the fused acquisition with explicit eps."""
    lib = hostemu.load()
    X, y, sel, Xt, yt = _train_set()
    n, d, cap, E = len(sel), X.shape[1], 16, 40
    eps0 = torch.randn(n, dtype=torch.double, generator=torch.Generator().manual_seed(3))
    st = ov.train_natural(Xt, yt, kind, E, 0.01, init_noise=eps0)
    Xp = np.zeros((cap, d)); Xp[:n] = Xt.numpy(); yp = np.zeros(cap); yp[:n] = yt.numpy()
    Z = np.zeros((cap, d)); hyp = np.zeros(d + 4); nat1 = np.zeros(cap); nat2 = np.zeros((cap, cap))
    am = np.zeros(cap * d + d + 4); av = np.zeros(cap * d + d + 4); lh = np.zeros(E)
    en = np.zeros(cap); en[:n] = eps0.numpy()
    rc = lib.emu_var_fit(cap, d, LIK[kind], n, E, ctypes.c_double(0.01), ctypes.c_double(0.01), P(Xp), P(yp), P(en),
                         P(np.array(varops.var_hyp0(d))), P(np.array(PRIOR)), P(np.array(GH)), 1, 0, P(Z), P(hyp), P(nat1),
                         P(nat2), P(am), P(av), P(lh))
    assert rc == 0
    assert rel(lh, st.losses) <= 1e-9
    assert rel(Z[:n], st.Z.detach().numpy()) <= 1e-7 and rel(nat2[:n, :n], st.nat2.numpy()) <= 1e-7
    assert rel(nat1[:n], st.nat1.numpy()) <= 1e-7 and rel(hyp[1:1 + d], st.ls.detach().numpy().ravel()) <= 1e-8
    N, S = len(X), 64
    mask = np.ones(N, dtype=np.uint8); mask[sel] = 0
    eps = torch.randn(N, S, dtype=torch.double, generator=torch.Generator().manual_seed(5))
    oi = ctypes.c_int(); oa = ctypes.c_double(); mean = np.zeros(N); var = np.zeros(N); acq = np.zeros(N)
    bf = float(yt.max())
    rc = lib.emu_var_acq(cap, d, LIK[kind], n, N, P(X.numpy()), P(mask), P(Z), P(hyp), P(nat1), P(nat2), ctypes.c_double(bf),
                         S, P(eps.numpy()), 0, 0, ctypes.byref(oi), ctypes.byref(oa), P(mean), P(var), P(acq))
    mo, vo = ov.latent_posterior(st, X)
    ao = ov.acquisition(st, X, bf, eps=eps)
    assert rc == 0 and rel(mean, mo.numpy()) <= 1e-7 and (np.abs(var - vo.numpy()) / vo.numpy()).max() <= 1e-6
    assert (np.abs(acq - ao.numpy()) / np.maximum(1, np.abs(ao.numpy()))).max() <= 1e-6
    ref = ao.clone(); ref[torch.from_numpy(mask) == 0] = -float("inf")
    assert oi.value == int(torch.argmax(ref))


def test_philox_normals_are_standard_and_reproducible():
    lib = hostemu.load()
    a = np.zeros(20000); b = np.zeros(20000); c = np.zeros(20000)
    lib.emu_philox_normal(ctypes.c_ulonglong(11), ctypes.c_ulonglong(3), ctypes.c_ulonglong(0), 10000, P(a))
    lib.emu_philox_normal(ctypes.c_ulonglong(11), ctypes.c_ulonglong(3), ctypes.c_ulonglong(0), 10000, P(b))
    lib.emu_philox_normal(ctypes.c_ulonglong(11), ctypes.c_ulonglong(4), ctypes.c_ulonglong(0), 10000, P(c))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 0.03 and abs(a.std() - 1) < 0.03 and abs((a ** 3).mean()) < 0.1 and abs((a ** 4).mean() - 3) < 0.3
    assert abs(np.corrcoef(a[::2], a[1::2])[0, 1]) < 0.03


def test_digamma():
    from scipy.special import digamma
    lib = hostemu.load()
    lib.emu_digamma.restype = ctypes.c_double
    lib.emu_digamma.argtypes = [ctypes.c_double]
    for x in (0.3, 1.0, 2.5, 3.5, 4.0, 9.99, 10.0, 57.3):
        assert abs(lib.emu_digamma(x) - digamma(x)) <= 1e-13 * max(1, abs(digamma(x)))


def test_elbo_backward_with_the_square_in_the_workspace():
    """n_cap above kVarSmemSquareMax (144): the working square lives in the campaign's global workspace instead of
    shared memory (cfg #5 shapes: d = 8, M = n up to 260).  Same routines, same parity bar."""
    from stpbenchmark_b200 import synthetic
    lib = hostemu.load()
    X, y = synthetic.cfg5_pool(0, N=400)
    n, d, cap, kind = 150, 8, 152, "VarTGP"
    Xt, yt = X[:n].contiguous(), y[:n].contiguous()
    st = perturbed_state(Xt, kind)
    Z, hyp, nat1, nat2 = pack(st, cap)
    Xp = np.zeros((cap, d)); Xp[:n] = Xt.numpy(); yp = np.zeros(cap); yp[:n] = yt.numpy()
    loss = np.zeros(1); grads = np.zeros(cap + cap * cap + cap * d + d + 4)
    rc = lib.emu_var_elbo_fg(cap, d, LIK[kind], n, P(Xp), P(yp), 1, P(Z), P(hyp), P(nat1), P(nat2), P(np.array(PRIOR)),
                             P(np.array(GH)), P(loss), P(grads))
    lo, ref = ov.epoch(st, Xt, ov.standardise(yt), lr=0.0)
    assert rc == 0 and abs(loss[0] - lo) <= 1e-11 * max(1.0, abs(lo))
    check_grads(split_grads(grads, n, d, cap), ref, d, 1e-8)
