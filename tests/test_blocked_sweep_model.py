"""Numerical model (numpy) of the two ways to take eight sweep pivots at once, against the pivot-by-pivot sweep in
extended precision.  The device routine `sweep_bordered_blocked` (csrc/linalg.cuh) uses the factored form
(W = C X^T, V = W / Delta, A -= V W^T); the explicit-inverse form (A += (C D^-1) C^T) is kept here to document why
it was rejected: on long-lengthscale kernels it loses 4 - 6 digits of the quadratic form."""
import numpy as np
import pytest


def _seq_sweep(A, npiv):
    A = A.copy(); piv = []
    for k in range(npiv):
        p = A[k, k]; piv.append(p)
        col = A[:, k].copy(); row = A[k, :].copy()
        A -= np.outer(col, row) / p
        A[:, k] = col / p; A[k, :] = row / p; A[k, k] = -1 / p
    return A, np.array(piv)


def _blocked_inverse_form(A, npiv, nb=8):
    A = A.copy(); piv = []
    for kb in range(0, npiv, nb):
        K = np.arange(kb, min(kb + nb, npiv)); C = A[:, K].copy()
        S, pv = _seq_sweep(C[K, :].copy(), len(K)); piv += list(pv)      # S = -D^-1
        Tn = C @ S
        A = A + Tn @ C.T
        A[:, K] = -Tn; A[K, :] = -Tn.T; A[np.ix_(K, K)] = S
    return A, np.array(piv)


def _blocked_factored_form(A, npiv, nb=8):
    A = A.copy(); piv = []
    for kb in range(0, npiv, nb):
        K = np.arange(kb, min(kb + nb, npiv)); C = A[:, K].copy(); r = len(K)
        Xm = np.eye(r); Am = C[K, :].copy(); d = np.zeros(r)
        for k in range(r):                                               # phase (A) of the device routine
            p = Am[k, k]; d[k] = p
            f = Am[:, k] / p; f[:k + 1] = 0
            Am -= np.outer(f, Am[k, :]); Xm -= np.outer(f, Xm[k, :])
        piv += list(d)
        W = C @ Xm.T; V = W / d; T = V @ Xm                              # phase (B)
        A = A - V @ W.T                                                  # phase (C)
        A[:, K] = T; A[K, :] = T.T; A[np.ix_(K, K)] = -(Xm.T / d) @ Xm
    return A, np.array(piv)


def _bordered(n, ls, os_, noise, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((n, 4))
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) / ls ** 2
    A = np.zeros((n + 1, n + 1))
    A[:n, :n] = os_ * np.exp(-0.5 * d2) + noise * np.eye(n)
    r = rng.standard_normal(n) * 5
    A[n, :n] = r; A[:n, n] = r
    return A


@pytest.mark.parametrize("n", [10, 16, 37, 80])
@pytest.mark.parametrize("ls,os_", [(0.3, 50.0), (3.0, 500.0), (10.0, 500.0)])
def test_factored_block_sweep_is_as_accurate_as_the_scalar_sweep(n, ls, os_):
    A = _bordered(n, ls, os_, 1e-4, n)
    ref, pref = _seq_sweep(A.astype(np.longdouble), n)
    seq, pseq = _seq_sweep(A, n)
    blk, pblk = _blocked_factored_form(A, n)
    scale = np.abs(ref).max()
    e_seq = float(np.abs(seq - ref).max() / scale); e_blk = float(np.abs(blk - ref).max() / scale)
    assert e_blk <= 50 * max(e_seq, 1e-15)
    assert float(np.abs(pblk - pref).max()) <= 50 * max(float(np.abs(pseq - pref).max()), 1e-15 * os_)
    assert pblk.min() > 0


def test_explicit_inverse_block_sweep_loses_digits():
    A = _bordered(80, 10.0, 500.0, 1e-4, 3)
    ref, _ = _seq_sweep(A.astype(np.longdouble), 80)
    q = lambda M: float(abs(M[80, 80] - ref[80, 80]) / abs(ref[80, 80]))
    assert q(_blocked_inverse_form(A, 80)[0]) > 1e3 * q(_blocked_factored_form(A, 80)[0])
