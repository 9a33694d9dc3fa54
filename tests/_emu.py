"""ctypes helpers around the host emulation (tests/hostemu) -- test scaffolding."""
import ctypes

import numpy as np

import hostemu

FACTR = 2.220446049250313e-09 / np.finfo(float).eps


def P(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def exact_mll_fg(X, y, theta, prior, n=None):
    lib = hostemu.load()
    X = np.ascontiguousarray(X, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    B, cap, d = X.shape
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    f = np.zeros(B); g = np.zeros((B, d + 3)); st = np.zeros(B, dtype=np.int32)
    nn = None if n is None else np.ascontiguousarray(n, dtype=np.int32)
    lib.emu_exact_mll_fg(B, cap, d, P(X), P(y), P(nn), P(theta), P(np.asarray(prior, dtype=np.float64)), P(f), P(g), P(st))
    return f, g, st


def exact_fit(X, y, theta0, lower, upper, prior, n=None, m=10, pgtol=1e-5, maxiter=15000, maxfun=15000, maxls=20):
    lib = hostemu.load()
    X = np.ascontiguousarray(X, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    B, cap, d = X.shape
    theta = np.ascontiguousarray(np.broadcast_to(np.asarray(theta0, dtype=np.float64), (B, d + 3))).copy()
    lb = np.array([-np.inf if v is None else v for v in lower], dtype=np.float64)
    ub = np.array([np.inf if v is None else v for v in upper], dtype=np.float64)
    f = np.zeros(B); nit = np.zeros(B, dtype=np.int32); nfev = np.zeros(B, dtype=np.int32); st = np.zeros(B, dtype=np.int32)
    nn = None if n is None else np.ascontiguousarray(n, dtype=np.int32)
    lib.emu_exact_fit(B, cap, d, P(X), P(y), P(nn), P(theta), P(lb), P(ub), P(np.asarray(prior, dtype=np.float64)), m,
                      ctypes.c_double(FACTR), ctypes.c_double(pgtol), maxiter, maxfun, maxls, P(f), P(nit), P(nfev), P(st))
    return theta, f, nit, nfev, st


def exact_acq(pool, X, y, theta, best_f=None, mask=None, n=None):
    lib = hostemu.load()
    pool = np.ascontiguousarray(pool, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    B, cap, d = X.shape
    N = pool.shape[-2]
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    idx = np.zeros(B, dtype=np.int32); best = np.zeros(B)
    mean = np.zeros((B, N)); var = np.zeros((B, N)); acq = np.zeros((B, N)); st = np.zeros(B, dtype=np.int32)
    bf = None if best_f is None else np.ascontiguousarray(best_f, dtype=np.float64)
    mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    nn = None if n is None else np.ascontiguousarray(n, dtype=np.int32)
    lib.emu_exact_acq(B, cap, d, N, P(pool), None, P(mk), P(X), P(y), P(nn), P(theta), P(bf), P(idx), P(best), P(mean),
                      P(var), P(acq), P(st))
    return dict(idx=idx, acq_best=best, mean=mean, var=var, acq=acq, status=st)


def lbfgsb_minimize(fun, x0, lower, upper, m=10, pgtol=1e-5, maxiter=15000, maxfun=15000, maxls=20):
    """Ask/tell loop over the device optimiser's state machine with a Python objective."""
    lib = hostemu.load()
    x0 = np.asarray(x0, dtype=np.float64); n = len(x0)
    lb = np.array([-np.inf if v is None else v for v in lower], dtype=np.float64)
    ub = np.array([np.inf if v is None else v for v in upper], dtype=np.float64)
    h = ctypes.c_void_p(lib.emu_lbfgsb_new(n, P(x0), P(lb), P(ub), m, ctypes.c_double(FACTR), ctypes.c_double(pgtol),
                                           maxiter, maxfun, maxls))
    x = np.zeros(n); xs = []
    while True:
        lib.emu_lbfgsb_x(h, P(x)); xs.append(x.copy())
        f, g = fun(x.copy())
        if not lib.emu_lbfgsb_advance(h, ctypes.c_double(f), P(np.ascontiguousarray(g, dtype=np.float64))):
            break
    lib.emu_lbfgsb_x(h, P(x))
    f = ctypes.c_double(); nit = ctypes.c_int(); nfev = ctypes.c_int(); st = ctypes.c_int()
    lib.emu_lbfgsb_result(h, ctypes.byref(f), ctypes.byref(nit), ctypes.byref(nfev), ctypes.byref(st))
    lib.emu_lbfgsb_free(h)
    return dict(x=x.copy(), fun=f.value, nit=nit.value, nfev=nfev.value, status=st.value, xs=xs)
