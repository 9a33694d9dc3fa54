"""ctypes binding of libstp_b200.so (C ABI: include/stp_b200.h).

The only compute path of the package.  There is NO CPU fallback: if the library is missing,
or no CUDA device is visible, every operator raises.  Arguments are torch CUDA tensors; the
raw device pointers and the current CUDA stream are handed to the library.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# STP_LIB: load another build of the same library (diagnostic variants, e.g. tools/gpu_phase_timers.py)
LIB_PATH = os.environ.get("STP_LIB") or os.path.join(_PKG, "libstp_b200.so")
_lib = None

MAX_D = 8
MAX_N = 144        # STP_MAX_N: exact-GP entry points with the campaign in shared memory
MAX_N_FIT = 142    # ... of which stp_exact_fit also keeps the optimiser state there
MAX_N_LARGE = 512  # STP_MAX_N_LARGE: stp_exact_*_large (working square in a global workspace; row f3)
MAX_N_VAR = 512    # STP_MAX_N_VAR: variational entry points (subject to the shared-memory check of the launch)

STATUS_TEXT = {
    0: "ok", 1: "matrix not positive definite", 2: "non-finite parameter or objective",
    3: "fit stopped on a non-finite objective", 4: "iteration/evaluation cap reached",
    5: "abnormal line-search termination",
}


class LbfgsbOpts(ctypes.Structure):
    """scipy.optimize.minimize(method='L-BFGS-B') defaults."""

    _fields_ = [("m", ctypes.c_int32), ("maxiter", ctypes.c_int32), ("maxfun", ctypes.c_int32),
                ("maxls", ctypes.c_int32), ("factr", ctypes.c_double), ("pgtol", ctypes.c_double)]

    @staticmethod
    def scipy_defaults() -> "LbfgsbOpts":
        eps = 2.220446049250313e-16
        return LbfgsbOpts(10, 15000, 15000, 20, 2.220446049250313e-09 / eps, 1e-5)


class ModelOpts(ctypes.Structure):
    """struct stp_model_opts: kernel / acquisition variants (None everywhere = RBF + LogEI, the reference's path)."""

    _fields_ = [("kernel", ctypes.c_int32), ("acquisition", ctypes.c_int32), ("beta", ctypes.c_double)]
    KERNELS = {"rbf": 0, "matern52": 1}
    ACQUISITIONS = {"logei": 0, "ucb": 1, "ei": 2}

    @staticmethod
    def make(kernel: str = "rbf", acquisition: str = "logei", beta: float = 0.0) -> Optional["ModelOpts"]:
        k, a = ModelOpts.KERNELS[str(kernel).lower()], ModelOpts.ACQUISITIONS[str(acquisition).lower()]
        if k == 0 and a == 0:
            return None
        return ModelOpts(k, a, float(beta))


def _m(model):
    return ctypes.byref(model) if model is not None else None


class Turnover(ctypes.Structure):
    """struct stp_turnover (continuous batching of stp_exact_bo_step); pointers are device pointers."""

    _fields_ = [("n_final", ctypes.c_int32), ("n_init", ctypes.c_int32), ("generations", ctypes.c_int32),
                ("sink_capacity", ctypes.c_int32), ("bank", ctypes.c_void_p), ("gen", ctypes.c_void_p),
                ("sink", ctypes.c_void_p), ("sink_count", ctypes.c_void_p)]


def lib():
    """Load libstp_b200.so; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m stpbenchmark_b200.build` "
                "(nvcc, sm_100a). stpbenchmark_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.stp_version.restype = ctypes.c_int
        L.stp_last_error_string.restype = ctypes.c_char_p
        L.stp_exact_large_workspace_bytes.restype = ctypes.c_size_t
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().stp_last_error_string().decode()}")


def _require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise RuntimeError("stpbenchmark_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors")


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _host(vals):
    arr = (ctypes.c_double * len(vals))(*[float(v) for v in vals])
    return arr


def _bounds(lower, upper, np_):
    lo = [(-math.inf if v is None else float(v)) for v in (lower if lower is not None else [None] * np_)]
    hi = [(math.inf if v is None else float(v)) for v in (upper if upper is not None else [None] * np_)]
    return _host(lo), _host(hi)


def exact_large_work(B: int, n_cap: int, device) -> torch.Tensor:
    """Scratch for the stp_exact_*_large entry points: the working squares of B campaigns of capacity n_cap."""
    nbytes = lib().stp_exact_large_workspace_bytes(int(B), int(n_cap))
    if nbytes == 0 and B > 0:
        raise ValueError(f"n_cap = {n_cap} is beyond STP_MAX_N_LARGE = {MAX_N_LARGE}")
    return torch.empty(max(nbytes // 8, 1), dtype=torch.double, device=device)


def exact_mll_fg(X, y, theta, prior, n=None, model: Optional[ModelOpts] = None, large: Optional[bool] = None):
    """f [B], g [B, d+3], status [B] for X [B, n_cap, d], y [B, n_cap], theta [B, d+3].  ``large``: run with the working
    square in a global workspace (default: when n_cap > MAX_N; the model-level entry points only)."""
    _require_cuda(X, y, theta, n)
    B, n_cap, d = X.shape
    if (n_cap > MAX_N) if large is None else large:
        f = torch.empty(B, dtype=torch.double, device=X.device)
        g = torch.empty(B, d + 3, dtype=torch.double, device=X.device)
        status = torch.zeros(B, dtype=torch.int32, device=X.device)
        work = exact_large_work(B, n_cap, X.device)
        _check(lib().stp_exact_mll_fg_large(B, n_cap, d, _p(X), _p(y), _p(n), _p(theta), _host(prior), _p(f), _p(g),
                                            _p(status), _m(model), _p(work), _stream()), "stp_exact_mll_fg_large")
        return f, g, status
    f = torch.empty(B, dtype=torch.double, device=X.device)
    g = torch.empty(B, d + 3, dtype=torch.double, device=X.device)
    status = torch.zeros(B, dtype=torch.int32, device=X.device)
    _check(lib().stp_exact_mll_fg(B, n_cap, d, _p(X), _p(y), _p(n), _p(theta), _host(prior), _p(f), _p(g),
                                  _p(status), _m(model), _stream()), "stp_exact_mll_fg")
    return f, g, status


def exact_fit(X, y, theta, lower, upper, prior, n=None, opts: Optional[LbfgsbOpts] = None,
              model: Optional[ModelOpts] = None, large: Optional[bool] = None):
    """In-place device L-BFGS-B fit of theta [B, d+3]; returns (f, nit, nfev, status).  ``large`` as in exact_mll_fg
    (default: when n_cap > MAX_N_FIT)."""
    _require_cuda(X, y, theta, n)
    B, n_cap, d = X.shape
    f = torch.empty(B, dtype=torch.double, device=X.device)
    nit = torch.zeros(B, dtype=torch.int32, device=X.device)
    nfev = torch.zeros(B, dtype=torch.int32, device=X.device)
    status = torch.zeros(B, dtype=torch.int32, device=X.device)
    lo, hi = _bounds(lower, upper, d + 3)
    o = opts or LbfgsbOpts.scipy_defaults()
    if (n_cap > MAX_N_FIT) if large is None else large:
        work = exact_large_work(B, n_cap, X.device)
        _check(lib().stp_exact_fit_large(B, n_cap, d, _p(X), _p(y), _p(n), _p(theta), lo, hi, _host(prior), ctypes.byref(o),
                                         _p(f), _p(nit), _p(nfev), _p(status), _m(model), _p(work), _stream()),
               "stp_exact_fit_large")
        return f, nit, nfev, status
    _check(lib().stp_exact_fit(B, n_cap, d, _p(X), _p(y), _p(n), _p(theta), lo, hi, _host(prior), ctypes.byref(o),
                               _p(f), _p(nit), _p(nfev), _p(status), _m(model), _stream()), "stp_exact_fit")
    return f, nit, nfev, status


def exact_acq(pool, X, y, theta, best_f=None, mask=None, pool_id=None, n=None, want_pointwise=False,
              model: Optional[ModelOpts] = None, large: Optional[bool] = None):
    """Fused predictive + LogEI + argmax.  pool [G, N, d].  Returns dict(idx, acq_best, status
    [, mean, var, acq]).  ``large`` as in exact_mll_fg."""
    _require_cuda(pool, X, y, theta, best_f, mask, pool_id, n)
    B, n_cap, d = X.shape
    N = pool.shape[-2]
    dev = X.device
    idx = torch.full((B,), -1, dtype=torch.int32, device=dev)
    best = torch.empty(B, dtype=torch.double, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    mean = var = acq = None
    if want_pointwise:
        mean = torch.empty(B, N, dtype=torch.double, device=dev)
        var = torch.empty(B, N, dtype=torch.double, device=dev)
        acq = torch.empty(B, N, dtype=torch.double, device=dev)
    if (n_cap > MAX_N) if large is None else large:
        work = exact_large_work(B, n_cap, dev)
        _check(lib().stp_exact_acq_large(B, n_cap, d, N, _p(pool), _p(pool_id), _p(mask), _p(X), _p(y), _p(n), _p(theta),
                                         _p(best_f), _p(idx), _p(best), _p(mean), _p(var), _p(acq), _p(status), _m(model),
                                         _p(work), _stream()), "stp_exact_acq_large")
    else:
        _check(lib().stp_exact_acq(B, n_cap, d, N, _p(pool), _p(pool_id), _p(mask), _p(X), _p(y), _p(n), _p(theta),
                                   _p(best_f), _p(idx), _p(best), _p(mean), _p(var), _p(acq), _p(status), _m(model),
                                   _stream()), "stp_exact_acq")
    return {"idx": idx, "acq_best": best, "status": status, "mean": mean, "var": var, "acq": acq}


def exact_bo_step(pool, pool_y, pool_id, mask, X, y, n, trace, theta0, lower, upper, prior, status,
                  theta_out=None, acq_out=None, nit=None, nfev=None, opts: Optional[LbfgsbOpts] = None, order=None,
                  turnover: Optional[Turnover] = None, work_acc=None, iters_acc=None, n_max: int = 0,
                  model: Optional[ModelOpts] = None):
    """One BO iteration for every campaign, in place (see stp_exact_bo_step).  ``n_max``: a host-known upper bound
    of ``n`` over the running campaigns (0 = unknown); sizes the launch's shared memory / thread count."""
    _require_cuda(pool, pool_y, pool_id, mask, X, y, n, trace, status, theta_out, acq_out, nit, nfev, order, work_acc,
                  iters_acc)
    B, n_cap, d = X.shape
    N = pool.shape[-2]
    lo, hi = _bounds(lower, upper, d + 3)
    o = opts or LbfgsbOpts.scipy_defaults()
    _check(lib().stp_exact_bo_step(B, n_cap, d, N, _p(pool), _p(pool_y), _p(pool_id), _p(mask), _p(X), _p(y), _p(n),
                                   _p(trace), _host(theta0), lo, hi, _host(prior), ctypes.byref(o), _p(theta_out),
                                   _p(acq_out), _p(nit), _p(nfev), _p(status), _p(order),
                                   ctypes.byref(turnover) if turnover is not None else None, _p(work_acc), _p(iters_acc),
                                   ctypes.c_int(int(n_max)), _m(model), _stream()), "stp_exact_bo_step")


def exact_bo_step_large(pool, pool_y, pool_id, mask, X, y, n, trace, theta0, lower, upper, prior, status, work,
                        theta_out=None, acq_out=None, nit=None, nfev=None, opts: Optional[LbfgsbOpts] = None, order=None,
                        work_acc=None, iters_acc=None, n_max: int = 0, model: Optional[ModelOpts] = None):
    """exact_bo_step for row capacities beyond MAX_N (stp_exact_bo_step_large); ``work`` = exact_large_work(B, n_cap)."""
    _require_cuda(pool, pool_y, pool_id, mask, X, y, n, trace, status, work, theta_out, acq_out, nit, nfev, order, work_acc,
                  iters_acc)
    B, n_cap, d = X.shape
    N = pool.shape[-2]
    lo, hi = _bounds(lower, upper, d + 3)
    o = opts or LbfgsbOpts.scipy_defaults()
    _check(lib().stp_exact_bo_step_large(B, n_cap, d, N, _p(pool), _p(pool_y), _p(pool_id), _p(mask), _p(X), _p(y), _p(n),
                                         _p(trace), _host(theta0), lo, hi, _host(prior), ctypes.byref(o), _p(theta_out),
                                         _p(acq_out), _p(nit), _p(nfev), _p(status), _p(order), None, _p(work_acc),
                                         _p(iters_acc), ctypes.c_int(int(n_max)), _m(model), _p(work), _stream()),
           "stp_exact_bo_step_large")


def exact_bo_run(pool, pool_y, pool_id, mask, X, y, n, trace, theta0, lower, upper, prior, status, iters_left, sched,
                 theta_out=None, acq_out=None, nit=None, nfev=None, opts: Optional[LbfgsbOpts] = None, order=None,
                 turnover: Optional[Turnover] = None, work_acc=None, iters_acc=None, n_max: int = 0, max_ctas: int = 0,
                 model: Optional[ModelOpts] = None):
    """Persistent work-queue form of exact_bo_step: every campaign b is advanced ``iters_left[b]`` iterations in ONE
    launch (see stp_exact_bo_run).  ``sched``: int32 [B + 1] scratch, zeroed once by the caller."""
    _require_cuda(pool, pool_y, pool_id, mask, X, y, n, trace, status, iters_left, sched, theta_out, acq_out, nit, nfev,
                  order, work_acc, iters_acc)
    B, n_cap, d = X.shape
    N = pool.shape[-2]
    if sched.numel() < B + 1 or sched.dtype != torch.int32 or iters_left.dtype != torch.int32:
        raise ValueError("sched must be int32 [B + 1] and iters_left int32 [B]")
    lo, hi = _bounds(lower, upper, d + 3)
    o = opts or LbfgsbOpts.scipy_defaults()
    _check(lib().stp_exact_bo_run(B, n_cap, d, N, _p(pool), _p(pool_y), _p(pool_id), _p(mask), _p(X), _p(y), _p(n),
                                  _p(trace), _host(theta0), lo, hi, _host(prior), ctypes.byref(o), _p(theta_out),
                                  _p(acq_out), _p(nit), _p(nfev), _p(status), _p(order),
                                  ctypes.byref(turnover) if turnover is not None else None, _p(work_acc), _p(iters_acc),
                                  ctypes.c_int(int(n_max)), _p(iters_left), _p(sched), ctypes.c_int(int(max_ctas)),
                                  _m(model), _stream()), "stp_exact_bo_run")


def design_snap(points: torch.Tensor, eligible: torch.Tensor) -> torch.Tensor:
    """int32 [Q]: index (into ``eligible`` [Ne, d]) of the first nearest eligible point of every row of ``points``
    [Q, d] (stp_design_snap; the nearest-point step of utils.get_initial_samples_sobol)."""
    points = points.double().contiguous(); eligible = eligible.double().contiguous()
    _require_cuda(points, eligible)
    out = torch.empty(points.shape[0], dtype=torch.int32, device=points.device)
    _check(lib().stp_design_snap(points.shape[0], eligible.shape[0], points.shape[1], _p(points), _p(eligible), _p(out),
                                 _stream()), "stp_design_snap")
    return out


def log_ei(mean: torch.Tensor, var: torch.Tensor, best_f: float) -> torch.Tensor:
    """Elementwise analytic LogEI on the device; output has the shape of ``mean``."""
    _require_cuda(mean, var)
    if mean.shape != var.shape:
        raise ValueError("mean and var must have the same shape")
    out = torch.empty_like(mean)
    _check(lib().stp_log_ei(ctypes.c_longlong(mean.numel()), _p(mean), _p(var), ctypes.c_double(best_f), _p(out),
                            _stream()), "stp_log_ei")
    return out


PHASE_NAMES = {0: "load", 1: "gram", 2: "sweep_A", 3: "sweep_B", 4: "sweep_C", 5: "gradient", 6: "reduce", 7: "advance",
               8: "factorise", 9: "pool_pass", 10: "update", 11: "adv_build_B", 12: "adv_cauchy", 13: "adv_subsm",
               14: "adv_ls_start", 15: "adv_ls_step", 19: "adv_entry", 20: "adv_copy_dot", 21: "adv_dcsrch", 22: "adv_projgr_tests", 16: "closures", 17: "sweep_blocks", 18: "campaigns"}


# the same counters as charged by var_epoch (k_var_fit) in a -DSTP_PHASE_TIMERS build
VAR_PHASE_NAMES = {33: "setup_sweep_P", 34: "Ct_Kzz_cholinv", 35: "write_LI_LT", 36: "c3_At", 37: "c4_Tt_qf",
                   38: "quadrature_loss", 39: "c6_GS_ngd", 40: "c8_CBt", 41: "c9_LB", 42: "c10_PH", 43: "c11_U", 44: "c12_KB",
                   45: "grad_pass", 46: "adam", 48: "epochs"}


def phase_timers(reset: bool = False) -> dict:
    """Per-phase cycle counters of stp_exact_bo_step (only in a -DSTP_PHASE_TIMERS build, see stp_phase_timers)."""
    out = (ctypes.c_ulonglong * 64)()
    _check(lib().stp_phase_timers(out, int(bool(reset))), "stp_phase_timers")
    return {name: int(out[i]) for i, name in list(PHASE_NAMES.items()) + list(VAR_PHASE_NAMES.items())}


def acq_values(mean: torch.Tensor, var: torch.Tensor, best_f: float, acquisition: str = "logei", beta: float = 0.0):
    """Elementwise LogEI / UCB / EI on the device (stp_acq_values); output has the shape of ``mean``."""
    _require_cuda(mean, var)
    if mean.shape != var.shape:
        raise ValueError("mean and var must have the same shape")
    out = torch.empty_like(mean)
    m = ModelOpts(0, ModelOpts.ACQUISITIONS[acquisition.lower()], float(beta))
    _check(lib().stp_acq_values(ctypes.c_longlong(mean.numel()), _p(mean), _p(var), ctypes.c_double(best_f), ctypes.byref(m),
                                _p(out), _stream()), "stp_acq_values")
    return out


def log_ei_t(mean: torch.Tensor, sd: torch.Tensor, df: torch.Tensor, best_f: float) -> torch.Tensor:
    """Elementwise Student-t LogEI on the device (stp_log_ei_t, acqf.py:7-52)."""
    _require_cuda(mean, sd, df)
    out = torch.empty_like(mean)
    _check(lib().stp_log_ei_t(ctypes.c_longlong(mean.numel()), _p(mean), _p(sd), _p(df), ctypes.c_double(best_f), _p(out),
                              _stream()), "stp_log_ei_t")
    return out


def fp64_probe(iters: int = 4096) -> float:
    """Measured FP64 FMA throughput of the current device in TFLOP/s (roofline denominator)."""
    _require_cuda()
    sink = torch.zeros(1, dtype=torch.double, device="cuda")
    L = lib()
    _check(L.stp_fp64_probe(64, _p(sink), _stream()), "stp_fp64_probe")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(L.stp_fp64_probe(iters, _p(sink), _stream()), "stp_fp64_probe")
    e1.record()
    torch.cuda.synchronize()
    flops = 148 * 8 * 256 * float(iters) * 16 * 8 * 2
    return flops / (e0.elapsed_time(e1) * 1e-3) / 1e12


def fp64_mma_probe(iters: int = 4096) -> float:
    """Measured FP64 tensor-core (DMMA m8n8k4) throughput of the current device in TFLOP/s."""
    _require_cuda()
    sink = torch.zeros(1, dtype=torch.double, device="cuda")
    L = lib()
    _check(L.stp_fp64_mma_probe(64, _p(sink), _stream()), "stp_fp64_mma_probe")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _check(L.stp_fp64_mma_probe(iters, _p(sink), _stream()), "stp_fp64_mma_probe")
    e1.record()
    torch.cuda.synchronize()
    flops = 148 * 8 * 8 * float(iters) * 16 * 8 * 256 * 2     # CTAs x warps x iters x 16 x 8 DMMA x 256 MAC x 2
    return flops / (e0.elapsed_time(e1) * 1e-3) / 1e12


# ---------------------------------------------------------------------------------------------
# variational models
# ---------------------------------------------------------------------------------------------
LIK = {"gaussian": 0, "student_t": 1, "laplace": 2, "VarGP": 0, "VarTGP": 1, "VarLGP": 2}


def var_workspace(B: int, n_cap: int, device) -> torch.Tensor:
    L = lib()
    L.stp_var_workspace_bytes.restype = ctypes.c_size_t
    nbytes = L.stp_var_workspace_bytes(B, n_cap)
    return torch.empty(nbytes // 8, dtype=torch.double, device=device)


def var_capacity(n: int, d: int) -> int:
    """Smallest n_cap >= n every variational entry point can launch with (stp_var_capacity): n itself except in the
    shared-memory gap just below 144, where it is 145 (the global-square path)."""
    cap = int(lib().stp_var_capacity(int(n), int(d)))
    if cap < 0:
        raise ValueError(f"no variational capacity holds n = {n} training points at d = {d} "
                         f"(STP_MAX_N_VAR = {MAX_N_VAR}, 227 KB of shared memory)")
    return cap


def _u64(v: int):
    return ctypes.c_ulonglong(int(v) & 0xFFFFFFFFFFFFFFFF)


def var_fit(X, y, n, state: dict, lik: int, epochs: int, lr_ngd: float, lr_adam: float, hyp0, prior, gh, status,
            work, fresh: bool = True, t0: int = 0, init_noise=None, seed: int = 0, loss_hist=None, order=None):
    """stp_var_fit on the state dict {Z, hyp, nat1, nat2, adam_m, adam_v} (in place)."""
    _require_cuda(X, y, n, init_noise, status, work, loss_hist, order, *state.values())
    B, n_cap, d = X.shape
    _check(lib().stp_var_fit(B, n_cap, d, int(lik), int(epochs), ctypes.c_double(lr_ngd), ctypes.c_double(lr_adam),
                             int(bool(fresh)), int(t0), _p(X), _p(y), _p(n), _p(init_noise), _u64(seed), _host(hyp0),
                             _host(prior), _host(gh), _p(state["Z"]), _p(state["hyp"]), _p(state["nat1"]),
                             _p(state["nat2"]), _p(state["adam_m"]), _p(state["adam_v"]), _p(loss_hist), _p(status),
                             _p(work), _p(order), _stream()), "stp_var_fit")


def var_elbo_fg(X, y, n, state: dict, lik: int, prior, gh, work, standardise: bool = True):
    """Loss [B] and raw gradients [B, G] at the given state (nothing updated)."""
    _require_cuda(X, y, n, work, *[state[k] for k in ("Z", "hyp", "nat1", "nat2")])
    B, n_cap, d = X.shape
    G = n_cap + n_cap * n_cap + n_cap * d + d + 4
    loss = torch.empty(B, dtype=torch.double, device=X.device)
    grads = torch.zeros(B, G, dtype=torch.double, device=X.device)
    status = torch.zeros(B, dtype=torch.int32, device=X.device)
    _check(lib().stp_var_elbo_fg(B, n_cap, d, int(lik), int(bool(standardise)), _p(X), _p(y), _p(n), _p(state["Z"]),
                                 _p(state["hyp"]), _p(state["nat1"]), _p(state["nat2"]), _host(prior), _host(gh),
                                 _p(loss), _p(grads), _p(status), _p(work), _stream()), "stp_var_elbo_fg")
    return loss, grads, status


def var_acq(pool, state: dict, n, lik: int, status, work, best_f=None, mask=None, pool_id=None, S: int = 2048, eps=None,
            seed: int = 0, want_pointwise: bool = False, update=None, order=None, model: Optional[ModelOpts] = None):
    """stp_var_acq.  ``update`` = dict(pool_y, X, y, trace) performs the append of runners.py:98-109."""
    _require_cuda(pool, n, status, work, best_f, mask, pool_id, eps, *[state[k] for k in ("Z", "hyp", "nat1", "nat2")])
    B, n_cap, d = state["Z"].shape
    N = pool.shape[-2]
    dev = pool.device
    idx = torch.full((B,), -1, dtype=torch.int32, device=dev)
    best = torch.empty(B, dtype=torch.double, device=dev)
    mean = var = acq = None
    if want_pointwise:
        mean = torch.empty(B, N, dtype=torch.double, device=dev)
        var = torch.empty(B, N, dtype=torch.double, device=dev)
        acq = torch.empty(B, N, dtype=torch.double, device=dev)
    u = update or {}
    _require_cuda(*u.values())
    _check(lib().stp_var_acq(B, n_cap, d, N, int(lik), _p(pool), _p(pool_id), _p(mask), _p(state["Z"]), _p(state["hyp"]),
                             _p(state["nat1"]), _p(state["nat2"]), _p(n), _p(best_f), int(S), _p(eps), _u64(seed),
                             _p(idx), _p(best), _p(mean), _p(var), _p(acq), int(update is not None), _p(u.get("pool_y")),
                             _p(u.get("X")), _p(u.get("y")), _p(u.get("trace")), _p(status), _p(work), _p(order), _m(model),
                             _stream()), "stp_var_acq")
    return {"idx": idx, "acq_best": best, "mean": mean, "var": var, "acq": acq}


def philox_normal(seed: int, stream_id: int, counter0: int, count: int, device="cuda") -> torch.Tensor:
    """The in-kernel N(0,1) stream: ``count`` (even) draws for counters counter0 .. counter0 + count/2 - 1."""
    _require_cuda()
    pairs = (count + 1) // 2
    out = torch.empty(2 * pairs, dtype=torch.double, device=device)
    _check(lib().stp_philox_normal(_u64(seed), _u64(stream_id), _u64(counter0), ctypes.c_longlong(pairs), _p(out),
                                   _stream()), "stp_philox_normal")
    return out[:count]
