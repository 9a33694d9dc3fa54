"""The C-ABI library loads on a CPU-only box and exports every symbol include/stp_b200.h declares;
the Python binding refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "stp_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|size_t|const char\*)\s+(stp_\w+)\s*\(", text, flags=re.M)))


@pytest.fixture(scope="module")
def lib():
    from stpbenchmark_b200 import build
    return ctypes.CDLL(build.build())


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert {"stp_exact_mll_fg", "stp_exact_fit", "stp_exact_acq", "stp_exact_bo_step", "stp_var_fit", "stp_var_acq",
            "stp_log_ei", "stp_version", "stp_last_error_string"} <= set(names)
    for n in names:
        assert getattr(lib, n) is not None


def test_version_and_argument_rejection(lib):
    assert lib.stp_version() == 3
    lib.stp_last_error_string.restype = ctypes.c_char_p
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.stp_exact_mll_fg(1, 10, 99, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0 and b"d out of range" in lib.stp_last_error_string()
    rc = lib.stp_exact_mll_fg(1, 100000, 3, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0 and b"n_cap out of range" in lib.stp_last_error_string()
    rc = lib.stp_exact_fit(1, 10, 3, None, None, None, None, None, None, None, None, None, None, None, None, None, None)
    assert rc < 0 and b"null argument" in lib.stp_last_error_string()
    assert lib.stp_exact_mll_fg(0, 10, 3, ctypes.c_void_p(8), ctypes.c_void_p(8), None, ctypes.c_void_p(8),
                                (ctypes.c_double * 6)(), ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), None, None) == 0


def test_library_is_device_code_only():
    """The shipped .so has no host implementation of the numerics: only kernels + launchers."""
    from stpbenchmark_b200 import build
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from stpbenchmark_b200 import _lib, models, optim
    from stpbenchmark_b200.engine import CampaignBatch
    X = torch.rand(12, 3, dtype=torch.double); y = torch.rand(12, dtype=torch.double)
    with pytest.raises(RuntimeError, match="CUDA"):
        _lib.exact_mll_fg(X[None], y[None], torch.zeros(1, 6, dtype=torch.double), [0.0] * 6)
    with pytest.raises(RuntimeError, match="CUDA"):
        CampaignBatch(X, y, [0], [[0, 1, 2]], n_cap=8)
    m = models.ExactGP(X, y).double()
    with pytest.raises(RuntimeError, match="CUDA"):
        optim.train_exact_model_botorch(m, X, y)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.posterior(X)


def test_struct_layouts_match_the_header():
    """ctypes mirrors of the two structs that cross the ABI: stp_lbfgsb_opts (4 x int32 + 2 x double) and
    stp_turnover (4 x int32 + 4 pointers)."""
    import ctypes as C
    from stpbenchmark_b200 import _lib
    assert C.sizeof(_lib.LbfgsbOpts) == 32 and _lib.LbfgsbOpts.factr.offset == 16
    assert C.sizeof(_lib.Turnover) == 48 and _lib.Turnover.bank.offset == 16 and _lib.Turnover.sink_count.offset == 40
    assert C.sizeof(_lib.ModelOpts) == 16 and _lib.ModelOpts.beta.offset == 8
    assert _lib.ModelOpts.make() is None and _lib.ModelOpts.make("matern52", "ucb", 4.0).acquisition == 1
    o = _lib.LbfgsbOpts.scipy_defaults()
    assert (o.m, o.maxiter, o.maxfun, o.maxls) == (10, 15000, 15000, 20) and abs(o.factr - 1e7) < 1e-6 and o.pgtol == 1e-5
