"""stpbenchmark_b200 -- the fit -> predict -> acquire hot path of pool-based Bayesian
optimisation (standw7/STPBenchmark), rebuilt for B200: PyTorch host code over hand-written
sm_100a CUDA behind a C ABI (include/stp_b200.h, libstp_b200.so).  Same Python surface as the
reference's models.py / optim.py / acqf.py / runners.py / utils.py; no gpytorch, no botorch,
no Triton, no CPU fallback.
"""
from . import utils  # noqa: F401  (host-only; imports without CUDA)

__version__ = "0.1.0"
__all__ = ["utils", "models", "optim", "acqf", "runners", "engine", "priors"]
