"""Synthetic pools of BASELINE.json configs #3-#5 (definitions: SURVEY.md section 8(d)).

Host-side data generation only (torch CPU generators, fixed draw order) so that the CUDA path
and the CPU oracle see bit-identical inputs.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .utils import get_initial_samples_sobol, set_seeds

# Hartmann-6 constants (botorch.test_functions.Hartmann(dim=6)).
_H6_ALPHA = torch.tensor([1.0, 1.2, 3.0, 3.2], dtype=torch.double)
_H6_A = torch.tensor([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8],
                      [17, 8, 0.05, 10, 0.1, 14]], dtype=torch.double)
_H6_P = 1e-4 * torch.tensor([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                             [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]],
                            dtype=torch.double)


def hartmann6(X: torch.Tensor, negate: bool = True) -> torch.Tensor:
    """f(x) = -sum_i alpha_i exp(-sum_j A_ij (x_j - P_ij)^2); ``negate`` flips the sign (test_hartmann.py:16)."""
    inner = (_H6_A * (X.double().unsqueeze(-2) - _H6_P) ** 2).sum(-1)
    f = -(_H6_ALPHA * torch.exp(-inner)).sum(-1)
    return -f if negate else f


def _student_t3(z: torch.Tensor) -> torch.Tensor:
    return z[:, 0] / torch.sqrt((z[:, 1] ** 2 + z[:, 2] ** 2 + z[:, 3] ** 2) / 3.0)


def cfg4_pool(g: int, N: int = 2048, d: int = 6) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pool g of config #4: U[0,1]^{N x 6}, negated Hartmann-6 + 0.05 t_3 noise (draw order X, then z)."""
    gen = torch.Generator().manual_seed(1234 + g)
    X = torch.rand(N, d, generator=gen, dtype=torch.double)
    z = torch.randn(N, 4, generator=gen, dtype=torch.double)
    return X, hartmann6(X) + 0.05 * _student_t3(z)


def cfg5_pool(g: int, N: int = 8192, d: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pool g of config #5: U[0,1]^{N x 8}, four Gaussian bumps + 0.05 t_3 noise."""
    gen = torch.Generator().manual_seed(5678 + g)
    X = torch.rand(N, d, generator=gen, dtype=torch.double)
    centres = 0.2 + 0.6 * torch.rand(4, d, generator=gen, dtype=torch.double)
    z = torch.randn(N, 4, generator=gen, dtype=torch.double)
    amp = torch.tensor([1.0, 0.8, 0.6, 0.4], dtype=torch.double)
    r2 = ((X.unsqueeze(1) - centres.unsqueeze(0)) ** 2).sum(-1)
    return X, (amp * torch.exp(-r2 / (2 * 0.25 ** 2))).sum(-1) + 0.05 * _student_t3(z)


def campaign_seed(g: int, k: int) -> int:
    """Seed of the k-th campaign on pool g (config #4: seed = 1000 g + k)."""
    return 1000 * g + k


def initial_designs(X: torch.Tensor, y: torch.Tensor, seeds: List[int], n_initial: int = 10) -> torch.Tensor:
    """[len(seeds), n_initial] int64: set_seeds(seed) then the Sobol initial design (runners.py:33-48)."""
    out = []
    for s in seeds:
        set_seeds(s)
        out.append(get_initial_samples_sobol(X, y, n_samples=n_initial, percentile=50))
    return torch.stack(out)
