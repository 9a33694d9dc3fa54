"""acqf.log_expected_improvement / log_expected_improvement_t (row a10) against the outputs of the reference's
own acqf.py (which imports here: numpy / torch / scipy only), stored in tests/golden/acqf_ref.npz by
tests/golden/make_fixtures.py.  These two functions are plain torch / scipy on whatever device the posterior
lives on, so they are testable without a GPU."""
import numpy as np
import torch

from _fixtures import _npz
from stpbenchmark_b200 import acqf
from stpbenchmark_b200.models import Posterior


class _Stub:
    def __init__(self, post):
        self._p = post
        self.num_outputs = 1

    def posterior(self, X, **kw):
        return self._p


def _fx():
    z = _npz("acqf_ref.npz")
    t = lambda k: torch.tensor(z[k], dtype=torch.double)
    return z, t


def test_gaussian_log_ei_matches_reference_file():
    z, t = _fx()
    Xq = torch.zeros(64, 3, dtype=torch.double)
    got = acqf.log_expected_improvement(_Stub(Posterior(t("mean"), t("var"))), Xq, t("best_f"))
    assert np.allclose(got.numpy(), z["logei"], rtol=1e-12, atol=1e-12)
    assert got[5].item() == 0.0                       # sigma <= eps -> 0 (acqf.py:107)


def test_mc_shaped_posterior_is_averaged_over_samples():
    z, t = _fx()
    Xq = torch.zeros(64, 3, dtype=torch.double)
    post = Posterior(t("mc_mean"), t("var").expand(16, 64))
    got = acqf.log_expected_improvement(_Stub(post), Xq, t("best_f"))
    assert np.allclose(got.numpy(), z["logei_mc"], rtol=1e-12, atol=1e-12)


def test_student_t_log_ei_matches_reference_file():
    z, t = _fx()
    Xq = torch.zeros(64, 3, dtype=torch.double)
    got = acqf.log_expected_improvement_t(_Stub(Posterior(t("mean"), t("var"), df=t("df"))), Xq, t("best_f"))
    ref = z["logei_t"]
    assert got.dtype == Xq.dtype and np.array_equal(np.isinf(got.numpy()), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.allclose(got.numpy()[fin], ref[fin], rtol=1e-12, atol=1e-12)
    assert got[5].item() == -float("inf")             # sigma <= eps -> -inf (acqf.py:50)
