"""Row f2 (initial design off the critical path): stp_design_snap against the host path of
utils.get_initial_samples_sobol (reference utils.py:239-281), through the C ABI."""
import time

import pytest
import torch

from _fixtures import DATASETS, dataset, seeds, trace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", DATASETS)
def test_device_snap_reproduces_every_golden_initial_design(name):
    from stpbenchmark_b200 import utils
    X, y = dataset(name)
    S = seeds()
    got = utils.initial_designs_batch(X, y, S, 10, 50.0, device="cuda")
    for k, s in enumerate(S):
        assert got[k].tolist() == trace("gold1", "ExactGP", name, s)[:10]


def test_device_snap_equals_the_host_snap_on_the_synthetic_pools_and_is_fast():
    """1024 + 1024 designs on BASELINE config #4 / #5 pools: device snap == vectorised host snap (== the serial
    reference path, tests/test_host_utils.py), and the whole batch takes well under the 1.6 s the serial path needs."""
    from stpbenchmark_b200 import synthetic, utils
    for pool, n0 in ((synthetic.cfg4_pool(2), 10), (synthetic.cfg5_pool(1), 256)):
        X, y = pool
        S = list(range(31000, 31000 + (1024 if n0 == 10 else 24)))
        host = utils.initial_designs_batch(X, y, S, n0, 50.0)
        utils.initial_designs_batch(X, y, S[:4], n0, 50.0, device="cuda")     # warm-up (library load)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev = utils.initial_designs_batch(X, y, S, n0, 50.0, device="cuda")
        dt = time.perf_counter() - t0
        assert torch.equal(host, dev)
        if n0 == 10:
            print(f"1024 cfg #4 designs: {dt:.3f} s")
            assert dt < 0.5
