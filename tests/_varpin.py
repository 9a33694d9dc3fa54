"""Jobs that pin the VARIATIONAL oracle against the reference's own SMOKE traces (SURVEY App. C.4-5,
results/Var*_traces_SMOKE.csv committed as tests/golden/traces.npz).  Shared by tests/test_oracle_var_golden.py
and tools/validate_oracle_var_smoke.py; each job runs in its own single-thread worker process."""
import os

SMOKE = dict(n_initial=10, n_trials=30, epochs=200, learning_rate=0.01)   # run_benchmark.py:36-40 with SMOKE_TEST

# the five VarGP SMOKE campaigns the restated oracle reproduces index-exactly over all 40 rows
VARGP_KAT = [("AutoAM", 45), ("AutoAM", 6185), ("P3HT", 45), ("AgNP", 6185), ("Perovskite", 359)]
# teacher-forced MC models (App. C.5): (kind, dataset, seed, min rank-first of 30, worst admissible rank)
MC_TEACHER = [("VarTGP", "AutoAM", 45, 18, 2), ("VarLGP", "AgNP", 45, 26, 2)]


def _single_thread():
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)
    import _fixtures
    _fixtures._cache.clear()      # never share an open .npz handle with the parent process


def vargp_free_running(job):
    """oracle.loop.run_single_loop('VarGP', SMOKE settings) -> (name, seed, selected indices)."""
    _single_thread()
    from _fixtures import dataset
    from oracle import loop as oloop
    name, seed = job
    X, y = dataset(name)
    sel, _ = oloop.run_single_loop(X, y, "VarGP", seed=seed, **SMOKE)
    return name, seed, [int(i) for i in sel]


def mc_teacher_forced(job):
    """Teacher-forced along the golden SMOKE trace: at every step fit a fresh model (200 epochs) on the golden
    training set and rank the golden pick by the oracle's MC LogEI (eps from the global generator, seeded like the
    reference).  -> (kind, name, seed, ranks[30], gaps[30]) with rank 0 = the oracle's own argmax."""
    _single_thread()
    import torch
    from _fixtures import dataset, trace
    from oracle import loop as oloop
    from stpbenchmark_b200.utils import get_initial_samples_sobol, set_seeds
    kind, name, seed = job[:3]
    X, y = dataset(name)
    gold = trace("smoke", kind, name, seed)
    set_seeds(seed)
    init = get_initial_samples_sobol(X, y, n_samples=SMOKE["n_initial"], percentile=50).tolist()
    assert init == gold[:10], "initial design differs from the golden trace"
    ranks, gaps = [], []
    for t in range(SMOKE["n_trials"]):
        picked = gold[: 10 + t]
        mask = torch.ones(len(X), dtype=torch.bool); mask[picked] = False
        _, acq = oloop.bo_iteration(kind, X[picked], y[picked], X[mask], epochs=SMOKE["epochs"], lr=SMOKE["learning_rate"])
        pos = int((torch.where(mask)[0] == gold[10 + t]).nonzero()[0, 0])
        ranks.append(int((acq > acq[pos]).sum()))
        gaps.append(float(acq.max() - acq[pos]))
    return kind, name, seed, ranks, gaps


def run_all(vargp=VARGP_KAT, mc=MC_TEACHER, processes=None):
    """Every job in a pool of single-thread processes -> (vargp results, mc results)."""
    import multiprocessing as mp
    jobs = [("v", j) for j in vargp] + [("m", j) for j in mc]
    procs = processes or min(len(jobs), os.cpu_count() or 1)
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_dispatch, jobs, chunksize=1)
    return [r for (k, _), r in zip(jobs, res) if k == "v"], [r for (k, _), r in zip(jobs, res) if k == "m"]


def _dispatch(job):
    return vargp_free_running(job[1]) if job[0] == "v" else mc_teacher_forced(job[1])
