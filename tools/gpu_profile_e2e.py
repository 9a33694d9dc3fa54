"""Where the wall time of the cfg4 job goes when it runs through the product API (runners.run_sweep): cProfile of
one job of 1024 ExactGP campaigns (8 pools x 128 seeds, 80 trials) after a small warm-up job.
usage (under gpurun): python tools/gpu_profile_e2e.py > gpurun_out/e2e_profile.txt"""
import cProfile, io, os, pstats, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stpbenchmark_b200 import models, runners, synthetic

def jobs_of(B, tmp, off):
    per = B // 8
    out = []
    for g in range(8):
        X, y = synthetic.cfg4_pool(g)
        out.append(dict(X=X, y=y, model_class=models.ExactGP, seeds=[off + 1000 * g + k for k in range(per)], n_initial=10,
                        n_trials=80, output_path=os.path.join(tmp, f"ExactGP_pool{g}_{off}_traces_.csv")))
    return out

with tempfile.TemporaryDirectory() as tmp:
    runners.run_sweep(jobs_of(64, tmp, 7))
    torch.cuda.synchronize()
    jobs = jobs_of(1024, tmp, 100000)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    runners.run_sweep(jobs)
    pr.disable()
    dt = time.perf_counter() - t0
    print(f"run_sweep(1024 campaigns x 80 trials): {dt:.3f} s  -> {1024 * 80 / dt:.0f} it/s")
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(45)
    print(s.getvalue())
