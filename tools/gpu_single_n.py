"""One stp_exact_bo_step launch of B campaigns that all hold n training points (cfg #4 pools): the unit that
tools/gpu_phase_timers.py times, as a plain program for `ncu` (profile the SECOND k_exact_bo_step launch:
  ncu --set full --clock-control none --import-source on -k regex:k_exact_bo_step -s 1 -c 1 -f -o gpurun_out/x \
      python tools/gpu_single_n.py --n 60 --B 296)."""
import argparse, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_batch(n, B, dev, cap=None):
    import torch
    from stpbenchmark_b200 import synthetic
    from stpbenchmark_b200.engine import CampaignBatch
    pools = [synthetic.cfg4_pool(g) for g in range(8)]
    PX = torch.stack([p[0] for p in pools]); PY = torch.stack([p[1] for p in pools])
    gen = torch.Generator().manual_seed(n)
    init, pid = [], []
    for b in range(B):
        g = b % 8
        d0 = synthetic.initial_designs(pools[g][0], pools[g][1], [7000 + b], 10)[0].tolist()
        rest = [i for i in torch.randperm(2048, generator=gen).tolist() if i not in set(d0)][: n - 10]
        init.append(d0 + rest); pid.append(g)
    return CampaignBatch(PX, PY, pid, init, n_cap=cap or n + 2, kind="ExactGP", device=dev)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=60)
    ap.add_argument("--B", type=int, default=296)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    warm = build_batch(args.n, args.B, dev)
    warm.step(); torch.cuda.synchronize()
    batch = build_batch(args.n, args.B, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); batch.step(); e1.record(); torch.cuda.synchronize()
    print({"n": args.n, "B": args.B, "ms": e0.elapsed_time(e1), "mean_closures": float(batch.nfev.double().mean())})
